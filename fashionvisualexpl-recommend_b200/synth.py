"""Synthetic Amazon-fashion-shaped implicit-feedback data.

Shape follows the reference's data-prep scripts: 5-core filtering
(src/create_urls_amazon_like.py:73-92) and temporal leave-one-out with exactly one
validation and one test item per user (src/split_dataset.py:16-33).  Files are
written in the formats ``DataLoader`` reads (src/dataset/dataset.py:41-81,
src/config/configs.py:9-17).

Per-user interaction count: n_u = 5 + (Geometric(0.25) - 1), clipped to 200
(mean 8, minimum 5).  Items: drawn without replacement per user from a Zipf(1.0)
popularity over a seeded random permutation of the item ids.  The last two
interactions by synthetic timestamp become the validation and test item.
Features: F_raw = max(0, N(0,1)) * Exp(1) (about half zeros, like post-ReLU
pooled CNN features); the loader applies the reference's global max-abs scale.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class Interactions:
    num_users: int
    num_items: int
    row_ptr: np.ndarray       # int64 [U+1]  train CSR
    col_file: np.ndarray      # int32 [N]    train items, file (timestamp) order per user
    val: np.ndarray           # int32 [U]
    test: np.ndarray          # int32 [U]


def make_interactions(num_users: int, num_items: int, seed: int = 1234, max_len: int = 200) -> Interactions:
    rng = np.random.default_rng(seed)
    n_u = np.minimum(5 + rng.geometric(0.25, size=num_users) - 1, min(max_len, num_items // 2))
    n_u = n_u.astype(np.int64)
    total = int(n_u.sum())
    owner = np.repeat(np.arange(num_users, dtype=np.int64), n_u)
    item_of_rank = rng.permutation(num_items)
    cdf = np.cumsum(1.0 / np.arange(1, num_items + 1))
    cdf /= cdf[-1]

    def draw(n):
        return item_of_rank[np.minimum(np.searchsorted(cdf, rng.random(n)), num_items - 1)]

    items = draw(total)
    for _ in range(200):                               # redraw duplicates until none are left
        key = owner * num_items + items
        order = np.argsort(key, kind="stable")
        dup_sorted = np.zeros(total, dtype=bool)
        dup_sorted[1:] = key[order][1:] == key[order][:-1]
        dup = np.zeros(total, dtype=bool)
        dup[order] = dup_sorted
        nd = int(dup.sum())
        if nd == 0:
            break
        items[dup] = draw(nd)
    else:
        raise RuntimeError("could not de-duplicate synthetic interactions")
    # `items` is in generation order per user == synthetic timestamp order
    row_all = np.zeros(num_users + 1, dtype=np.int64)
    np.cumsum(n_u, out=row_all[1:])
    last = row_all[1:] - 1
    test = items[last].astype(np.int32)
    val = items[last - 1].astype(np.int32)
    keep = np.ones(total, dtype=bool)
    keep[last] = False
    keep[last - 1] = False
    row_ptr = np.zeros(num_users + 1, dtype=np.int64)
    np.cumsum(n_u - 2, out=row_ptr[1:])
    return Interactions(num_users, num_items, row_ptr, items[keep].astype(np.int32), val, test)


def make_features(num_items: int, dim: int, seed: int = 4321, dtype=np.float32, chunk: int = 16384):
    """Raw (un-normalised) features, generated in row chunks to bound host memory."""
    rng = np.random.default_rng(seed)
    F = np.empty((num_items, dim), dtype=dtype)
    for s in range(0, num_items, chunk):
        e = min(num_items, s + chunk)
        F[s:e] = (np.maximum(rng.standard_normal((e - s, dim)), 0.0)
                  * rng.exponential(1.0, (e - s, dim))).astype(dtype)
    return F


def write_dataset(root: str, name: str, inter: Interactions, features=None,
                  cnn_model: str = "resnet50", output_layer: str = "avg_pool") -> str:
    """Writes ``<root>/data/<name>/`` in the reference's layout and returns that path."""
    d = os.path.join(root, "data", name)
    os.makedirs(os.path.join(d, "original"), exist_ok=True)
    U, I = inter.num_users, inter.num_items
    n_train = int(inter.row_ptr[-1])
    with open(os.path.join(d, "stats_after_downloading"), "w") as f:   # lines 2,3 = Users, Items
        f.write("Statistics (after downloading images):\n")
        f.write("Lowest number of positive items per user: 5\n")
        f.write("Users: %d\nItems: %d\nInteractions: %d\n" % (U, I, n_train + 2 * U))
    owner = np.repeat(np.arange(U, dtype=np.int64), np.diff(inter.row_ptr))
    t_in_user = np.arange(n_train) - inter.row_ptr[owner]
    with open(os.path.join(d, "trainingset.tsv"), "w") as f:
        f.write("".join("%d\t%d\t%d\t1.0\n" % (u, i, t) for u, i, t in
                        zip(owner.tolist(), inter.col_file.tolist(), t_in_user.tolist())))
    lens = np.diff(inter.row_ptr)
    with open(os.path.join(d, "validationset.tsv"), "w") as f:
        f.write("".join("%d\t%d\t%d\t1.0\n" % (u, inter.val[u], lens[u]) for u in range(U)))
    with open(os.path.join(d, "testset.tsv"), "w") as f:
        f.write("".join("%d\t%d\t%d\t1.0\n" % (u, inter.test[u], lens[u] + 1) for u in range(U)))
    if features is not None:
        # the reference's extractor saves float64 (OLD_classify_extract.py:73,109)
        np.save(os.path.join(d, "original", "cnn_features_%s_%s.npy" % (cnn_model, output_layer)),
                np.asarray(features, dtype=np.float64))
    return d
