#!/usr/bin/env python
"""Golden fixture for the per-user evaluation metrics: the REFERENCE's own ``_eval_by_user``
(/root/reference/src/recommender/Evaluator.py:82-128, imported, never copied) on a randomised BPRMF model.

    python tests/golden/make_golden_evaluator_random.py      # rewrites tests/golden/evaluator_random.npz

2 000 users x 3 000 items, K = 16.  A tenth of the catalog are exact DUPLICATES of other items (same Gi row,
same Bi: bitwise equal scores on any implementation - the '>=' tie rule of :98 and heapq.nlargest's candidate-order
tie rule :108 both bite), users hold 0, 1 or 3 validation / test items and 1-12 train items; some held-out items
are duplicates of each other or of high-scoring items.  Stored: the parameters, the lists, k, the reference's five
metrics per user and split, and a mask of the users whose metrics cannot flip under 1e-5 relative score noise
(a different but equally valid fp32 summation order): the GPU test compares those users exactly.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
import recommender.Evaluator as RefEval  # noqa: E402

U, I, K, k = 2000, 3000, 16, 10
rng = np.random.default_rng(2026)
Gu = (rng.standard_normal((U, K)) * 0.3).astype(np.float32)
Gi = (rng.standard_normal((I, K)) * 0.3).astype(np.float32)
Bi = (rng.standard_normal(I) * 0.1).astype(np.float32)
dup = rng.choice(I, I // 10, replace=False)
src = rng.choice(np.setdiff1d(np.arange(I), dup), len(dup))
Gi[dup], Bi[dup] = Gi[src], Bi[src]
scores = (Bi[None, :].astype(np.float64) + Gu.astype(np.float64) @ Gi.astype(np.float64).T).astype(np.float32)
scores[:, dup] = scores[:, src]                                  # exact ties, whatever the BLAS did

train, val, test = [], [], []
for u in range(U):
    items = rng.permutation(I)
    n_tr = int(rng.integers(1, 13))
    n_v, n_t = int(rng.choice([0, 1, 3], p=[0.1, 0.6, 0.3])), int(rng.choice([0, 1, 3], p=[0.05, 0.65, 0.3]))
    tr = sorted(items[:n_tr].tolist())
    rest = items[n_tr:]
    if u % 7 == 0:                                               # held-out items among the user's best: hits and ties at the top
        best = [int(i) for i in np.argsort(-scores[u]) if i not in set(tr)][:40]
        rest = np.array(best + [int(i) for i in rest if i not in set(best)])
    v = rest[:n_v].tolist()
    t = rest[n_v:n_v + n_t].tolist()
    if u % 11 == 0 and n_t == 3:                                 # a duplicate pair inside the held-out list
        m = np.nonzero(dup == t[0])[0]
        if len(m) and src[m[0]] not in tr and src[m[0]] not in v and src[m[0]] not in t:
            t[1] = int(src[m[0]])
    train.append(tr); val.append(v); test.append(t)


class Data:
    num_users, num_items = U, I
    training_list, validation_list, test_list = train, val, test


RefEval._dataset = Data
RefEval._K = k
RefEval._feed_dict_test = [RefEval._evaluate_input_list_test(u) for u in range(U)]
RefEval._feed_dict_validation = [RefEval._evaluate_input_list_validation(u) for u in range(U)]
out = {"v": np.full((U, 5), np.nan), "t": np.full((U, 5), np.nan)}
for u in range(U):
    for name, is_val in (("t", False), ("v", True)):
        r = RefEval._eval_by_user(u, scores[u], val=is_val)
        if r:
            out[name][u] = r

# users whose metrics are stable under 1e-5 relative noise on the scores of non-duplicate items
robust = np.ones(U, dtype=bool)
group = np.arange(I)
group[dup] = src                                                 # items of one group score identically
for u in range(U):
    s = scores[u].astype(np.float64)
    for h in val[u] + test[u]:
        near = np.abs(s - s[h]) <= 1e-5 * max(1.0, abs(s[h]))
        near &= group != group[h]
        if near.any():
            robust[u] = False
    # the top-k boundary: the k-th and (k+1)-th candidates must not be a near-tie of different groups
    cand = np.setdiff1d(np.arange(I), train[u])
    o = cand[np.argsort(-s[cand], kind="stable")]
    if len(o) > k and abs(s[o[k - 1]] - s[o[k]]) <= 1e-5 * max(1.0, abs(s[o[k]])) and group[o[k - 1]] != group[o[k]]:
        robust[u] = False


def csr(lists):
    ptr = np.zeros(U + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(x) for x in lists])
    return ptr, np.array([i for x in lists for i in x], dtype=np.int32)


tp, tc = csr(train); vp, vc = csr(val); sp, sc = csr(test)
np.savez_compressed(os.path.join(HERE, "evaluator_random.npz"), Gu=Gu, Gi=Gi, Bi=Bi, k=k, train_ptr=tp, train_col=tc,
                    val_ptr=vp, val_col=vc, test_ptr=sp, test_col=sc, metrics_v=out["v"], metrics_t=out["t"], robust=robust)
print("users with metrics: val %d test %d; robust %d of %d; mean test metrics %s"
      % ((~np.isnan(out["v"][:, 0])).sum(), (~np.isnan(out["t"][:, 0])).sum(), robust.sum(), U,
         np.nanmean(out["t"], axis=0).round(4)))
