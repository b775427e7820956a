"""Oracle (TEST INFRASTRUCTURE): the BPR triple sampler.

Two generators over the same enumeration order and rejection rule
(``DataLoader.all_triple_batches``, src/dataset/dataset.py:83-114):

* ``reference_stream_triples`` - the reference's own RNG streams: CPython
  ``random.shuffle`` for the per-epoch user order (:95) and legacy NumPy
  ``np.random.randint(I)`` for the negatives (:100-103), both seeded with 0 as the
  model modules do at import (BPRMF.py:15-16).  Pure-Python loop like the
  reference; small cases only.
* ``philox_*`` - the counter-based generator of the device path: Philox4x32-10
  keyed by ``seed``, counter = (global triple index, attempt block).  The user
  permutation of an epoch is a Philox-keyed Feistel network with cycle walking.
"""
from __future__ import annotations

import random

import numpy as np

# ----------------------------------------------------------------------------------------
# reference streams


def reference_stream_triples(training_list, num_items, batch_size, epochs, seed=0,
                             py_rng=None, np_rng=None):
    """Restates dataset.py:83-114 with private, identically-seeded RNG objects."""
    py_rng = py_rng or random.Random(seed)
    np_rng = np_rng or np.random.RandomState(seed)
    num_users = len(training_list)
    n_train = sum(len(p) for p in training_list)
    actual = (n_train // batch_size) * batch_size * epochs     # :89-91
    users, pos, neg = [], [], []
    if actual == 0:
        return (np.zeros(0, np.int64),) * 3
    for _ in range(epochs):
        order = list(range(num_users))
        py_rng.shuffle(order)                                   # :94-95
        for u in order:
            uis = training_list[u]
            for i in uis:                                       # file order (:99)
                j = int(np_rng.randint(num_items))              # :100
                while j in uis:                                 # train items only (:101)
                    j = int(np_rng.randint(num_items))
                users.append(u)
                pos.append(i)
                neg.append(j)
                if len(users) == actual:                        # :109-110
                    return (np.array(users, np.int64), np.array(pos, np.int64),
                            np.array(neg, np.int64))
    return np.array(users, np.int64), np.array(pos, np.int64), np.array(neg, np.int64)


# ----------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11), vectorised over counters

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Returns the four 32-bit output words for arrays of counters (uint32 each)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


STREAM_NEG = 0      # counter word 3: negative-sampling stream
STREAM_PERM = 1     # counter word 3: epoch permutation stream
MAX_ATTEMPTS = 256  # the device kernel's bound; never reached on real data


def philox_negatives(row_ptr, col_sorted, users, num_items, seed, offset):
    """Negatives for triples with global indices offset .. offset+len(users)-1.

    Attempt ``a`` of triple ``g`` uses word ``a & 3`` of Philox(ctr=(g_lo, g_hi,
    a >> 2, STREAM_NEG), key=(seed_lo, seed_hi)); the candidate is
    ``(word * num_items) >> 32`` and is rejected while it is one of the user's
    *training* items (dataset.py:101-103).
    """
    users = np.asarray(users, dtype=np.int64)
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n = len(users)
    # (user, item) membership keys; CSR rows are sorted so the keys are sorted
    owner = np.repeat(np.arange(len(row_ptr) - 1, dtype=np.int64), np.diff(row_ptr))
    keys = owner * np.int64(num_items) + np.asarray(col_sorted, dtype=np.int64)
    g = np.arange(n, dtype=np.uint64) + np.uint64(offset)
    out = np.full(n, -1, dtype=np.int64)
    pending = np.arange(n)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for a in range(MAX_ATTEMPTS):
        if pending.size == 0:
            break
        gp = g[pending]
        w = philox4x32(gp & _MASK, gp >> np.uint64(32), np.full(gp.shape, a >> 2), STREAM_NEG,
                       k0, k1)[a & 3]
        cand = ((w.astype(np.uint64) * np.uint64(num_items)) >> np.uint64(32)).astype(np.int64)
        q = users[pending] * np.int64(num_items) + cand
        idx = np.searchsorted(keys, q)
        rej = (idx < len(keys)) & (keys[np.minimum(idx, len(keys) - 1)] == q) if len(keys) else \
            np.zeros(len(q), dtype=bool)
        out[pending] = cand                 # the kernel keeps the last candidate at the bound
        pending = pending[rej]
    return out


def feistel_user_permutation(num_users, seed, epoch):
    """Epoch user order of the device path (fvx_epoch_perm): perm[p] = user at position p.

    A 4-round Feistel network over 2h bits (h = ceil(bits(U) / 2), so 2^(2h) >= U) with the
    round function F(R, r) = word 0 of Philox(ctr=(R, r, epoch, STREAM_PERM), key=seed) masked
    to h bits, (L, R) <- (R, L ^ F(R, r)); positions that land outside [0, U) walk the cycle
    again.  A bijection of [0, U) computed independently per position - no sort."""
    bits = 1
    while bits < 31 and (1 << bits) < num_users:
        bits += 1
    h = (bits + 1) // 2
    mask = np.uint64((1 << h) - 1)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    x = np.arange(num_users, dtype=np.uint64)
    out = np.empty(num_users, dtype=np.int64)
    pending = np.arange(num_users)
    while pending.size:
        L, R = x[pending] >> np.uint64(h), x[pending] & mask
        for r in range(4):
            f = philox4x32(R, np.full(R.shape, r), np.full(R.shape, epoch), STREAM_PERM, k0, k1)[0]
            L, R = R, L ^ (f.astype(np.uint64) & mask)
        x[pending] = (L << np.uint64(h)) | R
        ok = x[pending] < num_users
        out[pending[ok]] = x[pending[ok]].astype(np.int64)
        pending = pending[~ok]
    return out


def enumerate_epoch(row_ptr, col_file, perm):
    """(user, pos) pairs of one epoch in the reference's order (dataset.py:96-99):
    users in ``perm`` order, each user's positives in file order."""
    users, pos = [], []
    for u in perm:
        a, b = row_ptr[u], row_ptr[u + 1]
        users.extend([u] * (b - a))
        pos.extend(col_file[a:b].tolist())
    return np.array(users, np.int64), np.array(pos, np.int64)
