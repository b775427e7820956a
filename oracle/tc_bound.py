"""Oracle (TEST INFRASTRUCTURE): the arithmetic behind the tensor-core evaluation sweep's exactness claim
(fashionvisualexpl-recommend_b200/csrc/fvx_eval_tc.cu), restated in NumPy.

The reference scores every (user, item) pair in fp32 (``predict_all``, BPRMF.py:78-85 / VBPR.py:88-97) and takes
the top-k of the unmasked items (Evaluator.py:225-239).  The CUDA sweep computes the same scores from bf16
operands and claims to return the SAME top-k.  What that rests on, and what ``tests/test_oracle_tc_bound.py``
checks on the CPU:

* operands: ``A[u] = [bf16(Gu|Tu) | 1 | 1 | eps_u]``, ``B[i] = [bf16(Gi|theta) | b_hi | b_lo | nb_i]`` with
  ``eps_u = up(c * |a_u| * 1.0001)``, ``nb_i = up(|b_i| * 1.0001)`` (``up`` = smallest bf16 not below),
  ``c = 1.003 * 2^-8 + KP * 2^-21`` (k_pack_users / k_pack_items);
* the product ``s_ub = <A[u], B[i]>`` accumulated in fp32 in any order is an upper bound of the fp32 score up to the
  bias residual, ``s_ub >= s - beta0``, and ``s_ub - 2.001 * eps_u * nb_i - beta0`` a LOWER bound,
  ``beta0 = (2^-17 + KP * 2^-21) * max|bias|``;
* selection (``topk_via_bounds``): per row, the maximum of ``s_ub`` over every group of 32 consecutive items minus the
  group's margin is a lower bound of the true score of the group's best item; a value tau that at least
  ``kk = k + #train`` group entries reach is therefore reached by kk distinct items, the exact top-k lies inside
  ``{s_ub >= tau - beta0}`` (a row publishes tau - a statement about true scores that item splits and shards
  combine with MAX - and the candidates sweep subtracts the beta0 of its own items), and re-scoring that set
  exactly gives the exact answer;
* ``enc`` / ``dec``: the int32 encoding of a bound whose signed order is the float's (``FvxEvalWs.thr``: combined
  over item splits with ``atomicMax`` and over ranks with an all-reduce MAX).
"""
from __future__ import annotations

import numpy as np


# ---- bf16 ------------------------------------------------------------------------------------
def bf16_rn(x):
    """float32 -> nearest bf16 (ties to even), returned as float32."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((b + 0x7FFF + ((b >> 16) & 1)) >> 16) << 16
    return (r & 0xFFFFFFFF).astype(np.uint32).view(np.float32)


def bf16_up(x):
    """smallest bf16 >= x for x >= 0 (bf16_up in fvx_eval_tc.cu)."""
    x = np.asarray(x, dtype=np.float32)
    h = bf16_rn(x)
    bump = ((h.view(np.uint32) >> 16) + 1).astype(np.uint32) << 16
    return np.where(h < x, bump.view(np.float32), h)


def split_hi_lo(x):
    hi = bf16_rn(x)
    return hi, bf16_rn(np.asarray(x, dtype=np.float32) - hi)


# ---- constants of the kernel -----------------------------------------------------------------------
def kp_of(kd):
    return (kd + 3 + 63) // 64 * 64


def c_rel(KP):
    return np.float32(1.003 * 0.00390625) + np.float32(KP) * np.float32(4.76837158e-7)


def beta_c(KP):
    return np.float32(7.6294e-6) + np.float32(KP) * np.float32(4.76837158e-7)


# ---- scores ----------------------------------------------------------------------------------
def exact_scores(a, b, bias):
    """fvx_score_one for every (user row, item row): one fmaf chain in index order, then the bias (fp32).
    ``bias`` is the pair (item bias, visual bias) added one after the other, or one array."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    s = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    for c in range(a.shape[1]):
        # fmaf: the product is exact in float64, one rounding to float32 after the add
        s = (a[:, c:c + 1].astype(np.float64) * b[None, :, c].astype(np.float64) + s.astype(np.float64)).astype(np.float32)
    for t in (bias if isinstance(bias, tuple) else (bias,)):
        s = (s + np.asarray(t, np.float32)[None, :]).astype(np.float32)
    return s


def pack(a, b, bias_total):
    """(A, B, eps_u, nb_i, beta0): the kernel's operands as float32 arrays holding bf16 values."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    kd = a.shape[1]
    KP = kp_of(kd)
    na = np.sqrt(np.sum(a * a, axis=1, dtype=np.float32), dtype=np.float32)
    nbn = np.sqrt(np.sum(b * b, axis=1, dtype=np.float32), dtype=np.float32)
    eps = bf16_up(c_rel(KP) * na * np.float32(1.0001))
    nb = bf16_up(nbn * np.float32(1.0001))
    bh, bl = split_hi_lo(bias_total)
    A = np.zeros((a.shape[0], KP), np.float32)
    B = np.zeros((b.shape[0], KP), np.float32)
    A[:, :kd], B[:, :kd] = bf16_rn(a), bf16_rn(b)
    A[:, kd] = A[:, kd + 1] = 1.0
    A[:, kd + 2] = eps
    B[:, kd], B[:, kd + 1], B[:, kd + 2] = bh, bl, nb
    beta0 = beta_c(KP) * np.float32(np.max(np.abs(bias_total))) + np.float32(1e-30)
    return A, B, eps, nb, np.float32(beta0)


def mma_scores(A, B, rng=None, block=16):
    """<A[u], B[i]> with fp32 accumulation.  The tensor core's summation order inside a K step is unspecified:
    with ``rng`` the K steps (blocks of 16 columns, each summed exactly) are added in a random order."""
    KP = A.shape[1]
    parts = [A[:, k:k + block].astype(np.float64) @ B[:, k:k + block].astype(np.float64).T for k in range(0, KP, block)]
    order = np.arange(len(parts)) if rng is None else rng.permutation(len(parts))
    s = np.zeros((A.shape[0], B.shape[0]), np.float32)
    for j in order:
        s = (s.astype(np.float64) + parts[j]).astype(np.float32)
    return s


def lower_bounds(s_ub, eps, nb, beta0):
    return s_ub - np.float32(2.001) * eps[:, None] * nb[None, :] - beta0


# ---- the bound's int encoding ---------------------------------------------------------------------
def enc(x):
    b = np.asarray(x, np.float32).view(np.int32)
    return np.where(b >= 0, b, b ^ np.int32(0x7FFFFFFF))


def dec(e):
    e = np.asarray(e, np.int32)
    return np.where(e >= 0, e, e ^ np.int32(0x7FFFFFFF)).view(np.float32)


ENC_NEG_INF = np.int32(-2139095041)         # 0x807FFFFF


# ---- selection ------------------------------------------------------------------------------------
def row_bound(s_ub_row, eps_u, nb, beta0, kk, group=32):
    """The bound tau one row publishes for one item range: bisection over the group entries as in k_topk_tc
    (<= 16 iterations, stops when the count of entries >= tau lies in [kk, kk + kk/4]); -inf when the range has
    fewer than kk groups."""
    n = len(s_ub_row)
    ng = (n + group - 1) // group
    pad = ng * group - n
    su = np.concatenate([s_ub_row, np.full(pad, -np.inf, np.float32)]).reshape(ng, group)
    nbg = np.concatenate([nb, np.zeros(pad, np.float32)]).reshape(ng, group)
    lb = (su.max(axis=1) - np.float32(2.001) * eps_u * nbg.max(axis=1) - beta0).astype(np.float32)
    fin = lb[np.isfinite(lb)]
    if len(fin) < kk:
        return np.float32(-np.inf)
    a, b = np.float32(fin.min()), np.float32(fin.max())
    for _ in range(16):
        if not a < b:
            break
        mid = np.float32(0.5) * a + np.float32(0.5) * b
        if not (mid > a) or not (mid < b):
            break
        cnt = int(np.sum(lb >= mid))
        if cnt >= kk:
            a = mid
            if cnt <= kk + (kk >> 2):
                break
        else:
            b = mid
    return np.float32(a)


def topk_via_bounds(a, b, bias, train_lists, k, shards=1, rng=None):
    """The whole procedure for every user: bounds (per item shard, combined with MAX through the int encoding),
    candidates {s_ub >= tau - beta0}, exact re-scoring, mask, top-k.  Returns (ids [U, k], scores [U, k], candidate counts)."""
    bias_total = (np.asarray(bias[0], np.float32) + np.asarray(bias[1], np.float32)).astype(np.float32) \
        if isinstance(bias, tuple) else np.asarray(bias, np.float32)
    A, B, eps, nb, beta0 = pack(a, b, bias_total)
    s_ub = mma_scores(A, B, rng)
    s_ex = exact_scores(a, b, bias)
    U, I = s_ub.shape
    ids = np.full((U, k), -1, np.int64)
    sc = np.full((U, k), -np.inf, np.float32)
    counts = np.zeros(U, np.int64)
    cuts = [I * r // shards for r in range(shards + 1)]
    for u in range(U):
        kk = k + len(train_lists[u])
        e = ENC_NEG_INF
        for r in range(shards):
            lo, hi = cuts[r], cuts[r + 1]
            e = max(e, int(enc(row_bound(s_ub[u, lo:hi], eps[u], nb[lo:hi], beta0, kk))))
        tau = dec(np.int32(e))
        cand = np.nonzero(s_ub[u] >= np.float32(tau - beta0))[0]
        counts[u] = len(cand)
        cand = np.setdiff1d(cand, np.asarray(train_lists[u], dtype=np.int64))
        order = np.lexsort((cand, -s_ex[u, cand].astype(np.float64)))[:k]       # score descending, id ascending
        ids[u, :len(order)] = cand[order]
        sc[u, :len(order)] = s_ex[u, cand[order]]
    return ids, sc, counts
