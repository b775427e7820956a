"""CPU: the arithmetic the tensor-core evaluation sweep's exactness rests on (oracle/tc_bound.py restates
fvx_eval_tc.cu's operand packing, bounds, bound encoding and selection rule): the bf16 product with the extra
eps_u * |b_i| column brackets the fp32 score (from above up to the bias residual beta0), and selecting through the
bounds returns the exact masked top-k."""
import numpy as np
import pytest
import torch

from oracle import tc_bound as tb


def test_bf16_helpers_match_torch():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(4096) * 10.0 ** rng.integers(-6, 6, 4096), [0.0, 1.0, 2.0 ** -126]]).astype(np.float32)
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(tb.bf16_rn(x), want)
    p = np.abs(x)
    up = tb.bf16_up(p)
    assert np.all(up >= p) and np.array_equal(tb.bf16_rn(up), up)            # a bf16 value, not below
    below = ((up.view(np.uint32) >> 16) - 1).astype(np.uint32) << 16      # the next bf16 down is below x
    nz = up > p
    assert np.all(below.view(np.float32)[nz] < p[nz])


def test_bound_encoding_orders_like_floats():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-20, 20, 2000),
                        [0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45]]).astype(np.float32)
    e = tb.enc(x)
    assert e.dtype == np.int32 and np.array_equal(tb.dec(e).view(np.uint32), x.view(np.uint32))
    o = np.argsort(e, kind="stable")
    assert np.all(np.diff(x[o].astype(np.float64)) >= 0)                    # signed int order = float order
    assert int(tb.enc(np.float32(-np.inf))) == int(tb.ENC_NEG_INF) and tb.dec(tb.ENC_NEG_INF) == -np.inf
    # the combination rule of item splits / ranks: MAX of the encodings = encoding of the MAX
    a, b = x[:1000], x[1000:2000]
    assert np.array_equal(tb.dec(np.maximum(tb.enc(a), tb.enc(b))), np.maximum(a, b))


def _problem(rng, U, I, kd, kind):
    a = (rng.standard_normal((U, kd)) * 0.3).astype(np.float32)
    b = (rng.standard_normal((I, kd)) * 0.3).astype(np.float32)
    bi = (rng.standard_normal(I) * 0.1).astype(np.float32)
    vb = (rng.standard_normal(I) * 0.1).astype(np.float32)
    if kind == "outliers":                      # a few heavy items and users, many orders of magnitude apart
        b[rng.integers(0, I, 5)] *= 300.0
        a[rng.integers(0, U, 3)] *= 1e3
        a[rng.integers(0, U, 3)] *= 1e-4
    elif kind == "bias":                        # scores dominated by the biases
        bi *= 1e3
        vb *= -7e2
    elif kind == "cancel":                      # near-identical items: scores differ in the last bits
        b[:] = b[0] + 1e-4 * rng.standard_normal((I, kd)).astype(np.float32)
        bi[:] = 0.25
    elif kind == "nonneg":                      # post-ReLU-like operands: no cancellation inside the dot product
        a, b = np.abs(a), np.abs(b)
    return a, b, bi, vb


@pytest.mark.parametrize("kd", [16, 84, 445])
@pytest.mark.parametrize("kind", ["plain", "outliers", "bias", "cancel", "nonneg"])
def test_bf16_sweep_brackets_the_fp32_score(kd, kind):
    rng = np.random.default_rng(kd * 7 + len(kind))
    a, b, bi, vb = _problem(rng, 48, 700, kd, kind)
    s = tb.exact_scores(a, b, (bi, vb))
    A, B, eps, nb, beta0 = tb.pack(a, b, (bi + vb).astype(np.float32))
    assert A.shape[1] == tb.kp_of(kd) and A.shape[1] % 64 == 0 and A.shape[1] >= kd + 3
    for order_rng in (None, np.random.default_rng(3)):
        s_ub = tb.mma_scores(A, B, order_rng)
        s_lb = tb.lower_bounds(s_ub, eps, nb, beta0)
        assert np.all(s_ub.astype(np.float64) + float(beta0) >= s), float(np.min(s_ub.astype(np.float64) - s))
        assert np.all(s_lb <= s), float(np.max(s_lb.astype(np.float64) - s))
    # the bracket is tight: its width is the rounding band, ~2^-7 |a| |b|, not a Cauchy-Schwarz bound on the score
    width = (s_ub - s_lb).astype(np.float64)
    na = np.linalg.norm(a.astype(np.float64), axis=1)[:, None] * np.linalg.norm(b.astype(np.float64), axis=1)[None, :]
    assert np.all(width <= 0.02 * na + 3 * float(beta0) + 1e-30)


@pytest.mark.parametrize("shards", [1, 3])
@pytest.mark.parametrize("kind,k", [("plain", 10), ("outliers", 20), ("bias", 5), ("cancel", 10), ("nonneg", 100)])
def test_selection_through_the_bounds_is_the_exact_masked_topk(kind, k, shards):
    rng = np.random.default_rng(len(kind) + k + shards)
    U, I, kd = 24, 6000, 36
    a, b, bi, vb = _problem(rng, U, I, kd, kind)
    train = [np.unique(rng.integers(0, I, int(rng.integers(0, 40)))).tolist() for _ in range(U)]
    train[3] = []                                                            # a user without train items
    ids, sc, counts = tb.topk_via_bounds(a, b, (bi, vb), train, k, shards=shards, rng=np.random.default_rng(9))
    s = tb.exact_scores(a, b, (bi, vb))
    for u in range(U):
        keep = np.setdiff1d(np.arange(I), np.asarray(train[u], dtype=np.int64))
        order = np.lexsort((keep, -s[u, keep].astype(np.float64)))[:k]
        assert ids[u].tolist() == keep[order].tolist(), (u, kind)
        assert np.array_equal(sc[u], s[u, keep[order]])
    if kind in ("plain", "nonneg") and shards == 1:
        # the lists stay within a small multiple of k + #train: what makes the sweep cheap, not what makes it exact
        # (a 2000-item shard has 63 groups for up to 50 wanted entries: its bound is loose, the MAX over shards helps)
        assert counts.max() <= 8 * (k + 40) + 64, counts.max()
