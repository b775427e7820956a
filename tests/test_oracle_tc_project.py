"""CPU: the hi/lo bf16 operands of the tensor-core projection (oracle/tc_project.py restates fvx_split_planes' layout
and the three-pass product of fvx_project_tc.cu): how far the product can be from fp32/fp64 arithmetic."""
import numpy as np
import pytest

from oracle import tc_project as tp


def test_plane_layout_round_trip():
    rng = np.random.default_rng(0)
    F = (np.maximum(rng.standard_normal((37, 256)), 0) * rng.exponential(1.0, (37, 256))).astype(np.float32)
    P = tp.split_planes(F)
    assert P.shape == (37, 4, 2, 64) and P.dtype == np.uint16
    hi, lo = tp.planes_to_float(P)
    assert np.array_equal(hi, tp.bf16_rn(F)) and np.array_equal(lo, tp.bf16_rn(F - hi))
    # one (row, chunk) is 256 contiguous bytes: 64 x hi then 64 x lo
    raw = P.reshape(37, -1)
    assert np.array_equal(raw[5, 128:192], P[5, 1, 0]) and np.array_equal(raw[5, 192:256], P[5, 1, 1])
    # hi + lo carries 16 significand bits: relative error <= 2^-16 (2^-17 typical), zeros stay zero
    err = np.abs((hi.astype(np.float64) + lo) - F)
    assert np.all(err <= 2.0 ** -16 * np.abs(F)) and np.all((hi + lo)[F == 0] == 0)


@pytest.mark.parametrize("D,d", [(256, 20), (2048, 20), (4096, 256), (128, 5)])
def test_three_pass_product_is_fp32_class(D, d):
    rng = np.random.default_rng(D + d)
    n = 96
    F = (np.maximum(rng.standard_normal((n, D)), 0) * rng.exponential(1.0, (n, D))).astype(np.float32)
    F /= np.abs(F).max()                                                     # visual_loader_mixin.py:30
    lim = np.sqrt(6.0 / (D + d))
    E = rng.uniform(-lim, lim, (D, d + 1)).astype(np.float32)                # Glorot (VBPR.py:44-54), last column = Bp
    want = F.astype(np.float64) @ E.astype(np.float64)
    mag = np.abs(F).astype(np.float64) @ np.abs(E).astype(np.float64)
    for order in (None, np.random.default_rng(1)):
        got = tp.project3(F, E, order)
        assert np.all(np.abs(got - want) <= 3e-5 * mag + 1e-12)              # the bar of tests/test_gpu_tc.py
    # a plain fp32 matmul (what the reference computes) is no closer to the fp64 product than a few 1e-7 of the
    # same magnitude: the three-pass product is within the same class, two orders of magnitude inside the 1e-4 bar
    f32 = F @ E
    assert np.max(np.abs(got - want) / mag) <= 2e-5 and np.max(np.abs(f32 - want) / mag) <= 2e-6
