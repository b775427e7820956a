"""Oracle (TEST INFRASTRUCTURE): the bf16 hi/lo operands of the tensor-core projection
(fashionvisualexpl-recommend_b200/csrc/fvx_project_tc.cu), restated in NumPy.

The reference computes ``theta_i = F[i] E`` and ``F[i] Bp`` in fp32 (VBPR.py:83-84) and their gradient
``dE = F[rows]^T W`` through the tape (VBPR.py:141).  The CUDA path keeps F as two bf16 planes,
``hi = bf16(F)`` and ``lo = bf16(F - hi)`` (4 bytes per element, the footprint of fp32), does the same with the
small operand (``E_ext^T`` forward, the coefficient rows W backward) and sums three bf16 products with fp32
accumulation, ``hi*hi + lo*hi + hi*lo`` (the ``lo*lo`` term, <= 2^-18 relative, is dropped):

* ``split_planes``      fp32 [n, D] -> the interleaved plane layout of ``FvxModel.F_pl`` / ``fvx_split_planes``:
                        ``[n][D/64][2][64]`` - per row and 64-feature chunk 64 x hi, then 64 x lo (256 contiguous bytes);
* ``project3``          the three-pass product; ``tests/test_oracle_tc_project.py`` bounds its distance from the fp64
                        product element by element: <= 3e-5 of sum_k |F_ik| |E_kn| - the bar the GPU tests hold
                        the kernels to (tests/test_gpu_tc.py).
"""
from __future__ import annotations

import numpy as np

from .tc_bound import bf16_rn, split_hi_lo


def split_planes(F):
    """fp32 [n, D] (D % 64 == 0) -> uint16 [n, D/64, 2, 64]: the bit patterns of the bf16 hi and lo planes."""
    F = np.asarray(F, np.float32)
    n, D = F.shape
    assert D % 64 == 0
    hi, lo = split_hi_lo(F)
    out = np.empty((n, D // 64, 2, 64), np.uint16)
    out[:, :, 0, :] = (hi.view(np.uint32) >> 16).astype(np.uint16).reshape(n, D // 64, 64)
    out[:, :, 1, :] = (lo.view(np.uint32) >> 16).astype(np.uint16).reshape(n, D // 64, 64)
    return out


def planes_to_float(P):
    """uint16 [n, D/64, 2, 64] -> (hi, lo) fp32 [n, D]."""
    n, c = P.shape[0], P.shape[1]
    f = (P.astype(np.uint32) << 16).view(np.float32)
    return f[:, :, 0, :].reshape(n, c * 64), f[:, :, 1, :].reshape(n, c * 64)


def project3(F, E, rng=None, block=64):
    """F @ E from the hi/lo planes of both operands: three bf16 products per 64-feature stage, stages accumulated in
    fp32 (in a random order with ``rng``: the K split / stream-K partials of the kernel are summed in no fixed order)."""
    Fh, Fl = split_hi_lo(F)
    Eh, El = split_hi_lo(E)
    D = F.shape[1]
    parts = []
    for k in range(0, D, block):
        s = slice(k, k + block)
        parts.append(Fh[:, s].astype(np.float64) @ Eh[s].astype(np.float64) + Fl[:, s].astype(np.float64) @ Eh[s].astype(np.float64)
                     + Fh[:, s].astype(np.float64) @ El[s].astype(np.float64))
    order = np.arange(len(parts)) if rng is None else rng.permutation(len(parts))
    out = np.zeros((F.shape[0], E.shape[1]), np.float32)
    for j in order:
        out = (out.astype(np.float64) + parts[j]).astype(np.float32)
    return out


__all__ = ["bf16_rn", "split_hi_lo", "split_planes", "planes_to_float", "project3"]
