"""Conditioning study: distance of fp32 / tensor-core GPU steps and of a perturbed fp64 oracle from the fp64 oracle."""
import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from oracle import bpr
from helpers import rel_err
from test_gpu_parity import _random_problem, _user_contiguous_batches, _dev, _engine

for (K, d, D, B) in [(64, 20, 256, 512), (16, 64, 128, 96), (32, 20, 2048, 1024)]:
    U, I, steps, lr, reg = 700, 900, 20, 0.001, 1e-3
    P, F, rng = _random_problem(U, I, K, d, D, seed=K + d)
    batches = _user_contiguous_batches(rng, U, I, B, steps)
    def run64(Fx):
        Q = {k: v.astype(np.float64) for k, v in P.items()}; S = bpr.init_adam(Q)
        for b in batches: bpr.train_step(Q, S, b, reg, lr, Fx)
        return Q
    P64 = run64(F.astype(np.float64))
    prng = np.random.default_rng(7)
    Ppert = run64(F.astype(np.float64) * (1 + 2.0 ** -15 * prng.standard_normal(F.shape)))
    Q32 = {k: v.copy() for k, v in P.items()}; S32 = bpr.init_adam(Q32)
    for b in batches: bpr.train_step(Q32, S32, b, reg, lr, F)
    res = {}
    for tc in (False, True):
        e = _engine(U, I, K, d=d, D=D, lr=lr, reg=reg, adam_mode="dense", max_batch=B, use_tensor_cores=tc)
        e.set_features(F); e.load_params(P)
        for b in batches: e.step(*(_dev(x) for x in b))
        res[tc] = e.params()
    print("config", (K, d, D, B))
    for k in P64:
        ref = P64[k]
        def frac(x):
            dlt = np.abs(x.reshape(ref.shape) - ref) / np.abs(ref).max()
            return "max %.2e  >1e-4: %d/%d" % (dlt.max(), int((dlt > 1e-4).sum()), dlt.size)
        print("  %-3s oracle32 %s | pert64 %s | gpu32 %s | gpuTC %s" % (k, frac(Q32[k]), frac(Ppert[k]), frac(res[False][k]), frac(res[True][k])))
