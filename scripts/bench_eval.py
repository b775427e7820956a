#!/usr/bin/env python
"""Times the full-catalog top-k sweep (fp32 CUDA-core kernel vs tcgen05 kernel)."""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvx.engine import Engine
from fvx import synth
from fvx.dataset.dataset import DataLoader

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=40000)
ap.add_argument("--items", type=int, default=100000)
ap.add_argument("--K", type=int, default=64)
ap.add_argument("--d", type=int, default=20)
ap.add_argument("--D", type=int, default=256)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--fp32_users", type=int, default=4096)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
inter = synth.make_interactions(a.users, a.items, seed=1234)
p = argparse.Namespace(dataset="x", batch_size=4096, epochs=1, sampler="device", seed=0)
data = DataLoader(p, interactions=inter)
e = Engine(a.users, a.items, a.K, d=a.d, D=a.D, max_batch=8)
if a.D:
    g = torch.Generator(device="cuda").manual_seed(1)
    F = torch.rand(a.items, a.D, generator=g, device="cuda")
    e.set_features(F)
st = data.device_state()
def timed(fn):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(a.reps):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(); fn(); t1.record(); torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    return best
e.theta()
nf = min(a.users, a.fp32_users)
ms32 = timed(lambda: e.score_topk(st["row_ptr"], st["col_sorted"], a.k, u0=0, u1=nf, tc=False))
mstc = timed(lambda: e.score_topk(st["row_ptr"], st["col_sorted"], a.k, tc=True))
flops = 2.0 * a.items * (a.K + a.d) + 2.0 * a.items
print(json.dumps({"users": a.users, "items": a.items, "k": a.k,
                  "fp32_users_per_s": nf / ms32 * 1e3, "tc_users_per_s": a.users / mstc * 1e3, "tc_ms": mstc,
                  "tc_tflops_algorithmic": a.users * flops / mstc / 1e9, "overflow_rows": e.tc_overflow_rows}))
