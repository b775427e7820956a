"""Times the projection kernels (fp32 CUDA-core vs tcgen05) on gathered rows at bench scale.
    python scripts/bench_proj.py [--items 100000] [--D 2048] [--rows 32768]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvx.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--items", type=int, default=100000)
ap.add_argument("--D", type=int, default=2048)
ap.add_argument("--d", type=int, default=20)
ap.add_argument("--rows", type=int, default=32768)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()

g = torch.Generator(device="cuda").manual_seed(1)
F = torch.rand(a.items, a.D, generator=g, device="cuda")
out = {}
for tc in (False, True):
    e = Engine(1000, a.items, 64, d=a.d, D=a.D, max_batch=a.rows // 2, use_tensor_cores=tc)
    e.set_features(F)
    rows = torch.randint(0, a.items, (a.rows,), generator=g, device="cuda", dtype=torch.int32)
    W = torch.randn(a.rows, e.de, generator=g, device="cuda")
    for name, fn in (("fwd", lambda: e.project_rows(rows)), ("bwd", lambda: e.grad_E_rows(rows, W)),
                     ("catalog", lambda: e.theta(refresh=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(a.iters):
            fn()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / a.iters
        nbytes = (a.items if name == "catalog" else a.rows) * a.D * 4.0
        out["%s_%s" % (name, "tc" if tc else "fp32")] = {"ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1)}
    del e
print(json.dumps(out))
