"""Profiling aid: per-tile clock64 stamps of CTA 0 / cluster 0 of k_step_fused (fvx_debug_fused_trace)."""
import ctypes as C
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvx import _lib, synth
from fvx.engine import Engine

U, I, K, d, D, B = 40000, 100000, 64, 20, 2048, int(os.environ.get("B", 65536))
dev = "cuda:0"
e = Engine(U, I, K, d=d, D=D, max_batch=B, use_tensor_cores=True, device=dev, fused_step=True)
g = torch.Generator(device=dev).manual_seed(1)
F = torch.rand(I, D, device=dev, generator=g)
e.set_features(F, keep_fp32=False); del F
rng = np.random.default_rng(0)
def batch():
    u = np.repeat(rng.integers(0, U, B // 6 + 1), 6)[:B]
    return [torch.as_tensor(x).to(dev, dtype=torch.int32) for x in (u, rng.integers(0, I, B), rng.integers(0, I, B))]
for _ in range(3):
    e.step(*batch())
torch.cuda.synchronize()
n = 4096 * 16
tr = torch.zeros(n, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.fvx_debug_fused_trace.argtypes = [C.c_void_p, C.c_longlong]
lib.fvx_debug_fused_trace(tr.data_ptr(), n)
e.step(*batch())
torch.cuda.synchronize()
lib.fvx_debug_fused_trace(None, 0)
t = tr.cpu().numpy().reshape(-1, 16)
t = t[t[:, 0] != 0]
names = ["prod_start", "prod_issued", "fwd_start", "fwd_issued", "bwd_start", "bwd_issued", "tfull_seen", "xb_pushed",
         "xb_seen", "scored", "w_pushed"]
t0 = t[0, 0]
print("tiles traced", len(t))
print("it " + " ".join("%11s" % n for n in names))
for i in list(range(0, 8)) + list(range(60, 66)):
    if i < len(t):
        print("%3d " % i + " ".join("%11d" % (t[i, j] - t0) for j in range(11)))
dd = np.diff(t[:, 5])
print("cycles per tile (bwd_issued deltas): mean %.0f median %.0f" % (dd.mean(), np.median(dd)))
for a, b in [(0, 1), (1, 2), (2, 3), (3, 6), (6, 7), (7, 8), (8, 9), (9, 10), (10, 4), (4, 5)]:
    x = t[5:, b] - t[5:, a]
    print("%12s -> %-12s mean %7.0f  median %7.0f" % (names[a], names[b], x.mean(), np.median(x)))
# next tile's producer start after this tile's bwd (same stage: tile i+2)
x = t[7:, 0] - t[5:-2, 5]
print("bwd_issued(i) -> prod_start(i+2): mean %.0f median %.0f" % (x.mean(), np.median(x)))
