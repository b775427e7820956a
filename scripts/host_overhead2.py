import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvx import synth, _lib
from fvx.dataset.dataset import DataLoader
U, I, B = 40000, 100000, 16384
inter = synth.make_interactions(U, I, seed=1234)
data = DataLoader(argparse.Namespace(dataset="x", batch_size=B, epochs=10**6, sampler="device", seed=0), interactions=inter)
for ep in range(3):
    t0 = time.perf_counter(); r = data.device_epoch(ep); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("device_epoch host %.1f us, +drain %.1f us" % ((t1 - t0) * 1e6, (t2 - t0) * 1e6))
st = data.device_state("cuda:0"); dv = torch.device("cuda:0")
def T(f, name):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = f(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("  %-28s host %.1f us  total %.1f us" % (name, (t1 - t0) * 1e6, (t2 - t0) * 1e6)); return out
perm = T(lambda: torch.empty(U, dtype=torch.int32, device=dv), "empty")
lens = torch.empty(U, dtype=torch.int64, device=dv)
T(lambda: _lib.call("fvx_epoch_perm", _lib.ptr(perm), _lib.ptr(lens), _lib.ptr(st["row_ptr"]), U, 0, 5, _lib.stream_ptr()), "fvx_epoch_perm")
offs = T(lambda: torch.cumsum(lens, 0), "cumsum")
N = data.num_train
u, p, n = (torch.empty(N, dtype=torch.int32, device=dv) for _ in range(3))
T(lambda: _lib.call("fvx_epoch_triples", _lib.ptr(st["row_ptr"]), _lib.ptr(st["col_file"]), _lib.ptr(st["col_sorted"]), _lib.ptr(perm), _lib.ptr(offs), U, I, 0, 0, _lib.ptr(u), _lib.ptr(p), _lib.ptr(n), _lib.stream_ptr()), "fvx_epoch_triples")
it = data.next_triple_batch("cuda:0")
ts = []
for i in range(40):
    t0 = time.perf_counter(); b = next(it); ts.append((time.perf_counter() - t0) * 1e6)
print("next(batches) us:", [round(x) for x in ts])
