"""Profiling aid: timeline of fvx_bpr_step_sharded (torchrun, one rank per GPU): where the pieces and the four
all-reduces of one step start and end on the two streams (fvx_debug_trace_sharded), averaged over steps.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 \
        scripts/sharded_phases.py [--config weak|c3] [--batch 65536]
"""
import argparse, ctypes as C, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvx import _lib, parallel, synth
from fvx.dataset.dataset import DataLoader

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="weak", choices=["weak", "c3"])
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
if a.config == "weak":
    U, I, B = 40000 * world, 100000 * world, a.batch * world
else:
    U, I, B = 1000000, 500000, 524288
K, d, D = 64, 20, 2048
inter = synth.make_interactions(U, I, seed=1234)
p = argparse.Namespace(dataset="synthetic", batch_size=B, epochs=10 ** 6, sampler="device", seed=0)
data = DataLoader(p, interactions=inter)
e = parallel.sharded_engine(world, rank, U, I, K, d=d, D=D, max_batch=B, device=str(dev), use_tensor_cores=True)
g = torch.Generator(device=dev).manual_seed(1 + rank)
e.set_features(torch.rand(e.Ic, D, device=dev, generator=g), keep_fp32=False)
lens = np.diff(inter.row_ptr)
max_runs = min(B // max(int(lens.min()), 1) + 2, int(1.3 * B / float(lens.mean())) + 1024)
grp = parallel.DistGroup()
ss = parallel.ShardedStep([e], grp, max_runs=max_runs)
batches = data.next_triple_batch(str(dev))
for _ in range(5):
    ss.step(*next(batches))
torch.cuda.synchronize(); dist.barrier()
lib = _lib.load()
names = ["begin", "p1 run ids+rows+claims", "p2 user catch-up+pack [side]", "all-gather WU [side]", "p3 owned list+projection",
         "p4 partial scores", "all-reduce S", "p5 grads+planes", "reduce-scatter RU [side]", "p7 scatter runs [side]",
         "p6 grad_E+pack", "all-reduce dE", "p8 update / end"]
acc = np.zeros(len(names))
lib.fvx_debug_trace_sharded(1)
out = (C.c_float * len(names))()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(a.steps):
    ss.step(*next(batches))
    assert lib.fvx_debug_trace_sharded_read(out) == 0
    acc += np.array(list(out)) / a.steps
lib.fvx_debug_trace_sharded(0)
dist.barrier()
t0.record()
for _ in range(a.steps):
    ss.step(*next(batches))
t1.record(); torch.cuda.synchronize()
if rank == 0:
    print("world %d, %d users x %d items, global batch %d, exchanged user rows %d x %d floats (%.1f MB per buffer)"
          % (world, U, I, B, max_runs, e.Su, max_runs * e.Su * 4 / 1e6))
    for n, v in zip(names, acc):
        print("%-32s %8.1f us" % (n, v))
    print("untraced: %.1f us per step" % (t0.elapsed_time(t1) / a.steps * 1e3))
grp.close()
dist.destroy_process_group()
