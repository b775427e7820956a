"""Profiling aid (ONE GPU): the sharded step with R emulated ranks in one process (fvx.parallel.LocalGroup: the
step cut at its collectives, collectives = tensor copies / sums) at the per-rank sizes of the weak-scaling job,
so that `ncu --metrics gpu__time_duration.sum` can list the kernels of one rank's step - ncu must not wrap a
multi-rank command.   python scripts/sharded_emul.py [--ranks 2] [--steps 3]"""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvx import parallel, synth
from fvx.dataset.dataset import DataLoader

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
R = a.ranks
dev = torch.device("cuda:0")
U, I, B, K, d, D = 40000 * R, 100000 * R, 65536 * R, 64, 20, 2048
inter = synth.make_interactions(U, I, seed=1234)
p = argparse.Namespace(dataset="synthetic", batch_size=B, epochs=10 ** 6, sampler="device", seed=0)
data = DataLoader(p, interactions=inter)
es = []
for r in range(R):
    e = parallel.sharded_engine(R, r, U, I, K, d=d, D=D, max_batch=B, device=str(dev), use_tensor_cores=True)
    g = torch.Generator(device=dev).manual_seed(1 + r)
    e.set_features(torch.rand(e.Ic, D, device=dev, generator=g), keep_fp32=False)
    es.append(e)
lens = np.diff(inter.row_ptr)
max_runs = min(B // max(int(lens.min()), 1) + 2, int(1.3 * B / float(lens.mean())) + 1024)
ss = parallel.ShardedStep(es, parallel.LocalGroup(R), max_runs=max_runs)
batches = data.next_triple_batch(str(dev))
for _ in range(a.steps):
    ss.step(*next(batches))
torch.cuda.synchronize()
print("loss", ss.read_loss())
