"""Where does the host time of one training step go?  (next(batches) vs Engine.step)"""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvx import synth
from fvx.dataset.dataset import DataLoader
from fvx.engine import Engine
U, I, B, D = 40000, 100000, 16384, 2048
inter = synth.make_interactions(U, I, seed=1234)
data = DataLoader(argparse.Namespace(dataset="x", batch_size=B, epochs=10**6, sampler="device", seed=0), interactions=inter)
e = Engine(U, I, 64, d=20, D=D, max_batch=B, use_tensor_cores=True)
g = torch.Generator(device="cuda").manual_seed(1)
e.set_features(torch.rand(I, D, generator=g, device="cuda"), keep_fp32=False)
it = data.next_triple_batch("cuda:0")
for _ in range(5): e.step(*next(it))
torch.cuda.synchronize()
tn = ts = 0.0
t_all = time.perf_counter()
for _ in range(60):
    t0 = time.perf_counter(); b = next(it); t1 = time.perf_counter(); e.step(*b); t2 = time.perf_counter()
    tn += t1 - t0; ts += t2 - t1
t_host = time.perf_counter() - t_all
torch.cuda.synchronize()
t_tot = time.perf_counter() - t_all
print("per step: next(batches) %.1f us, Engine.step %.1f us, host loop %.1f us, incl. GPU drain %.1f us" % (tn / 60 * 1e6, ts / 60 * 1e6, t_host / 60 * 1e6, t_tot / 60 * 1e6))
b = next(it)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(60): e.step(*b)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("same batch x60: host %.1f us/step, total %.1f us/step" % ((t1 - t0) / 60 * 1e6, (t2 - t0) / 60 * 1e6))
