#!/usr/bin/env python
"""Two-stream timeline of the unique-row train step (fvx_debug_trace): where the kernels of one step
start and end on the main and the side stream.  usage: python scripts/step_trace.py [--batch B]"""
import argparse, ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fvx import _lib, synth
from fvx.dataset.dataset import DataLoader
from fvx.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--users", type=int, default=40000)
ap.add_argument("--items", type=int, default=100000)
a = ap.parse_args()
dev = torch.device("cuda", 0)
inter = synth.make_interactions(a.users, a.items, seed=1234)
p = argparse.Namespace(dataset="synthetic", batch_size=a.batch, epochs=10 ** 6, sampler="device", seed=0)
data = DataLoader(p, interactions=inter)
e = Engine(a.users, a.items, 64, d=20, D=2048, lr=1e-3, reg=1e-5, max_batch=a.batch, use_tensor_cores=True)
e.set_features(bench.make_features_device(a.items, 2048, dev), keep_fp32=False)
batches = data.next_triple_batch("cuda:0")
lib = _lib.load()
names = ["begin", "uniq_rows", "fwd", "prep_start", "prep_end", "score", "w_planes", "grad_E", "upd_start", "upd_end", "end"]
for _ in range(10):
    e.step(*next(batches))
torch.cuda.synchronize()
acc = np.zeros(len(names))
N = 20
lib.fvx_debug_trace(1)
for _ in range(N):
    for _ in range(3):
        e.step(*next(batches))          # the traced step runs behind queued work, like in the bench
    out = (C.c_float * len(names))()
    assert lib.fvx_debug_trace_read(out) == 0
    acc += np.array(list(out))
lib.fvx_debug_trace(0)
for n, v in zip(names, acc / N):
    print("%-11s %8.1f us" % (n, v))
