#!/usr/bin/env python
"""Diagnostics of the two-sweep top-k kernel on a trained model: candidates per row, flagged rows."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvx.engine import Engine
from fvx import synth
from fvx.dataset.dataset import DataLoader
ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=40000); ap.add_argument("--items", type=int, default=100000)
ap.add_argument("--train_steps", type=int, default=300); ap.add_argument("--k", type=int, default=100)
a = ap.parse_args()
dev = torch.device("cuda:0")
inter = synth.make_interactions(a.users, a.items, seed=1234)
p = argparse.Namespace(dataset="x", batch_size=65536, epochs=10 ** 6, sampler="device", seed=0)
data = DataLoader(p, interactions=inter)
e = Engine(a.users, a.items, 64, d=20, D=2048, lr=1e-3, reg=1e-5, max_batch=65536, use_tensor_cores=True)
g = torch.Generator(device=dev).manual_seed(4321)
F = torch.randn(a.items, 2048, generator=g, device=dev).clamp_(min=0) * torch.empty(a.items, 2048, device=dev).exponential_(1.0, generator=g)
F /= F.abs().max(); e.set_features(F, keep_fp32=False); del F
it = data.next_triple_batch(str(dev))
st = data.device_state(str(dev))
for n in (0, 60, a.train_steps):
    while e.steps_done() < n:
        e.step(*next(it))
    e.flush(); e.theta(refresh=True)
    ids, sc = e.score_topk(st["row_ptr"], st["col_sorted"], a.k)
    torch.cuda.synchronize()
    ws = e._ws
    cc = ws["ccount"].cpu().numpy(); fl = ws["flags"].cpu().numpy()
    nb = ws["nb"].cpu().numpy(); ea = ws["epsa"].cpu().numpy()
    ntr = np.diff(inter.row_ptr)
    bad = np.nonzero(fl)[0]
    print(json.dumps({"steps": n, "lists": int(cc.size), "cand_mean": float(cc.mean()), "cand_p50": float(np.percentile(cc, 50)),
                      "cand_p99": float(np.percentile(cc, 99)), "cand_max": int(cc.max()), "flagged": int(fl.sum()),
                      "nb_mean": float(nb.mean()), "nb_max": float(nb.max()), "nb_p99": float(np.percentile(nb, 99)),
                      "eps_mean": float(ea.mean()), "eps_max": float(ea.max()),
                      "bad_users_ntrain": ntr[bad[:10]].tolist(), "bad_eps": ea[bad[:10]].tolist(), "splits": ws["splits"]}))
