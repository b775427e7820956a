"""Diagnostics of the tensor-core top-k sweep: time of the C call, candidate-list statistics, flagged rows."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from fvx.engine import Engine
from fvx import synth, _lib
from fvx._lib import call, ptr, stream_ptr
from fvx.dataset.dataset import DataLoader
ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=40000); ap.add_argument("--items", type=int, default=100000)
ap.add_argument("--k", type=int, default=100); ap.add_argument("--D", type=int, default=256)
ap.add_argument("--norm", type=int, default=1); ap.add_argument("--train_steps", type=int, default=0)
a = ap.parse_args()
inter = synth.make_interactions(a.users, a.items, seed=1234)
data = DataLoader(argparse.Namespace(dataset="x", batch_size=16384, epochs=100, sampler="device", seed=0), interactions=inter)
e = Engine(a.users, a.items, 64, d=20, D=a.D, max_batch=16384, use_tensor_cores=True)
g = torch.Generator(device="cuda").manual_seed(1)
F = torch.randn(a.items, a.D, generator=g, device="cuda").clamp_(min=0) * torch.empty(a.items, a.D, device="cuda").exponential_(1.0, generator=g)
if a.norm: F /= F.abs().max()
e.set_features(F)
if a.train_steps:
    it = data.next_triple_batch("cuda:0")
    for _ in range(a.train_steps): e.step(*next(it))
st = data.device_state()
n = a.users
ws = e._eval_ws(n)
ids = torch.empty(n, a.k, dtype=torch.int32, device="cuda"); sc = torch.empty(n, a.k, device="cuda")
e.flush(); th = e.theta()
def run():
    call("fvx_score_topk_tc", C.byref(e.struct()), ptr(th), 0, n, ptr(st["row_ptr"]), ptr(st["col_sorted"]), a.k,
         ptr(ids), ptr(sc), C.byref(ws["struct"]), stream_ptr())
run(); torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); run(); t1.record(); torch.cuda.synchronize()
fl = ws["flags"][:n].cpu().numpy(); cc = ws["ccount"].cpu().numpy().reshape(-1, 1)
lens = np.diff(data.train_ptr)
print(json.dumps({"ms_c_call": t0.elapsed_time(t1), "users_per_s": n / t0.elapsed_time(t1) * 1e3, "splits": int(ws["splits"]), "KP": ws["KP"],
                  "flagged": int(fl.sum()), "stat": ws["stat"].tolist(), "epsa_mean": float(ws["epsa"][:n].mean()),
                  "ccount_mean": float(cc.mean()), "ccount_max": int(cc.max()), "cand_per_user_mean": float(cc.sum(1).mean()),
                  "train_len_max": int(lens.max()), "flagged_train_len": lens[np.nonzero(fl)[0][:8]].tolist()}))
ids2, sc2 = e.score_topk(st["row_ptr"], st["col_sorted"], a.k, u0=0, u1=512, tc=False)
ok = (fl[:512] != 0) | ((ids[:512] == ids2).all(1).cpu().numpy() & (sc[:512] == sc2).all(1).cpu().numpy())
print("first 512 rows equal to the fp32 kernel:", bool(ok.all()), int((~ok).sum()))
