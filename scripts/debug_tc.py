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
a = ap.parse_args()
inter = synth.make_interactions(a.users, a.items, seed=1234)
data = DataLoader(argparse.Namespace(dataset="x", batch_size=4096, epochs=1, sampler="device", seed=0), interactions=inter)
e = Engine(a.users, a.items, 64, d=20, D=a.D, max_batch=8)
g = torch.Generator(device="cuda").manual_seed(1)
e.set_features(torch.rand(a.items, a.D, generator=g, device="cuda"))
st = data.device_state()
n = a.users
ws = e._eval_ws(n)
ids = torch.empty(n, a.k, dtype=torch.int32, device="cuda"); sc = torch.empty(n, a.k, device="cuda")
e.flush()
call("fvx_score_topk_tc", C.byref(e.struct()), ptr(e.theta()), 0, n, ptr(st["row_ptr"]), ptr(st["col_sorted"]), a.k,
     ptr(ids), ptr(sc), C.byref(ws["struct"]), stream_ptr())
torch.cuda.synchronize()
fl = ws["flags"][:n].cpu().numpy(); cc = ws["ccount"].cpu().numpy().reshape(n, -1)
print("lists/user", cc.shape[1], "KP", ws["KP"], "flagged", int(fl.sum()), "bmax", float(ws["bmax"].item()),
      "unorm", ws["unorm"][:4].tolist())
print("ccount: mean %.1f max %d  per-list hist" % (cc.mean(), cc.max()), np.bincount(np.minimum(cc.reshape(-1) // 32, 8)))
bad = np.nonzero(fl)[0][:3]
lens = np.diff(data.train_ptr)
for r in bad:
    print("row", r, "train len", lens[r], "counts", cc[r].tolist())
print("rows flagged by train len>60:", int((lens > 60).sum()))
