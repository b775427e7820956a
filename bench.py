#!/usr/bin/env python
"""Benchmark of the BPR train step (+ full-catalog top-100 evaluation) on B200.

    python bench.py --gpus 1 --steps 50 --warmup 5            # this repo's CUDA path, BASELINE configs[1]
    python bench.py --config c3 --steps 20                     # 1 M users x 500 k items (configs[2]) on one GPU
    torchrun ... bench.py --gpus 8 --config c3                 # the same job, strong scaling over 8 GPUs
    python bench.py --batch 256 --steps 400                    # the reference's default batch (CUDA-graph replay)
    python bench.py --impl reference --steps 5 --warmup 1      # the reference's CPU path (oracle port)

Prints ONE JSON line (contract in the task statement / DESIGN.md "Measurement").

Workloads (synthetic Amazon-fashion-shaped data, random-init weights):
  c1      BPRMF K=64, 20 k users x 10 k items                                  (BASELINE configs[0])
  c2      VBPR K=64 d=20 D=2048, 40 k x 100 k, B=65 536                        (configs[1]; the N=1 default)
  c3      VBPR K=64 d=20 D=2048, 1 M x 500 k, global batch 524 288             (configs[2]; strong scaling)
  c4      = c3, evaluation only: top-100 sweep over all users                  (configs[3]; strong scaling)
  c5      VBPR K=256 d=20 D=4096, 40 k x 100 k                                 (configs[4])
  c5d256  ... with embed_d = 256 (257 columns of E_ext: two column slices on the tensor cores)
With --gpus N > 1 and no --config the job is the WEAK scaling of c2: per GPU 40 k users, 100 k catalog rows and
65 536 triples per step (N=8: 320 k users x 800 k items, 524 288 triples per step); every rank's item shard is
the size of c2's catalog (819 MB of features, far beyond L2).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "BPR triples/s (train)"
UNIT = "triples/s"

CONFIGS = {
    # users, items, K, d, D, global batch, scaling when N > 1
    "c1": dict(users=20000, items=10000, embed_k=64, embed_d=0, feat_dim=0, batch=4096, scaling="strong"),
    "c2": dict(users=40000, items=100000, embed_k=64, embed_d=20, feat_dim=2048, batch=65536, scaling="weak"),
    "c3": dict(users=1000000, items=500000, embed_k=64, embed_d=20, feat_dim=2048, batch=524288, scaling="strong"),
    "c4": dict(users=1000000, items=500000, embed_k=64, embed_d=20, feat_dim=2048, batch=524288, scaling="strong",
               eval_only=True),
    "c5": dict(users=40000, items=100000, embed_k=256, embed_d=20, feat_dim=4096, batch=65536, scaling="weak"),
    "c5d256": dict(users=40000, items=100000, embed_k=256, embed_d=256, feat_dim=4096, batch=32768, scaling="weak"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="fvx", choices=["fvx", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--users", type=int, default=None)
    ap.add_argument("--items", type=int, default=None)
    ap.add_argument("--embed_k", type=int, default=None)
    ap.add_argument("--embed_d", type=int, default=None)
    ap.add_argument("--feat_dim", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)     # weak: per GPU; strong: the global batch
    ap.add_argument("--top_k", type=int, default=100)
    ap.add_argument("--adam_mode", default="auto", choices=["auto", "deferred", "dense", "lazy"])
    ap.add_argument("--tensor_cores", type=int, default=1)
    ap.add_argument("--unique_rows", type=int, default=1)  # 0: one projection per (triple, side) slot
    ap.add_argument("--no_eval", action="store_true")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_parity", action="store_true")
    ap.add_argument("--cpu_seconds", type=float, default=12.0)
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config or "c2"])
    a.config = a.config or "c2"
    for k in ("users", "items", "embed_k", "embed_d", "feat_dim", "batch"):
        if getattr(a, k) is None:
            setattr(a, k, cfg[k])
        elif not a.config.endswith("*"):
            a.config = a.config + "*"                       # a shape flag overrides the named configuration
    a.scaling = a.scaling or cfg["scaling"]
    a.eval_only = bool(cfg.get("eval_only"))
    return a


def bytes_per_triple(K, d, D, B):
    """algorithmic HBM bytes of one triple of one train step."""
    if D == 0:
        return 24 * (3 * K + 2) + 12
    return 24 * (3 * K + d + 2) + 8 * D + 12 + 24.0 * (D * d + D) / B


def load_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def load_peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def global_shape(args, world):
    """(users, items, global batch) of the job on `world` GPUs."""
    if args.scaling == "weak":
        return args.users * world, args.items * world, args.batch * world
    return args.users, args.items, args.batch


def make_features_device(I, D, device, seed=4321, lo=0, cnt=None):
    """Rows [lo, lo+cnt) of the synthetic feature matrix, already divided by the global max |F|
    (visual_loader_mixin.py:30).  Every rank draws the whole matrix in fixed chunks (one generator stream) and
    keeps its rows, so the shards of an N-GPU job are slices of the matrix the one-GPU job holds."""
    import torch
    cnt = I - lo if cnt is None else cnt
    F = torch.empty(cnt, D, dtype=torch.float32, device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    chunk, mx = 65536, torch.zeros((), device=device)
    for s in range(0, I, chunk):
        e = min(I, s + chunk)
        n = torch.randn(e - s, D, generator=g, device=device).clamp_(min=0)
        x = torch.empty(e - s, D, device=device).exponential_(1.0, generator=g)
        n *= x
        mx = torch.maximum(mx, n.max())
        a, b = max(s, lo), min(e, lo + cnt)
        if a < b:
            F[a - lo:b - lo] = n[a - s:b - s]
    F /= mx
    return F


def make_batches(inter, U, I, B, rng):
    """Sampler-shaped host batches for the CPU arms: runs of one user, positives in file order, uniform negatives."""
    owner = np.repeat(np.arange(U), np.diff(inter.row_ptr))
    N = len(owner)

    def batch(i):
        s = (i * B) % max(N - B, 1)
        return owner[s:s + B], inter.col_file[s:s + B].astype(np.int64), rng.integers(0, I, B)
    return batch


# ------------------------------------------------------------------------------------------
def cpu_oracle_rate(U, I, K, d, D, inter, F_host, seconds, B):
    """triples/s of the CPU oracle (NumPy restatement of the reference step, dense Keras-Adam)."""
    from oracle import bpr
    P = bpr.init_params(U, I, K, d, D, seed=0)
    S = bpr.init_adam(P)
    batch = make_batches(inter, U, I, B, np.random.default_rng(0))
    bpr.train_step(P, S, batch(0), 1e-5, 1e-3, F_host)          # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        bpr.train_step(P, S, batch(n + 1), 1e-5, 1e-3, F_host)
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds or n >= 200:
            break
    return n * B / el, n, el


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  TensorFlow 2.3.1 is
    not installable here, so this times the oracle port (NumPy, all host cores through BLAS)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fvx import synth
    from oracle import bpr
    world = int(os.environ.get("WORLD_SIZE", "1"))
    U, I, Bg = global_shape(args, world)
    B = min(Bg, 65536)                  # one per-GPU batch per step: a bounded sample of the N-GPU workload
    K, d, D = args.embed_k, args.embed_d, args.feat_dim
    inter = synth.make_interactions(U, I, seed=1234)
    F = bpr.normalise_features(synth.make_features(I, D)) if D else None
    P = bpr.init_params(U, I, K, d, D, seed=0)
    S = bpr.init_adam(P)
    batch = make_batches(inter, U, I, B, np.random.default_rng(0))
    for i in range(args.warmup):
        bpr.train_step(P, S, batch(i), 1e-5, 1e-3, F)
    t0 = time.perf_counter()
    for i in range(args.steps):
        bpr.train_step(P, S, batch(args.warmup + i), 1e-5, 1e-3, F)
    el = time.perf_counter() - t0
    v = args.steps * B / el
    cores = len(os.sched_getaffinity(0))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps of B=%d triples (one per-GPU batch) on the full tables (NumPy oracle, "
                                       "dense Keras-Adam sweep like TF 2.3; TensorFlow itself not installable)"
                                       % (args.steps, B)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, n):
    U, I, Bg = global_shape(args, n)
    K, d, D = args.embed_k, args.embed_d, args.feat_dim
    name = {"c1": "BASELINE configs[0]", "c2": "BASELINE configs[1]", "c3": "BASELINE configs[2]",
            "c4": "BASELINE configs[3]", "c5": "BASELINE configs[4]", "c5d256": "BASELINE configs[4], embed_d=256"}
    return {"workload": "%s train step, K=%d d=%d D=%d, %d users x %d items (%s%s), %d triples/step%s, on-device Philox "
                        "sampler, %s Adam%s"
                        % ("VBPR" if D else "BPRMF", K, d, D, U, I, name.get(args.config.rstrip("*"), "custom"),
                           "" if not args.config.endswith("*") else ", shape flags overridden",
                           Bg, "" if n == 1 else " over %d GPUs" % n, getattr(args, "adam_resolved", args.adam_mode),
                           ", each distinct catalog row of a batch projected once" if args.tensor_cores and args.unique_rows and D else ""),
            "name": args.config, "users": U, "items": I, "K": K, "d": d, "D": D, "batch": Bg,
            "adam_mode": getattr(args, "adam_resolved", args.adam_mode), "tensor_cores": bool(args.tensor_cores),
            "scaling": None if n == 1 else (
                "weak: per GPU %d users, %d catalog rows, %d triples per step" % (args.users, args.items, args.batch)
                if args.scaling == "weak" else "strong: users, catalog and global batch fixed"),
            "parallelism": "1 GPU" if n == 1 else
            "item catalog (Gi, Bi, F) row-sharded and users block-owned over %d GPUs, E replicated; one C call per "
            "step with 4 exchanges inside (fresh user rows, partial scores, user-row gradient shares, dE), transport %s; "
            "eval: per-shard bounds, all-reduce(MAX), per-shard candidates + top-k, all-to-all merge"
            % (n, getattr(args, "transport_resolved", "p2p (peer stores over NVLink) unless FVX_SHARDED_TRANSPORT=nccl")),
            "l2": "F (%.2f GB%s) and the tables exceed the 126 MB L2; rows are gathered at random, no flush needed"
                  % (I * D * 4 / 1e9 / n, " per GPU" if n > 1 else "") if D else
                  "tables: %.0f MB; batches touch random rows" % ((U + I) * K * 16 / 1e6)}


# ------------------------------------------------------------------------------------------
def parity_full(e, data, F_host, reg, lr, dev, top_k, n_steps=3, n_eval_users=256):
    """The engine against the fp64 oracle on the ACTUAL tables of this run (one GPU): n_steps batches of the timed
    sampler stream (loss <= 1e-4 relative per step), then the masked top-k of n_eval_users users (ids exact at
    tie-free scores, scores <= 1e-4).  Runs before the warm-up; the engine simply continues from there."""
    import torch
    from oracle import bpr, evaluator as oe
    P = {k: v.astype(np.float64) for k, v in e.params().items()}
    S = bpr.init_adam(P)
    F64 = F_host.astype(np.float64) if F_host is not None else None
    it = data.next_triple_batch(str(dev))
    worst = 0.0
    for s in range(n_steps):
        b = next(it)
        hb = tuple(x.cpu().numpy().astype(np.int64) for x in b)
        want = bpr.train_step(P, S, hb, reg, lr, F64)
        e.step(*b, loss_slot=0)
        got = e.read_loss(0)
        err = abs(got - want) / abs(want)
        worst = max(worst, err)
        assert err <= 1e-4, ("parity: loss of step %d" % s, got, want)
    users = np.sort(np.random.default_rng(7).choice(e.U, min(n_eval_users, e.U), replace=False))
    st = data.device_state(str(dev))
    ids, sc = e.score_topk(st["row_ptr"], st["col_sorted"], top_k)
    ids, sc = ids[torch.as_tensor(users, device=dev)].cpu().numpy(), sc[torch.as_tensor(users, device=dev)].cpu().numpy()
    # (the evaluation kernels are checked on the parameters the GPU holds: after Adam steps of size lr on weights of
    # size ~1e-2, a gradient at rounding-noise level moves a weight differently in fp32 and fp64 - the training
    # arithmetic is pinned by the losses above and by tests/)
    Pg = {k: v.astype(np.float64) for k, v in e.params().items()}
    Sc = bpr.predict_all(Pg, F64, users=users)
    tr = [data.training_list[int(u)] for u in users]
    o_ids, o_sc = oe.masked_topk(Sc, tr, top_k)
    for j in range(len(users)):
        # scores: <= 1e-4 of the row's score scale (north_star; SURVEY 8(c): per tensor, absolute floor 1e-7); ids:
        # equal wherever the oracle's scores are not tied within 1e-5
        scale = max(float(np.max(np.abs(Sc[j]))), 1e-3)
        assert np.max(np.abs(sc[j].astype(np.float64) - o_sc[j])) <= 1e-4 * scale + 1e-7, ("parity: scores of user %d" % users[j])
        ok, msg = oe.topk_matches(ids[j], sc[j], o_ids[j], o_sc[j], rel_tol=1e-2)
        assert ok, ("parity: top-%d of user %d" % (top_k, users[j]), msg)
    return {"kind": "full", "what": "%d steps of the timed sampler stream on the run's own tables vs the fp64 oracle (loss <= 1e-4 "
                                   "rel per step, worst %.2e); masked top-%d of %d users vs the oracle (ids at tie-free scores, "
                                   "scores <= 1e-4)" % (n_steps, worst, top_k, len(users)), "worst_loss_rel_err": worst}


def parity_small(world, rank, dev, grp, K, d, D, tc, top_k):
    """The SAME code path as the timed run (sharded: fvx_bpr_step_sharded over real NCCL; one GPU: fvx_bpr_step) on
    a problem small enough for the fp64 oracle on every rank: 6 steps (loss <= 1e-4 rel on every rank), parameters
    (user rows gathered from their owners), item-sharded top-k + merge against the oracle."""
    import torch
    import torch.distributed as dist
    from fvx import parallel
    from fvx.engine import Engine
    from oracle import bpr, evaluator as oe
    U, I, B, reg, lr = 3000, 4001, 4096, 1e-4, 1e-3
    Ds = min(D, 256)
    rng = np.random.default_rng(11)
    P = bpr.init_params(U, I, K, d if Ds else 0, Ds, seed=3)
    F = bpr.normalise_features(np.maximum(rng.standard_normal((I, Ds)), 0)) if Ds else None
    kw = dict(d=d if Ds else 0, D=Ds, lr=lr, reg=reg, max_batch=B, device=str(dev), use_tensor_cores=tc)
    e = parallel.sharded_engine(world, rank, U, I, K, **kw) if world > 1 else Engine(U, I, K, **kw)
    if Ds:
        e.set_features(F[e.item_lo:e.item_lo + e.Ic])
    e.load_params(P)
    step = parallel.ShardedStep([e], grp, max_runs=B // 5 + 3) if world > 1 else None
    Q = {k: v.astype(np.float64) for k, v in P.items()}
    S = bpr.init_adam(Q)
    F64 = F.astype(np.float64) if Ds else None
    worst = 0.0
    for s in range(6):
        u = np.repeat(rng.permutation(U)[:B // 5 + 1], 5)[:B]
        b = (u, rng.integers(0, I, B), rng.integers(0, I, B))
        want = bpr.train_step(Q, S, b, reg, lr, F64)
        db = tuple(torch.as_tensor(x, dtype=torch.int32).to(dev) for x in b)
        if step is not None:
            step.step(*db, loss_slot=0)
            got = step.read_loss(0)
        else:
            e.step(*db, loss_slot=0)
            got = e.read_loss(0)
        worst = max(worst, abs(got - want) / abs(want))
        assert abs(got - want) <= 1e-4 * abs(want), ("parity: loss of step %d on rank %d" % (s, rank), got, want)
    if step is not None:
        step.sync_users()
    R_ = e.params()
    for name, ref in Q.items():
        if name in ("Gi", "Bi"):
            ref = ref[e.item_lo:e.item_lo + e.Ic]
        dlt = np.abs(R_[name].reshape(ref.shape) - ref) / np.abs(ref).max()
        assert (dlt > 1e-4).mean() <= 2e-3 and dlt.max() <= 5e-3, ("parity: %s on rank %d" % (name, rank), float(dlt.max()))
    tr = [sorted(rng.choice(I, 6, replace=False).tolist()) for _ in range(U)]
    rp = torch.as_tensor(np.arange(U + 1) * 6, dtype=torch.int64).to(dev)
    cs = torch.as_tensor(np.array(tr).reshape(-1), dtype=torch.int32).to(dev)
    k = min(top_k, 100)
    Pg = {}
    for name in Q:                                                # the model as the GPUs hold it, assembled
        t = torch.as_tensor(R_[name]).to(dev)
        if world > 1 and name in ("Gi", "Bi"):
            cnts = [parallel.shard_bounds(I, world, r)[1] for r in range(world)]
            pad = torch.zeros((max(cnts),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            pad[:t.shape[0]] = t
            outs = [torch.zeros_like(pad) for _ in range(world)]
            dist.all_gather(outs, pad)
            t = torch.cat([o[:c] for o, c in zip(outs, cnts)])
        Pg[name] = t.cpu().numpy().astype(np.float64)
    o_ids, o_sc = oe.masked_topk(bpr.predict_all(Pg, F64), tr, k)
    if world > 1:
        ids, sc = parallel.sharded_topk([e], grp, rp, cs, k)[0]
        per, _ = parallel.user_slices(U, world)
        u0 = rank * per
    else:
        ids, sc = e.score_topk(rp, cs, k)
        per, u0 = U, 0
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    for j in range(0, min(per, U - u0), 7):
        scale = max(float(np.max(np.abs(o_sc[u0 + j]))), 1e-3)
        assert np.max(np.abs(sc[j].astype(np.float64) - o_sc[u0 + j])) <= 1e-4 * scale + 1e-7, ("parity: scores", u0 + j, rank)
        ok, msg = oe.topk_matches(ids[j], sc[j], o_ids[u0 + j], o_sc[u0 + j], rel_tol=1e-2)
        assert ok, ("parity: top-k of user %d on rank %d" % (u0 + j, rank), msg)
    if world > 1:
        t = torch.tensor([worst], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = float(t.item())
    return {"kind": "small", "what": "the timed code path (%s) on a %d x %d problem with the run's K/d and D=%d vs the fp64 oracle on "
                                    "every rank: 6 steps (loss <= 1e-4 rel, worst %.2e), parameters after them, masked top-%d "
                                    "(%s)" % ("fvx_bpr_step_sharded over NCCL, %d ranks" % world if world > 1 else "fvx_bpr_step",
                                              U, I, Ds, worst, k, "per-shard sweep + all-to-all + merge" if world > 1 else "one GPU"),
            "worst_loss_rel_err": worst}


# ------------------------------------------------------------------------------------------
def run_fvx(args):
    import torch
    import torch.distributed as dist
    from fvx import parallel, synth
    from fvx.build import build
    from fvx.dataset.dataset import DataLoader
    from fvx.engine import Engine, HostStepper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build()
    if world > 1:
        dist.barrier()
    U, I, B = global_shape(args, world)
    K, d, D = args.embed_k, args.embed_d, args.feat_dim
    tc = bool(args.tensor_cores)
    lr, reg = 1e-3, 1e-5
    grp = parallel.DistGroup() if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- parity first: the timed code path against the oracle ------------------------------------------------
    parity = None
    if not args.no_parity and (world > 1 or U * (K + d) + I * max(D, K) > 500e6):
        parity = parity_small(world, rank, dev, grp, K, d, D, tc, args.top_k)

    inter = synth.make_interactions(U, I, seed=1234)
    p = argparse.Namespace(dataset="synthetic", batch_size=B, epochs=10 ** 6, sampler="device", seed=0, rec="vbpr" if D else "bprmf",
                           embed_k=K, embed_d=d, lr=lr, reg=reg, top_k=args.top_k, verbose=-1, restore_epochs=1,
                           adam_mode=args.adam_mode, tensor_cores=tc, device=str(dev), batch_eval=128, best_metric="ndcg",
                           validation=True, cnn_model="resnet50", output_layer="avg_pool")
    data = DataLoader(p, interactions=inter)
    model = None
    F_host = None
    if world == 1:
        # the reference-facing class owns the engine: the kernel-only loop, the e2e loops and the evaluator below all
        # run on the engine a user of `VBPR(data, params)` gets
        from fvx.recommender.models.BPRMF import BPRMF
        from fvx.recommender.models.VBPR import VBPR
        if D:
            F_host = make_features_device(I, D, dev).cpu().numpy()
            data.cnn_features_raw = F_host                      # already max-normalised: the loader's scale is 1
            model = VBPR(data, p)
        else:
            model = BPRMF(data, p)
        e = model.engine
        if D and tc and e.use_tensor_cores:
            e.F = None                                          # the tensor-core path reads the planes only
            e._struct = None
        if not args.unique_rows:
            e.upos_t = e.W_sum = e.uslot_t = None
            e._struct = None
    else:
        e = parallel.sharded_engine(world, rank, U, I, K, d=d, D=D, lr=lr, reg=reg, adam_mode=args.adam_mode, max_batch=B,
                                    device=str(dev), seed=0, use_tensor_cores=tc, unique_rows=bool(args.unique_rows))
        if D:
            e.set_features(make_features_device(I, D, dev, lo=e.item_lo, cnt=e.Ic), keep_fp32=not e.use_tensor_cores)
    uniq = e.upos_t is not None and os.environ.get("FVX_STEP_DEDUP", "1") != "0"   # unique-row step in use
    args.adam_resolved = {0: "dense", 1: "deferred", 2: "lazy"}[e.adam_mode]
    if not args.no_parity and parity is None:
        parity = parity_full(e, data, F_host, reg, lr, dev, args.top_k)
    # rows of the exchanged user buffers (WU, RU): runs of equal users in a batch.  The hard bound is B / (shortest
    # train list) + 2; the batches hold B / (mean list) runs with a relative spread of a few per mille at this size,
    # so 1.3 x the mean + 1024 is used: a batch beyond it poisons the loss with NaN on every rank (checked below)
    lens = np.diff(inter.row_ptr)
    min_len, mean_len = int(lens.min()), float(lens.mean())
    max_runs = min(B // max(min_len, 1) + 2, int(1.3 * B / mean_len) + 1024)
    sharded = parallel.ShardedStep([e], grp, max_runs=max_runs) if world > 1 else None
    if sharded is not None:
        args.transport_resolved = {"p2p": "p2p (the kernels store into their peers' buffers over NVLink, one-warp barriers)",
                                   "nccl": "nccl (all-gather / all-reduce / reduce-scatter / all-reduce)"}[sharded.transport]
    graph = world == 1 and B <= 8192                 # small batches: 8 steps per CUDA-graph launch (fvx_bpr_steps)
    runs = data.next_batch_run(str(dev))
    cur = {"bufs": None, "n": 0, "pos": 0}

    def next_run():
        cur["bufs"], cur["n"] = next(runs)
        cur["pos"] = 0

    def next_batch():
        if cur["pos"] >= cur["n"]:
            next_run()
        s = cur["pos"]
        cur["pos"] += 1
        return tuple(x[s * B:(s + 1) * B] for x in cur["bufs"])

    def do_step(b, slot=0):
        if sharded is not None:
            sharded.step(*b, loss_slot=slot)
        else:
            e.step(*b, loss_slot=slot)

    def do_steps(n):
        """n consecutive steps of the sampler stream"""
        while n > 0:
            if cur["pos"] >= cur["n"]:
                next_run()
            if graph:
                m = min(n, cur["n"] - cur["pos"])
                e.steps(cur["bufs"][0], cur["bufs"][1], cur["bufs"][2], cur["pos"], m, B, loss_slot=0)
                cur["pos"] += m
                n -= m
            else:
                do_step(next_batch())
                n -= 1

    def loss(slot=0):
        return sharded.read_loss(slot) if sharded is not None else e.read_loss(slot)

    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world)}
    hbm, tf_burst, tf_sus, src = load_peaks()

    if not args.eval_only:
        do_steps(max(args.warmup, 8 if graph else 0))        # (the first graph launch captures)
        clocks = ClockSampler(local)
        barrier()
        clocks.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        do_steps(args.steps)
        ev1.record()
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        # the timed region can be shorter than one nvidia-smi sampling period: keep the identical load
        # running (untimed) until the sampler has seen it for ~0.5 s
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 0.5:
            do_steps(16)
            torch.cuda.synchronize()
        clk = clocks.stop()
        clk["note"] = "sampled during the timed steps and %.1f s of the same steps run untimed right after" % 0.5
        l_ = loss(0)
        assert np.isfinite(l_) and l_ > 0, l_
        value = args.steps * B / ms * 1e3
        step_ms = ms / args.steps
        epochs_in_region = (args.steps * B) / max(data.num_train, 1)
        # our kernels per step.  One GPU, VBPR, unique-row step, deferred Adam: rows+claims, catch-up, projection,
        # score+grad, coefficient planes, grad_E, E update (= 7; + batch fetch and cursor when graph-replayed);
        # sharded: + run ids (2), fresh-row pack, owned-slot list, partial scores, run scatter, dE pack (= 14)
        merged = e.adam_mode == 1                            # DEFERRED
        per_step = (((8 if uniq else 7) if D else 3) - (1 if merged and D else 0)) + (2 if graph else 0) if world == 1 else (14 if D else 10)
        line.update({"value": value, "ms_per_step": step_ms, "clocks": clk,
                     "gpu_launches": args.steps * per_step + int(np.ceil(epochs_in_region)) * 2,
                     "launch_mode": "8 steps per CUDA-graph launch (fvx_bpr_steps)" if graph else
                     ("one C call per step (fvx_bpr_step_sharded), NCCL inside" if world > 1 else "one C call per step (fvx_bpr_step)")})

        # ---- end to end through the public API: pinned host batches in, one float loss out per step ----
        # fvx.engine.HostStepper keeps the reference's per-step contract (BPRMF.py:125 returns float(loss) from every
        # train_step) with ONE step in flight: the upload of batch s+1 and the launch of step s+1 are issued before
        # the host blocks on the loss of step s.  Every step's batch crosses PCIe inside the timed region and every
        # step's loss is read back inside it.  One GPU: the step is `VBPR.train_step(batch, sync=False)` of the
        # reference-facing class; N GPUs: ShardedStep.step.
        hb = []
        for _ in range(min(args.steps, 20)):
            hb.append(torch.stack([x.cpu() for x in next_batch()]).to(torch.int32).pin_memory())
        if model is not None:
            step_fn = lambda u, i, j, loss_slot=0: model.train_step((u, i, j), sync=False, loss_slot=loss_slot)   # noqa: E731
        else:
            step_fn = lambda u, i, j, loss_slot=0: do_step((u, i, j), loss_slot)                                   # noqa: E731
        stepper = HostStepper(step_fn, (sharded.take_loss if sharded is not None else e.take_loss), B, dev)
        e.loss_t.zero_()
        for b in hb[:2]:
            stepper.submit(b); stepper.collect()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        losses = []
        t0.record()
        for b in hb:
            stepper.submit(b)
            if stepper.pending() > 1:
                losses.append(stepper.collect())                 # D2H of a batch loss (BPRMF.py:125)
        while stepper.pending():
            losses.append(stepper.collect())
        t1.record()
        barrier()
        assert len(losses) == len(hb) and all(np.isfinite(x) and x > 0 for x in losses), losses[:4]
        e2e_ms = max_over_ranks(t0.elapsed_time(t1))
        line["e2e"] = {"value": len(hb) * B / e2e_ms * 1e3, "unit": UNIT, "h2d_bytes_per_step": 12 * B, "d2h_bytes_per_step": 8,
                       "steps": len(hb),
                       "how": "fvx.engine.HostStepper over %s: pinned [3,B] int32 host batch -> device every step, float loss "
                              "read back every step, one step in flight (the host blocks on step s after launching s+1)"
                              % ("VBPR.train_step(batch, sync=False)" if model is not None else "ShardedStep.step")}
        if model is not None:
            # the reference's literal contract: int64 NumPy batch in, float(loss) out, nothing in flight
            nb = [tuple(x.numpy().astype(np.int64) for x in b) for b in hb[:10]]
            model.train_step(nb[0]); torch.cuda.synchronize()
            tw = time.perf_counter()
            ls = [model.train_step(b) for b in nb]
            tw = time.perf_counter() - tw
            assert all(np.isfinite(x) and x > 0 for x in ls)
            line["e2e"]["train_step_sync"] = {"value": len(nb) * B / tw, "unit": UNIT, "h2d_bytes_per_step": 24 * B,
                                              "d2h_bytes_per_step": 8, "steps": len(nb),
                                              "how": "loss = VBPR.train_step((user, pos, neg) int64 NumPy arrays) -> Python float, "
                                                     "synchronous like the reference (BPRMF.py:87-125), host wall clock"}

        # ---- per-kernel shares (profiling entry point; single rank; separate from the timed region) ----
        if world == 1 and D:
            phases = {}
            n_rows = []
            for _ in range(8):
                b_ = next_batch()
                # rows one launch of the projection kernels gathers: every slot, or (unique-row step) the
                # distinct catalog rows of the batch
                n_rows.append(int(torch.unique(torch.cat([b_[1], b_[2]])).numel()) if uniq else 2 * B)
                for k_, v in e.step_timed(*b_).items():
                    phases[k_] = phases.get(k_, 0.0) + v / 8
            dom = max(phases, key=phases.get)
            rows_launch = float(np.mean(n_rows))
            rows_bytes = rows_launch * D * 4.0
            tbl = 4.0 * (3 * K + d + 2)                      # floats of the three rows of a triple, once
            kern_bytes = {"project": rows_bytes + rows_launch * e.de * 4.0, "grad_E": rows_bytes + rows_launch * e.de * 4.0,
                          "score_grad": B * 2 * tbl, "update": B * 4 * tbl + (28.0 * D * e.de if D else 0),
                          "prep": B * (12.0 + 3 * tbl)}
            kname = {"project": "k_proj_fwd_tc" if e.use_tensor_cores else "k_project",
                     "grad_E": "k_grad_E_tc" if e.use_tensor_cores else "k_grad_E",
                     "score_grad": "k_score_grad_v4" if K % 4 == 0 else "k_score_grad",
                     "update": "k_update", "prep": "k_prep"}[dom]
            ach = kern_bytes.get(dom, 0.0) / (phases[dom] * 1e-3) / 1e9 if phases[dom] > 0 else 0.0
            traffic = load_traffic(kname) if args.config == "c2" else None
            line["roofline"] = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": hbm, "unit": "GB/s",
                                "frac": ach / hbm, "traffic": traffic, "peak_source": src,
                                "algorithmic_bytes_per_launch": kern_bytes.get(dom, 0.0),
                                "kernel_ms": phases[dom], "phase_ms": phases,
                                "phase_gbs": {k_: (kern_bytes.get(k_, 0.0) / (v * 1e-3) / 1e9 if v > 0 else None)
                                              for k_, v in phases.items()},
                                "rows_per_projection_launch": rows_launch, "slots_per_step": 2 * B,
                                "unique_row_step": bool(uniq),
                                # the same kernel under SURVEY 8(d)'s per-triple accounting (2 feature rows per triple,
                                # no credit for rows a batch repeats): what the step "asked for", not what the kernel moved
                                "survey_accounting": ({"bytes_per_launch": 2.0 * B * D * 4.0,
                                                       "achieved": 2.0 * B * D * 4.0 / (phases[dom] * 1e-3) / 1e9,
                                                       "frac": 2.0 * B * D * 4.0 / (phases[dom] * 1e-3) / 1e9 / hbm}
                                                      if dom in ("project", "grad_E") and phases[dom] > 0 else None),
                                "dram_achieved": (traffic / (phases[dom] * 1e-3) / 1e9) if traffic and phases[dom] > 0 else None,
                                "note": "achieved = bytes the launch is asked to move (projection kernels: gathered rows x 4D "
                                        "- with the unique-row step the DISTINCT catalog rows of the batch, not the 2B slots; "
                                        "the per-triple accounting of step_roofline gives no such credit) / CUDA-event kernel "
                                        "time (single-stream timed entry point); dram_achieved = ncu DRAM bytes of the same "
                                        "launch / the same time (profiles/traffic.json, BASELINE configs[1] only)"}
        bpt = bytes_per_triple(K, d, D, B)
        step_roof = {"achieved": B * bpt / world / (step_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "bound": "hbm",
                     "bytes_per_triple": bpt, "note": "algorithmic bytes of the whole step per GPU (SURVEY 8(d), no credit "
                                                      "for rows a batch repeats) / step time"}
        step_roof["frac"] = step_roof["achieved"] / hbm
        line["step_roofline"] = step_roof
        if "roofline" not in line:
            line["roofline"] = dict(step_roof, kernel="whole step", traffic=None,
                                    note="per-kernel split is measured on one GPU with VBPR only; " + step_roof["note"])

    # ---- evaluation: users/s full-catalog top-k, train items masked (all users) --------------------
    if not args.no_eval:
        st = data.device_state(str(dev))
        e.flush()
        torch.cuda.synchronize()

        def sweep_item_shards():
            # every rank: all users x its item shard -> all-to-all by user slice -> merge
            e.theta(refresh=True)
            if world > 1:
                return parallel.sharded_topk([e], grp, st["row_ptr"], st["col_sorted"], args.top_k)[0]
            return e.score_topk(st["row_ptr"], st["col_sorted"], args.top_k)

        def sweep_user_slices():
            # theta of the own shard -> all-gather of (Gi|Bi, theta) -> every rank: its users x the catalog
            e.theta(refresh=True)
            return parallel.user_sliced_topk([e], grp, st["row_ptr"], st["col_sorted"], args.top_k)[0]

        def timed(fn, d2h=False, reps=3):
            for _ in range(2):
                fn()                                      # warm-up (workspace allocation)
            barrier()
            best = 1e30
            for _ in range(reps):
                a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ids_, sc_ = fn()
                if d2h:
                    hi_, hs_ = ids_.cpu(), sc_.cpu()    # the recommendation lists leave the device (Evaluator.py:233-239)
                b2.record()
                barrier()
                best = min(best, max_over_ranks(a.elapsed_time(b2)))
            return best

        if os.environ.get("FVX_BENCH_EVAL_TRACE") and world == 1:
            # where the sweep's time goes on the host's clock and on the device's (diagnostics, stderr)
            for rep in range(3):
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ev[0].record()
                e.theta(refresh=True)
                t1 = time.perf_counter()
                ev[1].record()
                e.score_topk(st["row_ptr"], st["col_sorted"], args.top_k)
                t2 = time.perf_counter()
                ev[2].record()
                torch.cuda.synchronize()
                t3 = time.perf_counter()
                if rep == 2:
                    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    ev2[0].record()
                    e.topk_bounds(st["row_ptr"], args.top_k)
                    ev2[1].record()
                    e.topk_select(st["row_ptr"], st["col_sorted"], args.top_k)
                    ev2[2].record()
                    torch.cuda.synchronize()
                    sys.stderr.write("eval trace: bounds %.3f ms, select %.3f ms\n"
                                     % (ev2[0].elapsed_time(ev2[1]), ev2[1].elapsed_time(ev2[2])))
                    w_ = e._eval_ws(U)
                    cc_ = w_["ccount"][:w_["lists"]].float()
                    sys.stderr.write("eval trace: KP %d splits %d n_ut %d lists %d a_stride %d cand mean %.1f max %d flags %d\n"
                                     % (w_["KP"], w_["splits"], w_["n_ut"], w_["lists"], w_["struct"].a_stride,
                                        cc_.mean().item(), int(cc_.max().item()), int(w_["flags"].sum().item())))
                sys.stderr.write("eval trace: theta %.3f ms (host %.3f), topk %.3f ms (host %.3f), host total %.3f\n"
                                 % (ev[0].elapsed_time(ev[1]), (t1 - t0) * 1e3, ev[1].elapsed_time(ev[2]),
                                    (t2 - t1) * 1e3, (t3 - t0) * 1e3))
        nu = U
        flops_user = 2.0 * I * (K + d) + 2.0 * I
        modes = {"item_shards": timed(sweep_item_shards)}
        if world > 1:
            modes["user_slices"] = timed(sweep_user_slices)
        mode = "item_shards"                              # the north-star decomposition is the one reported
        best = modes[mode]
        e2e_eval = timed(sweep_item_shards, d2h=True, reps=2)
        tfl = nu * flops_user / (best * 1e-3) / 1e12
        line["eval"] = {"metric": "users/s full-catalog top-%d eval" % args.top_k, "value": nu / best * 1e3,
                        "unit": "users/s", "users": nu, "items": I, "ms": best,
                        "scaling": None if world == 1 else args.scaling,
                        "kernel": "k_topk_tc (tcgen05 bf16 bounds sweep + candidates sweep, exact fp32 re-scoring)" if (e.use_tensor_cores and e.tc_eval_eligible())
                        else "k_score_topk (fp32 CUDA cores)",
                        "fallback_rows": getattr(e, "tc_overflow_rows", 0),
                        "decomposition": "item shards: per-shard sweep of all users, all-to-all by user slice, merge" if world > 1 else "1 GPU",
                        "ms_by_decomposition": modes,
                        "e2e": {"value": nu / e2e_eval * 1e3, "unit": "users/s", "ms": e2e_eval,
                                "d2h_bytes": 8 * args.top_k * (nu // world), "h2d_bytes": 0,
                                "how": "the same sweep with the [users, k] ids and scores of this rank's users copied to the host "
                                       "inside the timed region (what Evaluator.store_recommendation writes out)"},
                        "includes": "theta = F*E projection of the catalog (F stays item-sharded), operand packing, both "
                                    "sweeps, exact re-scoring, mask, top-k" + ("; all-to-all exchange + merge of per-shard lists"
                                                                                if world > 1 else ""),
                        "roofline": {"bound": "tensor", "achieved": tfl, "peak": tf_burst * world, "unit": "TFLOP/s",
                                     "frac": tfl / (tf_burst * world),
                                     "note": "algorithmic flops (one bf16 pass over U x I x (K+d+1)) / time; the kernel sweeps "
                                             "the catalog twice (bounds, candidates)"}}
        if args.eval_only:
            line.update({"metric": line["eval"]["metric"], "unit": "users/s", "value": line["eval"]["value"],
                         "ms_per_step": best, "steps": 3, "warmup": 2, "gpu_launches": 3 * 6,
                         "e2e": dict(line["eval"]["e2e"], h2d_bytes_per_step=0, d2h_bytes_per_step=8 * args.top_k * (nu // world)),
                         "roofline": dict(line["eval"]["roofline"], kernel="k_topk_tc", traffic=None)})

    # ---- Evaluator.eval's device work: rank counts of two held-out items per user (validation + test)
    if not args.no_eval and world == 1 and U <= 200000:
        held = torch.randint(0, I, (U, 2), device=dev, dtype=torch.int32)
        who = torch.arange(U, device=dev, dtype=torch.int32).repeat_interleave(2)
        thr = e.score_pairs(who, held.reshape(-1).contiguous()).reshape(U, 2).contiguous()

        def t_ms(fn, reps):
            fn(); torch.cuda.synchronize()
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                out = fn()
            b2.record(); torch.cuda.synchronize()
            return a.elapsed_time(b2) / reps, out

        ms_new, c_new = t_ms(lambda: e.rank_counts(st["row_ptr"], st["col_sorted"], thr), 3)
        fl = 2.0 * U * I * (K + d)
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12            # CUDA-core fp32 FMA peak at the boost clock
        line["eval"]["rank_counts"] = {
            "what": "Evaluator.eval device work: #items scoring >= each of 2 held-out items per user, train items "
                    "excluded (fvx_rank_counts, exact fp32 register-tiled sweep)",
            "ms": ms_new, "users_per_s": U / ms_new * 1e3, "tflops_fp32": fl / (ms_new * 1e-3) / 1e12,
            "frac_of_fp32_fma_peak": fl / (ms_new * 1e-3) / 1e12 / fp32_peak}

    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.eval_only:
        v, n, el = cpu_oracle_rate(U, I, K, d, D, inter, F_host, args.cpu_seconds, min(B, 65536))
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
                                "sample": "%d steps of B=%d in %.1f s on the full tables (NumPy oracle of the "
                                          "reference step with TF-2.3 dense Keras-Adam)" % (n, min(B, 65536), el)}
    # ---- the reference's evaluation on the host cores (port): predict_all + store_recommendation's
    # mask / top-k (Evaluator.py:225-239) and _eval_by_user (:82-128) on a slice of users (the full [U, I]
    # matrix of this config is 16 GB and the per-user Python loops take minutes) --------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_eval:
        try:
            from oracle import bpr as obpr, evaluator as oev
            n_s = 256
            Pq = {k_: v for k_, v in e.params().items()}
            tr_ptr, tr_col = inter.row_ptr, inter.col_file
            tr = [np.sort(tr_col[tr_ptr[u]:tr_ptr[u + 1]]).tolist() for u in range(n_s)]
            held = [[int(inter.test[u])] if np.ndim(inter.test) == 1 else list(inter.test[u]) for u in range(n_s)]
            t0_ = time.perf_counter()
            Sc = obpr.predict_all(Pq, F_host, users=np.arange(n_s))
            oev.masked_topk(Sc, tr, args.top_k)
            t1_ = time.perf_counter()
            for u in range(n_s):
                oev.eval_by_user(Sc[u], I, tr[u], held[u], args.top_k)
            t2_ = time.perf_counter()
            line["eval"]["cpu_baseline"] = {
                "kind": "port", "cores": len(os.sched_getaffinity(0)), "users": n_s,
                "topk_users_per_s": n_s / (t1_ - t0_), "eval_users_per_s": n_s / (t2_ - t1_),
                "sample": "%d users x %d items: NumPy predict_all + masked top-%d (store_recommendation), then "
                          "_eval_by_user per user for one held-out item" % (n_s, I, args.top_k)}
            if args.eval_only:
                line["cpu_baseline"] = {"value": n_s / (t1_ - t0_), "unit": "users/s", "cores": len(os.sched_getaffinity(0)),
                                        "kind": "port", "sample": line["eval"]["cpu_baseline"]["sample"]}
        except Exception as ex:  # a reported baseline must never cost the bench line
            line["eval"]["cpu_baseline"] = {"kind": "port", "error": repr(ex)}

    line["parity_checked"] = parity is not None
    line["parity"] = parity
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        grp.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    # ONE JSON line on stdout, whatever the libraries print (NCCL writes its version banner to stdout when a
    # communicator is created): file descriptor 1 is pointed at stderr for the run and the line goes to the real stdout
    sys.stdout.flush()
    _real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    _print = print

    def print(*x, **k):      # noqa: A001  (the run functions print the line through this)
        _print(*x, file=_real, **k)
        _real.flush()

    if a.impl == "reference":
        run_reference(a)
    else:
        run_fvx(a)
