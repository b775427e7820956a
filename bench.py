#!/usr/bin/env python
"""Benchmark of the BPR train step (+ full-catalog top-100 evaluation) on B200.

    python bench.py --gpus 1 --steps 50 --warmup 5            # this repo's CUDA path
    python bench.py --impl reference --steps 5 --warmup 1      # the reference's CPU path (oracle port)

Prints ONE JSON line (contract in the task statement / DESIGN.md "Measurement").
Workload at N=1: BASELINE.json configs[1] - VBPR K=64, d=20, 2048-d features,
40k users x 100k items, synthetic Amazon-fashion-shaped data, random-init weights.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="fvx", choices=["fvx", "reference"])
    ap.add_argument("--users", type=int, default=40000)
    ap.add_argument("--items", type=int, default=100000)
    ap.add_argument("--embed_k", type=int, default=64)
    ap.add_argument("--embed_d", type=int, default=20)
    ap.add_argument("--feat_dim", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=65536)   # per GPU; BASELINE C3 / SURVEY 8(d) batch
    ap.add_argument("--top_k", type=int, default=100)
    ap.add_argument("--adam_mode", default="deferred", choices=["deferred", "dense", "lazy"])
    ap.add_argument("--tensor_cores", type=int, default=1)
    ap.add_argument("--unique_rows", type=int, default=1)  # 0: one projection per (triple, side) slot
    ap.add_argument("--no_eval", action="store_true")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--cpu_seconds", type=float, default=12.0)
    return ap.parse_args()


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


def make_problem(args, device=None):
    """Synthetic interactions (host) + normalised features (device tensor or numpy)."""
    from fvx import synth
    inter = synth.make_interactions(args.users, args.items, seed=1234)
    return inter


def make_features_device(I, D, device, seed=4321):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    F = torch.empty(I, D, dtype=torch.float32, device=device)
    chunk = 65536
    for s in range(0, I, chunk):
        e = min(I, s + chunk)
        n = torch.randn(e - s, D, generator=g, device=device).clamp_(min=0)
        x = torch.empty(e - s, D, device=device).exponential_(1.0, generator=g)
        F[s:e] = n * x
    F /= F.abs().max()                       # visual_loader_mixin.py:30
    return F


# ------------------------------------------------------------------------------------------
def cpu_oracle_rate(args, inter, F_host, seconds, B):
    """triples/s of the CPU oracle (NumPy restatement of the reference step, dense Keras-Adam)."""
    from oracle import bpr
    rng = np.random.default_rng(0)
    P = bpr.init_params(args.users, args.items, args.embed_k, args.embed_d, args.feat_dim, seed=0)
    S = bpr.init_adam(P)
    owner = np.repeat(np.arange(args.users), np.diff(inter.row_ptr))
    N = len(owner)

    def batch(i):
        s = (i * B) % max(N - B, 1)
        return owner[s:s + B], inter.col_file[s:s + B].astype(np.int64), rng.integers(0, args.items, B)

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
    B = args.batch                      # one per-GPU batch per step: a bounded sample of the N-GPU workload
    args.users = args.users * world     # the same problem size as the fvx arm at this N
    inter = make_problem(args)
    F = None
    if args.feat_dim:
        F = bpr.normalise_features(synth.make_features(args.items, args.feat_dim))
    rng = np.random.default_rng(0)
    P = bpr.init_params(args.users, args.items, args.embed_k, args.embed_d, args.feat_dim, seed=0)
    S = bpr.init_adam(P)
    owner = np.repeat(np.arange(args.users), np.diff(inter.row_ptr))
    N = len(owner)

    def batch(i):
        s = (i * B) % max(N - B, 1)
        return owner[s:s + B], inter.col_file[s:s + B].astype(np.int64), rng.integers(0, args.items, B)

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
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps of B=%d triples (one per-GPU batch) on the full tables (NumPy oracle, "
                                       "dense Keras-Adam sweep like TF 2.3; TensorFlow itself not installable)"
                                       % (args.steps, B)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, n):
    return {"workload": "VBPR train step, K=%d d=%d D=%d, %d users x %d items (BASELINE configs[1]), "
                        "B=%d triples/step per GPU, on-device Philox sampler, %s Adam%s%s"
                        % (args.embed_k, args.embed_d, args.feat_dim, args.users, args.items, args.batch,
                           args.adam_mode,
                           ", each distinct catalog row of a batch projected once" if args.tensor_cores and args.unique_rows else "",
                           "" if n == 1 else "; weak scaling: 40 000 users and one batch per GPU"),
            "users": args.users, "items": args.items, "K": args.embed_k, "d": args.embed_d, "D": args.feat_dim,
            "batch": args.batch * n, "adam_mode": args.adam_mode, "tensor_cores": bool(args.tensor_cores),
            "parallelism": "1 GPU" if n == 1 else
            "item catalog row-sharded over %d GPUs (users, E replicated); NCCL all-reduce of partial scores, "
            "user-row gradients and dE; eval: per-shard top-k + all-to-all merge" % n,
            "l2": "F (%.2f GB) and the tables exceed the 126 MB L2; rows are gathered at random, no flush needed"
                  % (args.items * args.feat_dim * 4 / 1e9)}


def run_fvx(args):
    import torch
    import torch.distributed as dist
    from fvx.build import build
    from fvx.dataset.dataset import DataLoader
    from fvx.engine import Engine
    from fvx import parallel

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
    # weak scaling: the per-GPU batch is fixed, the global batch grows with the number of ranks;
    # the item catalog (Gi, Bi, F + Adam state) is row-sharded, users and E are replicated
    B, K, d, D = args.batch * world, args.embed_k, args.embed_d, args.feat_dim

    # ... and so does the number of users (40 000 per GPU): training and evaluation work per GPU stay
    # fixed while the catalog stays at `items` rows, sharded
    args.users = args.users * world
    inter = make_problem(args)
    p = argparse.Namespace(dataset="synthetic", batch_size=B, epochs=10 ** 6, sampler="device", seed=0)
    data = DataLoader(p, interactions=inter)
    lo, cnt = parallel.shard_bounds(args.items, world, rank)
    e = Engine(args.users, args.items, K, d=d, D=D, lr=1e-3, reg=1e-5, adam_mode=args.adam_mode,
               max_batch=B, device=str(dev), seed=0, use_tensor_cores=bool(args.tensor_cores),
               item_lo=lo, item_cnt=cnt, unique_rows=bool(args.unique_rows))
    uniq = e.upos_t is not None and os.environ.get("FVX_STEP_DEDUP", "1") != "0"   # unique-row step in use
    if D:
        F = make_features_device(args.items, D, dev)          # same generator seed on every rank
        e.set_features(F[lo:lo + cnt].contiguous(), keep_fp32=not args.tensor_cores)
        del F
    # rows of the all-reduced user-gradient buffer: runs of equal users in a batch.  The hard bound is
    # B / (shortest train list); the batches hold B / (mean list) runs with a relative spread of a few
    # per mille at this size, so 1.3 x the mean + 1024 is used (an overflow is detected on the device and
    # raised by read_loss - it cannot pass silently)
    lens = np.diff(inter.row_ptr)
    min_len, mean_len = int(lens.min()), float(lens.mean())
    max_runs = min(B // max(min_len, 1) + 2, int(1.3 * B / mean_len) + 1024)
    sharded = parallel.ShardedStep([e], parallel.DistGroup(), max_runs=max_runs) if world > 1 else None
    batches = data.next_triple_batch(str(dev))

    def do_step(b, slot=0):
        if sharded is not None:
            sharded.step(*b, loss_slot=slot)
        else:
            e.step(*b, loss_slot=slot)

    def loss(slot=0):
        return sharded.read_loss(slot) if sharded is not None else e.read_loss(slot)

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

    for _ in range(args.warmup):
        do_step(next(batches))
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        do_step(next(batches))
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    # the timed region can be shorter than one nvidia-smi sampling period: keep the identical load
    # running (untimed) until the sampler has seen it for ~0.5 s
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.5:
        for _ in range(20):
            do_step(next(batches))
        torch.cuda.synchronize()
    clk = clocks.stop()
    clk["note"] = "sampled during the timed steps and %.1f s of the same steps run untimed right after" % 0.5
    value = args.steps * B / ms * 1e3
    # kernels of the timed region: 5 per step (prep, projection, score+grad, grad_E, update), 8 on the
    # sharded path (+ partial scores, reduce, scatter), 2 per generated epoch
    epochs_in_region = (args.steps * B) / max(data.num_train, 1)
    # 1 GPU, VBPR: rows+planes(+item claims), claims/catch-up, projection, score+grad, row update,
    # (coefficient planes: unique-row step), grad_E, E update
    merged = args.adam_mode == "deferred" and os.environ.get("FVX_STEP_MERGED_UPDATE", "1") != "0"
    per_step = (((8 if uniq else 7) if D else 3) - (1 if merged and D else 0)) if world == 1 else (8 if D else 5)
    gpu_launches = args.steps * per_step + int(np.ceil(epochs_in_region)) * 2

    # ---- end to end through the public API: pinned host batches in, one float loss out per step ----
    # fvx.engine.HostStepper keeps the reference's per-step contract (BPRMF.py:125 returns float(loss)
    # from every train_step) with ONE step in flight: the upload of batch s+1 and the launch of step s+1
    # are issued before the host blocks on the loss of step s.  Every step's batch crosses PCIe inside
    # the timed region and every step's loss is read back inside it.
    from fvx.engine import HostStepper
    hb = []
    for _ in range(min(args.steps, 20)):
        hb.append(torch.stack([x.cpu() for x in next(batches)]).to(torch.int32).pin_memory())
    stepper = HostStepper(lambda u, i, j, loss_slot=0: do_step((u, i, j), loss_slot),
                          (sharded.take_loss if sharded is not None else e.take_loss), B, dev)
    loss(0); loss(1)                                       # clear the two slots the stepper uses
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
    if sharded is not None:
        sharded.check_runs()
    assert len(losses) == len(hb) and all(np.isfinite(x) and x > 0 for x in losses), losses[:4]
    e2e_ms = max_over_ranks(t0.elapsed_time(t1))
    e2e_value = len(hb) * B / e2e_ms * 1e3

    hbm, tf_burst, tf_sus, src = load_peaks()
    step_ms = ms / args.steps
    bpt = bytes_per_triple(K, d, D, B)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clk, "gpu_launches": gpu_launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 12 * B, "d2h_bytes_per_step": 8,
                    "steps": len(hb),
                    "how": "fvx.engine.HostStepper: pinned [3,B] int32 host batch -> device every step, float loss "
                           "read back every step, one step in flight (the host blocks on step s after launching s+1)"}}

    # ---- per-kernel shares (profiling entry point; single rank; separate from the timed region) ----
    if world == 1:
        phases = {}
        n_rows = []
        for _ in range(8):
            b_ = next(batches)
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
        kname = {"project": "k_proj_fwd_tc" if args.tensor_cores else "k_project",
                 "grad_E": "k_grad_E_tc" if args.tensor_cores else "k_grad_E",
                 "score_grad": "k_score_grad_v4" if K % 4 == 0 else "k_score_grad",
                 "update": "k_update", "prep": "k_prep"}[dom]
        ach = kern_bytes.get(dom, 0.0) / (phases[dom] * 1e-3) / 1e9 if phases[dom] > 0 else 0.0
        traffic = load_traffic(kname)
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
                            # DRAM bytes of the same launch from the committed ncu capture / the same time
                            "dram_achieved": (traffic / (phases[dom] * 1e-3) / 1e9) if traffic and phases[dom] > 0 else None,
                            "note": "achieved = bytes the launch is asked to move (projection kernels: gathered rows x 4D "
                                    "- with the unique-row step the DISTINCT catalog rows of the batch, not the 2B slots; "
                                    "the per-triple accounting of step_roofline gives no such credit) / CUDA-event kernel "
                                    "time (single-stream timed entry point); dram_achieved = ncu DRAM bytes of the same "
                                    "launch / the same time"}
    step_roof = {"achieved": B * bpt / world / (step_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                 "bytes_per_triple": bpt, "note": "algorithmic bytes of the whole step per GPU / step time"}
    step_roof["frac"] = step_roof["achieved"] / hbm
    line["step_roofline"] = step_roof

    # ---- evaluation: users/s full-catalog top-k, train items masked (all users) --------------------
    if not args.no_eval:
        st = data.device_state(str(dev))
        e.flush()
        torch.cuda.synchronize()

        grp = parallel.DistGroup() if world > 1 else None

        def sweep_item_shards():
            # every rank: all users x its item shard -> all-to-all by user slice -> merge
            e.theta(refresh=True)
            if world > 1:
                return parallel.sharded_topk([e], grp, st["row_ptr"], st["col_sorted"], args.top_k)
            return e.score_topk(st["row_ptr"], st["col_sorted"], args.top_k)

        def sweep_user_slices():
            # theta of the own shard -> all-gather of (Gi|Bi, theta) -> every rank: its users x the catalog
            e.theta(refresh=True)
            return parallel.user_sliced_topk([e], grp, st["row_ptr"], st["col_sorted"], args.top_k)

        def timed(fn):
            for _ in range(2):
                fn()                                      # warm-up (workspace allocation)
            barrier()
            best = 1e30
            for _ in range(3):
                a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b2.record()
                barrier()
                best = min(best, max_over_ranks(a.elapsed_time(b2)))
            return best

        nu = args.users
        flops_user = 2.0 * args.items * (K + d) + 2.0 * args.items
        modes = {"item_shards": timed(sweep_item_shards)}
        if world > 1:
            modes["user_slices"] = timed(sweep_user_slices)
        mode = min(modes, key=modes.get)
        best = modes[mode]
        tfl = nu * flops_user / (best * 1e-3) / 1e12
        line["eval"] = {"metric": "users/s full-catalog top-%d eval" % args.top_k, "value": nu / best * 1e3,
                        "unit": "users/s", "users": nu, "ms": best, "scaling": "weak (users per GPU fixed)",
                        "kernel": "k_topk_tc (tcgen05 bf16 filter + exact fp32 re-scoring)" if args.tensor_cores
                        else "k_score_topk (fp32 CUDA cores)",
                        "fallback_rows": getattr(e, "tc_overflow_rows", 0),
                        "decomposition": mode if world > 1 else "1 GPU",
                        "ms_by_decomposition": modes,
                        "includes": "theta = F*E projection of the catalog (F stays item-sharded), score sweep, "
                                    "mask, top-k" + ("; item_shards: all-to-all exchange + merge of per-shard lists; "
                                                     "user_slices: all-gather of the item operands" if world > 1 else ""),
                        "roofline": {"bound": "tensor", "achieved": tfl, "peak": tf_burst * world, "unit": "TFLOP/s",
                                     "frac": tfl / (tf_burst * world)}}

    # ---- Evaluator.eval's device work: rank counts of two held-out items per user (validation + test)
    if not args.no_eval and world == 1:
        held = torch.randint(0, args.items, (args.users, 2), device=dev, dtype=torch.int32)
        who = torch.arange(args.users, device=dev, dtype=torch.int32).repeat_interleave(2)
        thr = e.score_pairs(who, held.reshape(-1).contiguous()).reshape(args.users, 2).contiguous()

        def t_ms(fn, reps):
            fn(); torch.cuda.synchronize()
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                out = fn()
            b2.record(); torch.cuda.synchronize()
            return a.elapsed_time(b2) / reps, out

        ms_new, c_new = t_ms(lambda: e.rank_counts(st["row_ptr"], st["col_sorted"], thr), 3)
        ms_old, c_old = t_ms(lambda: e.score_topk(st["row_ptr"], st["col_sorted"], 1, thr_scores=thr)[2], 1)
        fl = 2.0 * args.users * args.items * (K + d)
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12            # CUDA-core fp32 FMA peak at the boost clock
        line["eval"]["rank_counts"] = {
            "what": "Evaluator.eval device work: #items scoring >= each of 2 held-out items per user, train items "
                    "excluded (fvx_rank_counts, exact fp32 register-tiled sweep)",
            "ms": ms_new, "users_per_s": args.users / ms_new * 1e3, "tflops_fp32": fl / (ms_new * 1e-3) / 1e12,
            "frac_of_fp32_fma_peak": fl / (ms_new * 1e-3) / 1e12 / fp32_peak,
            "ms_list_kernel": ms_old, "identical_counts": bool(torch.equal(c_new, c_old))}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        F_host = None
        if D:
            F_host = make_features_device(args.items, D, dev).cpu().numpy()
        v, n, el = cpu_oracle_rate(args, inter, F_host, args.cpu_seconds, B)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
                                "sample": "%d steps of B=%d in %.1f s on the full tables (NumPy oracle of the "
                                          "reference step with TF-2.3 dense Keras-Adam)" % (n, B, el)}
    # ---- the reference's evaluation on the host cores (port): predict_all + store_recommendation's
    # mask / top-k (Evaluator.py:225-239) and _eval_by_user (:82-128) on a slice of users (the full [U, I]
    # matrix of this config is 16 GB and the per-user Python loops take minutes) --------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_eval:
        try:
            from oracle import bpr as obpr, evaluator as oev
            n_s = 256
            Pq = {k_: v for k_, v in e.params().items()}
            Fh = F_host if D else None
            tr_ptr, tr_col = inter.row_ptr, inter.col_file
            tr = [np.sort(tr_col[tr_ptr[u]:tr_ptr[u + 1]]).tolist() for u in range(n_s)]
            held = [[int(inter.test[u])] if np.ndim(inter.test) == 1 else list(inter.test[u]) for u in range(n_s)]
            t0_ = time.perf_counter()
            Sc = obpr.predict_all(Pq, Fh, users=np.arange(n_s))
            oev.masked_topk(Sc, tr, args.top_k)
            t1_ = time.perf_counter()
            for u in range(n_s):
                oev.eval_by_user(Sc[u], args.items, tr[u], held[u], args.top_k)
            t2_ = time.perf_counter()
            line["eval"]["cpu_baseline"] = {
                "kind": "port", "cores": len(os.sched_getaffinity(0)), "users": n_s,
                "topk_users_per_s": n_s / (t1_ - t0_), "eval_users_per_s": n_s / (t2_ - t1_),
                "sample": "%d users x %d items: NumPy predict_all + masked top-%d (store_recommendation), then "
                          "_eval_by_user per user for one held-out item" % (n_s, args.items, args.top_k)}
        except Exception as ex:  # a reported baseline must never cost the bench line
            line["eval"]["cpu_baseline"] = {"kind": "port", "error": repr(ex)}

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_fvx(a)
