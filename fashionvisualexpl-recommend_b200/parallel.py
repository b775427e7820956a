"""One process per GPU: the item catalog (Gi, Bi, F and their Adam state) row-sharded over the
ranks of a ``torch.distributed`` group, the USERS owned in contiguous blocks (the owner keeps a user's Adam
state and publishes the user's row to the others for the steps that touch it), E replicated.

* ``shard_bounds`` / ``user_bounds``   contiguous item block / user block of a rank;
* ``sharded_engine``          an ``Engine`` holding one rank's part;
* ``ShardedStep``             the BPR step: ONE C-ABI call per rank per step (``fvx_bpr_step_sharded``) that
                              issues its four exchanges itself - WU (fresh user rows) beside the projection
                              and RU (user-row gradient shares) beside grad_E on a side stream, S (partial
                              scores) and dE on the caller's - either as NCCL collectives over communicators
                              created from an id that ``torch.distributed`` broadcasts, or (default) as stores
                              into the peers' buffers of a CUDA-IPC arena with one-warp barrier kernels;
* ``exchange_topk`` / ``sharded_topk``   evaluation: every rank sweeps all users over its shard, the per-shard
                              top-k lists are exchanged by user slice (all-to-all) and merged there
                              (``fvx_topk_merge``).

``world`` may also be *emulated* on one GPU (``LocalGroup``): the ranks are engines in one process, the step
runs phase by phase (``fvx_bpr_step_sharded_phase``) and the collectives are plain tensor sums - used by the
GPU tests, since several ranks of one job must never be separate launches on one GPU.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def shard_bounds(num_items, world, rank):
    """(item_lo, item_cnt) of ``rank``: contiguous blocks, the first ``num_items % world`` one longer."""
    base, rem = divmod(int(num_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, base + (1 if rank < rem else 0)


def user_bounds(num_users, world, rank):
    """(user_lo, user_cnt, user_rows) of ``rank``: equal blocks of ceil(U / world) users (the last ones shorter or
    empty); ``user_rows`` = world * block, the rows every rank allocates so that the blocks gather in place."""
    per = (int(num_users) + world - 1) // world
    lo = min(int(num_users), rank * per)
    return lo, max(0, min(int(num_users), lo + per) - lo), per * world


def sharded_engine(world, rank, num_users, num_items, K, **kw):
    """The ``Engine`` of one rank of a ``world``-rank job (item shard + user block of that rank)."""
    from .engine import Engine
    lo, cnt = shard_bounds(num_items, world, rank)
    ulo, ucnt, urows = user_bounds(num_users, world, rank)
    return Engine(num_users, num_items, K, item_lo=lo, item_cnt=cnt, user_lo=ulo, user_cnt=ucnt, user_rows=urows,
                  sharded=True, **kw)


def run_ids(user):
    """run_id[b] of the batch (int32): runs of equal consecutive users, counted from 0."""
    start = torch.ones_like(user, dtype=torch.int32)
    start[1:] = (user[1:] != user[:-1]).to(torch.int32)
    return (torch.cumsum(start, 0, dtype=torch.int32) - 1).contiguous()


class DistGroup:
    """Collectives of a real ``torch.distributed`` process group (one engine per process)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self._comm = None

    def comm(self, device):
        """The ``FvxComm`` of this rank (two NCCL communicators owned by libfvx), created on first use: rank 0
        draws the id, ``torch.distributed`` broadcasts its 256 bytes, every rank joins (collective)."""
        if self._comm is None:
            buf = (C.c_uint8 * _lib.COMM_ID_BYTES)()
            if self.rank == 0:
                call("fvx_comm_unique_id", buf)
            t = torch.tensor(list(buf), dtype=torch.uint8, device=device)
            self.dist.broadcast(t, src=self.dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                group=self.group)
            raw = bytes(t.cpu().tolist())
            idb = (C.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(raw)
            out = C.c_void_p()
            with torch.cuda.device(device):
                call("fvx_comm_create", idb, self.rank, self.world, C.byref(out))
            self._comm = out
        return self._comm

    def close(self):
        if self._comm is not None:
            call("fvx_comm_destroy", self._comm)
            self._comm = None

    def all_reduce(self, tensors):
        for t in tensors:
            self.dist.all_reduce(t, group=self.group)

    def all_reduce_max(self, tensors):
        for t in tensors:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)

    def all_reduce_start(self, tensors):
        """Starts the all-reduce on NCCL's own stream (it waits for the work already queued on the
        current stream); ``all_reduce_finish`` makes the current stream wait for the result."""
        return [self.dist.all_reduce(t, group=self.group, async_op=True) for t in tensors]

    def all_reduce_finish(self, works):
        for w in works:
            w.wait()

    def all_to_all(self, outs, ins):
        self.dist.all_to_all_single(outs[0], ins[0], group=self.group)

    def all_gather(self, outs, ins):
        """outs[0]: [R, ...] receives ins[0] of every rank."""
        self.dist.all_gather_into_tensor(outs[0], ins[0], group=self.group)


class LocalGroup:
    """R emulated ranks in one process: ``tensors[r]`` is rank r's buffer."""

    def __init__(self, world):
        self.world, self.rank = int(world), None

    def all_reduce(self, tensors):
        total = torch.stack(list(tensors)).sum(0)
        for t in tensors:
            t.copy_(total)

    def all_reduce_max(self, tensors):
        total = torch.stack(list(tensors)).max(0).values
        for t in tensors:
            t.copy_(total)

    def all_reduce_start(self, tensors):
        self.all_reduce(tensors)
        return []

    def all_reduce_finish(self, works):
        pass

    def all_gather(self, outs, ins):
        full = torch.stack(list(ins))
        for o in outs:
            o.copy_(full)

    def all_to_all(self, outs, ins):
        R = self.world
        for dst in range(R):
            chunks = [ins[src].reshape(R, -1)[dst] for src in range(R)]
            outs[dst].copy_(torch.stack(chunks).reshape(outs[dst].shape))


class ShardedStep:
    """``engines``: this process's engines - ONE with a ``DistGroup``, or all R with a ``LocalGroup``."""

    def __init__(self, engines, group, max_runs=None, transport=None):
        """``transport`` (real process groups): "p2p" (default: the kernels of the step store into their peers'
        exchange buffers over NVLink - a CUDA-IPC arena - and one-warp barrier kernels replace the collectives) or
        "nccl" (all-gather / all-reduce / reduce-scatter / all-reduce inside the C call).

        ``max_runs``: upper bound of the number of runs of equal users in a batch.  Default: the batch size
        (the hard bound).  The user rows (WU) and user-gradient shares (RU) exchanged per step are laid out in one
        segment per owner of ``run_cap`` = max_runs / R (+ 15 % + 64 for the spread of the owners' shares) rows, so
        a tight bound - B // (shortest train list) + 2 is exact for batches of the reference's sampler - saves
        NVLink bytes; a batch in which an owner has more runs poisons that step's loss with NaN on every rank."""
        self.engines, self.group = list(engines), group
        e = self.engines[0]
        B, R = e.max_batch, group.world
        if R > 8:
            raise _lib.FvxError("the sharded step supports up to 8 ranks")
        self.max_runs = int(max_runs) if max_runs else B
        self.run_cap = self.max_runs if R == 1 else min(self.max_runs, int(1.15 * self.max_runs / R) + 64)
        rows = R * self.run_cap
        per = e.U_rows // R
        self.ws = []
        for e in self.engines:
            dv = e.device
            f32, i32 = dict(dtype=torch.float32, device=dv), dict(dtype=torch.int32, device=dv)
            t = {"S": torch.zeros(2 * B, **f32), "run_id": torch.zeros(B, **i32),
                 "run_scratch": torch.zeros(8 * (B // 1024 + 2), **i32), "WU": torch.zeros(rows, e.Su, **f32),
                 "RU": torch.zeros(rows, e.Su, **f32), "dE": torch.zeros(e.D * e.de + 4, **f32),
                 "loss_part": torch.zeros(1, dtype=torch.float64, device=dv)}
            t["run_user"] = torch.zeros(self.run_cap, **i32)
            t["run_counts"] = torch.zeros(8, **i32)
            w = _lib.FvxShardWs(ptr(t["S"]), ptr(t["run_id"]), ptr(t["run_scratch"]), ptr(t["WU"]), ptr(t["RU"]),
                                ptr(t["dE"]), ptr(t["loss_part"]), rows, self.run_cap, R, max(per, 1))
            w.run_user, w.run_counts = ptr(t["run_user"]), ptr(t["run_counts"])
            t["struct"] = w
            self.ws.append(t)
        self._comm = group.comm(self.engines[0].device) if isinstance(group, DistGroup) else None
        import os
        asked = transport or os.environ.get("FVX_SHARDED_TRANSPORT")
        self.transport = asked or ("p2p" if self._comm is not None and R > 1 else "nccl")
        if self.transport == "p2p" and not asked:
            # the default: every rank tries to map its peers' arenas; unless ALL succeed (CUDA IPC can be refused,
            # e.g. across containers) every rank takes the NCCL transport - the ranks must agree
            err = None
            try:
                self._map_arena(rows, B, R)
            except _lib.FvxError as ex:
                err = ex
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.engines[0].device)
            group.dist.all_reduce(ok, op=group.dist.ReduceOp.MIN, group=group.group)
            if int(ok.item()) == 0:
                import sys
                print("fvx: peer-mapped exchange buffers unavailable (%s); the sharded step uses NCCL collectives"
                      % (err or "a peer failed"), file=sys.stderr)
                w = self.ws[0]["struct"]
                w.p2p = 0
                for name in ("WU", "S"):
                    setattr(w, name, ptr(self.ws[0][name]))
                self.transport = "nccl"
        elif self.transport == "p2p":
            self._map_arena(rows, B, R)

    def _map_arena(self, rows, B, R):
        if self._comm is None:
            raise _lib.FvxError("the peer-to-peer transport needs a real process group")
        # the exchange buffers inside the peer-mapped arena, at the same offsets on every rank
        e, w = self.engines[0], self.ws[0]["struct"]
        sizes = (("WU", rows * e.Su * 4), ("S", 2 * B * 4), ("RUin", R * self.run_cap * e.Su * 4),
                 ("dEall", R * max(e.D * e.de, 1) * 4), ("tails", R * 16), ("flags", 4 * 8 * 4))
        off, total = {}, 0
        for name, nbytes in sizes:
            off[name] = total
            total += (nbytes + 255) // 256 * 256
        base = C.c_void_p()
        with torch.cuda.device(e.device):
            call("fvx_comm_arena", self._comm, total, C.byref(base))
        for name, _ in sizes:
            setattr(w, name, base.value + off[name])
        w.p2p = 1          # (the torch copies of WU and S stay allocated: the NCCL transport uses them)

    def step(self, user, pos, neg, loss_slot=0):
        """One optimiser step on every (local) rank; asynchronous."""
        B = user.numel()
        if self._comm is not None:
            e, w = self.engines[0], self.ws[0]
            call("fvx_bpr_step_sharded", C.byref(e.struct()), C.byref(w["struct"]), self._comm, ptr(user), ptr(pos),
                 ptr(neg), B, loss_slot, stream_ptr())
            return
        # emulated ranks: the step cut at its collectives, which are done here
        sums = ((), ("S",), ("RU", "dE"), ())
        for phase in range(4):
            for e, w in zip(self.engines, self.ws):
                call("fvx_bpr_step_sharded_phase", C.byref(e.struct()), C.byref(w["struct"]), ptr(user), ptr(pos),
                     ptr(neg), B, loss_slot, phase, stream_ptr())
            if phase == 0:                       # all-gather: every rank receives every owner's segment of WU
                c = self.run_cap
                segs = [w["WU"][r * c:(r + 1) * c].clone() for r, w in enumerate(self.ws)]
                for w in self.ws:
                    for r, seg in enumerate(segs):
                        w["WU"][r * c:(r + 1) * c].copy_(seg)
            for name in sums[phase]:             # (the sum of RU over the ranks contains every owner's reduced segment)
                self.group.all_reduce([w[name] for w in self.ws])

    def take_loss(self, slot=0):
        """Batch loss as a 1-element device tensor of the first local rank (every rank holds the whole loss: the
        ranks' shares travel with dE); the accumulators are cleared.  Stream-ordered, no host synchronisation.
        NaN: the batch had more runs of equal users than ``max_runs``."""
        parts = [e.take_loss(slot) for e in self.engines]
        return parts[0]

    def read_loss(self, slot=0, clear=True):
        """Batch loss (synchronises); raises if the batch had more runs than ``max_runs``."""
        vals = [float(e.loss_t[slot].item()) for e in self.engines]
        if clear:
            for e in self.engines:
                e.loss_t[slot] = 0
        if any(v != v for v in vals):
            raise _lib.FvxError("a batch had more runs of equal users per owner than run_cap=%d (max_runs=%d)"
                                % (self.run_cap, self.max_runs))
        return vals[0]

    def sync_users(self):
        """Flushes deferred optimiser state and gathers the user rows from their owners, so that every rank
        holds every user's current row (evaluation sweeps all users on every rank; ``Engine.params()``)."""
        gather_users(self.engines, self.group)


def gather_users(engines, group):
    R = group.world
    for e in engines:
        e.flush()
    if R == 1:
        return
    per = engines[0].U_rows // R
    if isinstance(group, DistGroup):
        e = engines[0]
        w = e.users["w"]
        own = w[group.rank * per:(group.rank + 1) * per].clone()
        group.dist.all_gather_into_tensor(w, own, group=group.group)
    else:
        blocks = [e.users["w"][r * per:(r + 1) * per].clone() for r, e in enumerate(engines)]
        for e in engines:
            for r, blk in enumerate(blocks):
                e.users["w"][r * per:(r + 1) * per].copy_(blk)


# ---- evaluation ---------------------------------------------------------------------------------
def user_slices(num_users, world):
    """Users padded to a multiple of ``world``: (slice length, padded count)."""
    per = (int(num_users) + world - 1) // world
    return per, per * world


def exchange_topk(ids, scores, group):
    """Per-shard lists ``[U, k]`` of every local rank -> for each local rank r the lists of ITS user
    slice from every shard: ``[U/R, R, k]`` (ids global; padding users carry id -1 / -inf)."""
    R = group.world
    outs_i, outs_s, ins_i, ins_s = [], [], [], []
    for i_, s_ in zip(ids, scores):
        U, k = i_.shape
        per, Up = user_slices(U, R)
        pi = torch.full((Up, k), -1, dtype=torch.int32, device=i_.device)
        ps = torch.full((Up, k), float("-inf"), dtype=torch.float32, device=i_.device)
        pi[:U], ps[:U] = i_, s_
        ins_i.append(pi.reshape(R, per, k).contiguous())
        ins_s.append(ps.reshape(R, per, k).contiguous())
        outs_i.append(torch.empty(R, per, k, dtype=torch.int32, device=i_.device))
        outs_s.append(torch.empty(R, per, k, dtype=torch.float32, device=i_.device))
    group.all_to_all(outs_i, ins_i)
    group.all_to_all(outs_s, ins_s)
    return ([o.permute(1, 0, 2).contiguous() for o in outs_i], [o.permute(1, 0, 2).contiguous() for o in outs_s])


def sharded_topk(engines, group, mask_row_ptr, mask_col, k, tc=None):
    """Full-catalog masked top-k with the catalog sharded: returns, per local rank, the merged
    ``(ids, scores)`` of that rank's user slice ``[U/R (padded), k]``.

    Tensor-core path: every rank first runs the BOUNDS sweep of all users over its shard; a shard's bound is a
    valid lower bound of the (k + #train)-th best score of the whole catalog, so one all-reduce (MAX, 4 bytes per
    user) gives every rank the best bound any shard found, and the candidates sweep then keeps, per shard, only
    what can reach the global top-k: candidate lists, exact re-scoring and list lengths shrink with the number of
    shards.  Then the per-shard lists are exchanged by user slice (all-to-all) and merged (``fvx_topk_merge``)."""
    from .engine import topk_merge
    gather_users(engines, group)
    ids, scores = [], []
    use_tc = all((e.use_tensor_cores if tc is None else tc) and e.tc_eval_eligible() for e in engines)
    if use_tc:
        thr = [e.topk_bounds(mask_row_ptr, k) for e in engines]
        group.all_reduce_max(thr)
        for e in engines:
            i_, s_ = e.topk_select(mask_row_ptr, mask_col, k)
            ids.append(i_)
            scores.append(s_)
    else:
        for e in engines:
            i_, s_ = e.score_topk(mask_row_ptr, mask_col, k, tc=False)
            ids.append(i_)
            scores.append(s_)
    xi, xs = exchange_topk(ids, scores, group)
    return [topk_merge(a, b) for a, b in zip(xi, xs)]


def gathered_view(engines, group):
    """Evaluation with the USERS split over the ranks instead of the items: the item-side operands of
    the score (Gi | Bi rows and theta = F*E of every shard: (Si + de) * 4 bytes per item - F itself
    stays sharded) are all-gathered once per sweep, then every rank scores its own user slice against
    the whole catalog and no merge is needed.  Returns one view per local engine for
    ``Engine.score_topk(view=...)``."""
    R = group.world
    gather_users(engines, group)
    e0 = engines[0]
    I = e0.I
    cnts = [shard_bounds(I, R, r)[1] for r in range(R)]
    mx = max(cnts)
    equal = min(cnts) == mx                 # equal shards: the gathered buffer IS the catalog-ordered operand
    views, ins_w, ins_t, outs_w, outs_t = [], [], [], [], []
    for e in engines:
        e.flush()
        if equal:
            w = e.items["w"]
        else:
            w = torch.zeros(mx, e.Si, dtype=torch.float32, device=e.device)
            w[:e.Ic] = e.items["w"]
        ins_w.append(w)
        outs_w.append(torch.empty(R, mx, e.Si, dtype=torch.float32, device=e.device))
        if e.D:
            if equal:
                t = e.theta()
            else:
                t = torch.zeros(mx, e.de, dtype=torch.float32, device=e.device)
                t[:e.Ic] = e.theta()
            ins_t.append(t)
            outs_t.append(torch.empty(R, mx, e.de, dtype=torch.float32, device=e.device))
    group.all_gather(outs_w, ins_w)
    if ins_t:
        group.all_gather(outs_t, ins_t)
    for i, e in enumerate(engines):
        if equal:
            W = outs_w[i].reshape(R * mx, e.Si)
            T = outs_t[i].reshape(R * mx, e.de) if e.D else None
        else:
            W = torch.cat([outs_w[i][r, :cnts[r]] for r in range(R)]).contiguous()
            T = torch.cat([outs_t[i][r, :cnts[r]] for r in range(R)]).contiguous() if e.D else None
        m = _lib.FvxModel()
        C.pointer(m)[0] = e.struct()
        m.item_lo, m.item_cnt = 0, I
        m.items.w, m.items.rows = ptr(W), I
        views.append({"struct": m, "theta": T, "Ic": I, "keep": [W, T]})
    return views


def user_sliced_topk(engines, group, mask_row_ptr, mask_col, k, tc=None):
    """(ids, scores) of each local rank's user slice [U/R (last slice shorter), k]; see gathered_view."""
    views = gathered_view(engines, group)
    out = []
    for i, (e, v) in enumerate(zip(engines, views)):
        r = group.rank if group.rank is not None else i
        per, _ = user_slices(e.U, group.world)
        u0, u1 = min(e.U, r * per), min(e.U, (r + 1) * per)
        out.append(e.score_topk(mask_row_ptr, mask_col, k, u0=u0, u1=u1, tc=tc, view=v))
    return out
