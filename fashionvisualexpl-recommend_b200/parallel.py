"""One process per GPU: the item catalog (Gi, Bi, F and their Adam state) row-sharded over the
ranks of a ``torch.distributed`` group, user tables and E replicated.

* ``shard_bounds``            contiguous item block of a rank;
* ``ShardedStep``             the BPR step over the three C-ABI phases of include/fvx.h with the
                              all-reduces between them (S: 2B floats; RU: packed user-row
                              gradients; dE: D x de) - NCCL over NVLink on a GPU box;
* ``exchange_topk`` / ``sharded_topk``   evaluation: every rank sweeps all users over its shard,
                              the per-shard top-k lists are exchanged by user slice
                              (all-to-all) and merged there (``fvx_topk_merge``).

``world`` may also be *emulated* on one GPU (``LocalGroup``): the ranks are engines in one
process and the collectives plain tensor sums - used by the GPU tests, since several ranks of one
job must never be separate launches on one GPU.
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

    def all_reduce(self, tensors):
        for t in tensors:
            self.dist.all_reduce(t, group=self.group)

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

    def __init__(self, engines, group, max_runs=None):
        """``max_runs``: upper bound of the number of runs of equal users in a batch (default: the batch
        size); the all-reduced user-gradient buffer has that many rows, so a tight bound - e.g.
        B // (shortest train list) + 2 for batches of the reference's sampler - saves NVLink bytes."""
        self.engines, self.group = list(engines), group
        e = self.engines[0]
        B, dv = e.max_batch, e.device
        self.max_runs = int(max_runs) if max_runs else B
        self.S = [torch.zeros(2 * B, dtype=torch.float32, device=dv) for _ in self.engines]
        self.RU = [torch.zeros(self.max_runs, e.Su, dtype=torch.float32, device=dv) for _ in self.engines]
        self.dE = [torch.zeros(e.D, e.de, dtype=torch.float32, device=dv) if e.D else None for _ in self.engines]
        self.rid = torch.zeros(B, dtype=torch.int32, device=dv)
        self.rid_scratch = torch.zeros(B // 4096 + 2, dtype=torch.int32, device=dv)

    def _rank_of(self, i):
        return self.group.rank if self.group.rank is not None else i

    def step(self, user, pos, neg, loss_slot=0):
        """One optimiser step on every (local) rank; asynchronous apart from the collectives."""
        B = user.numel()
        # run ids of the batch (needed from phase B on): two small launches through the C ABI - the step is
        # ~20 host calls and every torch op saved shortens the launch-bound tail on a busy host
        rid = self.rid[:B]
        call("fvx_run_ids", ptr(user), B, ptr(rid), ptr(self.rid_scratch), stream_ptr())
        for e, S in zip(self.engines, self.S):
            call("fvx_bpr_step_sharded_a", C.byref(e.struct()), ptr(user), ptr(pos), ptr(neg), B, ptr(S), stream_ptr())
        self.group.all_reduce(self.S)
        for e, S, RU in zip(self.engines, self.S, self.RU):
            call("fvx_bpr_step_sharded_b1", C.byref(e.struct()), ptr(user), B, ptr(S), ptr(rid), ptr(RU), RU.shape[0],
                 loss_slot, stream_ptr())
        # the all-reduce of the user-row gradients (the only large message) travels while grad_E runs
        pending = self.group.all_reduce_start(self.RU)
        if self.dE[0] is not None:
            for e, dE in zip(self.engines, self.dE):
                call("fvx_bpr_step_sharded_b2", C.byref(e.struct()), B, ptr(dE), stream_ptr())
            self.group.all_reduce(self.dE)
        self.group.all_reduce_finish(pending)
        for i, (e, RU, dE) in enumerate(zip(self.engines, self.RU, self.dE)):
            call("fvx_bpr_step_sharded_c", C.byref(e.struct()), ptr(user), B, ptr(rid), ptr(RU), RU.shape[0],
                 ptr(dE), loss_slot if self._rank_of(i) == 0 else -1, stream_ptr())

    def take_loss(self, slot=0):
        """Batch loss (sum of the per-rank parts) as a 1-element device tensor of the first local rank;
        the accumulators are cleared.  Stream-ordered, no host synchronisation (run overflow is reported
        by ``read_loss`` / ``check_runs``)."""
        parts = [e.take_loss(slot) for e in self.engines]
        self.group.all_reduce(parts)
        return parts[0]

    def check_runs(self):
        if any(int(e.sync_t[2].item()) for e in self.engines):
            raise _lib.FvxError("a batch had more runs of equal users than max_runs=%d" % self.max_runs)

    def read_loss(self, slot=0, clear=True):
        """Batch loss = sum of the per-rank partial losses (synchronises)."""
        if any(int(e.sync_t[2].item()) for e in self.engines):
            raise _lib.FvxError("a batch had more runs of equal users than max_runs=%d" % self.max_runs)
        parts = [e.loss_t[slot:slot + 1].clone() for e in self.engines]
        if clear:
            for e in self.engines:
                e.loss_t[slot] = 0
        self.group.all_reduce(parts)
        return float(parts[0].item())


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
    ``(ids, scores)`` of that rank's user slice ``[U/R (padded), k]``."""
    from .engine import topk_merge
    ids, scores = [], []
    for e in engines:
        i_, s_ = e.score_topk(mask_row_ptr, mask_col, k, tc=tc)
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
    e0 = engines[0]
    I = e0.I
    cnts = [shard_bounds(I, R, r)[1] for r in range(R)]
    mx = max(cnts)
    views, ins_w, ins_t, outs_w, outs_t = [], [], [], [], []
    for e in engines:
        e.flush()
        w = torch.zeros(mx, e.Si, dtype=torch.float32, device=e.device)
        w[:e.Ic] = e.items["w"]
        ins_w.append(w)
        outs_w.append(torch.empty(R, mx, e.Si, dtype=torch.float32, device=e.device))
        if e.D:
            t = torch.zeros(mx, e.de, dtype=torch.float32, device=e.device)
            t[:e.Ic] = e.theta()
            ins_t.append(t)
            outs_t.append(torch.empty(R, mx, e.de, dtype=torch.float32, device=e.device))
    group.all_gather(outs_w, ins_w)
    if ins_t:
        group.all_gather(outs_t, ins_t)
    for i, e in enumerate(engines):
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
