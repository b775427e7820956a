"""CPU (gloo, world_size 2): the host-side plumbing of the item-sharded paths - shard bounds, run
ids, the all-to-all layout of the per-shard top-k exchange - with real process groups."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_the_catalog():
    from fvx.parallel import shard_bounds
    for I, R in ((10, 3), (100000, 8), (7, 8), (500000, 4)):
        b = [shard_bounds(I, R, r) for r in range(R)]
        assert b[0][0] == 0 and sum(c for _, c in b) == I
        assert all(b[r][0] + b[r][1] == b[r + 1][0] for r in range(R - 1))
        assert max(c for _, c in b) - min(c for _, c in b) <= 1


def test_run_ids():
    from fvx.parallel import run_ids
    u = torch.tensor([5, 5, 5, 2, 2, 9, 5, 5], dtype=torch.int32)
    assert run_ids(u).tolist() == [0, 0, 0, 1, 1, 2, 3, 3]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, U, I, k, seed, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fvx.parallel import DistGroup, exchange_topk, shard_bounds, user_slices
    g = torch.Generator().manual_seed(seed)
    S = torch.rand(U, I, generator=g)                                   # the same score matrix on every rank
    lo, cnt = shard_bounds(I, world, rank)
    sc, idx = torch.topk(S[:, lo:lo + cnt], k, dim=1)                   # this shard's list, global ids
    ids = (idx + lo).to(torch.int32)
    xi, xs = exchange_topk([ids], [sc], DistGroup())
    per, _ = user_slices(U, world)
    assert xi[0].shape == (per, world, k)
    # merge on the host and compare with the global top-k of this rank's user slice
    flat_s = xs[0].reshape(per, world * k)
    flat_i = xi[0].reshape(per, world * k)
    top_s, pos = torch.topk(flat_s, k, dim=1)
    top_i = torch.gather(flat_i, 1, pos)
    u0, u1 = rank * per, min(U, (rank + 1) * per)
    want_s, want_i = torch.topk(S[u0:u1], k, dim=1)
    ok = bool(torch.equal(top_i[:u1 - u0].long(), want_i) and torch.equal(top_s[:u1 - u0], want_s))
    pad_ok = bool((xi[0][u1 - u0:] == -1).all())
    t = torch.tensor([1.0 if ok and pad_ok else 0.0])
    dist.all_reduce(t)                                                  # all ranks agree
    if rank == 0:
        out.put(float(t.item()))
    dist.destroy_process_group()


def test_topk_exchange_layout_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, 400, 10, 3, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 2.0
