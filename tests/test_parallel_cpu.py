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


def _step_worker(rank, world, port, out):
    """Two gloo ranks run the item-sharded step of oracle/sharded.py (first half: replicated users, the phases of
    fvx_bpr_step_sharded_phase with S, RU and dE summed) with real all-reduces; every rank must end with the parameters of the
    single-rank oracle step, the item rows on their owner."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fvx.parallel import shard_bounds
    from oracle import bpr, sharded
    U, I, K, d, D, B, steps, reg, lr = 40, 61, 8, 5, 12, 48, 4, 1e-3, 0.01
    rng = np.random.default_rng(5)                                      # the same problem on every rank
    P0 = bpr.init_params(U, I, K, d, D, seed=1, dtype=np.float64)
    P0["Bi"] = rng.standard_normal(I) * 0.1
    F = np.maximum(rng.standard_normal((I, D)), 0)
    lo, cnt = shard_bounds(I, world, rank)
    P = {k: v.copy() for k, v in P0.items()}                            # users, E replicated; items: own rows matter
    Q = {k: v.copy() for k, v in P0.items()}                            # single-rank oracle
    S_ad, SQ = bpr.init_adam(P), bpr.init_adam(Q)
    ok = True
    for s in range(steps):
        u = np.repeat(rng.integers(0, U, B // 4), 4)
        batch = (u, rng.integers(0, I, B), rng.integers(0, I, B))
        want_loss = bpr.train_step(Q, SQ, batch, reg, lr, F)
        St = torch.from_numpy(sharded.phase_a(P, lo, cnt, batch, F))
        dist.all_reduce(St)
        n_runs = int(sharded.run_ids(u)[-1]) + 1
        G_items, RU, dE, dBp, loss = sharded.phase_b(P, lo, cnt, batch, St.numpy(), reg, n_runs, F)
        RUt, dEt, dBt = torch.from_numpy(RU), torch.from_numpy(dE), torch.from_numpy(dBp)
        for t in (RUt, dEt, dBt):
            dist.all_reduce(t)
        G, extra = sharded.phase_c_grads(P, batch, G_items, RUt.numpy(), dEt.numpy(), dBt.numpy(), reg,
                                         add_e_reg=(rank == 0))
        lt = torch.tensor([float(loss + extra)], dtype=torch.float64)
        dist.all_reduce(lt)
        ok = ok and abs(float(lt.item()) - want_loss) <= 1e-9 * abs(want_loss)
        bpr.adam_apply(P, S_ad, G, lr)        # dense-semantics Adam: foreign item rows see a zero gradient here
    own = slice(lo, lo + cnt)
    for k in ("Gu", "Tu", "E", "Bp"):
        ok = ok and np.allclose(P[k], Q[k], rtol=1e-10, atol=1e-13)
    for k in ("Gi", "Bi"):
        ok = ok and np.allclose(P[k][own], Q[k][own], rtol=1e-10, atol=1e-13)
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t)
    if rank == 0:
        out.put(float(t.item()))
    dist.destroy_process_group()


def test_sharded_step_decomposition_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_step_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 2.0


def test_user_bounds_and_run_slots_layout():
    """Host logic of the owned-users layout: fvx.parallel.user_bounds and the oracle's restatement of fvx_run_slots."""
    from fvx.parallel import user_bounds
    from oracle import sharded
    for U, R in ((10, 3), (40000, 8), (7, 8), (1000001, 4)):
        b = [user_bounds(U, R, r) for r in range(R)]
        per = b[0][2] // R
        assert all(x[2] == per * R for x in b) and per * R >= U
        assert sum(c for _, c, _ in b) == U and all(lo == min(U, r * per) for r, (lo, _, _) in enumerate(b))
        assert [sharded.user_bounds(U, R, r)[:2] for r in range(R)] == [x[:2] for x in b]
    slot = sharded.run_slots([5, 5, 1, 1, 9, 5], per=5, owners=2, cap=3)
    assert slot.tolist() == [3, 3, 0, 0, 4, 5]                 # owner 1: runs (5), (9), (5 again); owner 0: run (1)
    try:
        sharded.run_slots([5, 6, 7, 8], per=5, owners=2, cap=3)
        assert False, "four runs of owner 1 do not fit a capacity of three"
    except OverflowError:
        pass


def _owned_worker(rank, world, port, out):
    """The sharded step with block-OWNED users (oracle/sharded.py, second half: the four exchanges of
    fvx_bpr_step_sharded) in two gloo ranks.  Every rank's copies of the users it does not own are NaN from the start:
    scores and gradients may only read the rows the owners publish."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fvx.parallel import shard_bounds
    from oracle import bpr, sharded
    U, I, K, d, D, B, steps, reg, lr = 41, 61, 8, 5, 12, 48, 4, 1e-3, 0.01
    rng = np.random.default_rng(6)
    P0 = bpr.init_params(U, I, K, d, D, seed=2, dtype=np.float64)
    P0["Bi"] = rng.standard_normal(I) * 0.1
    F = np.maximum(rng.standard_normal((I, D)), 0)
    lo, cnt = shard_bounds(I, world, rank)
    ulo, ucnt, per = sharded.user_bounds(U, world, rank)
    P = {k: v.copy() for k, v in P0.items()}
    Q = {k: v.copy() for k, v in P0.items()}
    foreign = np.ones(U, dtype=bool)
    foreign[ulo:ulo + ucnt] = False
    P["Gu"][foreign] = np.nan
    P["Tu"][foreign] = np.nan
    S_ad, SQ = bpr.init_adam(P), bpr.init_adam(Q)
    cap = B // 4 + 2
    ok = True
    for s in range(steps):
        u = np.repeat(rng.integers(0, U, B // 4), 4)
        batch = (u, rng.integers(0, I, B), rng.integers(0, I, B))
        want_loss = bpr.train_step(Q, SQ, batch, reg, lr, F)
        slot = sharded.run_slots(u, per, world, cap)
        WU = torch.from_numpy(sharded.publish_users(P, ulo, ucnt, batch, slot, world, cap, True))
        dist.all_reduce(WU)                              # disjoint segments: the sum is the all-gather
        ok = ok and bool(torch.isfinite(WU[np.unique(slot)]).all())
        St = torch.from_numpy(sharded.owned_scores(P, lo, cnt, batch, WU.numpy(), slot, F))
        dist.all_reduce(St)
        G_items, RU, dE, dBp, loss = sharded.owned_grads(P, lo, cnt, batch, St.numpy(), WU.numpy(), slot, reg, F)
        RUt, dEt, dBt = torch.from_numpy(RU), torch.from_numpy(dE), torch.from_numpy(dBp)
        for t in (RUt, dEt, dBt):
            dist.all_reduce(t)                           # (RU: every owner only reads its own segment)
        gGu, gTu = sharded.owned_user_grads(P, ulo, ucnt, batch, slot, RUt.numpy(), True)
        two = 2.0
        G = {"Gu": gGu, "Tu": gTu, "Gi": G_items["Gi"], "Bi": G_items["Bi"],
             "E": dEt.numpy() + two * reg * P["E"], "Bp": dBt.numpy() + two * reg * P["Bp"]}
        extra = reg * (np.sum(P["E"] * P["E"]) + np.sum(P["Bp"] * P["Bp"])) if rank == 0 else 0.0
        lt = torch.tensor([float(loss + extra)], dtype=torch.float64)
        dist.all_reduce(lt)
        ok = ok and abs(float(lt.item()) - want_loss) <= 1e-9 * abs(want_loss)
        with np.errstate(all="ignore"):                  # the NaN rows of the foreign users stay NaN
            bpr.adam_apply(P, S_ad, G, lr)
    ok = ok and bool(np.isnan(P["Gu"][foreign]).all())   # nobody ever wrote a foreign user's row either
    # the owners' rows, gathered (parallel.gather_users), are the single-rank oracle's
    for k in ("Gu", "Tu"):
        mine = torch.from_numpy(np.where(foreign[:, None], 0.0, P[k]))
        dist.all_reduce(mine)
        ok = ok and np.allclose(mine.numpy(), Q[k], rtol=1e-10, atol=1e-13)
    for k in ("E", "Bp"):
        ok = ok and np.allclose(P[k], Q[k], rtol=1e-10, atol=1e-13)
    own = slice(lo, lo + cnt)
    for k in ("Gi", "Bi"):
        ok = ok and np.allclose(P[k][own], Q[k][own], rtol=1e-10, atol=1e-13)
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t)
    if rank == 0:
        out.put(float(t.item()))
    dist.destroy_process_group()


def test_owned_users_decomposition_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_owned_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 2.0
