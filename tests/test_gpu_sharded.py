"""GPU: the sharded paths (fvx_bpr_step_sharded_phase, per-shard top-k + merge) against the single-rank
path and the oracle.  The R ranks are EMULATED on one GPU (fvx.parallel.LocalGroup: R engines in one
process, the step cut at its collectives, collectives = tensor sums), because mutually waiting ranks must
not be separate launches on one GPU; the one-call NCCL path (fvx_bpr_step_sharded) is checked against the
oracle by bench.py --gpus N (parity_checked) and by tests/test_gpu_multi.py on a multi-GPU box."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import bpr, evaluator as oe
from test_gpu_parity import REL, _dev, _engine, _oracle_pair, _random_problem, _user_contiguous_batches

pytestmark = pytest.mark.gpu


def _shards(U, I, K, d, D, R, P, F, **kw):
    from fvx.parallel import sharded_engine
    es = []
    for r in range(R):
        e = sharded_engine(R, r, U, I, K, d=d, D=D, **kw)
        if D:
            e.set_features(F[e.item_lo:e.item_lo + e.Ic])
        e.load_params(P)
        es.append(e)
    return es


def _gather_params(es, R):
    from fvx.parallel import LocalGroup, gather_users
    gather_users(es, LocalGroup(R))                 # user rows live on their owners until gathered
    Ps = [e.params() for e in es]
    out = {k: Ps[0][k] for k in Ps[0] if k not in ("Gi", "Bi")}
    out["Gi"] = np.concatenate([p["Gi"] for p in Ps], 0)
    out["Bi"] = np.concatenate([p["Bi"] for p in Ps], 0)
    return out, Ps


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("R", [2, 3])
@pytest.mark.parametrize("K,d,D,B,mode", [(64, 20, 256, 512, "deferred"), (16, 0, 0, 256, "dense"),
                                            (32, 20, 128, 200, "dense")])
def test_sharded_step_matches_oracle_and_single_rank(K, d, D, B, mode, R, tc):
    from fvx.parallel import LocalGroup, ShardedStep
    if tc and D == 0:
        pytest.skip("BPRMF has no projection")
    U, I, steps, lr, reg = 500, 701, 12, 0.001, 1e-3
    P, F, rng = _random_problem(U, I, K, d, D, seed=K + d + R)
    es = _shards(U, I, K, d, D, R, P, F, lr=lr, reg=reg, adam_mode=mode, max_batch=B, use_tensor_cores=tc)
    step = ShardedStep(es, LocalGroup(R))
    # runs of users as the reference's sampler emits them; now and then a user has two runs in a batch (as
    # across an epoch boundary): both runs must see the row its owner published
    batches = []
    for s_ in range(steps):
        order = rng.permutation(U)[:B // 6 + 1]
        if s_ % 3 == 1:
            order[-1] = order[0]
        u = np.repeat(order, 6)[:B]
        batches.append((u, rng.integers(0, I, B), rng.integers(0, I, B)))
    P32, P64, l32, l64 = _oracle_pair(P, F, batches, reg, lr)
    for s, b in enumerate(batches):
        step.step(*(_dev(x) for x in b), loss_slot=s % 5)
        per_rank = [float(e.loss_t[s % 5].item()) for e in es]
        assert max(per_rank) == min(per_rank), per_rank            # the whole loss on every rank
        got = step.read_loss(s % 5)
        assert got == pytest.approx(l64[s], rel=REL), (s, mode, R)
    Q, Ps = _gather_params(es, R)
    for k in P64:
        ref = P64[k]
        dlt = np.abs(Q[k].reshape(ref.shape) - ref) / np.abs(ref).max()
        assert (dlt > REL).mean() <= 2e-3, (k, float((dlt > REL).mean()))
        assert dlt.max() <= max(20 * REL, 3 * rel_err(P32[k], ref)), (k, float(dlt.max()))
    for name in (("E", "Bp") if D else ()):       # replicated state is bit-identical on every rank
        for p in Ps[1:]:
            assert np.array_equal(p[name], Ps[0][name]), name
    # only the owner keeps a user's optimiser state
    for r, e in enumerate(es):
        own = torch.zeros(e.U_rows, dtype=torch.bool, device=e.device)
        own[e.user_lo:e.user_lo + e.user_cnt] = True
        assert float(e.users["m"][~own].abs().max()) == 0.0 and float(e.users["g"][~own].abs().max()) == 0.0


def test_sharded_step_run_overflow_poisons_the_loss():
    """A batch with more runs of equal users than max_runs: the step's loss is NaN on every rank and
    read_loss raises (nothing is dropped silently); the next, well-formed batch is clean again."""
    from fvx import _lib
    from fvx.parallel import LocalGroup, ShardedStep
    U, I, K, d, D, B, R = 300, 401, 16, 8, 128, 128, 2
    P, F, rng = _random_problem(U, I, K, d, D, seed=4)
    es = _shards(U, I, K, d, D, R, P, F, max_batch=B, use_tensor_cores=True)
    step = ShardedStep(es, LocalGroup(R), max_runs=B // 4 + 2)
    ok = (np.repeat(rng.permutation(U)[:B // 4], 4), rng.integers(0, I, B), rng.integers(0, I, B))
    bad = (rng.permutation(U)[:B], rng.integers(0, I, B), rng.integers(0, I, B))      # B runs of one triple
    step.step(*(_dev(x) for x in ok))
    assert np.isfinite(step.read_loss())
    step.step(*(_dev(x) for x in bad))
    assert all(np.isnan(float(e.loss_t[0].item())) for e in es)
    with pytest.raises(_lib.FvxError):
        step.read_loss()
    step.step(*(_dev(x) for x in ok))
    assert np.isfinite(step.read_loss())


@pytest.mark.parametrize("R", [2, 4])
def test_sharded_topk_equals_single_rank(R):
    from fvx.parallel import LocalGroup, sharded_topk, user_slices
    U, I, K, d, D, k = 300, 9000, 32, 12, 64, 20
    P, F, rng = _random_problem(U, I, K, d, D, seed=9)
    one = _engine(U, I, K, d=d, D=D, max_batch=64)
    one.set_features(F)
    one.load_params(P)
    tr = [sorted(rng.choice(I, int(rng.integers(1, 9)), replace=False).tolist()) for _ in range(U)]
    rp = torch.as_tensor(np.concatenate([[0], np.cumsum([len(t) for t in tr])]), dtype=torch.int64).cuda()
    cs = torch.as_tensor(np.concatenate(tr), dtype=torch.int32).cuda()
    ids1, sc1 = one.score_topk(rp, cs, k)
    es = _shards(U, I, K, d, D, R, P, F, max_batch=64)
    for tc in (False, True):     # fp32 sweep per shard; tcgen05 sweeps with the bounds maximised over the shards
        merged = sharded_topk(es, LocalGroup(R), rp, cs, k, tc=tc)
        per, _ = user_slices(U, R)
        ids = torch.cat([m[0] for m in merged])[:U]
        sc = torch.cat([m[1] for m in merged])[:U]
        assert torch.equal(ids, ids1) and torch.equal(sc, sc1), tc
    o_ids, o_sc = oe.masked_topk(bpr.predict_all(P, F), tr, k)
    for u in range(U):
        ok, msg = oe.topk_matches(ids[u].cpu().numpy(), sc[u].cpu().numpy(), o_ids[u], o_sc[u])
        assert ok, (u, msg)
    # the other decomposition: users split over the ranks, item operands all-gathered
    from fvx.parallel import user_sliced_topk
    for tc in (False, True):
        sl = user_sliced_topk(es, LocalGroup(R), rp, cs, k, tc=tc)
        ids2 = torch.cat([a for a, _ in sl])
        sc2 = torch.cat([b for _, b in sl])
        assert ids2.shape[0] == U and torch.equal(ids2, ids1) and torch.equal(sc2, sc1), tc


@pytest.mark.parametrize("n", [1, 31, 256, 4096, 4097, 70001, 262144])
def test_run_ids_kernel_equals_torch(n):
    """fvx_run_ids (two launches) against the torch restatement the CPU test pins."""
    from fvx import _lib
    from fvx.parallel import run_ids
    g = torch.Generator().manual_seed(n)
    lens = torch.randint(1, 9, (n,), generator=g)
    user = torch.repeat_interleave(torch.randint(0, 50, (n,), generator=g), lens)[:n].to(torch.int32).cuda()
    out = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    scratch = torch.zeros(n // 4096 + 2, dtype=torch.int32, device="cuda")
    _lib.call("fvx_run_ids", _lib.ptr(user), n, _lib.ptr(out), _lib.ptr(scratch), _lib.stream_ptr())
    assert torch.equal(out, run_ids(user))


@pytest.mark.parametrize("n,R,cap", [(1, 1, 4), (300, 2, 40), (4097, 3, 400), (70001, 8, 1200), (262144, 8, 3000), (5000, 4, 20)])
def test_run_slots_kernel_equals_numpy(n, R, cap):
    """fvx_run_slots: run slot = owner * cap + index of the run among its owner's runs (0x7fffffff past cap)."""
    from fvx import _lib
    g = torch.Generator().manual_seed(n + R)
    U = 997
    per = (U + R - 1) // R
    lens = torch.randint(1, 9, (n,), generator=g)
    user = torch.repeat_interleave(torch.randint(0, U, (n,), generator=g), lens)[:n].to(torch.int32)
    u = user.numpy()
    start = np.ones(n, dtype=bool)
    start[1:] = u[1:] != u[:-1]
    owner = np.minimum(u // per, R - 1)
    want = np.zeros(n, dtype=np.int64)
    cnt = np.zeros(R, dtype=np.int64)
    cur = 0
    for b in range(n):
        if start[b]:
            o = owner[b]
            cur = o * cap + cnt[o] if cnt[o] < cap else 0x7fffffff
            cnt[o] += 1
        want[b] = cur
    out = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    scratch = torch.zeros(8 * (n // 1024 + 2), dtype=torch.int32, device="cuda")
    _lib.call("fvx_run_slots", _lib.ptr(user.cuda()), n, per, R, cap, _lib.ptr(out), _lib.ptr(scratch), _lib.stream_ptr())
    assert np.array_equal(out.cpu().numpy().astype(np.int64), want)
