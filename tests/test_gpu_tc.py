"""GPU: the tcgen05 / TMA-gather projection kernels (fvx_project_tc.cu) against fp64 NumPy and
against the fp32 CUDA-core kernels, through the C ABI (fvx_project_rows, fvx_grad_e_rows,
fvx_project, fvx_bpr_step with use_tensor_cores=1).  Tolerance: 1e-4 relative (north_star);
the hi/lo bf16 split keeps the tensor-core result within ~1e-5 of the fp32 kernels."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import bpr
from test_gpu_parity import (REL, _assert_params_close, _dev, _engine, _oracle_pair, _random_problem,
                             _user_contiguous_batches)

pytestmark = pytest.mark.gpu


def _pair(U, I, K, d, D, B, seed=0, **kw):
    P, F, rng = _random_problem(U, I, K, d, D, seed=seed)
    es = []
    for tc in (False, True):
        e = _engine(U, I, K, d=d, D=D, max_batch=B, use_tensor_cores=tc, **kw)
        e.set_features(F)
        e.load_params(P)
        es.append(e)
    return P, F, rng, es[0], es[1]


@pytest.mark.parametrize("I,d,D,n", [(900, 20, 256, 1000), (300, 20, 2048, 517), (2000, 5, 128, 4096),
                                       (500, 40, 512, 700), (400, 64, 256, 300),
                                       (400, 256, 256, 300),      # 257 columns: two column slices (192 + 128)
                                       (300, 270, 512, 1100)])
def test_project_rows_tensor_cores(I, d, D, n):
    P, F, rng, e32, etc = _pair(50, I, 8, d, D, B=max(n // 2 + 1, 64), seed=d + D)
    rows = rng.integers(0, I, n)
    rows[::7] = rows[0]                                   # duplicates
    want = F[rows].astype(np.float64) @ np.concatenate([P["E"], P["Bp"].reshape(D, 1)], 1).astype(np.float64)
    r = _dev(rows)
    got32 = e32.project_rows(r).cpu().numpy()[:, :d + 1]
    gottc = etc.project_rows(r).cpu().numpy()[:, :d + 1]
    assert rel_err(got32, want) <= 1e-5
    assert rel_err(gottc, want) <= 2e-5, rel_err(gottc, want)
    # per-element: absolute floor from the hi/lo split (2^-16 of the magnitudes entering the sum)
    mag = np.abs(F[rows].astype(np.float64)) @ np.abs(np.concatenate([P["E"], P["Bp"].reshape(D, 1)], 1))
    assert np.all(np.abs(gottc - want) <= 3e-5 * mag + 1e-12)


@pytest.mark.parametrize("I,d,D,n", [(900, 20, 256, 1000), (300, 20, 2048, 517), (2000, 5, 128, 4096),
                                       (500, 40, 512, 96), (400, 64, 256, 300),
                                       (400, 256, 512, 300), (300, 270, 256, 1100)])    # column slices
def test_grad_e_rows_tensor_cores(I, d, D, n):
    P, F, rng, e32, etc = _pair(50, I, 8, d, D, B=max(n // 2 + 1, 64), seed=d + D + 1)
    rows = rng.integers(0, I, n)
    rows[::5] = rows[1]
    W = (rng.standard_normal((n, e32.de)) * rng.exponential(1.0, (n, 1))).astype(np.float32)
    W[:, d + 1:] = 0
    skip = rng.random(n) < 0.1                            # slots owned by another rank
    rows_m = np.where(skip, -1, rows)
    want = (F[rows].astype(np.float64) * (~skip)[:, None]).T @ W.astype(np.float64)
    r, w = _dev(rows_m), torch.as_tensor(W).cuda()
    got32 = e32.grad_E_rows(r, w).cpu().numpy()
    gottc = etc.grad_E_rows(r, w).cpu().numpy()
    assert rel_err(got32[:, :d + 1], want[:, :d + 1]) <= 1e-5
    assert rel_err(gottc[:, :d + 1], want[:, :d + 1]) <= 2e-5, rel_err(gottc, want)


def test_catalog_projection_tensor_cores():
    """fvx_project over the whole catalog (theta for evaluation): identity-row TMA tiles."""
    P, F, rng, e32, etc = _pair(40, 3001, 8, 20, 256, B=64, seed=3)
    want = F.astype(np.float64) @ np.concatenate([P["E"], P["Bp"].reshape(-1, 1)], 1).astype(np.float64)
    assert rel_err(e32.theta().cpu().numpy()[:, :21], want) <= 1e-5
    assert rel_err(etc.theta().cpu().numpy()[:, :21], want) <= 2e-5


def _perturbed_oracle(P, F, batches, reg, lr, rel=2.0 ** -16, seed=7):
    """fp64 oracle fed features perturbed at the representation level of the bf16 hi/lo planes.
    Adam's update m/(sqrt(v)+eps) is discontinuous where a gradient crosses zero, so a handful
    of elements of an EXACT computation move by O(lr) under such a perturbation (DESIGN.md
    "Parity"); the tensor-core path is accepted when it is no farther from the exact result."""
    prng = np.random.default_rng(seed)
    Q = {k: v.astype(np.float64) for k, v in P.items()}
    S = bpr.init_adam(Q)
    Fp = F.astype(np.float64) * (1.0 + rel * prng.standard_normal(F.shape))
    for b in batches:
        bpr.train_step(Q, S, b, reg, lr, Fp)
    return Q


@pytest.mark.parametrize("mode", ["dense", "deferred"])
@pytest.mark.parametrize("K,d,D,B", [(64, 20, 256, 512), (16, 64, 128, 96), (8, 5, 128, 33), (32, 20, 2048, 1024),
                                       (256, 20, 4096, 512),      # BASELINE configs[4]: K = 256, 4096-d features
                                       (256, 255, 4096, 96),      # ... the widest single-launch operand (NP = 256)
                                       (256, 256, 4096, 96),      # ... and configs[4] with embed_d = 256: two column slices
                                       (16, 256, 256, 300)])
def test_train_steps_tensor_cores_match_oracle(K, d, D, B, mode):
    U, I, steps, lr, reg = 700, 900, 20, 0.001, 1e-3
    P, F, rng = _random_problem(U, I, K, d, D, seed=K + d)
    e = _engine(U, I, K, d=d, D=D, lr=lr, reg=reg, adam_mode=mode, max_batch=B, use_tensor_cores=True)
    e.set_features(F, keep_fp32=False)                    # the step must not need the fp32 copy
    assert e.struct().use_tensor_cores == 1
    e.load_params(P)
    batches = _user_contiguous_batches(rng, U, I, B, steps)
    P32, P64, l32, l64 = _oracle_pair(P, F, batches, reg, lr)
    for s, b in enumerate(batches):
        e.step(*(_dev(x) for x in b), loss_slot=s % 7)
        got = e.read_loss(s % 7)
        assert got == pytest.approx(l64[s], rel=REL), (s, mode)
    Pp = _perturbed_oracle(P, F, batches, reg, lr)
    Q = e.params()
    for k in P64:
        ref = P64[k]
        dlt = np.abs(Q[k].reshape(ref.shape) - ref) / np.abs(ref).max()
        e_f32, e_pert = rel_err(P32[k], ref), rel_err(Pp[k], ref)
        assert (dlt > REL).mean() <= 1e-3, (mode, k, "elements beyond 1e-4", int((dlt > REL).sum()), dlt.size)
        # the perturbed oracle is ONE draw of the discontinuity noise (the largest of ~1e5 elements, each of them an
        # Adam update whose sign a rounding-level gradient decides): a factor 3 on it; systematic errors are caught
        # by the fraction above and by the per-step losses
        assert dlt.max() <= max(REL, 3 * e_f32, 3 * e_pert), (mode, k, dlt.max(), e_f32, e_pert,
                                                               np.unravel_index(dlt.argmax(), dlt.shape))


@pytest.mark.parametrize("K,d,D,B,tc", [(64, 20, 256, 256, True), (64, 20, 2048, 300, True), (32, 0, 0, 256, False),
                                          (16, 8, 128, 64, False)])
def test_graph_replayed_steps_equal_single_steps(K, d, D, B, tc):
    """fvx_bpr_steps (small-batch regime: 8 steps captured once into a CUDA graph and replayed over batches
    that lie back to back in epoch-long index arrays, plus ungraphed remainder steps) against the same
    batches fed one fvx_bpr_step at a time: same losses, same parameters; and against the fp64 oracle."""
    U, I, lr, reg, steps = 500, 800, 1e-3, 1e-4, 21            # 2 graph launches + 5 plain steps
    P, F, rng = _random_problem(U, I, K, d, D, seed=K + B)
    es = []
    for _ in range(2):
        e = _engine(U, I, K, d=d, D=D, lr=lr, reg=reg, max_batch=B, use_tensor_cores=tc)
        if D:
            e.set_features(F)
        e.load_params(P)
        es.append(e)
    batches = _user_contiguous_batches(rng, U, I, B, steps + 3)
    eu, ep, en = (_dev(np.concatenate([b[c] for b in batches])) for c in range(3))
    first = 2                                                # the replay starts at batch 2 of the arrays
    _, P64, _, l64 = _oracle_pair(P, F, batches[first:first + steps], reg, lr)
    for s in range(steps):
        es[0].step(*(_dev(x) for x in batches[first + s]), loss_slot=0)
    es[1].steps(eu, ep, en, first, steps, B, loss_slot=0)
    la, lb = es[0].read_loss(0), es[1].read_loss(0)
    assert la == pytest.approx(float(np.sum(l64)), rel=REL)
    assert lb == pytest.approx(la, rel=1e-6)
    assert es[0].steps_done() == es[1].steps_done() == steps
    # a second call replays the cached graph from another position
    es[0].step(*(_dev(x) for x in batches[0]), loss_slot=1)
    es[1].steps(eu, ep, en, 0, 1, B, loss_slot=1)
    assert es[1].read_loss(1) == pytest.approx(es[0].read_loss(1), rel=1e-6)
    for _ in range(2):
        for s in range(8):
            es[0].step(*(_dev(x) for x in batches[s]), loss_slot=0)
        es[1].steps(eu, ep, en, 0, 8, B, loss_slot=0)
    assert es[1].read_loss(0) == pytest.approx(es[0].read_loss(0), rel=1e-6)
    Qa, Qb = es[0].params(), es[1].params()
    for k in Qa:
        dlt = np.abs(Qb[k] - Qa[k]) / np.abs(Qa[k]).max()
        assert dlt.max() <= 5e-3 and (dlt > 2e-5).mean() <= 1e-3, (k, dlt.max())


@pytest.mark.parametrize("mode", ["deferred", "dense"])
@pytest.mark.parametrize("K,d,D,B,I", [(64, 20, 2048, 3000, 400),     # ~every catalog row repeats 15 times
                                        (64, 20, 256, 4097, 50000),    # hardly any duplicates, ragged batch
                                        (32, 63, 512, 777, 300),       # NP = 64
                                        (16, 100, 256, 500, 200),      # NP = 128 (three-pass operands)
                                        (16, 256, 256, 500, 200)])     # NP = 320: two column slices
def test_unique_row_step_equals_per_slot_step(K, d, D, B, I, mode):
    """fvx_bpr_step projects each DISTINCT catalog row of the batch once (k_uniq_rows / upos / W_sum,
    DESIGN.md section 3).  Against the per-slot path on the same batches - including triples whose
    item id lies outside the catalog - and against the fp64 oracle; the scratch the unique-row step
    leaves behind (row list, coefficient sums) must be clean after every step."""
    U, steps, lr, reg = 600, 8, 1e-3, 1e-4
    P, F, rng = _random_problem(U, I, K, d, D, seed=K + d + 5)
    es = []
    for uniq in (False, True):
        e = _engine(U, I, K, d=d, D=D, lr=lr, reg=reg, adam_mode=mode, max_batch=B, use_tensor_cores=True,
                    unique_rows=uniq)
        e.set_features(F, keep_fp32=False)
        e.load_params(P)
        es.append(e)
    assert es[0].struct().upos is None and es[1].struct().upos is not None
    batches = _user_contiguous_batches(rng, U, I, B, steps)
    clean = [(u, i, j) for (u, i, j) in batches]
    P32, P64, l32, l64 = _oracle_pair(P, F, clean[:4], reg, lr)
    for s, (u, i, j) in enumerate(batches):
        i = i.copy(); j = j.copy()
        if s >= 4:
            i[5::97] = I + 3                               # outside the catalog: triple ignored
            j[11::89] = -1
        losses = []
        for e in es:
            e.step(_dev(u), _dev(i), _dev(j), loss_slot=0)
            losses.append(e.read_loss(0))
        assert losses[1] == pytest.approx(losses[0], rel=2e-6), s
        if s < 4:
            assert losses[1] == pytest.approx(l64[s], rel=REL), s
        assert int(es[1].items["count"].item()) == 0
        assert float(es[1].W_sum.abs().max()) == 0.0
    Qa, Qb = es[0].params(), es[1].params()
    for k in Qa:
        dlt = np.abs(Qb[k] - Qa[k]) / np.abs(Qa[k]).max()
        assert dlt.max() <= 5e-3 and (dlt > 2e-5).mean() <= 1e-3, (k, dlt.max(), (dlt > 2e-5).mean())


def test_unique_row_step_timed_entry_point_and_hook():
    """The profiling entry point runs the same unique-row kernels on one stream; the debug hook
    (FVX_STEP_DEDUP=0 equivalent) selects the per-slot path on an engine that carries upos."""
    from fvx import _lib
    lib = _lib.load()
    U, I, K, d, D, B, steps = 300, 500, 64, 20, 2048, 2048, 5
    P, F, rng = _random_problem(U, I, K, d, D, seed=3)
    es = []
    for _ in range(3):
        e = _engine(U, I, K, d=d, D=D, lr=1e-3, reg=1e-4, max_batch=B, use_tensor_cores=True)
        e.set_features(F, keep_fp32=False)
        e.load_params(P)
        es.append(e)
    batches = _user_contiguous_batches(rng, U, I, B, steps)
    for s, b in enumerate(batches):
        db = [_dev(x) for x in b]
        es[0].step(*db, loss_slot=0)
        ph = es[1].step_timed(*db, loss_slot=0)
        assert set(ph) == set(_lib.PHASES) and all(v >= 0 for v in ph.values())
        old = lib.fvx_debug_set_dedup(0)
        try:
            es[2].step(*db, loss_slot=0)
        finally:
            lib.fvx_debug_set_dedup(1 if old != 0 else 0)
        la, lb, lc = (e.read_loss(0) for e in es)
        assert la == pytest.approx(lb, rel=2e-6) and la == pytest.approx(lc, rel=2e-6), s
    Qa, Qb, Qc = (e.params() for e in es)
    for k in Qa:
        for Q in (Qb, Qc):
            dlt = np.abs(Q[k] - Qa[k]) / np.abs(Qa[k]).max()
            assert dlt.max() <= 5e-3 and (dlt > 2e-5).mean() <= 1e-3, (k, dlt.max())
