"""CPU: the oracle against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py).  This is the pin of the checker, not a product test."""
import numpy as np
import pytest

from oracle import bpr, evaluator, sampler
from helpers import csr, golden, rel_err, tiny_dataset


def test_sampler_reference_stream_matches_reference():
    g = golden("sampler_ref.npz")
    U, I, tr, _, _, _ = tiny_dataset()
    u, p, n = sampler.reference_stream_triples(tr, I, int(g["batch_size"]), int(g["epochs"]), seed=0)
    assert len(u) == len(g["users"]) == (sum(map(len, tr)) // 16) * 16 * 3
    assert (u == g["users"]).all() and (p == g["pos"]).all() and (n == g["neg"]).all()


def test_evaluator_metrics_match_reference():
    g = golden("evaluator_ref.npz")
    U, I, tr, va, te, _ = tiny_dataset()
    res = evaluator.evaluate(g["scores"], I, tr, va, te, int(g["k"]))
    for key, want in zip(g["results_keys"].tolist(), g["results"].tolist()):
        assert res[key] == pytest.approx(want, rel=1e-12, abs=1e-15), key


def test_evaluator_topk_matches_reference_tsv():
    g = golden("evaluator_ref.npz")
    U, I, tr, _, _, _ = tiny_dataset()
    k = int(g["k"])
    ids, val = evaluator.masked_topk(g["scores"], tr, k)
    rows = [l.split("\t") for l in str(g["recs_tsv"]).strip().split("\n")]
    assert len(rows) == U * k
    ref_ids = np.array([int(r[1]) for r in rows]).reshape(U, k)
    ref_val = np.array([float(r[2]) for r in rows]).reshape(U, k)
    for u in range(U):
        ok, msg = evaluator.topk_matches(ids[u], val[u], ref_ids[u], ref_val[u])
        assert ok, (u, msg)
        assert not set(ids[u].tolist()) & set(tr[u])


@pytest.mark.parametrize("name,vis", [("bprmf_ref.npz", False), ("vbpr_ref.npz", True)])
def test_train_steps_match_reference_model_code(name, vis):
    g = golden(name)
    names = ["Bi", "Gu", "Gi"] + (["Tu", "E", "Bp"] if vis else [])
    P = {k: g["init_" + k].astype(np.float32).copy() for k in names}
    F = g["init_F"].astype(np.float32) if vis else None
    if vis:   # the reference normalises by the global max-abs (visual_loader_mixin.py:30)
        raw = tiny_dataset()[5]
        assert np.array_equal(bpr.normalise_features(raw), F)
    S = bpr.init_adam(P)
    lr, reg = float(g["hyper"][0]), float(g["hyper"][1])
    steps = len(g["losses"])
    for s in range(steps):
        loss = bpr.train_step(P, S, (g["users"][s], g["pos"][s], g["neg"][s]), reg, lr, F)
        assert loss == pytest.approx(float(g["losses"][s]), rel=2e-5), s
        if s + 1 in (1, 5):
            for k in names:
                assert rel_err(P[k], g["step%d_%s" % (s + 1, k)]) < 2e-5, (s, k)
    for k in names:
        assert rel_err(P[k], g["final_" + k]) < 1e-4, k
    assert rel_err(bpr.predict_all(P, F), g["predict_all_final"]) < 1e-4


@pytest.mark.parametrize("name,vis", [("bprmf_ref.npz", False), ("vbpr_ref.npz", True)])
def test_golden_batches_are_the_reference_sampler_stream(name, vis):
    g, s = golden(name), golden("sampler_ref.npz")
    assert (g["users"].reshape(-1) == s["users"]).all()
    assert (g["neg"].reshape(-1) == s["neg"]).all()


def test_closed_form_gradients_against_autograd():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)
    U, I, K, d, D, B = 9, 13, 5, 3, 7, 32
    P = bpr.init_params(U, I, K, d, D, seed=1, dtype=np.float64)
    P["Bi"] = rng.standard_normal(I)
    F = np.abs(rng.standard_normal((I, D)))
    batch = (rng.integers(0, U, B), rng.integers(0, I, B), rng.integers(0, I, B))
    reg = 0.01
    loss, G, _ = bpr.loss_and_grads(P, batch, reg, F)
    T = {k: torch.tensor(v, requires_grad=True) for k, v in P.items()}
    Ft = torch.tensor(F)
    u, i, j = (torch.tensor(b) for b in batch)

    def s(it):
        return (T["Bi"][it] + (T["Gu"][u] * T["Gi"][it]).sum(1)
                + (T["Tu"][u] * (Ft[it] @ T["E"])).sum(1) + (Ft[it] @ T["Bp"])[:, 0])
    x = torch.clamp(s(i) - s(j), -80.0, 1e8)
    L = torch.nn.functional.softplus(-x).sum()
    L = L + reg * ((T["Gu"][u] ** 2).sum() + (T["Gi"][i] ** 2).sum() + (T["Gi"][j] ** 2).sum()
                   + (T["Tu"][u] ** 2).sum()) + reg * (T["Bi"][i] ** 2).sum() \
        + reg * (T["Bi"][j] ** 2).sum() / 10 + reg * ((T["E"] ** 2).sum() + (T["Bp"] ** 2).sum())
    L.backward()
    assert float(L.detach()) == pytest.approx(float(loss), rel=1e-12)
    for k in P:
        assert np.max(np.abs(T[k].grad.numpy() - G[k])) < 1e-12, k


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    out = sampler.philox4x32([0], [0], [0], [0], 0, 0)
    assert [int(o[0]) for o in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    out = sampler.philox4x32([0xFFFFFFFF], [0xFFFFFFFF], [0xFFFFFFFF], [0xFFFFFFFF], 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(o[0]) for o in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = sampler.philox4x32([0x243F6A88], [0x85A308D3], [0x13198A2E], [0x03707344], 0xA4093822, 0x299F31D0)
    assert [int(o[0]) for o in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_negatives_never_in_train_and_deterministic():
    U, I, tr, _, _, _ = tiny_dataset()
    row_ptr, col_file, col_sorted = csr(tr)
    perm = sampler.feistel_user_permutation(U, seed=5, epoch=0)
    assert sorted(perm.tolist()) == list(range(U))
    for n in (1, 2, 3, 17, 1000, 4097):                      # a bijection for every domain size, every epoch
        p0, p1 = sampler.feistel_user_permutation(n, 9, 0), sampler.feistel_user_permutation(n, 9, 1)
        assert sorted(p0.tolist()) == list(range(n)) and sorted(p1.tolist()) == list(range(n))
        assert n < 17 or (p0 != p1).any()
    users, pos = sampler.enumerate_epoch(row_ptr, col_file, perm)
    neg = sampler.philox_negatives(row_ptr, col_sorted, users, I, seed=5, offset=100)
    assert all(int(j) not in tr[int(u)] for u, j in zip(users, neg))
    assert (0 <= neg).all() and (neg < I).all()
    again = sampler.philox_negatives(row_ptr, col_sorted, users[7:], I, seed=5, offset=107)
    assert (again == neg[7:]).all()


def test_gradfashion_oracle_matches_reference_model_code():
    """oracle/gradfashion.py against the reference's own GradFashion.py (call, train_step, predict_all)
    run over the tensorflow shim (tests/golden/make_golden_gradfashion.py): the two-stage visual
    projection v_i = [Fc Ec | Fe Ee], its regulariser (no /10 on the negative bias, unlike VBPR) and the
    closed-form gradients of all eight variables."""
    from oracle import gradfashion as gf
    g = golden("gradfashion_ref.npz")
    names = ["Bi", "Gu", "Gi", "Ec", "Ee", "Tu", "E", "Bp"]
    lr, reg = float(g["hyper"][0]), float(g["hyper"][1])
    Fc, Fe = g["Fc"].astype(np.float32), g["Fe"].astype(np.float32)
    P = {k: g["init_" + k].astype(np.float32).copy() for k in names}
    assert rel_err(gf.score(P, g["users"][0], g["pos"][0], Fc, Fe), g["call_x0"]) < 1e-5
    S = bpr.init_adam(P)
    for s in range(len(g["losses"])):
        loss = gf.train_step(P, S, (g["users"][s], g["pos"][s], g["neg"][s]), reg, lr, Fc, Fe)
        assert loss == pytest.approx(float(g["losses"][s]), rel=1e-5), s
        if s == 0:
            for k in names:
                assert rel_err(P[k].reshape(g["step1_" + k].shape), g["step1_" + k]) < 1e-4, k
    for k in names:
        assert rel_err(P[k].reshape(g["final_" + k].shape), g["final_" + k]) < 1e-4, k
    assert rel_err(gf.predict_all(P, Fc, Fe), g["predict_all_final"]) < 1e-4
    # gradients against central differences in float64 on one batch (independent of the shim's autodiff)
    P64 = {k: g["init_" + k].astype(np.float64).copy() for k in names}
    b = (g["users"][0], g["pos"][0], g["neg"][0])
    _, G, _ = gf.loss_and_grads(P64, b, reg, Fc.astype(np.float64), Fe.astype(np.float64))
    rng = np.random.default_rng(0)
    for k in ("Ec", "Ee", "E", "Bp", "Tu", "Gi"):
        idx = tuple(rng.integers(0, n) for n in P64[k].shape)
        if k in ("Tu",):
            idx = (int(b[0][0]),) + idx[1:]
        if k in ("Gi",):
            idx = (int(b[1][0]),) + idx[1:]
        h = 1e-6
        old = P64[k][idx]
        P64[k][idx] = old + h
        lp = gf.loss_and_grads(P64, b, reg, Fc.astype(np.float64), Fe.astype(np.float64))[0]
        P64[k][idx] = old - h
        lm = gf.loss_and_grads(P64, b, reg, Fc.astype(np.float64), Fe.astype(np.float64))[0]
        P64[k][idx] = old
        assert G[k][idx] == pytest.approx((lp - lm) / (2 * h), rel=1e-5, abs=1e-8), (k, idx)


def _evaluator_random():
    g = golden("evaluator_random.npz")
    U = len(g["train_ptr"]) - 1

    def lists(ptr, col):
        return [col[ptr[u]:ptr[u + 1]].tolist() for u in range(U)]
    return g, U, lists(g["train_ptr"], g["train_col"]), lists(g["val_ptr"], g["val_col"]), lists(g["test_ptr"], g["test_col"])


def test_oracle_eval_by_user_matches_reference_on_random_model():
    """oracle.evaluator.eval_by_user against the reference's own _eval_by_user (fixture generated by
    tests/golden/make_golden_evaluator_random.py): 2 000 users, exact ties from duplicated items, 0 / 1 / 3 held-out
    items - every user, every metric."""
    g, U, tr, va, te = _evaluator_random()
    Gu, Gi, Bi, k = g["Gu"], g["Gi"], g["Bi"], int(g["k"])
    I = Gi.shape[0]
    scores = (Bi[None, :].astype(np.float64) + Gu.astype(np.float64) @ Gi.astype(np.float64).T).astype(np.float32)
    _, first = np.unique(np.concatenate([Gi, Bi[:, None]], 1), axis=0, return_index=True)   # duplicates score identically
    rep = {tuple(np.concatenate([Gi[i], Bi[i:i + 1]])): i for i in sorted(first)}
    for i in range(I):
        scores[:, i] = scores[:, rep[tuple(np.concatenate([Gi[i], Bi[i:i + 1]]))]]
    for name, held in (("metrics_v", va), ("metrics_t", te)):
        want = g[name]
        for u in range(U):
            got = evaluator.eval_by_user(scores[u], I, tr[u], held[u], k)
            if not held[u]:
                assert got == () and np.isnan(want[u, 0])
            else:
                assert np.allclose(got, want[u], rtol=0, atol=1e-12), (name, u, got, want[u])
