"""GPU: GradFashion (two-stage visual projection, src/recommender/models/GradFashion.py) through the C ABI -
``Engine(two_stage=...)`` - against the golden vectors the reference's own GradFashion.py produced over the
tensorflow shim (tests/golden/make_golden_gradfashion.py) and against oracle/gradfashion.py at the tensor-core sizes."""
import argparse

import numpy as np
import pytest
import torch

from helpers import golden, rel_err
from oracle import bpr, evaluator as oe, gradfashion as gf
from test_gpu_parity import REL, _dev, _engine, _user_contiguous_batches

pytestmark = pytest.mark.gpu
NAMES = ("Gu", "Gi", "Bi", "Tu", "Ec", "Ee", "E", "Bp")


def _feat(Fc, Fe):
    return np.concatenate([Fc, Fe], axis=1).astype(np.float32)


@pytest.mark.parametrize("mode", ["dense", "deferred"])
def test_gradfashion_steps_match_reference_golden(mode):
    g = golden("gradfashion_ref.npz")
    P0 = {k: g["init_" + k] for k in NAMES}
    Fc, Fe = g["Fc"], g["Fe"]
    U, K = P0["Gu"].shape
    I, d = P0["Gi"].shape[0], P0["Tu"].shape[1]
    lr, reg = (float(x) for x in g["hyper"])
    e = _engine(U, I, K, d=d, D=Fc.shape[1] + Fe.shape[1], lr=lr, reg=reg, adam_mode=mode, max_batch=g["users"].shape[1],
                two_stage=(Fc.shape[1], Fe.shape[1], P0["Ec"].shape[1], P0["Ee"].shape[1]))
    e.set_features(_feat(Fc, Fe))
    e.load_params(P0)
    x0 = e.score_pairs(_dev(g["users"][0]), _dev(g["pos"][0])).cpu().numpy()
    assert rel_err(x0, g["call_x0"]) <= REL                                   # GradFashion.call (:122-125)
    for s in range(g["users"].shape[0]):
        e.step(_dev(g["users"][s]), _dev(g["pos"][s]), _dev(g["neg"][s]))
        assert e.read_loss(0) == pytest.approx(float(g["losses"][s]), rel=REL), s
        if s == 0:
            Q = e.params()
            for k in NAMES:
                assert rel_err(Q[k].reshape(g["step1_" + k].shape), g["step1_" + k]) <= 2e-4, k
    Q = e.params()
    for k in NAMES:
        ref = g["final_" + k]
        dlt = np.abs(Q[k].reshape(ref.shape) - ref) / np.abs(ref).max()
        assert dlt.max() <= 5e-3 and (dlt > REL).mean() <= 2e-2, (k, float(dlt.max()), float((dlt > REL).mean()))
    assert rel_err(e.predict_all().cpu().numpy(), g["predict_all_final"]) <= 5e-3   # predict_all (:304-320)


@pytest.mark.parametrize("tc", [False, True])
def test_gradfashion_tensor_core_sizes_match_oracle(tc):
    """Dc + De = 1024 features (tcgen05 projection on the composed matrix), batches of the sampler's shape: losses per
    step, parameters, predict_all and the masked top-k against the fp64 oracle."""
    U, I, K, d, Dc, De, ec, ee, B, steps, lr, reg = 400, 700, 32, 12, 640, 384, 10, 6, 512, 8, 1e-3, 1e-3
    rng = np.random.default_rng(17)
    P = bpr.init_params(U, I, K, d, ec + ee, seed=2)                      # Gu, Gi, Bi, Tu, E [ec+ee, d], Bp [ec+ee, 1]
    P["Bi"] = (0.05 * rng.standard_normal(I)).astype(np.float32)
    lim = np.sqrt(6.0 / (Dc + ec))
    P["Ec"] = rng.uniform(-lim, lim, (Dc, ec)).astype(np.float32)
    lim = np.sqrt(6.0 / (De + ee))
    P["Ee"] = rng.uniform(-lim, lim, (De, ee)).astype(np.float32)
    Fc = bpr.normalise_features(np.maximum(rng.standard_normal((I, Dc)), 0))
    Fe = bpr.normalise_features(np.maximum(rng.standard_normal((I, De)), 0))
    e = _engine(U, I, K, d=d, D=Dc + De, lr=lr, reg=reg, max_batch=B, use_tensor_cores=tc, two_stage=(Dc, De, ec, ee))
    assert e.use_tensor_cores == tc
    e.set_features(_feat(Fc, Fe))
    e.load_params(P)
    Q = {k: v.astype(np.float64) for k, v in P.items()}
    S = bpr.init_adam(Q)
    Fc64, Fe64 = Fc.astype(np.float64), Fe.astype(np.float64)
    for s, b in enumerate(_user_contiguous_batches(rng, U, I, B, steps)):
        want = gf.train_step(Q, S, b, reg, lr, Fc64, Fe64)
        e.step(*(_dev(x) for x in b))
        assert e.read_loss(0) == pytest.approx(want, rel=REL), s
    R = e.params()
    for k in NAMES:
        ref = Q[k]
        dlt = np.abs(R[k].reshape(ref.shape) - ref) / np.abs(ref).max()
        assert dlt.max() <= 5e-3 and (dlt > REL).mean() <= 5e-3, (k, float(dlt.max()), float((dlt > REL).mean()))
    want = gf.predict_all(Q, Fc64, Fe64)
    assert rel_err(e.predict_all().cpu().numpy(), want) <= 1e-3
    tr = [sorted(rng.choice(I, 5, replace=False).tolist()) for _ in range(U)]
    rp = torch.as_tensor(np.arange(U + 1) * 5, dtype=torch.int64).cuda()
    cs = torch.as_tensor(np.array(tr).reshape(-1), dtype=torch.int32).cuda()
    ids, sc = e.score_topk(rp, cs, 10)
    Rp = {k: v.astype(np.float64) for k, v in R.items()}
    o_ids, o_sc = oe.masked_topk(gf.predict_all(Rp, Fc64, Fe64), tr, 10)
    for u in range(U):
        ok, msg = oe.topk_matches(ids[u].cpu().numpy(), sc[u].cpu().numpy(), o_ids[u], o_sc[u])
        assert ok, (u, msg)


def test_gradfashion_model_class_trains_and_evaluates(tmp_path):
    """GradFashion(data, params): train_step losses against the oracle from the model's own initial values, call()
    tuple order, evaluator."""
    from fvx import synth
    from fvx.dataset.dataset import DataLoader
    from fvx.recommender.models.GradFashion import GradFashion
    U, I, B = 300, 900, 256
    inter = synth.make_interactions(U, I, seed=9)
    p = argparse.Namespace(dataset="gf", batch_size=B, epochs=2, sampler="device", seed=0, rec="gradfashion", embed_k=16,
                           embed_d=8, embed_color=6, embed_edges=5, lr=1e-3, reg=1e-4, top_k=10, verbose=-1,
                           restore_epochs=1, batch_eval=128, best_metric="ndcg", validation=True, cnn_model="resnet50",
                           output_layer="avg_pool")
    data = DataLoader(p, interactions=inter)
    rng = np.random.default_rng(4)
    data.color_features_raw = rng.random((I, 192)).astype(np.float32)
    data.edge_features_raw = np.maximum(rng.standard_normal((I, 64)), 0).astype(np.float32)
    m = GradFashion(data, p)
    assert m.engine.use_tensor_cores and m.engine.two_stage == (192, 64, 6, 5)
    Q = {k: v.astype(np.float64) for k, v in m.engine.params().items()}
    S = bpr.init_adam(Q)
    Fc, Fe = m.color_features.astype(np.float64), m.edge_features.astype(np.float64)
    it = data.next_triple_batch("cuda:0")
    for s in range(4):
        b = next(it)
        want = gf.train_step(Q, S, tuple(x.cpu().numpy().astype(np.int64) for x in b), 1e-4, 1e-3, Fc, Fe)
        assert m.train_step(b) == pytest.approx(want, rel=REL), s
    out = m.call((np.arange(5), np.arange(5)))
    assert len(out) == 8 and out[3].shape == (5, 192) and out[4].shape == (5, 64) and out[6].shape == (5, 8)
    want = gf.score(Q, np.arange(5), np.arange(5), Fc, Fe)
    assert rel_err(out[0].numpy(), want) <= 1e-3
    res = {}
    m.evaluator.eval(1, res, "gf", 0.0)
    assert 0.0 <= res[1]["auc_t_fixed"] <= 1.0
