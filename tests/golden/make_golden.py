#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ by running the REFERENCE's own
code (read from /root/reference, never copied) in the build container.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz, tiny/

What runs unmodified from /root/reference/src:
  * dataset/dataset.py       DataLoader.__init__/load_list/all_triple_batches   (:13-114)
  * recommender/Evaluator.py Evaluator.__init__/eval/store_recommendation       (:17-239)
  * recommender/models/BPRMF.py, VBPR.py (+ RecommenderModel.py, visual_loader_mixin.py,
    utils/write.py, config/configs.py): __init__, call, predict_all, train_step, train
over the torch-backed ``tensorflow`` shim in tf_shim.py (TensorFlow 2.3.1 itself is not
installable here; see the shim's docstring for what that does and does not pin).

/root/reference does not exist on the GPU box; tests only read the committed fixtures.
"""
import argparse
import os
import pickle
import random
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import tf_shim  # noqa: E402
from fvx import synth  # noqa: E402


def params_ns(rec, **kw):
    d = dict(gpu=-1, best_metric="ndcg", dataset="tiny", rec=rec, batch_size=16, top_k=5, epochs=3,
             verbose=-1, batch_eval=128, lr=0.01, validation=True, restore_epochs=1,
             list_of_regs=[1e-3], cnn_model="resnet50", output_layer="avg_pool", embed_k=8, embed_d=4,
             reg=1e-3)
    d.update(kw)
    return argparse.Namespace(**d)


def snapshot(model, names):
    return {n: getattr(model, n).numpy().copy() for n in names}


def run_model(cls, rec, names, data_cls):
    """Runs the reference's model.train() end to end and records everything."""
    random.seed(0)
    np.random.seed(0)
    tf_shim._set_seed(0)
    args = params_ns(rec)
    data = data_cls(params=args)
    model = cls(data, args)
    rec_out = {"init_" + k: v for k, v in snapshot(model, names).items()}
    batches, losses, snaps = [], [], {}
    orig = model.train_step

    def recording_step(batch):
        batches.append([b.numpy().copy() for b in batch])
        loss = orig(batch)
        losses.append(float(loss))
        if len(losses) in (1, 5):
            snaps[len(losses)] = snapshot(model, names)
        return loss

    model.train_step = recording_step
    os.makedirs("../results/rec_results/tiny/%s" % rec, exist_ok=True)
    os.makedirs("../results/rec_model_weights/tiny/%s" % rec, exist_ok=True)
    model.train()
    rec_out.update({"final_" + k: v for k, v in snapshot(model, names).items()})
    for s, sn in snaps.items():
        rec_out.update({"step%d_%s" % (s, k): v for k, v in sn.items()})
    b = np.array(batches)                                   # [steps, 3, B]
    rec_out["users"], rec_out["pos"], rec_out["neg"] = b[:, 0], b[:, 1], b[:, 2]
    rec_out["losses"] = np.array(losses)
    rec_out["predict_all_final"] = model.predict_all().numpy()
    rdir = "../results/rec_results/tiny/%s" % rec
    files = sorted(os.listdir(rdir))
    with open(os.path.join(rdir, [f for f in files if f.startswith("results-metrics")][0]), "rb") as f:
        results = pickle.load(f)
    epochs = sorted(results)
    keys = sorted(results[epochs[0]])
    rec_out["results_keys"] = np.array(keys)
    rec_out["results"] = np.array([[float(results[e][k]) for k in keys] for e in epochs])
    last = [f for f in files if f.startswith("recs-")][0]
    rec_out["recs_name"] = np.array(last)
    rec_out["recs_tsv"] = np.array(open(os.path.join(rdir, last)).read())
    rec_out["files"] = np.array(files)
    rec_out["hyper"] = np.array([args.lr, args.reg, args.batch_size, args.top_k, args.epochs,
                                 args.embed_k, args.embed_d])
    return rec_out


def main():
    tmp = tempfile.mkdtemp(prefix="fvx_golden_")
    os.makedirs(os.path.join(tmp, "src"))
    inter = synth.make_interactions(48, 120, seed=7)
    feats = synth.make_features(120, 32, seed=11, dtype=np.float64)
    synth.write_dataset(tmp, "tiny", inter, feats)
    # the fixture dataset itself is committed so that tests read the same bytes
    dst = os.path.join(HERE, "tiny")
    shutil.rmtree(dst, ignore_errors=True)
    shutil.copytree(os.path.join(tmp, "data", "tiny"), dst)

    os.chdir(os.path.join(tmp, "src"))              # the reference's paths are relative to src/
    sys.path.insert(0, REF_SRC)
    tf_shim.install()

    # ---- A. sampler: reference all_triple_batches, seeds as set at model-module import
    from dataset.dataset import DataLoader
    args = params_ns("bprmf")
    random.seed(0)
    np.random.seed(0)
    data = DataLoader(params=args)
    u, p, n = data.all_triple_batches()
    sampler = dict(users=np.array([int(x) for x in u]), pos=np.array([int(x) for x in p]),
                   neg=np.array([int(x) for x in n]), batch_size=args.batch_size, epochs=args.epochs,
                   num_users=data.num_users, num_items=data.num_items,
                   train_len=np.array([len(t) for t in data.training_list]))
    np.savez_compressed(os.path.join(HERE, "sampler_ref.npz"), **sampler)

    # ---- B. evaluator on a fixed score matrix (incl. exact ties)
    import recommender.Evaluator as RefEval
    rng = np.random.default_rng(3)
    scores = rng.standard_normal((data.num_users, data.num_items)).astype(np.float32)
    scores[:, 10:20] = np.round(scores[:, 10:20], 1)         # ties inside rows

    class StubModel:
        def __init__(self):
            self.data = data

        def predict_all(self):
            class R:
                @staticmethod
                def numpy():
                    return scores.copy()
            return R()

    ev = RefEval.Evaluator(StubModel(), data, 5)
    results = {}
    ev.eval(1, results, "golden", 0.0)
    ev.store_recommendation(path="recs.tsv")
    keys = sorted(results[1])
    np.savez_compressed(os.path.join(HERE, "evaluator_ref.npz"), scores=scores, k=5,
                        results_keys=np.array(keys),
                        results=np.array([float(results[1][k]) for k in keys]),
                        recs_tsv=np.array(open("recs.tsv").read()))

    # ---- C. models over the shim
    from recommender.models.BPRMF import BPRMF
    from recommender.models.VBPR import VBPR
    np.savez_compressed(os.path.join(HERE, "bprmf_ref.npz"),
                        **run_model(BPRMF, "bprmf", ["Bi", "Gu", "Gi"], DataLoader))
    out = run_model(VBPR, "vbpr", ["Bi", "Gu", "Gi", "Tu", "E", "Bp", "F"], DataLoader)
    np.savez_compressed(os.path.join(HERE, "vbpr_ref.npz"), **out)
    print("golden fixtures written to", HERE)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
