#!/usr/bin/env python
"""Golden vectors for the GradFashion oracle (oracle/gradfashion.py): runs the REFERENCE's own
src/recommender/models/GradFashion.py (read from /root/reference, never copied) over the torch-backed
tensorflow shim, on the committed tiny dataset plus seeded colour / edge descriptors.

    python tests/golden/make_golden_gradfashion.py      # writes tests/golden/gradfashion_ref.npz

The reference's CLI never defines ``embed_color`` / ``embed_edges`` (train_rec.py:17-46), so the model
cannot be reached through train_rec.py as shipped; the namespace below adds them.  ``train()`` needs
Evaluator.store_recommendation_grads and TF checkpoints - only __init__, call, train_step and
predict_all are exercised.
"""
import argparse
import os
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

NAMES = ["Bi", "Gu", "Gi"]


def snap(model):
    out = {n: getattr(model, n).numpy().copy() for n in NAMES}
    out["Ec"] = model.color_weights["Ec"].numpy().copy()
    out["Ee"] = model.edges_weights["Ee"].numpy().copy()
    for n in ("Tu", "E", "Bp"):
        out[n] = model.visual_profile[n].numpy().copy()
    return out


def main():
    tmp = tempfile.mkdtemp(prefix="fvx_golden_gf_")
    os.makedirs(os.path.join(tmp, "src"))
    data_dir = os.path.join(tmp, "data", "tiny")
    shutil.copytree(os.path.join(HERE, "tiny"), data_dir)
    rng = np.random.default_rng(23)
    I = 120
    color = rng.random((I, 24)) * (rng.random((I, 24)) < 0.5)          # sparse histogram-like
    edges = np.maximum(rng.standard_normal((I, 16)), 0) * 3.0
    os.makedirs(os.path.join(data_dir, "original", "features"), exist_ok=True)
    np.save(os.path.join(data_dir, "original", "features", "histograms.npy"), color)
    np.save(os.path.join(data_dir, "original", "edge_features_resnet50_avg_pool.npy"), edges)

    os.chdir(os.path.join(tmp, "src"))
    sys.path.insert(0, REF_SRC)
    tf_shim.install()
    from dataset.dataset import DataLoader
    from recommender.models.GradFashion import GradFashion

    args = argparse.Namespace(gpu=-1, best_metric="ndcg", dataset="tiny", rec="grad_fashion", batch_size=16,
                              top_k=5, epochs=1, verbose=-1, batch_eval=128, lr=0.01, validation=True,
                              restore_epochs=1, list_of_regs=[1e-3], cnn_model="resnet50",
                              output_layer="avg_pool", embed_k=8, embed_d=4, reg=1e-3, embed_color=6,
                              embed_edges=5)
    random.seed(0)
    np.random.seed(0)
    tf_shim._set_seed(0)
    data = DataLoader(params=args)
    model = GradFashion(data, args)
    out = {"init_" + k: v for k, v in snap(model).items()}
    out["Fc"] = model.color_weights["Fc"].numpy().copy()
    out["Fe"] = model.edges_weights["Fe"].numpy().copy()
    users, pos, neg, losses = [], [], [], []
    x0 = None
    for s, batch in enumerate(data.next_triple_batch()):
        if s == 0:
            x0 = model(inputs=(batch[0], batch[1]), training=True)[0].numpy().copy()
        b = [t.numpy().copy() for t in batch]
        users.append(b[0]); pos.append(b[1]); neg.append(b[2])
        losses.append(float(model.train_step(batch)))
        if len(losses) == 1:
            out.update({"step1_" + k: v for k, v in snap(model).items()})
        if len(losses) == 10:
            break
    out.update({"final_" + k: v for k, v in snap(model).items()})
    out["users"], out["pos"], out["neg"] = np.array(users), np.array(pos), np.array(neg)
    out["losses"] = np.array(losses)
    out["call_x0"] = x0
    out["predict_all_final"] = model.predict_all().numpy()
    out["hyper"] = np.array([args.lr, args.reg])
    np.savez_compressed(os.path.join(HERE, "gradfashion_ref.npz"), **out)
    print("written", os.path.join(HERE, "gradfashion_ref.npz"), "losses", losses[:3])
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
