"""Oracle (TEST INFRASTRUCTURE): BPRMF / VBPR train step and predict_all on the CPU.

Restates, in NumPy, the arithmetic of

* ``BPRMF.call``        src/recommender/models/BPRMF.py:55-76
* ``BPRMF.predict_all`` src/recommender/models/BPRMF.py:78-85
* ``BPRMF.train_step``  src/recommender/models/BPRMF.py:87-125
* ``VBPR.call``         src/recommender/models/VBPR.py:59-86
* ``VBPR.predict_all``  src/recommender/models/VBPR.py:88-97
* ``VBPR.train_step``   src/recommender/models/VBPR.py:99-144

and of the third-party optimiser those files call (``tf.optimizers.Adam`` of
``tensorflow==2.3.1``, requirements.txt:42): sparse gradients are summed per
unique row, then *every* row of every table takes the Adam step (it is dense
Adam on the scatter-summed gradient, eps added to sqrt(v), both bias
corrections folded into the step size).  That last part is "parity unpinned"
(see oracle/__init__.py).

All arithmetic is done in ``dtype`` (float32 like the reference, or float64 to
measure which of two fp32 implementations is closer to the real number).
"""
from __future__ import annotations

import numpy as np

BETA1 = 0.9
BETA2 = 0.999
EPS = 1e-7          # Keras default epsilon (tf.keras.backend.epsilon())
CLIP_LO = -80.0     # BPRMF.py:104 / VBPR.py:117
CLIP_HI = 1e8

SPARSE = ("Gu", "Gi", "Bi", "Tu")
DENSE = ("E", "Bp")


def glorot_uniform(rng: np.random.Generator, shape, dtype=np.float32):
    """tf.initializers.GlorotUniform on a 2-D shape: U(-L, L), L = sqrt(6/(r+c)).

    BPRMF.py:35,48-50.  TF's own RNG stream cannot be reproduced offline, so the
    initial values are oracle *inputs*; this helper only reproduces the law.
    """
    r, c = shape
    lim = np.sqrt(6.0 / (r + c))
    return rng.uniform(-lim, lim, size=shape).astype(dtype)


def init_params(num_users, num_items, K, d=0, D=0, seed=0, dtype=np.float32):
    """Parameter set with the reference's shapes (BPRMF.py:48-50, VBPR.py:44-54)."""
    rng = np.random.default_rng(seed)
    P = {
        "Bi": np.zeros(num_items, dtype=dtype),
        "Gu": glorot_uniform(rng, (num_users, K), dtype),
        "Gi": glorot_uniform(rng, (num_items, K), dtype),
    }
    if D > 0:
        P["Bp"] = glorot_uniform(rng, (D, 1), dtype)
        P["Tu"] = glorot_uniform(rng, (num_users, d), dtype)
        P["E"] = glorot_uniform(rng, (D, d), dtype)
    return P


def init_adam(P):
    S = {"t": 0}
    for k, v in P.items():
        S["m" + k] = np.zeros_like(v)
        S["v" + k] = np.zeros_like(v)
    return S


def softplus(x):
    """tf.nn.softplus (Eigen): x for large x, exp(x) for very negative x."""
    dt = x.dtype
    thr = dt.type(np.log(np.finfo(np.float32).eps) + 2.0)  # about -13.94
    out = np.log1p(np.exp(np.minimum(x, -thr)))
    out = np.where(x > -thr, x, out)
    out = np.where(x < thr, np.exp(np.minimum(x, dt.type(0))), out)
    return out.astype(dt)


def score(P, user, item, F=None):
    """x_ui of ``call`` (BPRMF.py:69-74; VBPR.py:73-84)."""
    x = P["Bi"][item] + np.sum(P["Gu"][user] * P["Gi"][item], axis=1)
    if F is not None:
        f = F[item]
        x = x + np.sum(P["Tu"][user] * (f @ P["E"]), axis=1) + (f @ P["Bp"])[:, 0]
    return x


def loss_and_grads(P, batch, reg, F=None):
    """Loss (BPRMF.py:104-115 / VBPR.py:117-130) and its gradient w.r.t. every
    trainable tensor, duplicates scatter-summed (what tape.gradient + the
    IndexedSlices de-duplication hand to Adam).  Closed forms: SURVEY.md App. A.
    """
    user, pos, neg = (np.asarray(a, dtype=np.int64) for a in batch)
    dt = P["Gu"].dtype
    reg = dt.type(reg)
    gu, gi, gj = P["Gu"][user], P["Gi"][pos], P["Gi"][neg]
    bi, bj = P["Bi"][pos], P["Bi"][neg]
    xi = bi + np.sum(gu * gi, axis=1)
    xj = bj + np.sum(gu * gj, axis=1)
    vis = F is not None
    if vis:
        tu = P["Tu"][user]
        fi, fj = F[pos].astype(dt), F[neg].astype(dt)
        thi, thj = fi @ P["E"], fj @ P["E"]
        xi = xi + np.sum(tu * thi, axis=1) + (fi @ P["Bp"])[:, 0]
        xj = xj + np.sum(tu * thj, axis=1) + (fj @ P["Bp"])[:, 0]
    x = xi - xj
    xc = np.clip(x, dt.type(CLIP_LO), dt.type(CLIP_HI))
    loss = np.sum(softplus(-xc), dtype=dt)
    two = dt.type(2)
    regl = reg * (np.sum(gu * gu, dtype=dt) + np.sum(gi * gi, dtype=dt) + np.sum(gj * gj, dtype=dt))
    if vis:
        regl += reg * np.sum(tu * tu, dtype=dt)
    regl += reg * np.sum(bi * bi, dtype=dt) + reg * np.sum(bj * bj, dtype=dt) / dt.type(10)
    if vis:
        regl += reg * (np.sum(P["E"] * P["E"], dtype=dt) + np.sum(P["Bp"] * P["Bp"], dtype=dt))
    loss = loss + regl

    inside = (x >= dt.type(CLIP_LO)) & (x <= dt.type(CLIP_HI))
    # c = d softplus(-x)/dx = -sigmoid(-x)
    c = np.where(inside, -1.0 / (1.0 + np.exp(x.astype(np.float64))), 0.0).astype(dt)
    G = {k: np.zeros_like(v) for k, v in P.items()}
    np.add.at(G["Gu"], user, c[:, None] * (gi - gj) + two * reg * gu)
    np.add.at(G["Gi"], pos, c[:, None] * gu + two * reg * gi)
    np.add.at(G["Gi"], neg, -c[:, None] * gu + two * reg * gj)
    np.add.at(G["Bi"], pos, c + two * reg * bi)
    np.add.at(G["Bi"], neg, -c + (two * reg / dt.type(10)) * bj)
    if vis:
        np.add.at(G["Tu"], user, c[:, None] * (thi - thj) + two * reg * tu)
        df = fi - fj
        G["E"] = df.T @ (c[:, None] * tu) + two * reg * P["E"]
        G["Bp"] = df.T @ c[:, None] + two * reg * P["Bp"]
    return loss, G, x


def adam_alpha(t, lr):
    """Keras Adam step size at 1-based step t (float64; callers cast)."""
    return lr * np.sqrt(1.0 - BETA2 ** t) / (1.0 - BETA1 ** t)


def adam_apply(P, S, G, lr):
    """Dense-semantics Keras Adam on every row of every variable."""
    S["t"] += 1
    t = S["t"]
    for k in P:
        dt = P[k].dtype
        a = dt.type(adam_alpha(t, lr))
        b1, b2, eps = dt.type(BETA1), dt.type(BETA2), dt.type(EPS)
        m, v, g = S["m" + k], S["v" + k], G[k]
        m *= b1
        m += (dt.type(1) - b1) * g
        v *= b2
        v += (dt.type(1) - b2) * (g * g)
        P[k] -= a * m / (np.sqrt(v) + eps)


def train_step(P, S, batch, reg, lr, F=None):
    """One ``train_step`` (BPRMF.py:87-125 / VBPR.py:99-144): returns the loss."""
    loss, G, _ = loss_and_grads(P, batch, reg, F)
    adam_apply(P, S, G, lr)
    return float(loss)


def predict_all(P, F=None, users=None):
    """``predict_all`` (BPRMF.py:85; VBPR.py:95-97), optionally on a user slice."""
    Gu = P["Gu"] if users is None else P["Gu"][users]
    S = P["Bi"][None, :] + Gu @ P["Gi"].T
    if F is not None:
        Tu = P["Tu"] if users is None else P["Tu"][users]
        dt = P["Gu"].dtype
        S = S + Tu @ (F.astype(dt) @ P["E"]).T + (F.astype(dt) @ P["Bp"])[:, 0][None, :]
    return S


def normalise_features(F_raw, dtype=np.float32):
    """visual_loader_mixin.py:30 : one global scale by max(abs(F)); VBPR.py:49-51 cast."""
    F_raw = np.asarray(F_raw)
    return (F_raw / np.max(np.abs(F_raw))).astype(dtype)
