"""Oracle (TEST INFRASTRUCTURE): the item-sharded BPR step as the three phases of
``fvx_bpr_step_sharded_a/b/c`` (include/fvx.h), restated in NumPy per rank.

The reference has no multi-GPU path; what is restated here is the DECOMPOSITION of its train step
(BPRMF.py:87-125 / VBPR.py:99-144, oracle/bpr.py) that the CUDA path uses: x_uij = s_ui - s_uj is linear in
the item-side terms, so each rank scores the (triple, side) slots whose item it owns, an all-reduce
assembles every x, and each rank then contributes the gradient terms tied to ITS items:

    phase A   S[slot] = Bi[i] + <Gu[u], Gi[i]> + <Tu[u], F[i] E> + F[i] Bp      for owned slots, else 0
    -- all-reduce(S) --
    phase B   c_b from x_b = S[b] - S[B+b]; gradients of the owned item rows; this rank's share of the
              user-row gradients, packed by RUN of equal users (RU[run_id[b]] += ...); its share of
              dE / dBp; the softplus loss and the user-side L2 of a triple belong to the rank that owns
              the POSITIVE item, an item-side L2 to the owner of that item
    -- all-reduce(RU), all-reduce(dE) --
    phase C   RU rows -> user gradients; Adam on users and E (identical on every rank), on the owned items

tests/test_parallel_cpu.py runs these phases in two gloo processes with real all-reduces and compares
with the single-rank oracle.
"""
from __future__ import annotations

import numpy as np

from .bpr import CLIP_HI, CLIP_LO, softplus


def run_ids(user):
    user = np.asarray(user)
    start = np.ones(len(user), dtype=np.int64)
    start[1:] = user[1:] != user[:-1]
    return np.cumsum(start) - 1


def phase_a(P, lo, cnt, batch, F=None):
    """Partial scores of the slots [pos(B) | neg(B)] whose item lies in [lo, lo+cnt)."""
    user, pos, neg = (np.asarray(a, dtype=np.int64) for a in batch)
    B = len(user)
    dt = P["Gu"].dtype
    S = np.zeros(2 * B, dtype=dt)
    for side, item in ((0, pos), (1, neg)):
        own = (item >= lo) & (item < lo + cnt)
        u, i = user[own], item[own]
        s = P["Bi"][i] + np.sum(P["Gu"][u] * P["Gi"][i], axis=1)
        if F is not None:
            f = F[i].astype(dt)
            s = s + np.sum(P["Tu"][u] * (f @ P["E"]), axis=1) + (f @ P["Bp"])[:, 0]
        S[side * B + np.nonzero(own)[0]] = s
    return S


def phase_b(P, lo, cnt, batch, S, reg, n_runs, F=None):
    """This rank's shares: (G_items dict of full-size arrays touched only on owned rows, RU [n_runs, K+d],
    dE, dBp, loss share)."""
    user, pos, neg = (np.asarray(a, dtype=np.int64) for a in batch)
    B = len(user)
    dt = P["Gu"].dtype
    reg, two = dt.type(reg), dt.type(2)
    K = P["Gu"].shape[1]
    vis = F is not None
    d = P["Tu"].shape[1] if vis else 0
    x = S[:B] - S[B:]
    inside = (x >= dt.type(CLIP_LO)) & (x <= dt.type(CLIP_HI))
    c = np.where(inside, -1.0 / (1.0 + np.exp(x.astype(np.float64))), 0.0).astype(dt)
    rid = run_ids(user)
    G = {"Gi": np.zeros_like(P["Gi"]), "Bi": np.zeros_like(P["Bi"])}
    RU = np.zeros((n_runs, K + d), dtype=dt)
    dE = np.zeros_like(P["E"]) if vis else None
    dBp = np.zeros_like(P["Bp"]) if vis else None
    loss = dt.type(0)
    for side, item in ((0, pos), (1, neg)):
        own = (item >= lo) & (item < lo + cnt)
        b = np.nonzero(own)[0]
        u, i = user[b], item[b]
        cs = c[b] if side == 0 else -c[b]
        gu, gi, bi = P["Gu"][u], P["Gi"][i], P["Bi"][i]
        breg = reg if side == 0 else reg / dt.type(10)
        np.add.at(G["Gi"], i, cs[:, None] * gu + two * reg * gi)
        np.add.at(G["Bi"], i, cs + two * breg * bi)
        ushare = cs[:, None] * gi
        loss = loss + reg * np.sum(gi * gi, dtype=dt) + breg * np.sum(bi * bi, dtype=dt)
        if side == 0:                                    # per-triple terms: owner of the positive item
            ushare = ushare + two * reg * gu
            xc = np.clip(x[b], dt.type(CLIP_LO), dt.type(CLIP_HI))
            loss = loss + np.sum(softplus(-xc), dtype=dt) + reg * np.sum(gu * gu, dtype=dt)
        np.add.at(RU[:, :K], rid[b], ushare)
        if vis:
            tu, f = P["Tu"][u], F[i].astype(dt)
            tshare = cs[:, None] * (f @ P["E"])
            if side == 0:
                tshare = tshare + two * reg * tu
                loss = loss + reg * np.sum(tu * tu, dtype=dt)
            np.add.at(RU[:, K:], rid[b], tshare)
            dE += f.T @ (cs[:, None] * tu)
            dBp += f.T @ cs[:, None]
    return G, RU, dE, dBp, loss


def phase_c_grads(P, batch, G_items, RU, dE, dBp, reg, add_e_reg=True):
    """All-reduced RU / dE -> the gradient dict oracle.bpr.adam_apply takes (item rows: this rank's own).
    Returns (G, loss term of the dense regulariser - counted once, by rank 0)."""
    user = np.asarray(batch[0], dtype=np.int64)
    dt = P["Gu"].dtype
    reg, two = dt.type(reg), dt.type(2)
    K = P["Gu"].shape[1]
    rid = run_ids(user)
    first = np.ones(len(user), dtype=bool)
    first[1:] = rid[1:] != rid[:-1]
    G = {"Gu": np.zeros_like(P["Gu"]), "Gi": G_items["Gi"], "Bi": G_items["Bi"]}
    np.add.at(G["Gu"], user[first], RU[rid[first], :K])
    extra = dt.type(0)
    if dE is not None:
        G["Tu"] = np.zeros_like(P["Tu"])
        np.add.at(G["Tu"], user[first], RU[rid[first], K:])
        G["E"] = dE + two * reg * P["E"]
        G["Bp"] = dBp + two * reg * P["Bp"]
        if add_e_reg:
            extra = reg * (np.sum(P["E"] * P["E"], dtype=dt) + np.sum(P["Bp"] * P["Bp"], dtype=dt))
    return G, extra
