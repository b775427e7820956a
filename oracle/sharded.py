"""Oracle (TEST INFRASTRUCTURE): the item-sharded BPR step of ``fvx_bpr_step_sharded`` /
``fvx_bpr_step_sharded_phase`` (include/fvx.h), restated in NumPy per rank - first with replicated user tables
(three phases, two sums), then (``run_slots`` ... ``owned_user_grads`` at the end of the file) with the users
block-OWNED as the CUDA path has them: four exchanges (WU, S, RU, dE).

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


# ---- users block-owned (FvxModel.user_lo / user_cnt; fashionvisualexpl-recommend_b200/parallel.py: user_bounds) ----
#
# The owner of a user keeps the user's Adam state and is the only rank whose copy of the row is current.  Per step:
#
#     phase 0   run_slots: the runs of equal users of the batch, grouped by owner (slot = owner * cap + index among
#               that owner's runs) - a function of the batch alone, so every rank computes the same layout;
#               the owner writes the CURRENT rows of its users into its segment of WU
#     -- all-gather(WU): every rank receives every owner's segment --
#     phase 1   partial scores of the owned slots, user rows read from WU (never from the local table)
#     -- sum(S) --
#     phase 2   as phase B above, user rows from WU, the user-gradient shares indexed by run slot
#     -- sum(RU) (each owner needs its own segment: a reduce-scatter), sum(dE) --
#     phase 3   the owner adds the RU rows of its segment into ITS users' gradients; Adam on the owned users, on E
#               (everywhere) and on the owned items
#
# tests/test_parallel_cpu.py::test_owned_users_decomposition_gloo_world2 runs it with the foreign user rows of every
# rank poisoned with NaN: nothing may ever read them.

def user_bounds(num_users, world, rank):
    per = (int(num_users) + world - 1) // world
    lo = min(int(num_users), rank * per)
    return lo, max(0, min(int(num_users), lo + per) - lo), per


def run_slots(user, per, owners, cap):
    """slot[b] = owner * cap + (index of b's run among the runs of that owner), owner = min(user // per, owners-1)
    (fvx_run_slots); raises when an owner has more than ``cap`` runs (the CUDA step poisons the loss instead)."""
    user = np.asarray(user, dtype=np.int64)
    start = np.ones(len(user), dtype=bool)
    start[1:] = user[1:] != user[:-1]
    owner = np.minimum(user // per, owners - 1)
    slot = np.zeros(len(user), dtype=np.int64)
    seen = np.zeros(owners, dtype=np.int64)
    cur = 0
    for b in range(len(user)):
        if start[b]:
            o = owner[b]
            if seen[o] >= cap:
                raise OverflowError("owner %d has more than %d runs" % (o, cap))
            cur = o * cap + seen[o]
            seen[o] += 1
        slot[b] = cur
    return slot


def publish_users(P, ulo, ucnt, batch, slot, owners, cap, vis):
    """This owner's contribution to WU [owners * cap, K + d]: the current rows of ITS users at their run slots
    (zero elsewhere, so that a sum over the ranks is the all-gather)."""
    user = np.asarray(batch[0], dtype=np.int64)
    K = P["Gu"].shape[1]
    d = P["Tu"].shape[1] if vis else 0
    WU = np.zeros((owners * cap, K + d), dtype=P["Gu"].dtype)
    mine = (user >= ulo) & (user < ulo + ucnt)
    WU[slot[mine], :K] = P["Gu"][user[mine]]
    if vis:
        WU[slot[mine], K:] = P["Tu"][user[mine]]
    return WU


def owned_scores(P, lo, cnt, batch, WU, slot, F=None):
    """phase 1: as phase_a, the user rows taken from WU."""
    user, pos, neg = (np.asarray(a, dtype=np.int64) for a in batch)
    B = len(user)
    dt = P["Gi"].dtype
    K = P["Gi"].shape[1]
    S = np.zeros(2 * B, dtype=dt)
    for side, item in ((0, pos), (1, neg)):
        own = (item >= lo) & (item < lo + cnt)
        b = np.nonzero(own)[0]
        i, w = item[b], WU[slot[b]]
        s = P["Bi"][i] + np.sum(w[:, :K] * P["Gi"][i], axis=1)
        if F is not None:
            f = F[i].astype(dt)
            s = s + np.sum(w[:, K:] * (f @ P["E"]), axis=1) + (f @ P["Bp"])[:, 0]
        S[side * B + b] = s
    return S


def owned_grads(P, lo, cnt, batch, S, WU, slot, reg, F=None):
    """phase 2: as phase_b with the user rows from WU and RU indexed by run slot ([owners * cap, K + d])."""
    user, pos, neg = (np.asarray(a, dtype=np.int64) for a in batch)
    B = len(user)
    dt = P["Gi"].dtype
    reg, two = dt.type(reg), dt.type(2)
    K = P["Gi"].shape[1]
    vis = F is not None
    x = S[:B] - S[B:]
    inside = (x >= dt.type(CLIP_LO)) & (x <= dt.type(CLIP_HI))
    c = np.where(inside, -1.0 / (1.0 + np.exp(x.astype(np.float64))), 0.0).astype(dt)
    G = {"Gi": np.zeros_like(P["Gi"]), "Bi": np.zeros_like(P["Bi"])}
    RU = np.zeros_like(WU)
    dE = np.zeros_like(P["E"]) if vis else None
    dBp = np.zeros_like(P["Bp"]) if vis else None
    loss = dt.type(0)
    for side, item in ((0, pos), (1, neg)):
        own = (item >= lo) & (item < lo + cnt)
        b = np.nonzero(own)[0]
        i, w = item[b], WU[slot[b]]
        gu = w[:, :K]
        cs = c[b] if side == 0 else -c[b]
        gi, bi = P["Gi"][i], P["Bi"][i]
        breg = reg if side == 0 else reg / dt.type(10)
        np.add.at(G["Gi"], i, cs[:, None] * gu + two * reg * gi)
        np.add.at(G["Bi"], i, cs + two * breg * bi)
        ushare = cs[:, None] * gi
        loss = loss + reg * np.sum(gi * gi, dtype=dt) + breg * np.sum(bi * bi, dtype=dt)
        if side == 0:
            ushare = ushare + two * reg * gu
            xc = np.clip(x[b], dt.type(CLIP_LO), dt.type(CLIP_HI))
            loss = loss + np.sum(softplus(-xc), dtype=dt) + reg * np.sum(gu * gu, dtype=dt)
        np.add.at(RU[:, :K], slot[b], ushare)
        if vis:
            tu, f = w[:, K:], F[i].astype(dt)
            tshare = cs[:, None] * (f @ P["E"])
            if side == 0:
                tshare = tshare + two * reg * tu
                loss = loss + reg * np.sum(tu * tu, dtype=dt)
            np.add.at(RU[:, K:], slot[b], tshare)
            dE += f.T @ (cs[:, None] * tu)
            dBp += f.T @ cs[:, None]
    return G, RU, dE, dBp, loss


def owned_user_grads(P, ulo, ucnt, batch, slot, RU, vis):
    """phase 3: the summed RU rows of the OWNED users' runs -> (Gu, Tu) gradients (zero rows for foreign users)."""
    user = np.asarray(batch[0], dtype=np.int64)
    K = P["Gu"].shape[1]
    first = np.ones(len(user), dtype=bool)
    first[1:] = slot[1:] != slot[:-1]
    mine = first & (user >= ulo) & (user < ulo + ucnt)
    gGu = np.zeros_like(P["Gu"])
    np.add.at(gGu, user[mine], RU[slot[mine], :K])
    gTu = None
    if vis:
        gTu = np.zeros_like(P["Tu"])
        np.add.at(gTu, user[mine], RU[slot[mine], K:])
    return gGu, gTu
