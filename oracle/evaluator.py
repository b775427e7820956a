"""Oracle (TEST INFRASTRUCTURE): per-user metrics and the masked top-k dump.

Restates src/recommender/Evaluator.py:

* candidate lists ``_evaluate_input_list_{test,validation}``   :36-79
* ``_eval_by_user`` (AUC, HR, nDCG, precision, recall)          :82-128
* ``Evaluator.eval`` means and result keys                      :149-223
* ``Evaluator.store_recommendation`` (mask train, top-k)        :225-239

The reference builds a Python dict over all candidates and calls
``heapq.nlargest`` (:104-108): on ties the candidate that comes first in
candidate order wins (ascending item id, held-out items last).  This file keeps
that tie rule; ``store_recommendation``'s ``argsort()[-k:][::-1]`` (:236) has no
defined tie order, so the top-k comparison helper below is tie-aware.
"""
from __future__ import annotations

import math

import numpy as np


def candidates(num_items, train_items, held_items):
    """Evaluator.py:36-53: all items minus train items, held-out items moved last."""
    mask = np.ones(num_items, dtype=bool)
    mask[np.asarray(train_items, dtype=np.int64)] = False
    held = list(held_items)
    for h in held:
        mask[h] = False
    return np.concatenate([np.nonzero(mask)[0], np.asarray(held, dtype=np.int64)])


def eval_by_user(row, num_items, train_items, held_items, k):
    """Evaluator.py:82-128 for one user and one split; ``row`` = predict_all()[u]."""
    n_held = len(held_items)
    if n_held == 0:
        return ()
    cand = candidates(num_items, train_items, held_items)
    preds = row[cand]
    neg, pos = preds[:-n_held], preds[-n_held:]
    position = 0
    for t in range(n_held):
        position += int((neg >= pos[t]).sum())                  # :96-98, '>=' tie rule
    auc = 1 - (position / (len(neg) * len(pos)))                 # :100
    # heapq.nlargest(K, dict, key=get): stable -> descending score, then candidate order
    order = np.argsort(-row[cand].astype(np.float64), kind="stable")[:k]
    top = cand[order]
    held_set = set(int(h) for h in held_items)
    r = [1 if int(i) in held_set else 0 for i in top]
    hr = 1.0 if sum(r) > 0 else 0.0
    ndcg = math.log(2) / math.log(position + 2) if position < k else 0    # :119
    prec = sum(r) / len(r)
    rec = sum(r) / len(pos)
    return hr, prec, rec, auc, ndcg


def evaluate(scores, num_items, training_list, validation_list, test_list, k):
    """Evaluator.eval's ``results[epoch]`` dict (:189-221), incl. the auc_t = auc_v quirk."""
    res_t, res_v = [], []
    for u in range(scores.shape[0]):
        res_t.append(eval_by_user(scores[u], num_items, training_list[u], test_list[u], k))
        if validation_list:
            res_v.append(eval_by_user(scores[u], num_items, training_list[u], validation_list[u], k))
    res_t = [r for r in res_t if r]
    hr_t, p_t, r_t, auc_t, ndcg_t = np.array(res_t).mean(axis=0).tolist()
    hr_v = p_v = r_v = auc_v = ndcg_v = "0"
    if validation_list:
        res_v = [r for r in res_v if r]
        hr_v, p_v, r_v, auc_v, ndcg_v = np.array(res_v).mean(axis=0).tolist()
    return {"hr_v": hr_v, "auc_v": auc_v, "p_v": p_v, "r_v": r_v, "ndcg_v": ndcg_v,
            "hr_t": hr_t, "auc_t": auc_v, "p_t": p_t, "r_t": r_t, "ndcg_t": ndcg_t,
            "auc_t_fixed": auc_t}


def masked_topk(scores, training_list, k):
    """store_recommendation (:231-237): mask *train* items to -inf, top-k descending.

    Ties are broken towards the smaller item id (the reference leaves it undefined).
    Returns (ids [U,k] int64, scores [U,k])."""
    U, I = scores.shape
    ids = np.zeros((U, k), dtype=np.int64)
    val = np.zeros((U, k), dtype=scores.dtype)
    for u in range(U):
        row = scores[u].copy()
        row[np.asarray(training_list[u], dtype=np.int64)] = -np.inf
        order = np.argsort(-row.astype(np.float64), kind="stable")[:k]
        ids[u], val[u] = order, row[order]
    return ids, val


def topk_matches(ids_a, val_a, ids_b, val_b, rel_gap=1e-5, rel_tol=1e-4):
    """Tie-aware comparison of two top-k lists of one user:
    ids must agree wherever the adjacent score gaps of ``b`` exceed
    rel_gap*max(1,|s|); inside a near-tie run only the id *sets* must agree,
    except for members tied with the k-th score (the boundary run may differ).
    Scores must agree within rel_tol.  Returns (ok, message)."""
    k = len(ids_b)
    va = np.asarray(val_a, dtype=np.float64)
    vb = np.asarray(val_b, dtype=np.float64)
    fin = np.isfinite(vb)
    if not np.allclose(va[fin], vb[fin], rtol=rel_tol, atol=rel_tol * 1e-3 + 1e-7):
        return False, "scores differ: max abs %g" % np.max(np.abs(va[fin] - vb[fin]))
    start = 0
    while start < k:
        end = start + 1
        while end < k and abs(vb[end - 1] - vb[end]) <= rel_gap * max(1.0, abs(vb[end])):
            end += 1
        sa, sb = set(ids_a[start:end].tolist()), set(ids_b[start:end].tolist())
        if sa != sb and end < k:
            return False, "ids differ in positions %d..%d: %s vs %s" % (start, end, sa, sb)
        start = end
    return True, ""
