"""Evaluator with the reference's protocol (src/recommender/Evaluator.py:131-239) whose
work happens in one fused device sweep instead of a dense ``[U, I]`` matrix plus
per-user Python loops.

* ``Evaluator(model, data, k)``; ``.eval(epoch, results, epoch_text, start_time)`` fills
  ``results[epoch]`` with the reference's keys (incl. its ``auc_t = auc_v`` quirk, :220)
  and returns / prints the same text block (:194-215);
* ``.store_recommendation(path)`` writes ``user \\t item \\t score`` for the top-k items
  that are not in the user's training list (:231-239).

The reference's candidate lists (all items minus train items, held-out items last,
:36-79) are never materialised: the sweep counts, per held-out item, how many non-train
items score >= it (``fvx_rank_counts``: a register-tiled sweep that keeps no lists), and the metrics of ``_eval_by_user``
(:82-128) follow from those counts - position (:96-98), AUC (:100), top-K membership
under ``heapq.nlargest``'s tie rule (held-out items come last, :104-114), nDCG (:119),
precision (:122), recall (:125).
"""
from __future__ import annotations

import datetime
import math
from time import time

import numpy as np
import torch


class Evaluator:
    def __init__(self, model, data, k):
        self.data = data
        self.batch_eval = getattr(self.data.params, "batch_eval", 128)
        self.k = k
        self.model = model
        self._held = None

    # ---- held-out items as a padded [U, T] matrix (validation columns, then test columns) ---
    def _held_matrix(self):
        if self._held is None:
            d = self.data
            vl = np.diff(d.val_ptr) if len(d.val_ptr) > 1 else np.zeros(d.num_users, np.int64)
            tl = np.diff(d.test_ptr)
            Tv, Tt = int(vl.max()) if vl.size else 0, int(tl.max()) if tl.size else 0
            H = np.full((d.num_users, Tv + Tt), -1, dtype=np.int32)
            for ptr_, col, off, T in ((d.val_ptr, d.val_col, 0, Tv), (d.test_ptr, d.test_col, Tv, Tt)):
                if T == 0:
                    continue
                lens = np.diff(ptr_)
                owner = np.repeat(np.arange(d.num_users), lens)
                within = np.arange(int(ptr_[-1])) - np.repeat(ptr_[:-1], lens)
                H[owner, off + within] = col
            in_train = np.zeros(H.shape, dtype=bool)
            for t in range(H.shape[1]):
                ok = H[:, t] >= 0
                in_train[ok, t] = d._member(np.nonzero(ok)[0].astype(np.int64), H[ok, t].astype(np.int64))
            self._held = (H, in_train, Tv, Tt)
        return self._held

    def _rank_counts(self):
        """thr scores [U,T] (NaN where there is no held-out item) and counts [U,T] of
        non-train items scoring >= them, in passes of at most 4 thresholds."""
        e, d = self.model.engine, self.data
        H, _, _, _ = self._held_matrix()
        U, T = H.shape
        st = d.device_state(e.device)
        Hd = torch.from_numpy(H).to(e.device)
        users = torch.arange(U, dtype=torch.int32, device=e.device).repeat_interleave(T)
        sc = e.score_pairs(users, torch.clamp(Hd.reshape(-1), min=0).contiguous()).reshape(U, T)
        sc = torch.where(Hd >= 0, sc, torch.full_like(sc, float("nan")))
        counts = e.rank_counts(st["row_ptr"], st["col_sorted"], sc.contiguous())
        return sc.cpu().numpy(), counts.cpu().numpy().astype(np.int64)

    def user_metrics(self):
        """Per-user (hr, prec, rec, auc, ndcg) arrays for the validation and the test split
        (NaN rows for users without held-out items): the vectorised ``_eval_by_user``."""
        d = self.data
        H, in_train, Tv, Tt = self._held_matrix()
        sc, cnt = self._rank_counts()
        n_train = np.diff(d.train_ptr)
        out = {}
        for name, cols in (("v", range(0, Tv)), ("t", range(Tv, Tv + Tt))):
            cols = list(cols)
            U = d.num_users
            valid = np.stack([H[:, t] >= 0 for t in cols], 1) if cols else np.zeros((U, 0), bool)
            h = valid.sum(1)
            counted = np.stack([valid[:, j] & ~in_train[:, t] for j, t in enumerate(cols)], 1) if cols \
                else np.zeros((U, 0), bool)
            n_neg = d.num_items - n_train - counted.sum(1)
            position = np.zeros(U, dtype=np.int64)
            hits = np.zeros(U, dtype=np.int64)
            with np.errstate(invalid="ignore"):
                for j, t in enumerate(cols):
                    neg_ge = cnt[:, t].copy()
                    rank = np.zeros(U, dtype=np.int64)
                    for j2, t2 in enumerate(cols):
                        ge = sc[:, t2] >= sc[:, t]
                        neg_ge -= (counted[:, j2] & ge).astype(np.int64)
                        if j2 < j:
                            rank += (valid[:, j2] & ge).astype(np.int64)
                        elif j2 > j:
                            rank += (valid[:, j2] & (sc[:, t2] > sc[:, t])).astype(np.int64)
                    neg_ge = np.where(valid[:, j], neg_ge, 0)
                    position += neg_ge
                    hits += (valid[:, j] & (neg_ge + rank < self.k)).astype(np.int64)
            ok = h > 0
            hs = np.maximum(h, 1)
            len_r = np.minimum(self.k, n_neg + h)
            auc = 1 - position / np.maximum(n_neg * hs, 1)
            ndcg = np.where(position < self.k, math.log(2) / np.log(position + 2.0), 0.0)
            res = np.stack([(hits > 0).astype(np.float64), hits / np.maximum(len_r, 1), hits / hs, auc, ndcg], 1)
            res[~ok] = np.nan
            out[name] = res
        return out

    def eval(self, epoch=0, results={}, epoch_text='', start_time=0, attentive=False):
        """Runtime evaluation of accuracy (top-k); same contract as Evaluator.py:149-223."""
        eval_start_time = time()
        m = self.user_metrics()
        hr_v, prec_v, rec_v, auc_v, ndcg_v = '0', '0', '0', '0', '0'
        rt = m["t"][~np.isnan(m["t"][:, 0])]
        hr_t, prec_t, rec_t, auc_t, ndcg_t = rt.mean(axis=0).tolist()
        has_val = bool(self.data.validation_list) and m["v"].shape[1] > 0 and (~np.isnan(m["v"][:, 0])).any()
        if has_val:
            rv = m["v"][~np.isnan(m["v"][:, 0])]
            hr_v, prec_v, rec_v, auc_v, ndcg_v = rv.mean(axis=0).tolist()
        print_results = \
            "%s \tTrain Time: %s \tEvaluation Time: %s" \
            "\nMetrics@%d (Validation)\n\t\tHR\tPrec\tRec\tAUC\tnDCG\n\t\t%f\t%f\t%f\t%f\t%f" \
            "\nMetrics@%d (Test)\n\t\tHR\tPrec\tRec\tAUC\tnDCG\n\t\t%f\t%f\t%f\t%f\t%f\n" % (
                epoch_text,
                datetime.timedelta(seconds=(time() - start_time)),
                datetime.timedelta(seconds=(time() - eval_start_time)),
                self.k, float(hr_v), float(prec_v), float(rec_v), float(auc_v), float(ndcg_v),
                self.k, hr_t, prec_t, rec_t, auc_t, ndcg_t)
        print(print_results)
        results[epoch] = {
            'hr_v': hr_v, 'auc_v': auc_v, 'p_v': prec_v, 'r_v': rec_v, 'ndcg_v': ndcg_v,
            'hr_t': hr_t, 'auc_t': auc_v, 'p_t': prec_t, 'r_t': rec_t, 'ndcg_t': ndcg_t,   # auc_t quirk :220
            'auc_t_fixed': auc_t,
        }
        return print_results

    def topk(self, k=None):
        """(ids [U,k] int32, scores [U,k] fp32) on the host; train items masked."""
        e, d = self.model.engine, self.data
        st = d.device_state(e.device)
        ids, sc = e.score_topk(st["row_ptr"], st["col_sorted"], self.k if k is None else k)
        return ids.cpu().numpy(), sc.cpu().numpy()

    def store_recommendation(self, path=""):
        """Top-k dump in the reference's TSV format (Evaluator.py:233-239)."""
        ids, sc = self.topk()
        write_recs_tsv(path, ids, sc)


def write_recs_tsv(path, ids, sc):
    U, k = ids.shape
    us = np.repeat(np.arange(U), k)
    flat_i, flat_s = ids.reshape(-1), sc.reshape(-1).astype(np.float32)
    keep = flat_i >= 0
    su, si, ss = us[keep].astype(str), flat_i[keep].astype(str), flat_s[keep].astype(str)
    with open(path, 'w') as out:
        out.write("".join("%s\t%s\t%s\n" % t for t in zip(su, si, ss)))
