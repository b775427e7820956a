"""Oracle (TEST INFRASTRUCTURE): the DEFERRED form of Keras Adam the CUDA step uses for the embedding tables
(``replay_row`` / ``k_prep`` / ``fvx_adam_flush`` in fashionvisualexpl-recommend_b200/csrc/fvx_train.cu), restated
in NumPy next to the literal form (oracle/bpr.py: ``adam_apply``).

TF 2.3.1's Adam on ``IndexedSlices`` (the gradients of ``tf.nn.embedding_lookup``, BPRMF.py:70-72 / VBPR.py:74-80,
applied at BPRMF.py:123 / VBPR.py:142) sums duplicate indices and then moves EVERY row of the table, every step:
``m <- b1 m + (1-b1) g``, ``v <- b2 v + (1-b2) g^2``, ``w <- w - alpha_t m / (sqrt(v) + eps)`` with g = 0 for the rows
the batch did not touch.  Sweeping whole tables per step is what dominates the reference at scale; the
deferred form gives the same numbers by bringing a row up to date only when it is next needed:

* a row carries ``last`` (the number of steps it is current to) and ``g`` (the gradient of its last touch, still
  to be applied as step ``last + 1``);
* ``catch_up(row, target)``: the pending step ``last + 1`` with ``g``, then the ``target - last - 1`` zero-gradient
  steps it skipped - at most ``REPLAY_MAX`` of them one by one (``m`` has decayed by 0.9^192 < 2e-9 by then: later
  steps move ``w`` by nothing representable), the rest as a closed-form decay of ``m`` and ``v``;
* a step first catches the rows it touches up to the previous step, then leaves their new gradient pending;
* ``flush`` catches every row up to the current step (before parameters are read).

tests/test_oracle_deferred_adam.py drives both forms with the same random touch pattern.
"""
from __future__ import annotations

import numpy as np

from .bpr import BETA1, BETA2, EPS, adam_alpha

REPLAY_MAX = 192        # FVX_REPLAY_MAX


class DeferredTable:
    """One embedding table [rows, cols] under deferred Adam."""

    def __init__(self, w, lr):
        self.w = np.array(w, dtype=np.float64)
        self.m = np.zeros_like(self.w)
        self.v = np.zeros_like(self.w)
        self.g = np.zeros_like(self.w)
        self.last = np.zeros(self.w.shape[0], dtype=np.int64)
        self.step = 0                       # completed optimiser steps
        self.lr = lr

    def catch_up(self, r, target):
        gap = target - self.last[r]
        if gap <= 0:
            return
        t = self.last[r] + 1                # the pending step
        w, m, v, g = self.w[r], self.m[r], self.v[r], self.g[r]
        m[:] = BETA1 * m + (1 - BETA1) * g
        v[:] = BETA2 * v + (1 - BETA2) * g * g
        w -= adam_alpha(t, self.lr) * m / (np.sqrt(v) + EPS)
        nz = gap - 1
        n = min(nz, REPLAY_MAX)
        for k in range(n):
            m *= BETA1
            v *= BETA2
            w -= adam_alpha(t + 1 + k, self.lr) * m / (np.sqrt(v) + EPS)
        rem = nz - n
        if rem > 0:
            m *= BETA1 ** rem
            v *= BETA2 ** rem
        g[:] = 0.0
        self.last[r] = target

    def train_step(self, rows, grads):
        """One optimiser step whose batch touches ``rows`` (duplicates allowed: their gradients are summed, as TF's
        sparse Adam does before squaring) with gradients ``grads`` [len(rows), cols]."""
        done = self.step
        for r in np.unique(rows):
            self.catch_up(r, done)          # (k_prep: claims + catch-up of the touched rows)
        np.add.at(self.g, rows, grads)      # pending gradient of step done + 1
        self.step = done + 1

    def flush(self):
        for r in range(self.w.shape[0]):
            self.catch_up(r, self.step)
        return self.w


def dense_reference(w, lr, touches):
    """The literal form on the same touches: [(rows, grads)] per step."""
    w = np.array(w, dtype=np.float64)
    m, v = np.zeros_like(w), np.zeros_like(w)
    for t, (rows, grads) in enumerate(touches, start=1):
        g = np.zeros_like(w)
        np.add.at(g, rows, grads)
        m = BETA1 * m + (1 - BETA1) * g
        v = BETA2 * v + (1 - BETA2) * g * g
        w = w - adam_alpha(t, lr) * m / (np.sqrt(v) + EPS)
    return w, m, v
