"""CPU: deferred Adam (oracle/deferred_adam.py restates replay_row / k_prep / fvx_adam_flush) against the literal
whole-table Keras Adam on the same touch pattern - the equivalence the CUDA step's default optimiser mode rests on."""
import numpy as np
import pytest

from oracle import deferred_adam as da


def _touches(rng, R, C, steps, p_touch, hot=()):
    out = []
    for _ in range(steps):
        n = max(1, rng.binomial(R, p_touch))
        rows = rng.integers(0, R, n)
        rows = np.concatenate([rows, np.asarray(hot, dtype=np.int64), rows[:n // 3]])    # duplicates inside a batch
        out.append((rows, rng.standard_normal((len(rows), C)) * 10.0 ** rng.uniform(-4, 0)))
    return out


@pytest.mark.parametrize("steps,p_touch", [(40, 0.3), (300, 0.02), (700, 0.004)])
def test_deferred_equals_dense(steps, p_touch):
    rng = np.random.default_rng(steps)
    R, C, lr = 60, 5, 1e-3
    w0 = rng.standard_normal((R, C)) * 0.1
    touches = _touches(rng, R, C, steps, p_touch, hot=(7,))      # row 7 is touched every step, some rows never
    T = da.DeferredTable(w0, lr)
    for rows, grads in touches:
        T.train_step(rows, grads)
    w = T.flush().copy()
    want, m, v = da.dense_reference(w0, lr, touches)
    # steps beyond REPLAY_MAX skipped zero-gradient steps move w by < 0.9^192 ~ 2e-9 of one update
    assert np.max(np.abs(w - want)) <= 1e-9 * max(1.0, np.max(np.abs(want)))
    assert np.allclose(T.m, m, rtol=1e-9, atol=1e-300) and np.allclose(T.v, v, rtol=1e-9, atol=1e-300)
    assert np.all(T.g == 0) and np.all(T.last == steps)


def test_flush_is_idempotent_and_mid_run_reads_are_current():
    rng = np.random.default_rng(3)
    R, C, lr = 20, 3, 1e-2
    w0 = rng.standard_normal((R, C))
    touches = _touches(rng, R, C, 25, 0.2)
    T = da.DeferredTable(w0, lr)
    for s, (rows, grads) in enumerate(touches, start=1):
        T.train_step(rows, grads)
        if s in (5, 17):                                          # evaluation in the middle of training
            w = T.flush().copy()
            assert np.array_equal(w, T.flush())                   # a second flush changes nothing
            want, _, _ = da.dense_reference(w0, lr, touches[:s])
            assert np.allclose(w, want, rtol=1e-12, atol=1e-14)
    want, _, _ = da.dense_reference(w0, lr, touches)
    assert np.allclose(T.flush(), want, rtol=1e-12, atol=1e-14)
