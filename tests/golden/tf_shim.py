"""A minimal torch-backed stand-in for the ``tensorflow`` API surface that the
reference's BPRMF.py / VBPR.py / RecommenderModel.py / dataset.py touch.

Purpose: ``tensorflow==2.3.1`` (requirements.txt:42) cannot be installed in the
build container, so ``tests/golden/make_golden.py`` imports the reference's model
files UNMODIFIED over this shim to generate golden vectors.  The forward pass, the
loss and the regulariser are then literally the reference's source; gradients come
from torch autograd; the optimiser below is a restatement of TF 2.3's Keras Adam
(dense semantics for IndexedSlices after duplicate-summing; eps outside the
bias-corrected sqrt) and is the one piece that remains unpinned.

This module is only used by the generator script and by the CPU test that
re-validates the fixtures when /root/reference is present; it is never imported
by the product or on the GPU box.
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np
import torch

_gen = torch.Generator().manual_seed(0)


def _t(x):
    return x.t if isinstance(x, TFTensor) else x


class TFTensor:
    """Wrapper so that ``.numpy()`` works on tensors that require grad."""

    __array_priority__ = 1000

    def __init__(self, t):
        self.t = t

    def numpy(self):
        return self.t.detach().cpu().numpy()

    @property
    def shape(self):
        return tuple(self.t.shape)

    def __add__(self, o): return TFTensor(self.t + _t(o))
    def __radd__(self, o): return TFTensor(_t(o) + self.t)
    def __sub__(self, o): return TFTensor(self.t - _t(o))
    def __rsub__(self, o): return TFTensor(_t(o) - self.t)
    def __mul__(self, o): return TFTensor(self.t * _t(o))
    def __rmul__(self, o): return TFTensor(_t(o) * self.t)
    def __truediv__(self, o): return TFTensor(self.t / _t(o))
    def __neg__(self): return TFTensor(-self.t)
    def __iadd__(self, o): return TFTensor(self.t + _t(o))
    def __float__(self): return float(self.t.detach())

    def __deepcopy__(self, memo):
        c = TFTensor(self.t.detach().clone().requires_grad_(self.t.requires_grad))
        memo[id(self)] = c
        return c


def Variable(initial_value, name=None, dtype=None, trainable=True):
    v = _t(initial_value)
    if isinstance(v, np.ndarray):
        v = torch.from_numpy(np.ascontiguousarray(v))
    v = v.detach().clone().to(torch.float32)
    v.requires_grad_(bool(trainable))
    out = TFTensor(v)
    out.name = name
    return out


def zeros(n):
    return TFTensor(torch.zeros(n, dtype=torch.float32))


def squeeze(x):
    return TFTensor(_t(x).squeeze())


def reduce_sum(x, axis=None):
    if isinstance(x, (list, tuple)):
        x = torch.stack([_t(e) for e in x])
    else:
        x = _t(x)
    return TFTensor(x.sum() if axis is None else x.sum(dim=axis))


def matmul(a, b, transpose_b=False):
    b = _t(b)
    return TFTensor(_t(a) @ (b.t() if transpose_b else b))


def concat(values, axis=0):
    return TFTensor(torch.cat([_t(v) for v in values], dim=axis))


def clip_by_value(x, lo, hi):
    return TFTensor(torch.clamp(_t(x), lo, hi))


def _embedding_lookup(params, ids):
    ids = _t(ids)
    if not isinstance(ids, torch.Tensor):
        ids = torch.as_tensor(np.asarray(ids))
    return TFTensor(_t(params)[ids.long()])


def _softplus(x):
    x = _t(x)
    thr = math.log(np.finfo(np.float32).eps) + 2.0
    mid = torch.log1p(torch.exp(torch.clamp(x, max=-thr)))
    out = torch.where(x > -thr, x, torch.where(x < thr, torch.exp(torch.clamp(x, max=0.0)), mid))
    return TFTensor(out)


def _l2_loss(x):
    x = _t(x)
    return TFTensor((x * x).sum() / 2)


class GradientTape:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, loss, params):
        gs = torch.autograd.grad(_t(loss), [_t(p) for p in params], allow_unused=True)
        return [None if g is None else TFTensor(g) for g in gs]


class Adam:
    """Keras Adam (TF 2.3): alpha_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= alpha_t*m/(sqrt(v)+eps)."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self.slots = {}

    def apply_gradients(self, grads_and_vars):
        self.iterations += 1
        t = self.iterations
        alpha = self.lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        with torch.no_grad():
            for g, var in grads_and_vars:
                if g is None:
                    continue
                w, g = var.t, _t(g)
                if id(var) not in self.slots:
                    self.slots[id(var)] = (torch.zeros_like(w), torch.zeros_like(w))
                m, v = self.slots[id(var)]
                m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                w.sub_(np.float32(alpha) * m / (v.sqrt() + np.float32(self.eps)))

    def __deepcopy__(self, memo):
        return self


class _Model:
    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, inputs, training=None, mask=None):
        return self.call(inputs, training=training, mask=mask)


class _Checkpoint:
    saved = []

    def __init__(self, **kw):
        self.kw = kw

    def save(self, path):
        _Checkpoint.saved.append(path)
        return path


class _GlorotUniform:
    def __call__(self, shape):
        r, c = shape
        lim = math.sqrt(6.0 / (r + c))
        return TFTensor((torch.rand(r, c, generator=_gen) * 2 - 1) * lim)


class _Dataset:
    def __init__(self, cols):
        self.cols = [np.asarray([int(x) for x in c], dtype=np.int64) for c in cols]
        self.bs = None

    @staticmethod
    def from_tensor_slices(cols):
        return _Dataset(cols)

    def batch(self, batch_size):
        self.bs = batch_size
        return self

    def prefetch(self, buffer_size=None):
        return self

    def map(self, *a, **k):
        raise NotImplementedError("image pipelines are out of scope")

    def __iter__(self):
        n = len(self.cols[0])
        for s in range(0, n, self.bs):
            yield tuple(TFTensor(torch.from_numpy(c[s:s + self.bs])) for c in self.cols)


def _set_seed(s):
    _gen.manual_seed(int(s))


def install():
    """Registers the shim as ``tensorflow`` in sys.modules."""
    tf = types.ModuleType("tensorflow")
    tf.float32 = torch.float32
    tf.Variable = Variable
    tf.zeros = zeros
    tf.squeeze = squeeze
    tf.reduce_sum = reduce_sum
    tf.matmul = matmul
    tf.clip_by_value = clip_by_value
    tf.concat = concat
    tf.GradientTape = GradientTape
    tf.nn = types.SimpleNamespace(embedding_lookup=_embedding_lookup, softplus=_softplus,
                                  l2_loss=_l2_loss)
    tf.optimizers = types.SimpleNamespace(Adam=Adam)
    tf.keras = types.SimpleNamespace(Model=_Model, optimizers=types.SimpleNamespace(Adam=Adam))
    tf.initializers = types.SimpleNamespace(GlorotUniform=_GlorotUniform)
    tf.random = types.SimpleNamespace(set_seed=_set_seed)
    tf.train = types.SimpleNamespace(Checkpoint=_Checkpoint)
    tf.data = types.SimpleNamespace(Dataset=_Dataset,
                                    experimental=types.SimpleNamespace(AUTOTUNE=-1))
    sys.modules["tensorflow"] = tf
    return tf
