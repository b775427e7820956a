"""VBPR with the reference's protocol (src/recommender/models/VBPR.py) over libfvx.

Adds to BPRMF: ``embed_d``, ``Tu [U,d]``, ``E [D,d]``, ``Bp [D,1]`` and the frozen,
max-abs-normalised CNN features ``F [I,D]`` (VBPR.py:41-54).  ``call`` returns the
reference's tuple order ``(xui, gamma_u, gamma_i, feature_i, theta_u, beta_i)`` (:86).
"""
from __future__ import annotations

from ...dataset.visual_loader_mixin import VisualLoader
from ..Evaluator import Evaluator
from ..RecommenderModel import DeviceArray, RecommenderModel
from .BPRMF import BPRMF, _Optimizer


class VBPR(BPRMF, VisualLoader):
    visual = True

    def __init__(self, data, params):
        RecommenderModel.__init__(self, data, params)
        self.embed_k = self.params.embed_k
        self.embed_d = self.params.embed_d
        self.learning_rate = self.params.lr
        self.reg = self.params.reg
        self.device = getattr(params, "device", "cuda:0")
        self.directory_parameters = f'batch_{self.params.batch_size}' \
                                    f'-D_{self.params.embed_d}' \
                                    f'-K_{self.params.embed_k}' \
                                    f'-lr_{self.params.lr}' \
                                    f'-reg_{self.params.reg}'
        self.process_cnn_visual_features()
        self._build_engine(d=self.embed_d, D=self.dim_cnn_features, features=self.cnn_features)
        self.evaluator = Evaluator(self, data, params.top_k)
        self.optimizer = _Optimizer(self.engine)

    @property
    def Tu(self): return self._current().Tu
    @property
    def E(self): return self.engine.Ew
    @property
    def Bp(self): return self.engine.Bp
    @property
    def F(self): return self.engine.F

    def call(self, inputs, training=None, mask=None):
        user, item = (self._idx(a) for a in inputs)
        xui = self.engine.score_pairs(user, item)
        u, i = user.long(), item.long()
        return (DeviceArray(xui), DeviceArray(self.Gu[u]), DeviceArray(self.Gi[i]), DeviceArray(self.F[i]),
                DeviceArray(self.Tu[u]), DeviceArray(self.Bi[i]))
