"""GradFashion with the reference's protocol (src/recommender/models/GradFashion.py) over libfvx.

VBPR whose visual feature is a learned two-stage projection of two handcrafted descriptors: the colour
histogram ``Fc [I, Dc]`` and the edge features ``Fe [I, De]`` (both max-abs normalised, frozen):

    v_i = [Fc[i] Ec | Fe[i] Ee]                               (:111-114)
    x_ui = Bi[i] + <Gu[u], Gi[i]> + <Tu[u], v_i E> + v_i Bp   (:122-125)

Trained tensors besides BPRMF's: ``Ec [Dc, embed_color]``, ``Ee [De, embed_edges]``, ``Tu [U, d]``,
``E [embed_color + embed_edges, d]``, ``Bp [embed_color + embed_edges, 1]`` (:60-80); the L2 term counts the negative
item's bias in full (:171-172) and includes Ec, Ee, E, Bp (:173-176).

On the device this is the VBPR step on the concatenated features ``F = [Fc | Fe]`` with the EFFECTIVE projection
matrix ``blockdiag(Ec, Ee) * [E | Bp]``, composed ahead of every step; the gradient of the effective matrix is carried
back to Ec, Ee, E, Bp before their Adam step (``Engine(two_stage=...)``, csrc/fvx_train.cu: k_gf_*).  ``call`` returns
the reference's tuple order ``(xui, gamma_u, gamma_i, color_i, edges_i, theta_u, theta_i, beta_i)`` (:126-133).
"""
from __future__ import annotations

import numpy as np
import torch

from ...dataset.visual_loader_mixin import VisualLoader
from ...engine import Engine
from ..Evaluator import Evaluator
from ..RecommenderModel import DeviceArray, RecommenderModel
from .BPRMF import BPRMF, _Optimizer


class GradFashion(BPRMF, VisualLoader):
    visual = True

    def __init__(self, data, params):
        RecommenderModel.__init__(self, data, params)
        p = self.params
        self.embed_k, self.embed_d = p.embed_k, p.embed_d
        self.embed_color, self.embed_edges = p.embed_color, p.embed_edges
        self.learning_rate, self.reg = p.lr, p.reg
        self.device = getattr(params, "device", "cuda:0")
        self.directory_parameters = f'batch_{p.batch_size}-D_{p.embed_d}-K_{p.embed_k}-lr_{p.lr}-reg_{p.reg}'
        self.process_edge_visual_features()
        self.process_color_visual_features()
        Dc, De = self.dim_color_features, self.dim_edge_features
        self.engine = Engine(self.num_users, self.num_items, self.embed_k, d=self.embed_d, D=Dc + De, lr=p.lr, reg=p.reg,
                             adam_mode=getattr(p, "adam_mode", "auto"), max_batch=p.batch_size, device=self.device,
                             seed=getattr(p, "seed", 0), use_tensor_cores=bool(getattr(p, "tensor_cores", True)),
                             two_stage=(Dc, De, self.embed_color, self.embed_edges))
        self.engine.set_features(np.concatenate([self.color_features, self.edge_features], axis=1))
        self.evaluator = Evaluator(self, data, params.top_k)
        self.optimizer = _Optimizer(self.engine)

    # the reference keeps these in dicts (color_weights, edges_weights, visual_profile): same keys, live device views
    @property
    def color_weights(self):
        e = self.engine
        return {"Fc": e.F[:, :self.dim_color_features] if e.F is not None else None, "Ec": e.gf["Ec"]}

    @property
    def edges_weights(self):
        e = self.engine
        return {"Fe": e.F[:, self.dim_color_features:] if e.F is not None else None, "Ee": e.gf["Ee"]}

    @property
    def visual_profile(self):
        e = self._current()
        return {"Bp": e.gf["E2"][:, e.d:e.d + 1], "E": e.gf["E2"][:, :e.d], "Tu": e.Tu}

    def call(self, inputs, training=True, mask=None):
        user, item = (self._idx(a) for a in inputs)
        xui = self.engine.score_pairs(user, item)
        u, i = user.long(), item.long()
        th = self.engine.theta()[i][:, :self.embed_d]
        return (DeviceArray(xui), DeviceArray(self.Gu[u]), DeviceArray(self.Gi[i]), DeviceArray(self.color_weights["Fc"][i]),
                DeviceArray(self.edges_weights["Fe"][i]), DeviceArray(self.visual_profile["Tu"][u]), DeviceArray(th),
                DeviceArray(self.Bi[i]))

    def state_dict(self):
        sd = super().state_dict()
        sd.update({"gf." + k: v.clone() for k, v in self.engine.gf.items()})
        return sd

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        for k, v in self.engine.gf.items():
            v.copy_(sd["gf." + k])
