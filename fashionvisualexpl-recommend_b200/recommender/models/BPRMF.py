"""BPR-MF with the reference's model/trainer protocol (src/recommender/models/BPRMF.py)
over the libfvx engine.

Same surface as the reference class: ``BPRMF(data, params)``; attributes ``Bi, Gu, Gi,
embed_k, learning_rate, reg, evaluator, directory_parameters, optimizer``; methods
``call((user, item))`` (:55-76), ``predict_all()`` (:78-85), ``train_step(batch)``
(:87-125) and ``train()`` (:127-192).  Every numeric operation runs in libfvx.so.

Extra knobs read from ``params`` when present (all optional, defaults keep the
reference behaviour): ``adam_mode`` (auto | deferred | dense | lazy; auto = dense for small batches on small tables), ``sampler``
(host_ref | device), ``device``, ``sync_loss``, ``tensor_cores`` (default on), ``graph_steps``
(default on: the steps of an epoch are replayed as CUDA graphs of 8, ``Engine.steps``).
"""
from __future__ import annotations

import os
from time import time

import numpy as np
import torch

from ...config import configs
from ...engine import Engine
from ...utils.write import save_obj
from ..Evaluator import Evaluator
from ..RecommenderModel import DeviceArray, RecommenderModel


class _Optimizer:
    """Stands where ``tf.optimizers.Adam`` is in the reference (BPRMF.py:52)."""

    def __init__(self, engine):
        self._e = engine
        self.learning_rate = engine.lr
        self.beta_1, self.beta_2, self.epsilon = 0.9, 0.999, 1e-7

    @property
    def iterations(self):
        return self._e.steps_done()


class BPRMF(RecommenderModel):
    visual = False

    def __init__(self, data, params):
        super().__init__(data, params)
        self.embed_k = self.params.embed_k
        self.learning_rate = self.params.lr
        self.reg = self.params.reg
        self.device = getattr(params, "device", "cuda:0")
        self.directory_parameters = f'batch_{self.params.batch_size}' \
                                    f'-K_{self.params.embed_k}' \
                                    f'-lr_{self.params.lr}' \
                                    f'-reg_{self.params.reg}'
        self._build_engine()
        self.evaluator = Evaluator(self, data, params.top_k)
        self.optimizer = _Optimizer(self.engine)

    def _build_engine(self, d=0, D=0, features=None):
        p = self.params
        self.engine = Engine(self.num_users, self.num_items, self.embed_k, d=d, D=D, lr=p.lr, reg=p.reg,
                             adam_mode=getattr(p, "adam_mode", "auto"), max_batch=p.batch_size,
                             device=self.device, seed=getattr(p, "seed", 0),
                             # tcgen05 projections / evaluation sweep wherever the geometry is eligible (the
                             # engine falls back to the exact fp32 CUDA-core kernels otherwise)
                             use_tensor_cores=bool(getattr(p, "tensor_cores", True)))
        if D:
            self.engine.set_features(features)

    # the reference's variables, as live views into the packed device tables.  In the default DEFERRED
    # Adam mode rows are brought up to date lazily, so every read flushes first:
    # what the caller sees is what a tf.Variable of the reference would hold after the same steps
    def _current(self):
        self.engine.flush()
        return self.engine

    @property
    def Bi(self): return self._current().Bi
    @property
    def Gu(self): return self._current().Gu
    @property
    def Gi(self): return self._current().Gi

    def _idx(self, x):
        return self.engine._i32(x, self.engine.device)

    def call(self, inputs, training=None, mask=None):
        """x_ui and the gathered rows for (user, item) index batches (BPRMF.py:69-76)."""
        user, item = (self._idx(a) for a in inputs)
        xui = self.engine.score_pairs(user, item)
        u, i = user.long(), item.long()
        return DeviceArray(xui), DeviceArray(self.Bi[i]), DeviceArray(self.Gu[u]), DeviceArray(self.Gi[i])

    def predict_all(self):
        """Dense ``[U, I]`` scores (BPRMF.py:85).  Small sizes only: the evaluator does not
        use it - it goes through the fused top-k sweep."""
        return DeviceArray(self.engine.predict_all())

    def train_step(self, batch, sync=True, loss_slot=0):
        """One optimiser step on ``(user, pos, neg)``; returns the batch loss as a float
        like the reference (BPRMF.py:125) - ``sync=False`` leaves it on the device, in
        ``engine.loss_t[loss_slot]`` (``engine.take_loss``)."""
        user, pos, neg = (self._idx(a) for a in batch)
        self.engine.step(user, pos, neg, loss_slot=loss_slot)
        return self.engine.read_loss(loss_slot) if sync else None

    # ---- checkpoints: tensors + optimiser state (the reference only ever saves) ----------
    def state_dict(self):
        e = self.engine
        e.flush()
        sd = {"step": e.step_t.clone()}
        for name, t in (("users", e.users), ("items", e.items)):
            for k in ("w", "m", "v", "last"):
                sd["%s.%s" % (name, k)] = t[k].clone()
        if e.D:
            sd.update({"E": e.E.clone(), "mE": e.mE.clone(), "vE": e.vE.clone()})
        return sd

    def load_state_dict(self, sd):
        e = self.engine
        e.step_t.copy_(sd["step"])
        for name, t in (("users", e.users), ("items", e.items)):
            for k in ("w", "m", "v", "last"):
                t[k].copy_(sd["%s.%s" % (name, k)])
            t["mark"].zero_()          # touch stamps are relative to the (restored) step counter
            t["g"].zero_()             # a checkpoint is taken after a flush: no pending gradient goes with it
        if e.D:
            e.E.copy_(sd["E"]); e.mE.copy_(sd["mE"]); e.vE.copy_(sd["vE"])
        e._theta_step = -1

    def save_weights(self, path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({k: v.cpu() for k, v in self.state_dict().items()}, path + ".pt")

    def restore_weights(self, path):
        sd = torch.load(path + ".pt", map_location=self.engine.device)
        self.load_state_dict(sd)

    # ---- the epoch loop of BPRMF.train (:127-192) ----------------------------------------
    def train(self):
        p = self.params
        max_metrics = {'hr': 0, 'p': 0, 'r': 0, 'auc': 0, 'ndcg': 0}
        best_state = None
        best_epoch = self.restore_epochs
        best_epoch_print = 'No best epoch found!'
        results = {}
        steps = 0
        it = 1
        steps_per_epoch = self.data.num_train // p.batch_size          # :137
        sync_loss = bool(getattr(p, "sync_loss", False))
        rdir = f'{configs.results_dir()}/{p.dataset}/{p.rec}'
        wdir = f'{configs.weight_dir()}/{p.dataset}/{p.rec}'
        os.makedirs(rdir, exist_ok=True)
        os.makedirs(wdir, exist_ok=True)
        start_ep = time()
        print('Start training...')
        graph = bool(getattr(p, "graph_steps", True)) and not sync_loss
        for bufs, n_batches in self.data.next_batch_run(self.device):
            first = 0
            while first < n_batches:
                # the steps up to the end of the epoch in one go (CUDA-graph replay), or one at a time when the
                # reference's per-step loss synchronisation is asked for (:125)
                n = min(n_batches - first, steps_per_epoch - steps) if graph else 1
                if graph:
                    self.engine.steps(bufs[0], bufs[1], bufs[2], first, n, p.batch_size, loss_slot=0)
                else:
                    B = p.batch_size
                    self.engine.step(*(x[first * B:(first + 1) * B] for x in bufs), loss_slot=0)
                    if sync_loss:
                        self.engine.loss_t[0].item()
                first += n
                steps += n
                if steps == steps_per_epoch:                  # epoch is over (:148)
                    loss = self.engine.read_loss(0)
                    epoch_text = 'Epoch {0}/{1} \tLoss: {2:.3f}'.format(it, p.epochs, loss / steps)
                    epoch_print = self.evaluator.eval(it, results, epoch_text, start_ep)
                    for metric in max_metrics.keys():
                        if max_metrics[metric] <= results[it][metric + '_v']:
                            max_metrics[metric] = results[it][metric + '_v']
                            if metric == p.best_metric:
                                best_epoch, best_state, best_epoch_print = it, self.state_dict(), epoch_print
                    if (it % self.verbose == 0 or it == 1) and self.verbose != -1:
                        self.save_weights(f'{wdir}/weights-{it}-{self.directory_parameters}')
                    start_ep = time()
                    it += 1
                    steps = 0
        print('Training end...')
        self.evaluator.store_recommendation(path=f'{rdir}/recs-{it - 1}-{self.directory_parameters}.tsv')
        save_obj(results, f'{rdir}/results-metrics-{self.directory_parameters}')
        print("Store Best Model at Epoch {0}".format(best_epoch))
        print(best_epoch_print)
        final_state = self.state_dict()
        if best_state is not None:
            self.load_state_dict(best_state)
        self.save_weights(f'{wdir}/best-weights-{best_epoch}-{self.directory_parameters}')
        self.evaluator.store_recommendation(
            path=f'{rdir}/best-recs-{best_epoch}-{self.directory_parameters}.tsv')
        self.load_state_dict(final_state)
        print('End Store Best Model!')
        print('Best Values for Each Metric:\nHR\tPrec\tRec\tAUC\tnDCG\n{}\t{}\t{}\t{}\t{}\n'.format(
            max_metrics['hr'], max_metrics['p'], max_metrics['r'], max_metrics['auc'], max_metrics['ndcg']))
        return results
