"""Device-resident model state and the thin host wrappers over the C ABI.

``Engine`` owns the torch tensors (torch is used for device memory and streams only),
fills the ``FvxModel`` struct of include/fvx.h with their raw pointers and exposes one
method per entry point.  The reference-facing classes (recommender/models/BPRMF.py,
VBPR.py, recommender/Evaluator.py, dataset/dataset.py) are built on it.

HBM layout (fp32, row-major):
  users table  UT[U, Su]   Su = round_up4(K+d):  cols [0,K) = Gu, [K,K+d) = Tu
  items table  IT[Ic, Si]  Si = round_up4(K+1):  cols [0,K) = Gi, col K = Bi   (Ic = owned rows)
  E_ext[D, de]             de = round_up4(d+1):  cols [0,d) = E,  col d = Bp
  each table has m, v (Adam) and g (gradient accumulator) twins of the same shape.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import FvxModel, FvxTable, call, ptr, stream_ptr


def _r4(x):
    return (x + 3) // 4 * 4


class Engine:
    def __init__(self, num_users, num_items, K, d=0, D=0, lr=1e-3, reg=0.0, adam_mode="deferred",
                 max_batch=4096, device="cuda:0", item_lo=0, item_cnt=None, loss_slots=4096,
                 ge_parts=80, seed=0, use_tensor_cores=False, unique_rows=True, user_lo=0, user_cnt=None,
                 user_rows=None, sharded=False, two_stage=None):
        """``item_lo / item_cnt``: the catalog rows this rank owns; ``user_lo / user_cnt``: the users it owns (their
        Adam state lives here; the user tables are still allocated and indexed globally - ``user_rows`` rows, at
        least ``num_users``, so that equal-sized blocks can be gathered in place).

        ``two_stage=(Dc, De, ec, ee)``: GradFashion (GradFashion.py:97-134) - the features are ``[Fc | Fe]`` (D = Dc +
        De) and the trained visual tensors are Ec [Dc, ec], Ee [De, ee], E [ec + ee, d], Bp [ec + ee, 1]; the
        negative item's bias enters the L2 term with weight 1 instead of VBPR's 1/10."""
        if not torch.cuda.is_available():
            raise _lib.FvxError("no CUDA device: the fvx engine has no CPU path")
        _lib.load()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.U, self.I, self.K, self.d, self.D = int(num_users), int(num_items), int(K), int(d), int(D)
        if self.D == 0:
            self.d = 0
        self.item_lo = int(item_lo)
        self.Ic = int(item_cnt) if item_cnt is not None else self.I - self.item_lo
        self.user_lo = int(user_lo)
        self.user_cnt = int(user_cnt) if user_cnt is not None else self.U - self.user_lo
        self.U_rows = max(int(user_rows) if user_rows is not None else self.U, self.U)
        self.Su, self.Si, self.de = _r4(self.K + self.d), _r4(self.K + 1), _r4(self.d + 1) if self.D else 0
        self.lr, self.reg = float(lr), float(reg)
        self.max_batch = int(max_batch)
        if adam_mode == "auto":
            # Both are the reference's dense Keras-Adam semantics.  DEFERRED replays, per touched row, the steps the
            # row skipped (up to 192 dependent iterations): right when a batch touches a large part of the tables.
            # With small batches on tables that fit a quick sweep the literal whole-table update is cheaper
            # (measured at 40 k x 100 k, K=64, d=20: B=256 87 vs 286 us per step, B=4096 107 vs 161 us; at
            # B=65536 the deferred step wins, 316 us, and at 1 M x 500 k the sweep alone would be 3.3 GB).
            sweep_bytes = 28.0 * (self.U * self.Su + (int(item_cnt) if item_cnt is not None else self.I) * self.Si)
            adam_mode = "dense" if (self.max_batch <= 8192 and sweep_bytes <= 1e9) else "deferred"
        self.adam_mode = _lib.ADAM_MODES[adam_mode] if isinstance(adam_mode, str) else int(adam_mode)
        self.loss_slots = int(loss_slots)
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)

        def table(rows, stride, cap):
            t = {k: torch.zeros(rows, stride, **f32) for k in ("w", "m", "v", "g")}
            t["last"] = torch.zeros(rows, **i32)
            t["mark"] = torch.zeros(rows, **i32)
            t["list"] = torch.zeros(cap, **i32)
            t["count"] = torch.zeros(1, **i32)
            return t

        self.users = table(self.U_rows, self.Su, self.max_batch)
        self.items = table(self.Ic, self.Si, 2 * self.max_batch)
        self.step_t = torch.zeros(1, dtype=torch.int64, device=dev)
        self.loss_t = torch.zeros(self.loss_slots, dtype=torch.float64, device=dev)
        self.rows_t = torch.zeros(2 * self.max_batch, **i32)
        self.sync_t = torch.zeros(4, **i32)
        self.stage_t = torch.zeros(3 * self.max_batch + 4, **i32)     # fvx_bpr_steps: batch in flight + cursor
        self.cmap_t = torch.zeros(6 * self.max_batch, **i32) if (sharded or self.item_lo or self.Ic != self.I) else None
        self.F = self.F_pl = None
        self.ET_hi = self.ET_lo = self.W_hi = self.W_lo = None
        self.upos_t = self.W_sum = self.uslot_t = None
        self.use_tensor_cores = bool(use_tensor_cores) and self.D > 0
        if self.use_tensor_cores and (self.D % 128 != 0 or (self.d + 1 + 63) // 64 * 64 > 320):
            # the tcgen05 projection kernels take D % 128 == 0 and d + 1 <= 320 padded columns (more than 256
            # columns - BASELINE configs[4] with embed_d = 256 - are cut into column slices, one launch each);
            # wider / ragged models run on the exact fp32 CUDA-core kernels - still this library, still the GPU
            self.use_tensor_cores = False
        if self.D:
            self.E = torch.zeros(self.D, self.de, **f32)
            self.mE, self.vE = torch.zeros_like(self.E), torch.zeros_like(self.E)
            self.ge_parts = int(ge_parts)
            self.NP = self.de
            if self.use_tensor_cores:
                # padded operand width of the tcgen05 kernels (fvx_tc_width) and up to 8 K-split partials
                self.NP = self.de if self.de <= 0 else (32 if self.de <= 32 else (self.de + 63) // 64 * 64)
                u16 = dict(dtype=torch.uint16, device=dev)
                self.ET_hi = torch.zeros(self.NP, self.D, **u16)
                self.ET_lo = torch.zeros(self.NP, self.D, **u16)
                # backward coefficients: the two bf16 planes interleaved row by row, [hi NP | lo NP]
                # (W_lo = W_hi + NP elements): the backward reads a row pair as one N = 2*NP operand
                self.W_il = torch.zeros(2 * self.max_batch, 2 * self.NP, **u16)
                self.W_hi = self.W_il
                self.W_lo = self.W_il.view(-1)[self.NP:]
                self.ge_parts = max(self.ge_parts, 148)
                th_rows = max(4 * 2 * self.max_batch, 1 << 17)
                self.TH = torch.zeros(th_rows, self.NP, **f32)
                self.W = None
                # unique-row step: each distinct (owned) catalog row of a batch is projected once;
                # upos = position of a row in the touched-row list, W_sum = per-row sums of the
                # backward coefficients (zero between steps).  unique_rows=False: one projection per slot.
                if unique_rows:
                    self.upos_t = torch.zeros(self.Ic, **i32)
                    self.W_sum = torch.zeros(2 * self.max_batch, self.NP, **f32)
                    self.uslot_t = torch.zeros(2 * self.max_batch, **i32)
            else:
                self.TH = torch.zeros(2 * self.max_batch, self.de, **f32)
                self.W = torch.zeros(2 * self.max_batch, self.de, **f32)
            self.gE_part = torch.zeros(self.ge_parts, self.D, self.NP, **f32)
        else:
            self.E = self.mE = self.vE = self.gE_part = self.TH = self.W = None
            self.ge_parts = 0
        self.two_stage = tuple(int(x) for x in two_stage) if two_stage else None
        self.bias_neg_scale = 1.0 if self.two_stage else 0.1
        if self.two_stage:
            Dc, De, ec, ee = self.two_stage
            if Dc + De != self.D or min(Dc, De, ec, ee) <= 0:
                raise ValueError("two_stage=(Dc, De, ec, ee) must satisfy Dc + De == D")
            self.gf = {}
            for name, shape in (("Ec", (Dc, ec)), ("Ee", (De, ee)), ("E2", (ec + ee, self.de))):
                for pre in ("", "m", "v"):
                    self.gf[pre + name] = torch.zeros(*shape, **f32)
            self.gf_scratch = torch.zeros(self.D * self.de + Dc * ec + De * ee + (ec + ee) * self.de, **f32)
        self._theta = None
        self._theta_step = -1
        self.init_glorot(seed)
        self._struct = None

    # ---- parameters ---------------------------------------------------------------------
    def init_glorot(self, seed=0):
        """GlorotUniform with the reference's 2-D shapes (BPRMF.py:35,48-50; VBPR.py:44-54);
        Bi = 0.  TF's RNG stream is not reproducible, so values differ from a TF run.

        One generator PER TABLE, keyed by (seed, table): every rank of an item-sharded job draws the
        same replicated tables (Gu, Tu, E, Bp) whatever its shard size, and the item table is drawn as
        the whole [I, K] matrix in fixed row chunks of which a rank keeps its own rows - the sharded
        model starts as the single-GPU model does (tests/test_host_logic.py checks the slices)."""
        dev = self.device

        def gen(table):
            return torch.Generator(device=dev).manual_seed(int(seed) * 1000003 + table)

        def glorot(g, fan_r, fan_c, rows, cols):
            lim = math.sqrt(6.0 / (fan_r + fan_c))
            return (torch.rand(rows, cols, generator=g, device=dev) * 2 - 1) * lim

        uw, iw = self.users["w"], self.items["w"]
        uw.zero_()
        iw.zero_()
        uw[:self.U, :self.K] = glorot(gen(1), self.U, self.K, self.U, self.K)
        g, chunk = gen(2), 65536
        lo, hi = self.item_lo, self.item_lo + self.Ic
        for s0 in range(0, self.I, chunk):                  # the same stream on every rank: all chunks are drawn
            s1 = min(self.I, s0 + chunk)
            blk = glorot(g, self.I, self.K, s1 - s0, self.K)
            a, b = max(s0, lo), min(s1, hi)
            if a < b:
                iw[a - lo:b - lo, :self.K] = blk[a - s0:b - s0]
        if self.D:
            uw[:self.U, self.K:self.K + self.d] = glorot(gen(3), self.U, self.d, self.U, self.d)
            self.E.zero_()
            self.E[:, :self.d] = glorot(gen(4), self.D, self.d, self.D, self.d)
            self.E[:, self.d:self.d + 1] = glorot(gen(5), self.D, 1, self.D, 1)
            if self.two_stage:                              # GradFashion.py:60-80: Ec, Ee, E [ec+ee, d], Bp [ec+ee, 1]
                Dc, De, ec, ee = self.two_stage
                self.gf["Ec"].copy_(glorot(gen(6), Dc, ec, Dc, ec))
                self.gf["Ee"].copy_(glorot(gen(7), De, ee, De, ee))
                self.gf["E2"].zero_()
                self.gf["E2"][:, :self.d] = glorot(gen(4), ec + ee, self.d, ec + ee, self.d)
                self.gf["E2"][:, self.d:self.d + 1] = glorot(gen(5), ec + ee, 1, ec + ee, 1)

    # reference attribute names as views into the packed tables
    @property
    def Gu(self): return self.users["w"][:self.U, :self.K]
    @property
    def Tu(self): return self.users["w"][:self.U, self.K:self.K + self.d]
    @property
    def Gi(self): return self.items["w"][:, :self.K]
    @property
    def Bi(self): return self.items["w"][:, self.K]
    @property
    def Ew(self): return self.E[:, :self.d]
    @property
    def Bp(self): return self.E[:, self.d:self.d + 1]

    def load_params(self, P):
        """P: dict with Gu, Gi, Bi (and Tu, E, Bp) as numpy arrays / tensors; Gi, Bi in
        GLOBAL item order (the owned rows are sliced out)."""
        def dv(x):
            return torch.as_tensor(np.asarray(x), dtype=torch.float32).to(self.device)
        lo, hi = self.item_lo, self.item_lo + self.Ic
        self.Gu.copy_(dv(P["Gu"]))
        self.Gi.copy_(dv(P["Gi"])[lo:hi])
        self.Bi.copy_(dv(P["Bi"])[lo:hi])
        if self.D:
            self.Tu.copy_(dv(P["Tu"]))
            if self.two_stage:
                self.gf["Ec"].copy_(dv(P["Ec"]))
                self.gf["Ee"].copy_(dv(P["Ee"]))
                self.gf["E2"][:, :self.d].copy_(dv(P["E"]))
                self.gf["E2"][:, self.d:self.d + 1].copy_(dv(P["Bp"]).reshape(-1, 1))
            else:
                self.Ew.copy_(dv(P["E"]))
                self.Bp.copy_(dv(P["Bp"]).reshape(self.D, 1))
        self._theta_step = -1

    def params(self):
        """Flushes deferred optimiser state and returns the parameters as numpy arrays."""
        self.flush()
        out = {"Gu": self.Gu, "Gi": self.Gi, "Bi": self.Bi}
        if self.D and self.two_stage:
            out.update({"Tu": self.Tu, "Ec": self.gf["Ec"], "Ee": self.gf["Ee"], "E": self.gf["E2"][:, :self.d],
                        "Bp": self.gf["E2"][:, self.d:self.d + 1]})
        elif self.D:
            out.update({"Tu": self.Tu, "E": self.Ew, "Bp": self.Bp})
        return {k: v.detach().cpu().numpy().copy() for k, v in out.items()}

    def set_features(self, F, keep_fp32=True):
        """F: [Ic, D] (owned rows) already normalised by the global max|F|."""
        F = torch.as_tensor(np.asarray(F) if not torch.is_tensor(F) else F)
        if tuple(F.shape) != (self.Ic, self.D):
            raise ValueError("features must be [%d, %d], got %s" % (self.Ic, self.D, tuple(F.shape)))
        self.F = F.to(self.device, dtype=torch.float32).contiguous()
        if self.use_tensor_cores:
            # two bf16 planes (hi, lo): the same 4 bytes per element as fp32, ~2^-17 relative
            self.F_pl = torch.empty(self.Ic, 2 * self.D, dtype=torch.uint16, device=self.device)
            call("fvx_split_planes", ptr(self.F), ptr(self.F_pl), self.Ic, self.D, stream_ptr())
            if not keep_fp32:
                torch.cuda.current_stream().synchronize()
                self.F = None
        self._struct = None
        self._theta_step = -1

    # ---- the struct handed to the C ABI ------------------------------------------------
    def _table_struct(self, t, rows, stride):
        return FvxTable(ptr(t["w"]), ptr(t["m"]), ptr(t["v"]), ptr(t["g"]), ptr(t["last"]), ptr(t["mark"]),
                        ptr(t["list"]), ptr(t["count"]), rows, stride, t["list"].numel())

    def struct(self):
        if self._struct is None:
            if self.D and self.F is None and self.F_pl is None:
                raise _lib.FvxError("VBPR engine: set_features() must be called before use")
            m = FvxModel()
            m.abi_version = _lib.ABI_VERSION
            m.num_users, m.num_items, m.item_lo, m.item_cnt = self.U, self.I, self.item_lo, self.Ic
            m.K, m.d, m.D, m.de, m.adam_mode = self.K, self.d, self.D, self.de, self.adam_mode
            m.lr, m.reg = self.lr, self.reg
            m.users = self._table_struct(self.users, self.U_rows, self.Su)
            m.user_lo, m.user_cnt = self.user_lo, self.user_cnt
            m.items = self._table_struct(self.items, self.Ic, self.Si)
            m.E, m.mE, m.vE, m.gE_part = ptr(self.E), ptr(self.mE), ptr(self.vE), ptr(self.gE_part)
            m.ge_parts = self.ge_parts
            m.F, m.F_pl = ptr(self.F), ptr(self.F_pl)
            m.ET_hi, m.ET_lo, m.W_hi, m.W_lo = ptr(self.ET_hi), ptr(self.ET_lo), ptr(self.W_hi), ptr(self.W_lo)
            m.step, m.loss, m.loss_slots = ptr(self.step_t), ptr(self.loss_t), self.loss_slots
            m.TH, m.W, m.rows = ptr(self.TH), ptr(self.W), ptr(self.rows_t)
            m.th_cap = self.TH.numel() if self.TH is not None else 0
            m.sync = ptr(self.sync_t)
            m.cmap = ptr(self.cmap_t)
            m.max_batch = self.max_batch
            m.use_tensor_cores = 1 if self.use_tensor_cores else 0
            m.upos, m.W_sum, m.uslot = ptr(self.upos_t), ptr(self.W_sum), ptr(self.uslot_t)
            m.batch_stage = ptr(self.stage_t)
            m.bias_neg_scale = self.bias_neg_scale
            if self.two_stage:
                m.two_stage = 1
                m.Dc, m.De, m.ec, m.ee = self.two_stage
                g = self.gf
                m.Ec, m.mEc, m.vEc = ptr(g["Ec"]), ptr(g["mEc"]), ptr(g["vEc"])
                m.Ee, m.mEe, m.vEe = ptr(g["Ee"]), ptr(g["mEe"]), ptr(g["vEe"])
                m.E2, m.mE2, m.vE2 = ptr(g["E2"]), ptr(g["mE2"]), ptr(g["vE2"])
                m.gf_scratch = ptr(self.gf_scratch)
            self._struct = m
        return self._struct

    def set_hyper(self, lr=None, reg=None):
        """Changes lr / reg for the steps that follow.  In DEFERRED mode the last Adam step of the touched rows
        (and their skipped zero-gradient steps) is still pending and would be replayed with the NEW lr, so the
        tables are flushed first: steps already taken keep the lr they were taken with."""
        if lr is not None and float(lr) != self.lr and self._struct is not None:
            self.flush()
        if lr is not None:
            self.lr = float(lr)
        if reg is not None:
            self.reg = float(reg)
        self._struct = None

    # ---- training ------------------------------------------------------------------------
    @staticmethod
    def _i32(x, dev):
        if torch.is_tensor(x):
            return x.to(device=dev, dtype=torch.int32).contiguous()
        return torch.as_tensor(np.ascontiguousarray(np.asarray(x), dtype=np.int32)).to(dev)

    def step(self, user, pos, neg, loss_slot=0):
        """One optimiser step (fvx_bpr_step); asynchronous.  Index tensors: int32 CUDA."""
        B = user.numel()
        call("fvx_bpr_step", C.byref(self.struct()), ptr(user), ptr(pos), ptr(neg), B, loss_slot, stream_ptr())

    def steps(self, user, pos, neg, first, n_steps, B, loss_slot=0):
        """``n_steps`` consecutive optimiser steps on the batches ``[(first + s) * B, (first + s + 1) * B)`` of
        epoch-long index tensors (fvx_bpr_steps: the launches of 8 steps replayed as one CUDA graph)."""
        if (first + n_steps) * B > user.numel():
            raise ValueError("steps(): batches run past the end of the index arrays")
        call("fvx_bpr_steps", C.byref(self.struct()), ptr(user), ptr(pos), ptr(neg), int(first), int(n_steps), int(B),
             loss_slot, stream_ptr())

    def step_timed(self, user, pos, neg, loss_slot=0):
        """Profiling step: returns {phase: ms} (synchronises)."""
        out = (C.c_float * _lib.N_PHASES)()
        call("fvx_bpr_step_timed", C.byref(self.struct()), ptr(user), ptr(pos), ptr(neg), user.numel(),
             loss_slot, out, stream_ptr())
        return dict(zip(_lib.PHASES, [float(x) for x in out]))

    def flush(self):
        call("fvx_adam_flush", C.byref(self.struct()), stream_ptr())

    def steps_done(self):
        return int(self.step_t.item())

    def read_loss(self, slot=0, clear=True):
        v = float(self.loss_t[slot].item())
        if clear:
            self.loss_t[slot] = 0
        return v

    def take_loss(self, slot=0):
        """The loss accumulated in ``slot`` as a 1-element device tensor; the accumulator is cleared
        (all stream-ordered, no synchronisation)."""
        out = self.loss_t[slot:slot + 1].clone()
        self.loss_t[slot:slot + 1].zero_()
        return out

    # ---- evaluation ----------------------------------------------------------------------
    def theta(self, refresh=False):
        """theta_ext = F * E_ext over the owned catalog rows, cached per optimiser step."""
        if not self.D:
            return None
        s = self.steps_done()
        if refresh or self._theta is None or self._theta_step != s:
            if self._theta is None:
                self._theta = torch.empty(self.Ic, self.de, dtype=torch.float32, device=self.device)
            call("fvx_project", C.byref(self.struct()), ptr(self._theta), stream_ptr())
            self._theta_step = s
        return self._theta

    def predict_all(self, u0=0, u1=None):
        u1 = self.U if u1 is None else u1
        self.flush()
        out = torch.empty(u1 - u0, self.Ic, dtype=torch.float32, device=self.device)
        call("fvx_predict_all", C.byref(self.struct()), ptr(self.theta()), u0, u1, ptr(out), stream_ptr())
        return out

    def project_rows(self, rows):
        """F[rows] * E_ext -> [n, de] (fvx_project_rows; VBPR.py:83-84)."""
        out = torch.empty(rows.numel(), self.de, dtype=torch.float32, device=self.device)
        call("fvx_project_rows", C.byref(self.struct()), ptr(rows), rows.numel(), ptr(out), stream_ptr())
        return out

    def grad_E_rows(self, rows, W):
        """sum_r F[rows[r]]^T W[r] -> [D, de] (fvx_grad_e_rows)."""
        out = torch.empty(self.D, self.de, dtype=torch.float32, device=self.device)
        call("fvx_grad_e_rows", C.byref(self.struct()), ptr(rows), rows.numel(), ptr(W.contiguous()), ptr(out),
             stream_ptr())
        return out

    def score_pairs(self, user, item):
        self.flush()
        out = torch.empty(user.numel(), dtype=torch.float32, device=self.device)
        call("fvx_score_pairs", C.byref(self.struct()), ptr(self.theta()), ptr(user), ptr(item), user.numel(),
             ptr(out), stream_ptr())
        return out

    def score_topk(self, mask_row_ptr, mask_col, k, u0=0, u1=None, thr_scores=None, tc=None, view=None):
        """Masked top-k (+ optional rank counts) for users [u0,u1): (ids, scores[, counts]).
        ``tc``: use the tcgen05 sweep (default: the engine's ``use_tensor_cores``); it returns
        the same ids / scores as the fp32 kernel and is only available without rank counts.
        ``view``: an item-side view other than this engine's own shard (parallel.gathered_view):
        dict(struct=FvxModel, theta=tensor, Ic=rows, keep=[tensors kept alive])."""
        u1 = self.U if u1 is None else u1
        self.flush()
        n = u1 - u0
        if view is not None:
            return self._score_topk_view(view, mask_row_ptr, mask_col, k, u0, u1, tc)
        ids = torch.empty(n, k, dtype=torch.int32, device=self.device)
        sc = torch.empty(n, k, dtype=torch.float32, device=self.device)
        tc = self.use_tensor_cores if tc is None else tc
        if tc and thr_scores is None and self.K + self.d + 3 <= 448 and n > 0:
            ws = self._eval_ws(n)
            call("fvx_score_topk_tc", C.byref(self.struct()), ptr(self.theta()), u0, u1, ptr(mask_row_ptr),
                 ptr(mask_col), k, ptr(ids), ptr(sc), C.byref(ws["struct"]), stream_ptr())
            self._last_flags = ws["flags"][:n]      # rows the exact kernel recomputed (inside the call)
            return ids, sc
        n_thr, counts = 0, None
        if thr_scores is not None:
            n_thr = thr_scores.shape[1]
            counts = torch.zeros(n, n_thr, dtype=torch.int32, device=self.device)
        call("fvx_score_topk", C.byref(self.struct()), ptr(self.theta()), u0, u1, ptr(mask_row_ptr),
             ptr(mask_col), k, ptr(ids), ptr(sc), n_thr, ptr(thr_scores), ptr(counts), stream_ptr())
        return (ids, sc, counts) if thr_scores is not None else (ids, sc)

    def topk_bounds(self, mask_row_ptr, k, u0=0, u1=None):
        """First half of the tensor-core sweep (fvx_score_topk_tc_bounds): returns the int32 tensor of per-user
        bounds (signed order = float order).  Item-sharded callers take the element-wise MAX over the ranks of it,
        in place, before ``topk_select``."""
        u1 = self.U if u1 is None else u1
        self.flush()
        n = u1 - u0
        ws = self._eval_ws(n)
        call("fvx_score_topk_tc_bounds", C.byref(self.struct()), ptr(self.theta()), u0, u1, ptr(mask_row_ptr), k,
             C.byref(ws["struct"]), stream_ptr())
        return ws["thr"][:n]

    def topk_select(self, mask_row_ptr, mask_col, k, u0=0, u1=None):
        """Second half (fvx_score_topk_tc_select): candidates sweep with the bounds of ``topk_bounds``, exact
        re-scoring, top-k: (ids, scores)."""
        u1 = self.U if u1 is None else u1
        n = u1 - u0
        ws = self._eval_ws(n)
        ids = torch.empty(n, k, dtype=torch.int32, device=self.device)
        sc = torch.empty(n, k, dtype=torch.float32, device=self.device)
        call("fvx_score_topk_tc_select", C.byref(self.struct()), ptr(self.theta()), u0, u1, ptr(mask_row_ptr),
             ptr(mask_col), k, ptr(ids), ptr(sc), C.byref(ws["struct"]), stream_ptr())
        self._last_flags = ws["flags"][:n]
        return ids, sc

    def tc_eval_eligible(self):
        return self.K + self.d + 3 <= 448

    def rank_counts(self, mask_row_ptr, mask_col, thr_scores, u0=0, u1=None):
        """counts[u, t] = number of owned, non-masked items scoring >= thr_scores[u, t] (NaN: unused) for
        users [u0,u1): the register-tiled sweep of fvx_rank_counts (at most 4 thresholds per launch)."""
        u1 = self.U if u1 is None else u1
        self.flush()
        n, T = u1 - u0, thr_scores.shape[1]
        counts = torch.zeros(n, T, dtype=torch.int32, device=self.device)
        theta = self.theta()
        for c0 in range(0, T, 4):
            thr = thr_scores[:, c0:c0 + 4].contiguous()
            out = torch.empty(n, thr.shape[1], dtype=torch.int32, device=self.device)
            call("fvx_rank_counts", C.byref(self.struct()), ptr(theta), u0, u1, ptr(mask_row_ptr), ptr(mask_col),
                 thr.shape[1], ptr(thr), ptr(out), stream_ptr())
            counts[:, c0:c0 + 4] = out
        return counts

    def _score_topk_view(self, view, mask_row_ptr, mask_col, k, u0, u1, tc):
        n = u1 - u0
        m, th = view["struct"], view["theta"]
        ids = torch.empty(n, k, dtype=torch.int32, device=self.device)
        sc = torch.empty(n, k, dtype=torch.float32, device=self.device)
        tc = self.use_tensor_cores if tc is None else tc
        if n <= 0:
            return ids, sc
        if tc and self.K + self.d + 3 <= 448:
            ws = self._eval_ws(n, struct=m, Ic=view["Ic"], key="_ws_view")
            call("fvx_score_topk_tc", C.byref(m), ptr(th), u0, u1, ptr(mask_row_ptr), ptr(mask_col), k, ptr(ids),
                 ptr(sc), C.byref(ws["struct"]), stream_ptr())
            self._last_flags = ws["flags"][:n]
            return ids, sc
        call("fvx_score_topk", C.byref(m), ptr(th), u0, u1, ptr(mask_row_ptr), ptr(mask_col), k, ptr(ids), ptr(sc),
             0, None, None, stream_ptr())
        return ids, sc

    @property
    def tc_overflow_rows(self):
        """Rows of the last tensor-core sweep that went through the exact fp32 kernel (synchronises)."""
        f = getattr(self, "_last_flags", None)
        return 0 if f is None else int((f != 0).sum().item())

    def _eval_ws(self, n_users, struct=None, Ic=None, key="_ws"):
        """Caller-owned workspace of fvx_score_topk_tc, sized by fvx_eval_ws_query and cached."""
        ws = getattr(self, key, None)
        q = _lib.FvxEvalWs()
        struct = self.struct() if struct is None else struct
        Ic = self.Ic if Ic is None else Ic
        call("fvx_eval_ws_query", C.byref(struct), n_users, C.byref(q))
        if ws is None or ws["n"] != n_users or ws["KP"] != q.KP or ws["splits"] != q.splits or ws["Ic"] != Ic \
                or ws["n_ut"] != q.n_ut:
            dv = self.device
            gm = getattr(self, "_ws_gmax", None)            # per-CTA group maxima: shared by all workspaces
            if gm is None or gm.numel() < q.gmax_elems:
                gm = self._ws_gmax = torch.empty(q.gmax_elems, dtype=torch.float32, device=dv)
            ws = {"n": n_users, "KP": q.KP, "splits": q.splits, "lists": q.lists, "Ic": Ic, "n_ut": q.n_ut,
                  "A": torch.empty(n_users * q.KP, dtype=torch.uint16, device=dv),
                  "Bm": torch.empty(Ic * q.KP, dtype=torch.uint16, device=dv),
                  "epsa": torch.empty(n_users, dtype=torch.float32, device=dv),
                  "nb": torch.empty(Ic, dtype=torch.float32, device=dv),
                  "nbc": torch.empty(Ic // 32 + 1, dtype=torch.float32, device=dv),
                  "stat": torch.zeros(2, dtype=torch.float32, device=dv),
                  "thr": torch.zeros(n_users, dtype=torch.int32, device=dv),
                  "cand": torch.empty(q.lists * q.cap, dtype=torch.int64, device=dv),
                  "ccount": torch.zeros(q.lists, dtype=torch.int32, device=dv),
                  "flags": torch.zeros(n_users, dtype=torch.int32, device=dv), "gmax": gm}
            q.A, q.Bm, q.epsa, q.nb, q.stat = ptr(ws["A"]), ptr(ws["Bm"]), ptr(ws["epsa"]), ptr(ws["nb"]), ptr(ws["stat"])
            q.cand, q.ccount, q.flags, q.thr = ptr(ws["cand"]), ptr(ws["ccount"]), ptr(ws["flags"]), ptr(ws["thr"])
            q.gmax = ptr(gm)
            q.nbc = ptr(ws["nbc"])
            ws["struct"] = q
            setattr(self, key, ws)
        ws["struct"].a_stride = int(getattr(self, "eval_a_stride", 1))
        return ws


class HostStepper:
    """Host batches in, one loss per step out, with one step in flight.

    The reference returns ``float(loss)`` from every ``train_step`` (BPRMF.py:125), i.e. the host waits
    for the device once per step and the device then waits for the host to stage the next batch.  This
    helper keeps that contract - one host batch uploaded and one loss read back per step - but
    pipelines it: the upload of batch s+1 (copy stream, double-buffered device slots) and the launch of
    step s+1 are issued before the host blocks on the loss of step s.

        st = HostStepper(engine.step, loss_of, B, device)     # or ShardedStep.step
        st.submit(batch)        # batch: pinned int32 [3, B] host tensor (user, pos, neg)
        loss = st.collect()     # loss of the oldest uncollected step (blocks on that step only)

    ``loss_of(slot)`` returns a 1-element device tensor with the step's loss (for an item-sharded step:
    after the all-reduce of the per-rank parts) and may clear the accumulator."""

    def __init__(self, step_fn, loss_of, max_batch, device, depth=2):
        self.step_fn, self.loss_of, self.depth = step_fn, loss_of, int(depth)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.dev = [torch.empty(3, max_batch, dtype=torch.int32, device=self.device) for _ in range(self.depth)]
        self.loss_host = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(self.depth)]
        self.copied = [torch.cuda.Event() for _ in range(self.depth)]
        self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self.n_submitted = self.n_collected = 0

    def pending(self):
        return self.n_submitted - self.n_collected

    def submit(self, host_batch):
        if self.pending() >= self.depth:
            raise RuntimeError("HostStepper: collect() before submitting more than %d steps" % self.depth)
        i = self.n_submitted % self.depth
        B = host_batch.shape[1]
        cur = torch.cuda.current_stream(self.device)
        buf = self.dev[i][:, :B]
        with torch.cuda.stream(self.copy_stream):
            # slot i was last read by step n_submitted - depth, whose `done` event the host has waited for
            buf.copy_(host_batch, non_blocking=True)
            self.copied[i].record(self.copy_stream)
        cur.wait_event(self.copied[i])
        self.step_fn(buf[0], buf[1], buf[2], loss_slot=i)
        self.loss_host[i].copy_(self.loss_of(i), non_blocking=True)
        self.done[i].record(cur)
        self.n_submitted += 1

    def collect(self):
        if self.pending() <= 0:
            raise RuntimeError("HostStepper: nothing to collect")
        i = self.n_collected % self.depth
        self.done[i].synchronize()
        self.n_collected += 1
        return float(self.loss_host[i].item())


def topk_merge(ids, scores):
    """ids/scores: [n_users, R, k] per-shard sorted lists -> merged [n_users, k]."""
    n, R, k = ids.shape
    out_i = torch.empty(n, k, dtype=torch.int32, device=ids.device)
    out_s = torch.empty(n, k, dtype=torch.float32, device=ids.device)
    call("fvx_topk_merge", ptr(ids.contiguous()), ptr(scores.contiguous()), n, R, k, ptr(out_i), ptr(out_s),
         stream_ptr())
    return out_i, out_s
