"""DataLoader: the reference's data protocol (src/dataset/dataset.py:8-122) over CSR
arrays, with the triple sampler either replayed in the reference's own RNG streams
on the host (``sampler='host_ref'``) or generated on the device (``sampler='device'``).

Kept from the reference: ``DataLoader(params)``; ``num_users``, ``num_items``,
``training_list``, ``validation_list``, ``test_list``, ``params``;
``all_triple_batches()``; ``next_triple_batch()`` yielding ``(user, pos, neg)`` batches
of ``params.batch_size`` triples, ``(N // B) * B * epochs`` triples in total, cut into
consecutive batches across epoch boundaries (dataset.py:89-91,109-110,116-122).
"""
from __future__ import annotations

import random

import numpy as np

from ..config import configs


class CSRLists:
    """List-of-lists view of a CSR (what the reference's ``training_list`` is)."""

    def __init__(self, row_ptr, col):
        self.row_ptr, self.col = row_ptr, col

    def __len__(self):
        return len(self.row_ptr) - 1

    def __getitem__(self, u):
        if u < 0:
            u += len(self)
        return self.col[self.row_ptr[u]:self.row_ptr[u + 1]].tolist()

    def __iter__(self):
        for u in range(len(self)):
            yield self[u]

    def __bool__(self):
        return len(self) > 0


def _read_pairs(path):
    """Fast TSV reader: first two columns (user, item) of ``u\\ti\\tt\\t1.0`` rows."""
    import pandas as pd
    df = pd.read_csv(path, sep="\t", header=None, usecols=[0, 1], dtype=np.int64, engine="c")
    return df[0].to_numpy(), df[1].to_numpy()


def _to_csr(users, items, num_users):
    order = np.argsort(users, kind="stable")          # keeps file order inside a user
    users, items = users[order], items[order]
    row_ptr = np.zeros(num_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(users, minlength=num_users), out=row_ptr[1:])
    return row_ptr, items.astype(np.int32)


def _sort_rows(row_ptr, col):
    owner = np.repeat(np.arange(len(row_ptr) - 1, dtype=np.int64), np.diff(row_ptr))
    order = np.lexsort((col, owner))
    return col[order]


_SHARED_STREAMS = {}     # seed -> (random.Random, RandomState) shared by the loaders of this process


class DataLoader(object):
    def __init__(self, params, interactions=None):
        """``params``: the argparse namespace of train_rec.py.  ``interactions``: an
        in-memory ``synth.Interactions`` instead of the TSV files (benchmarks)."""
        self.params = params
        self.sampler = getattr(params, "sampler", "host_ref")
        self.seed = int(getattr(params, "seed", 0))
        if interactions is None:
            self.num_users, self.num_items = self.get_length()
            self.train_ptr, self.train_col = self.load_list("train")
            if getattr(params, "validation", True):
                self.val_ptr, self.val_col = self.load_list("val")
            else:
                self.val_ptr, self.val_col = np.zeros(1, np.int64), np.zeros(0, np.int32)
            self.test_ptr, self.test_col = self.load_list("test")
        else:
            it = interactions
            self.num_users, self.num_items = it.num_users, it.num_items
            self.train_ptr, self.train_col = it.row_ptr.astype(np.int64), it.col_file.astype(np.int32)
            one = np.arange(it.num_users + 1, dtype=np.int64)
            self.val_ptr, self.val_col = one, it.val.astype(np.int32)
            self.test_ptr, self.test_col = one.copy(), it.test.astype(np.int32)
        self.train_col_sorted = _sort_rows(self.train_ptr, self.train_col)
        self.training_list = CSRLists(self.train_ptr, self.train_col)
        self.validation_list = CSRLists(self.val_ptr, self.val_col)
        self.test_list = CSRLists(self.test_ptr, self.test_col)
        self.num_train = int(self.train_ptr[-1])
        # host_ref streams: seeded like the model modules do at import (BPRMF.py:15-16).  The reference seeds the
        # GLOBAL streams once and keeps consuming them from one DataLoader to the next (train_rec.py builds one per
        # regulariser): ``params.share_sampler_streams`` gives the loaders of one process that one stream pair.
        if getattr(params, "share_sampler_streams", False):
            self._py_rng, self._np_rng = _SHARED_STREAMS.setdefault(
                self.seed, (random.Random(self.seed), np.random.RandomState(self.seed)))
        else:
            self._py_rng = random.Random(self.seed)
            self._np_rng = np.random.RandomState(self.seed)
        self._dev = None

    # ---- files (dataset.py:41-81) --------------------------------------------------------
    def get_length(self):
        with open(configs.dataset_info(self.params.dataset)) as f:
            lines = f.readlines()
        return int(lines[2].split(": ")[1]), int(lines[3].split(": ")[1])

    def load_list(self, which):
        path = {"train": configs.training_path, "val": configs.validation_path,
                "test": configs.test_path}[which](self.params.dataset)
        u, i = _read_pairs(path)
        return _to_csr(u, i, self.num_users)

    # ---- reference-stream sampler (dataset.py:83-114) ---------------------------------------
    def _total_triples(self):
        B = self.params.batch_size
        return (self.num_train // B) * B * self.params.epochs

    def _epoch_pairs(self, order):
        order = np.asarray(order, dtype=np.int64)
        lens = (self.train_ptr[order + 1] - self.train_ptr[order])
        users = np.repeat(order, lens)
        start = np.repeat(self.train_ptr[order], lens)
        offs = np.zeros(len(order) + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        within = np.arange(int(offs[-1]), dtype=np.int64) - np.repeat(offs[:-1], lens)
        return users, self.train_col[start + within].astype(np.int64)

    def _member(self, users, items):
        if not hasattr(self, "_train_keys"):
            owner = np.repeat(np.arange(self.num_users, dtype=np.int64), np.diff(self.train_ptr))
            self._train_keys = owner * self.num_items + self.train_col_sorted.astype(np.int64)
        q = users * np.int64(self.num_items) + items
        idx = np.searchsorted(self._train_keys, q)
        idx = np.minimum(idx, len(self._train_keys) - 1)
        return self._train_keys[idx] == q

    def _negatives_ref(self, users):
        """Sequential rejection sampling of dataset.py:100-103, consuming the legacy
        NumPy stream draw for draw but evaluated in vectorised runs between rejections."""
        n, I = len(users), self.num_items
        draws = self._np_rng.randint(I, size=n).astype(np.int64)
        neg = np.empty(n, dtype=np.int64)
        t = shift = 0
        while t < n:
            need = n + shift - len(draws)
            if need > 0:
                draws = np.concatenate([draws, self._np_rng.randint(I, size=need).astype(np.int64)])
            cand = draws[t + shift:n + shift]
            rej = np.flatnonzero(self._member(users[t:], cand))
            if rej.size == 0:
                neg[t:] = cand
                break
            f = int(rej[0])
            neg[t:t + f] = cand[:f]
            p = t + f + shift + 1
            u1 = users[t + f:t + f + 1]
            while True:
                if p >= len(draws):
                    draws = np.concatenate([draws, self._np_rng.randint(I, size=1).astype(np.int64)])
                if not self._member(u1, draws[p:p + 1])[0]:
                    break
                p += 1
            neg[t + f] = draws[p]
            shift = p - (t + f)
            t = t + f + 1
        return neg

    def all_triple_batches(self):
        """All triples of all epochs in the reference's order and RNG streams (int64 arrays)."""
        total = self._total_triples()
        U, P, Ng, have = [], [], [], 0
        for _ in range(self.params.epochs):
            if have >= total:
                break
            order = list(range(self.num_users))
            self._py_rng.shuffle(order)                                   # dataset.py:94-95
            u, p = self._epoch_pairs(order)
            if have + len(u) > total:
                u, p = u[:total - have], p[:total - have]
            U.append(u)
            P.append(p)
            Ng.append(self._negatives_ref(u))
            have += len(u)
        if not U:
            z = np.zeros(0, np.int64)
            return z, z.copy(), z.copy()
        return np.concatenate(U), np.concatenate(P), np.concatenate(Ng)

    # ---- device-side generation ------------------------------------------------------------
    def device_state(self, device="cuda:0"):
        """CSR arrays resident on the device (uploaded once)."""
        import torch
        if self._dev is None or self._dev["device"] != str(device):
            dv = torch.device(device)
            self._dev = {
                "device": str(device),
                "row_ptr": torch.from_numpy(self.train_ptr).to(dv),
                "col_file": torch.from_numpy(self.train_col).to(dv),
                "col_sorted": torch.from_numpy(np.ascontiguousarray(self.train_col_sorted)).to(dv),
                "lens": torch.from_numpy(np.diff(self.train_ptr)).to(dv),
            }
        return self._dev

    def device_epoch(self, epoch, device="cuda:0", out=None):
        """(user, pos, neg) int32 CUDA tensors of one whole epoch (N triples), generated on
        the device with no host synchronisation: Feistel/Philox user permutation ->
        prefix sum of the list lengths -> CSR expansion with the negatives drawn in the same
        pass.  ``out``: optional (user, pos, neg) tensors of N elements to write into."""
        import torch
        from .. import _lib
        st = self.device_state(device)
        dv = torch.device(device)
        U, N = self.num_users, self.num_train
        ws = st.get("epoch_ws")
        if ws is None:
            ws = st["epoch_ws"] = (torch.empty(U, dtype=torch.int32, device=dv),
                                   torch.empty(U, dtype=torch.int64, device=dv),
                                   torch.empty(U, dtype=torch.int64, device=dv))
        perm, lens, offs_incl = ws
        _lib.call("fvx_epoch_perm", _lib.ptr(perm), _lib.ptr(lens), _lib.ptr(st["row_ptr"]), U, self.seed, epoch,
                  _lib.stream_ptr())
        torch.cumsum(lens, 0, out=offs_incl)
        if out is None:
            out = tuple(torch.empty(N, dtype=torch.int32, device=dv) for _ in range(3))
        users, pos, neg = out
        _lib.call("fvx_epoch_triples", _lib.ptr(st["row_ptr"]), _lib.ptr(st["col_file"]), _lib.ptr(st["col_sorted"]),
                  _lib.ptr(perm), _lib.ptr(offs_incl), U, self.num_items, self.seed, epoch * N, _lib.ptr(users),
                  _lib.ptr(pos), _lib.ptr(neg), _lib.stream_ptr())
        return users, pos, neg

    def next_triple_batch(self, device="cuda:0", reuse=True):
        """Iterator over ``(user, pos, neg)`` int32 CUDA tensors of ``batch_size`` triples.
        With the device sampler the batches are views into two alternating epoch buffers: a
        batch stays valid until the epoch after the next one is generated (``reuse=False``
        allocates a fresh buffer per epoch for callers that keep every batch)."""
        B = self.params.batch_size
        for bufs, n in self.next_batch_run(device, reuse=reuse):
            for s in range(n):
                yield tuple(x[s * B:(s + 1) * B] for x in bufs)

    def next_batch_run(self, device="cuda:0", reuse=True):
        """Iterator over ``((user, pos, neg), n)``: index tensors holding ``n`` consecutive batches back to back
        (batch s = elements [s*B, (s+1)*B)) - what ``Engine.steps`` replays as CUDA graphs.  The same triples, in
        the same order, as ``next_triple_batch``: ``(N // B) * B * epochs`` in total, batches cut across epoch
        boundaries (dataset.py:89-91,109-110,116-122)."""
        import torch
        B = self.params.batch_size
        total = self._total_triples()
        if self.sampler == "host_ref":
            u, p, n = self.all_triple_batches()
            dv = torch.device(device)
            u, p, n = (torch.from_numpy(a.astype(np.int32)).to(dv) for a in (u, p, n))
            if total:
                yield (u, p, n), total // B
            return
        if self.sampler != "device":
            raise ValueError("unknown sampler %r (host_ref | device)" % self.sampler)
        # two fixed buffers of N + B triples: the tail of one epoch that does not fill a batch is
        # copied to the front of the other buffer and the next epoch is generated behind it, so
        # batches are consecutive slices across epoch boundaries and nothing is allocated per epoch
        N = self.num_train
        dv = torch.device(device)
        bufs = [tuple(torch.empty(N + B, dtype=torch.int32, device=dv) for _ in range(3)) for _ in range(2)]
        done, epoch, carry = 0, 0, 0
        prev, prev_s = None, 0
        while done < total:
            if not reuse and epoch >= 2:
                bufs[epoch & 1] = tuple(torch.empty(N + B, dtype=torch.int32, device=dv) for _ in range(3))
            cur = bufs[epoch & 1]
            if carry:
                for c, x in zip(cur, prev):
                    c[:carry].copy_(x[prev_s:prev_s + carry])
            self.device_epoch(epoch, device, out=tuple(c[carry:carry + N] for c in cur))
            epoch += 1
            n_av = carry + N
            n = min(n_av // B, (total - done) // B)
            if n > 0:
                yield cur, n
            done += n * B
            prev, prev_s, carry = cur, n * B, n_av - n * B
