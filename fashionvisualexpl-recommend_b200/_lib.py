"""ctypes binding of csrc/libfvx.so (the C ABI declared in include/fvx.h).

The library is loaded lazily on first use; loading or calling it without the built
shared object or without a CUDA device raises - there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libfvx.so")

ABI_VERSION = 5
ADAM_DENSE, ADAM_DEFERRED, ADAM_LAZY = 0, 1, 2
N_PHASES = 5
PHASES = ("prep", "project", "score_grad", "grad_E", "update")
ADAM_MODES = {"dense": ADAM_DENSE, "deferred": ADAM_DEFERRED, "lazy": ADAM_LAZY}

_p = C.c_void_p


class FvxTable(C.Structure):
    _fields_ = [("w", _p), ("m", _p), ("v", _p), ("g", _p), ("last", _p), ("mark", _p), ("list", _p),
                ("count", _p), ("rows", C.c_int64), ("stride", C.c_int32), ("list_cap", C.c_int32)]


class FvxModel(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("num_users", C.c_int32), ("num_items", C.c_int32),
                ("item_lo", C.c_int32), ("item_cnt", C.c_int32), ("K", C.c_int32), ("d", C.c_int32),
                ("D", C.c_int32), ("de", C.c_int32), ("adam_mode", C.c_int32), ("lr", C.c_float),
                ("reg", C.c_float), ("users", FvxTable), ("items", FvxTable), ("E", _p), ("mE", _p),
                ("vE", _p), ("gE_part", _p), ("ge_parts", C.c_int32), ("_pad0", C.c_int32), ("F", _p),
                ("F_pl", _p), ("ET_hi", _p), ("ET_lo", _p), ("W_hi", _p), ("W_lo", _p),
                ("step", _p), ("loss", _p), ("loss_slots", C.c_int32), ("_pad1", C.c_int32), ("TH", _p),
                ("th_cap", C.c_int64), ("W", _p), ("rows", _p), ("sync", _p), ("cmap", _p), ("max_batch", C.c_int32),
                ("use_tensor_cores", C.c_int32), ("upos", _p), ("W_sum", _p), ("uslot", _p), ("user_lo", C.c_int32),
                ("user_cnt", C.c_int32), ("two_stage", C.c_int32), ("Dc", C.c_int32), ("De", C.c_int32),
                ("ec", C.c_int32), ("ee", C.c_int32), ("bias_neg_scale", C.c_float), ("Ec", _p), ("mEc", _p), ("vEc", _p),
                ("Ee", _p), ("mEe", _p), ("vEe", _p), ("E2", _p), ("mE2", _p), ("vE2", _p), ("gf_scratch", _p),
                ("batch_stage", _p)]


class FvxShardWs(C.Structure):
    _fields_ = [("S", _p), ("run_id", _p), ("run_scratch", _p), ("WU", _p), ("RU", _p), ("dE", _p), ("loss_part", _p),
                ("max_runs", C.c_int32), ("run_cap", C.c_int32), ("owners", C.c_int32), ("users_per_owner", C.c_int32),
                ("p2p", C.c_int32), ("_pad", C.c_int32), ("RUin", _p), ("dEall", _p), ("tails", _p), ("flags", _p),
                ("run_user", _p), ("run_counts", _p)]


COMM_ID_BYTES = 256


class FvxEvalWs(C.Structure):
    _fields_ = [("A", _p), ("Bm", _p), ("epsa", _p), ("nb", _p), ("stat", _p), ("cand", _p), ("ccount", _p),
                ("flags", _p), ("thr", _p), ("gmax", _p), ("nbc", _p), ("lists", C.c_int64), ("gmax_elems", C.c_int64),
                ("u_cap", C.c_int32), ("i_cap", C.c_int32), ("KP", C.c_int32), ("splits", C.c_int32),
                ("cap", C.c_int32), ("n_ut", C.c_int32), ("a_stride", C.c_int32), ("_pad", C.c_int32)]


# name -> (restype, argtypes); exactly the prototypes of include/fvx.h
_i32, _i64, _u32, _u64 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
_MP = C.POINTER(FvxModel)
PROTOTYPES = {
    "fvx_abi_version": (C.c_int, []),
    "fvx_last_error": (C.c_char_p, []),
    "fvx_sizeof_model": (C.c_int, []),
    "fvx_sizeof_table": (C.c_int, []),
    "fvx_sizeof_shard_ws": (C.c_int, []),
    "fvx_sizeof_eval_ws": (C.c_int, []),
    "fvx_enumerate_epoch": (C.c_int, [_p, _p, _p, _p, _i32, _p, _p, _p]),
    "fvx_epoch_perm": (C.c_int, [_p, _p, _p, _i32, _u64, _u32, _p]),
    "fvx_epoch_triples": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _u64, _u64, _p, _p, _p, _p]),
    "fvx_sample_negatives": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _u64, _u64, _p]),
    "fvx_bpr_step": (C.c_int, [_MP, _p, _p, _p, _i32, _i32, _p]),
    "fvx_bpr_steps": (C.c_int, [_MP, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "fvx_bpr_step_timed": (C.c_int, [_MP, _p, _p, _p, _i32, _i32, C.POINTER(C.c_float), _p]),
    "fvx_run_ids": (C.c_int, [_p, _i64, _p, _p, _p]),
    "fvx_run_slots": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p]),
    "fvx_bpr_step_sharded": (C.c_int, [_MP, C.POINTER(FvxShardWs), _p, _p, _p, _p, _i32, _i32, _p]),
    "fvx_bpr_step_sharded_phase": (C.c_int, [_MP, C.POINTER(FvxShardWs), _p, _p, _p, _i32, _i32, _i32, _p]),
    "fvx_comm_unique_id": (C.c_int, [_p]),
    "fvx_comm_create": (C.c_int, [_p, _i32, _i32, C.POINTER(_p)]),
    "fvx_comm_destroy": (C.c_int, [_p]),
    "fvx_comm_arena": (C.c_int, [_p, _i64, C.POINTER(_p)]),
    "fvx_comm_all_reduce_f32": (C.c_int, [_p, _p, _i64, _p]),
    "fvx_adam_flush": (C.c_int, [_MP, _p]),
    "fvx_project": (C.c_int, [_MP, _p, _p]),
    "fvx_predict_all": (C.c_int, [_MP, _p, _i32, _i32, _p, _p]),
    "fvx_score_topk": (C.c_int, [_MP, _p, _i32, _i32, _p, _p, _i32, _p, _p, _i32, _p, _p, _p]),
    "fvx_rank_counts": (C.c_int, [_MP, _p, _i32, _i32, _p, _p, _i32, _p, _p, _p]),
    "fvx_score_topk_users": (C.c_int, [_MP, _p, _p, _i32, _p, _p, _i32, _p, _p, _p]),
    "fvx_eval_ws_query": (C.c_int, [_MP, _i32, C.POINTER(FvxEvalWs)]),
    "fvx_score_topk_tc": (C.c_int, [_MP, _p, _i32, _i32, _p, _p, _i32, _p, _p, C.POINTER(FvxEvalWs), _p]),
    "fvx_score_topk_tc_bounds": (C.c_int, [_MP, _p, _i32, _i32, _p, _i32, C.POINTER(FvxEvalWs), _p]),
    "fvx_score_topk_tc_select": (C.c_int, [_MP, _p, _i32, _i32, _p, _p, _i32, _p, _p, C.POINTER(FvxEvalWs), _p]),
    "fvx_score_pairs": (C.c_int, [_MP, _p, _p, _p, _i64, _p, _p]),
    "fvx_topk_merge": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p, _p]),
    "fvx_split_planes": (C.c_int, [_p, _p, _i64, _i32, _p]),
    "fvx_tc_width": (C.c_int, [_i32]),
    "fvx_project_rows": (C.c_int, [_MP, _p, _i64, _p, _p]),
    "fvx_grad_e_rows": (C.c_int, [_MP, _p, _i64, _p, _p, _p]),
    "fvx_debug_set_dedup": (C.c_int, [C.c_int]),
    "fvx_debug_trace": (C.c_int, [C.c_int]),
    "fvx_debug_trace_read": (C.c_int, [C.POINTER(C.c_float)]),
    "fvx_debug_trace_sharded": (C.c_int, [C.c_int]),
    "fvx_debug_trace_sharded_read": (C.c_int, [C.POINTER(C.c_float)]),
}

_lib = None


class FvxError(RuntimeError):
    pass


def load():
    """Loads libfvx.so (once) and checks the ABI version and struct layouts."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FvxError("libfvx.so is not built (%s missing): run `python -m fvx.build`; "
                       "this package has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.fvx_abi_version() != ABI_VERSION:
        raise FvxError("libfvx ABI %d != binding %d" % (lib.fvx_abi_version(), ABI_VERSION))
    if lib.fvx_sizeof_model() != C.sizeof(FvxModel) or lib.fvx_sizeof_table() != C.sizeof(FvxTable) or \
            lib.fvx_sizeof_shard_ws() != C.sizeof(FvxShardWs) or lib.fvx_sizeof_eval_ws() != C.sizeof(FvxEvalWs):
        raise FvxError("struct layout mismatch between include/fvx.h and fvx/_lib.py")
    _lib = lib
    return lib


def call(name, *args):
    """Calls an entry point; non-zero return -> FvxError(fvx_last_error())."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise FvxError("%s failed (%d): %s" % (name, rc, lib.fvx_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL); must be a contiguous CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise FvxError("libfvx entry points take CUDA tensors only (got a %s tensor)" % t.device)
    if not t.is_contiguous():
        raise FvxError("libfvx entry points take contiguous tensors")
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
