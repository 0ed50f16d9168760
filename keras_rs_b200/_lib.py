"""ctypes binding of libkrs_b200.so (C ABI declared in include/krs_b200.h).

The product path has NO CPU or eager-PyTorch fallback: if the CUDA extension is missing, import
fails loudly.  Tensors are only containers — every call hands raw device pointers and the current
CUDA stream to the hand-written sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# KRS_B200_LIB: alternate build of the same library (e.g. the -DKRS_TC_TRACE=1 build tests/tc_trace.py needs); debug only
LIB_PATH = os.environ.get("KRS_B200_LIB") or os.path.join(_HERE, "lib", "libkrs_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"keras_rs_b200: CUDA extension {LIB_PATH} is missing. Build it with "
        "`python -c 'import __graft_entry__ as g; g.build()'` (or keras_rs_b200/csrc/build.sh). "
        "There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

c_f32p = C.c_void_p
i64 = C.c_int64
i32 = C.c_int


class KrsFeature(C.Structure):
    """krs_feature_t (include/krs_b200.h)."""
    _fields_ = [
        ("table", C.c_void_p), ("ids", C.c_void_p), ("weights", C.c_void_p), ("grad", C.c_void_p),
        ("touched", C.c_void_p), ("vocab", C.c_int64), ("ids_stride", C.c_int64),
        ("hotness", C.c_int32), ("dim", C.c_int32), ("out_offset", C.c_int32),
        ("combiner", C.c_int32), ("ids_i64", C.c_int32), ("reduce", C.c_int32),
        ("shard_tables", C.c_void_p), ("shard_grads", C.c_void_p), ("shard_touched", C.c_void_p),
        ("num_shards", C.c_int32),
        ("shard_mode", C.c_int32),
    ]


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


# every symbol include/krs_b200.h declares (tests/test_abi.py checks this list against the header)
_sig("krs_version", C.c_int)
_sig("krs_last_error", C.c_char_p)
_sig("krs_device_sm_count", C.c_int)
_sig("krs_set_gemm_engine", C.c_int, i32)
_sig("krs_get_gemm_engine", C.c_int)
_sig("krs_gemm_tc_launch_count", C.c_longlong)
_sig("krs_gemm_tc_set_trace", C.c_int, C.c_void_p)
_sig("krs_gemm_set_workspace", C.c_int, C.c_void_p, C.c_size_t)
_sig("krs_gemm_split_launch_count", C.c_longlong)
_sig("krs_gather_fwd", C.c_int, C.POINTER(KrsFeature), i32, i64, c_f32p, i64, i32, C.c_void_p)
_sig("krs_gather_bwd", C.c_int, C.POINTER(KrsFeature), i32, i64, c_f32p, i64, C.c_void_p)
_sig("krs_cross_fwd", C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_float, i32, c_f32p, c_f32p,
     c_f32p, c_f32p, i64, i32, i32, C.c_void_p)
_sig("krs_cross_bwd", C.c_int, *([c_f32p] * 8), C.c_float, i32, *([c_f32p] * 7), i64, i32, i32, i32,
     C.c_void_p)
_sig("krs_cross_combine_fwd", C.c_int, c_f32p, c_f32p, c_f32p, C.c_float, c_f32p, i64, C.c_void_p)
_sig("krs_cross_combine_bwd", C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, C.c_float, c_f32p, c_f32p,
     c_f32p, i64, C.c_void_p)
_sig("krs_dense_fwd", C.c_int, c_f32p, c_f32p, c_f32p, i32, c_f32p, i64, i32, i32, C.c_void_p)
_sig("krs_dense_bwd", C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, i32, c_f32p, c_f32p, c_f32p, c_f32p,
     i64, i32, i32, C.c_void_p)
_sig("krs_sgemm", C.c_int, c_f32p, c_f32p, c_f32p, i64, i64, i64, i32, i32, i32, C.c_void_p)
_sig("krs_dot_fwd", C.c_int, C.POINTER(C.c_void_p), C.POINTER(i64), i32, i32, i64, i32, i32, c_f32p,
     C.c_void_p)
_sig("krs_dot_bwd", C.c_int, C.POINTER(C.c_void_p), C.POINTER(i64), c_f32p, C.POINTER(C.c_void_p),
     C.POINTER(i64), i32, i32, i64, i32, i32, C.c_void_p)
_sig("krs_topk_workspace_bytes", C.c_size_t, i64, i64, i32, i32)
_sig("krs_set_topk_engine", C.c_int, i32)
_sig("krs_topk_tc_launch_count", C.c_longlong)
_sig("krs_topk", C.c_int, c_f32p, c_f32p, C.c_void_p, c_f32p, C.c_void_p, i64, i64, i32, i32,
     C.c_void_p, C.c_size_t, C.c_void_p)
_sig("krs_row_topk", C.c_int, c_f32p, i64, i32, i64, c_f32p, i64, C.c_float, i32, c_f32p, C.c_void_p, c_f32p, i64, c_f32p,
     C.c_void_p, i64, C.c_void_p, C.c_void_p)
_sig("krs_row_scatter", C.c_int, c_f32p, C.c_void_p, i64, i32, i32, c_f32p, C.c_void_p)
_sig("krs_remove_accidental_hits", C.c_int, c_f32p, c_f32p, C.c_void_p, i32, i32, i64, i32, C.c_float, c_f32p, C.c_void_p)
_sig("krs_sampling_prob_correction", C.c_int, c_f32p, c_f32p, i64, i64, C.c_float, c_f32p, C.c_void_p)
_sig("krs_loss_fwd_bwd", C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, i64, i32, i64, C.c_void_p)
_sig("krs_adamw", C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p, i64, i32, C.c_float,
     C.c_float, C.c_float, C.c_float, C.c_float, i64, c_f32p, C.c_void_p)
_sig("krs_adamw_cold", C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_void_p, i64, i32, C.c_float,
     C.c_float, C.c_float, C.c_float, C.c_float, i64, c_f32p, C.c_void_p)
_sig("krs_adam_hyper_advance", C.c_int, c_f32p, C.c_void_p)
_sig("krs_sgd_adagrad", C.c_int, c_f32p, c_f32p, c_f32p, C.c_void_p, i64, i32, C.c_float, C.c_float,
     i32, C.c_void_p)
_sig("krs_mod_route", C.c_int, C.c_void_p, i32, i64, i32, C.c_void_p, C.c_void_p, C.c_void_p,
     C.c_void_p)
XCHG_MAX_SHARDS = 16


class KrsXchg(C.Structure):
    """krs_xchg_t (include/krs_b200.h): one rank's view of the row-sharded exchange regions."""
    _fields_ = [
        ("S", C.c_int32), ("me", C.c_int32), ("F", C.c_int32), ("E", C.c_int32), ("B", C.c_int64),
        ("peer_base", C.c_void_p * XCHG_MAX_SHARDS),
        ("off_flags", C.c_int64), ("off_hdr", C.c_int64), ("off_rows", C.c_int64), ("off_pos", C.c_int64),
        ("off_x0", C.c_int64), ("off_grad", C.c_int64),
    ]


_sig("krs_xchg_route_workspace_bytes", C.c_size_t, i64, i32, i32)
_sig("krs_xchg_route", C.c_int, C.POINTER(KrsXchg), i32, C.c_void_p, i32, i64, C.c_void_p, C.c_void_p, C.c_void_p, i32,
     C.c_void_p)
_sig("krs_xchg_barrier", C.c_int, C.POINTER(KrsXchg), C.c_uint32, C.c_double, C.c_void_p)
_sig("krs_xchg_gather_push", C.c_int, C.POINTER(KrsXchg), i32, c_f32p, C.c_void_p, C.c_void_p)
_sig("krs_slot_scan_blocks", C.c_size_t, i64)
_sig("krs_slot_scan", C.c_int, C.c_void_p, i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
_sig("krs_xchg_grad_pull", C.c_int, C.POINTER(KrsXchg), i32, C.c_void_p, C.c_void_p, C.c_void_p, c_f32p, C.c_void_p, i64,
     C.c_void_p)
_sig("krs_rows_apply", C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_void_p, i64, i32, i32,
     C.POINTER(C.c_float), C.c_void_p, i64, C.c_void_p)
_sig("krs_opt_apply", C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p, i64, i32, i32, C.POINTER(C.c_float), C.c_void_p)
_sig("krs_adamw_compact", C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i64, i32,
     C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i64, C.c_void_p)
_sig("krs_ipc_alloc", C.c_int, C.POINTER(C.c_void_p), C.c_size_t, C.c_void_p)
_sig("krs_ipc_open", C.c_int, C.c_void_p, C.POINTER(C.c_void_p))
_sig("krs_ipc_close", C.c_int, C.c_void_p)
_sig("krs_ipc_free", C.c_int, C.c_void_p)
_sig("krs_enable_peer_access", C.c_int, i32)

EXPORTED = [
    "krs_version", "krs_last_error", "krs_device_sm_count", "krs_set_gemm_engine",
    "krs_get_gemm_engine", "krs_gemm_tc_launch_count", "krs_gemm_tc_set_trace", "krs_gemm_set_workspace", "krs_gemm_split_launch_count", "krs_gather_fwd", "krs_gather_bwd", "krs_cross_fwd", "krs_cross_bwd",
    "krs_cross_combine_fwd", "krs_cross_combine_bwd", "krs_dense_fwd", "krs_dense_bwd", "krs_sgemm",
    "krs_dot_fwd", "krs_dot_bwd", "krs_topk_workspace_bytes", "krs_topk", "krs_set_topk_engine", "krs_topk_tc_launch_count", "krs_row_topk", "krs_row_scatter",
    "krs_remove_accidental_hits", "krs_sampling_prob_correction", "krs_loss_fwd_bwd",
    "krs_adamw", "krs_adamw_cold", "krs_adam_hyper_advance", "krs_sgd_adagrad", "krs_mod_route",
    "krs_xchg_route_workspace_bytes", "krs_xchg_route", "krs_xchg_barrier", "krs_xchg_gather_push", "krs_slot_scan_blocks",
    "krs_slot_scan", "krs_xchg_grad_pull", "krs_rows_apply", "krs_opt_apply", "krs_adamw_compact", "krs_ipc_alloc", "krs_ipc_open",
    "krs_ipc_close", "krs_ipc_free", "krs_enable_peer_access",
]

ACT = {None: 0, "linear": 0, "relu": 1, "sigmoid": 2, "tanh": 3, "swish": 4, "silu": 4}
COMBINER = {"sum": 0, "mean": 1, "sqrtn": 2}
CROSS_ACC_DX0 = 1
CROSS_SAME_INPUT = 2


class KrsError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = lib.krs_last_error().decode("utf-8", "replace")
        raise KrsError(f"libkrs_b200 error {rc}: {msg}")


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    """The product path only takes CUDA tensors — no silent CPU fallback."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise KrsError(f"{name} must live on a CUDA device (keras_rs_b200 has no CPU path); got {t.device}")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t


