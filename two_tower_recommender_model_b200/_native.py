"""ctypes binding of libtt_b200.so (the C ABI declared in include/tt_b200.h).

There is NO fallback: if the library is missing, was built for another
architecture, or a call fails, this module raises.  PyTorch is used only to own
device memory and streams; every pointer handed to the library is
``tensor.data_ptr()`` and every call is enqueued on torch's current stream.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p
from typing import Optional

import torch

TT_MAX_FEATURES = 32
TT_ABI_VERSION = 3

POOL_SUM, POOL_MEAN = 0, 1
OPT_DENSE_GRAD, OPT_ROWWISE_ADAGRAD, OPT_ROWWISE_ADAM, OPT_SGD = 0, 1, 2, 3

# TT_B200_LIB selects another build of the SAME library (kernel A/B experiments); there is still no fallback.
_LIB_PATH = os.environ.get("TT_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libtt_b200.so")


class NativeLibraryError(RuntimeError):
    pass


class EbcPlan(Structure):
    _fields_ = [
        ("num_slots", c_int32), ("batch_size", c_int32), ("num_kjt_keys", c_int32), ("out_stride", c_int32),
        ("total_rows", c_int64),
        ("weights", c_void_p * TT_MAX_FEATURES), ("state0", c_void_p * TT_MAX_FEATURES),
        ("state1", c_void_p * TT_MAX_FEATURES), ("row_base", c_int64 * TT_MAX_FEATURES),
        ("num_rows", c_int64 * TT_MAX_FEATURES), ("dim", c_int32 * TT_MAX_FEATURES),
        ("kjt_index", c_int32 * TT_MAX_FEATURES), ("out_col", c_int32 * TT_MAX_FEATURES),
        ("pooling", c_int32 * TT_MAX_FEATURES), ("slot_of_kjt", c_int32 * TT_MAX_FEATURES),
    ]


TT_MAX_PEERS = 16
TT_MAX_TOWERS = 2
TT_PEER_SCATTER_ADD = 1


class TowerForward(Structure):
    _fields_ = [("x", c_void_p), ("ldx", c_int64), ("w1_bf16", c_void_p), ("ldw1", c_int64), ("b1", c_void_p),
                ("w2_bf16", c_void_p), ("ldw2", c_int64), ("b2", c_void_p), ("xb", c_void_p), ("hb", c_void_p),
                ("yb", c_void_p), ("y", c_void_p), ("ldy", c_int64)]


class TowerBackward(Structure):
    _fields_ = [("dy", c_void_p), ("lddy", c_int64), ("w1_bf16", c_void_p), ("ldw1", c_int64), ("w2_bf16", c_void_p),
                ("ldw2", c_int64), ("xb", c_void_p), ("hb", c_void_p), ("yb", c_void_p), ("dx", c_void_p), ("lddx", c_int64),
                ("dw1", c_void_p), ("dw2", c_void_p), ("db1", c_void_p), ("db2", c_void_p)]


class PeerBuffers(Structure):
    _fields_ = [("world", c_int32), ("rows_per_peer", c_int32), ("flags", c_int32), ("reserved", c_int32),
                ("ptr", c_void_p * TT_MAX_PEERS)]


class SparseOptimizer(Structure):
    _fields_ = [
        ("kind", c_int32), ("lr", c_float), ("eps", c_float), ("beta1", c_float), ("beta2", c_float),
        ("bias_correction1", c_float), ("bias_correction2", c_float), ("grad_scale", c_float),
        ("step_dev", c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol include/tt_b200.h declares.
_P = c_void_p
SIGNATURES = {
    "tt_abi_version": (c_int32, []),
    "tt_last_error": (c_char_p, []),
    "tt_build_arch": (c_char_p, []),
    "tt_kernel_launch_count": (ctypes.c_uint64, []),
    "tt_kjt_offsets_workspace_bytes": (c_size_t, [c_int64]),
    "tt_kjt_lengths_to_offsets": (c_int32, [_P, _P, c_int64, _P, c_size_t, _P]),
    "tt_kjt_from_columns_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "tt_kjt_from_columns": (c_int32, [_P, _P, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "tt_kjt_from_columns_range": (c_int32, [_P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "tt_kjt_permute_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "tt_kjt_permute_2d": (c_int32, [_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "tt_kjt_bucketize_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64]),
    "tt_kjt_gathered_range_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "tt_kjt_gathered_range": (c_int32, [_P, c_int64, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "tt_kjt_block_bucketize": (c_int32, [_P, _P, _P, c_int64, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, c_size_t, _P]),
    "tt_ebc_forward": (c_int32, [POINTER(EbcPlan), _P, _P, _P, _P]),
    "tt_ebc_forward_peer": (c_int32, [POINTER(EbcPlan), _P, _P, POINTER(PeerBuffers), _P]),
    "tt_ebc_backward_fused_peer": (c_int32, [POINTER(EbcPlan), POINTER(SparseOptimizer), _P, c_int64, _P, POINTER(PeerBuffers), _P, c_size_t, _P]),
    "tt_ebc_backward_workspace_bytes": (c_size_t, [c_int64]),
    "tt_ebc_backward_fused": (c_int32, [POINTER(EbcPlan), POINTER(SparseOptimizer), _P, c_int64, _P, _P, _P, c_size_t, _P]),
    "tt_ebc_dedup_workspace_bytes": (c_size_t, [c_int64]),
    "tt_ebc_dedup": (c_int32, [POINTER(EbcPlan), _P, c_int64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "tt_sort_pairs_workspace_bytes": (c_size_t, [c_int64]),
    "tt_sort_pairs_u32": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, _P, c_size_t, _P]),
    "tt_linear_forward_f32": (c_int32, [_P, c_int64, _P, _P, _P, c_int64, c_int64, c_int64, c_int32, _P]),
    "tt_linear_backward_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "tt_linear_backward_f32": (c_int32, [_P, c_int64, _P, _P, _P, _P, c_int64, _P, _P, c_int64, c_int64, c_int64, c_int32, _P, c_size_t, _P]),
    "tt_cast_f32_to_bf16": (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int64, _P, c_int64, _P, c_int64, _P]),
    "tt_gemm_bf16": (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int32, _P, c_int64, _P, c_int64,
                               _P, c_int64, _P, c_int64, _P, c_int64, _P]),
    "tt_gemm_bf16_splitk_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "tt_gemm_bf16_splitk": (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "tt_gemm_bf16_splitk_mn": (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "tt_colsum_bf16_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "tt_colsum_bf16": (c_int32, [_P, c_int64, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "tt_dot_bce_workspace_bytes": (c_size_t, [c_int64]),
    "tt_dot_bce": (c_int32, [_P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, c_float, _P, c_size_t, _P]),
    "tt_inbatch_softmax_workspace_bytes": (c_size_t, [c_int64]),
    "tt_inbatch_softmax_forward_f32": (c_int32, [_P, _P, c_int64, c_int64, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "tt_inbatch_softmax_backward_f32": (c_int32, [_P, _P, _P, c_int64, c_int64, c_float, c_float, _P, _P, _P]),
    "tt_towers_forward_fused": (c_int32, [POINTER(TowerForward), c_int32, c_int64, c_int32, c_int32, c_int32, _P]),
    "tt_towers_backward_workspace_bytes": (c_size_t, [c_int64]),
    "tt_towers_backward_fused": (c_int32, [POINTER(TowerBackward), c_int32, c_int64, c_int32, c_int32, c_int32, _P, c_size_t, _P]),
    "tt_set_softmax_backward_mode": (c_int32, [c_int32]),
    "tt_set_softmax_wide_mode": (c_int32, [c_int32]),
    "tt_inbatch_softmax_bf16_workspace_bytes": (c_size_t, [c_int64]),
    "tt_inbatch_softmax_forward_bf16": (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int64, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "tt_inbatch_softmax_backward_bf16": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64,
                                                   _P, c_int64, c_int64, c_float, c_float, c_int32, _P, c_int64, _P, c_int64, _P, _P]),
    "tt_adam_flat":(c_int32, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_float, _P]),
    "tt_adam_flat_devstep": (c_int32, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, _P, _P]),
    "tt_debug_read_counters": (c_int32, [_P, c_int32, c_int32]),
    "tt_topk_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "tt_topk_bf16_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "tt_score_topk_bf16": (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, c_size_t, _P]),
    "tt_score_topk_f32": (c_int32, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, c_size_t, _P]),
}

_lib: Optional[ctypes.CDLL] = None
launch_count = 0  # number of library calls that enqueue kernels (bench.py reads it)


def library_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Loads the shared library (once).  Raises NativeLibraryError if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise NativeLibraryError(
            f"{_LIB_PATH} not found. Build it with `python -m two_tower_recommender_model_b200.build` "
            "(needs nvcc; targets sm_100a). There is no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{_LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    if lib.tt_abi_version() != TT_ABI_VERSION:
        raise NativeLibraryError(f"ABI version mismatch: library {lib.tt_abi_version()}, binding {TT_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tt_last_error()
        raise NativeLibraryError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise NativeLibraryError(
            f"{name} is on {t.device}; this package computes on CUDA (sm_100a) only -- there is no CPU path")


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


_timing = None  # name -> list of (start_event, end_event) while enabled


def enable_timing(on: bool) -> None:
    """Brackets every library call with CUDA events on torch's current stream
    (bench.py's instrumented pass).  Off by default: zero overhead."""
    global _timing
    _timing = {} if on else None


def timing_summary():
    """name -> {"ms": mean device time per call, "calls": n}; synchronise first."""
    out = {}
    for name, evs in (_timing or {}).items():
        ms = [a.elapsed_time(b) for a, b in evs]
        out[name] = {"ms": sum(ms) / len(ms), "calls": len(ms)}
    return out


def call(name: str, *args) -> None:
    """Invokes a status-returning entry point and raises on failure."""
    global launch_count
    lib = load()
    launch_count += 1
    if _timing is None:
        check(getattr(lib, name)(*args), name)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(getattr(lib, name)(*args), name)
    b.record()
    _timing.setdefault(name, []).append((a, b))
