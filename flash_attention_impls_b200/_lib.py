"""ctypes binding of libfa_b200.so (the C ABI declared in include/fa_b200.h).

There is no CPU or PyTorch fallback: if the CUDA library is missing, importing the symbols raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfa_b200.so")

FA_B200_FP16 = 0
FA_B200_BF16 = 1

STATUS = {
    0: "FA_B200_OK", 1: "FA_B200_ERR_NULL", 2: "FA_B200_ERR_SHAPE", 3: "FA_B200_ERR_HEAD_DIM",
    4: "FA_B200_ERR_DTYPE", 5: "FA_B200_ERR_ALIGNMENT", 6: "FA_B200_ERR_ARCH", 7: "FA_B200_ERR_CUDA",
    8: "FA_B200_ERR_DRIVER",
}

# every symbol include/fa_b200.h declares (tests/test_abi.py checks the library exports all of them)
EXPORTED_SYMBOLS = (
    "fa_b200_forward", "fa_b200_workspace_bytes", "fa_b200_forward_legacy", "fa_b200_forward_fp16", "fa_b200_merge_partial",
    "fa_b200_combine_partials", "fa_b200_cast_output", "fa_b200_backward", "fa_b200_host_ctx_create", "fa_b200_forward_host", "fa_b200_host_ctx_sync",
    "fa_b200_host_ctx_elapsed_ms", "fa_b200_host_ctx_destroy", "fa_b200_peer_alloc", "fa_b200_peer_free", "fa_b200_peer_open",
    "fa_b200_peer_close", "fa_b200_copy_async", "fa_b200_work_item", "fa_b200_launch_count", "fa_b200_last_error", "fa_b200_status_string",
    "fa_b200_version", "fa_b200_tmap_cache_stats", "fa_b200_set_group_heads",
    "fa_b200_ring_create", "fa_b200_ring_export", "fa_b200_ring_connect", "fa_b200_ring_forward", "fa_b200_ring_kv_buffers",
    "fa_b200_ring_wait_consumed", "fa_b200_ring_device_bytes", "fa_b200_ring_set_profile", "fa_b200_ring_timeline",
    "fa_b200_ring_destroy",
)
FA_B200_RING_EXPORT_BYTES = 512


class FaB200Params(Structure):
    """Mirror of `struct fa_b200_params` (include/fa_b200.h)."""
    _fields_ = [
        ("Q", c_void_p), ("K", c_void_p), ("V", c_void_p), ("O", c_void_p),
        ("lse", c_void_p), ("l", c_void_p), ("m", c_void_p),
        ("B", c_int), ("H", c_int), ("N", c_int), ("d", c_int),
        ("N_kv", c_int), ("dtype", c_int), ("causal", c_int), ("softmax_scale", c_float),
        ("q_stride_b", c_int64), ("q_stride_h", c_int64), ("q_stride_n", c_int64),
        ("kv_stride_b", c_int64), ("kv_stride_h", c_int64), ("kv_stride_n", c_int64),
        ("o_stride_b", c_int64), ("o_stride_h", c_int64), ("o_stride_n", c_int64),
        ("stat_stride_b", c_int64), ("stat_stride_h", c_int64),
        ("stream", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("precise", c_int),
    ]


class FaB200BwdParams(Structure):
    """Mirror of `struct fa_b200_bwd_params` (include/fa_b200.h)."""
    _fields_ = [
        ("Q", c_void_p), ("K", c_void_p), ("V", c_void_p), ("O", c_void_p), ("dO", c_void_p), ("lse", c_void_p),
        ("dQ", c_void_p), ("dK", c_void_p), ("dV", c_void_p), ("delta", c_void_p),
        ("B", c_int), ("H", c_int), ("N", c_int), ("d", c_int), ("dtype", c_int), ("causal", c_int),
        ("softmax_scale", c_float), ("stream", c_void_p),
        ("N_kv", c_int),
        ("q_stride", c_int64 * 3), ("kv_stride", c_int64 * 3), ("o_stride", c_int64 * 3), ("do_stride", c_int64 * 3),
        ("dq_stride", c_int64 * 3), ("dkv_stride", c_int64 * 3),
    ]


class FaB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


_lib = None


def load() -> ctypes.CDLL:
    """Load libfa_b200.so; raises (loudly) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make lib` (or `python -c 'import __graft_entry__ as g; "
            "g.build()'`). flash_attention_impls_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.fa_b200_forward.argtypes = [POINTER(FaB200Params)]
    lib.fa_b200_forward.restype = c_int
    lib.fa_b200_workspace_bytes.argtypes = [c_int] * 5
    lib.fa_b200_workspace_bytes.restype = ctypes.c_size_t
    lib.fa_b200_forward_legacy.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.fa_b200_forward_legacy.restype = c_int
    lib.fa_b200_forward_fp16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                         c_void_p]
    lib.fa_b200_forward_fp16.restype = c_int
    lib.fa_b200_backward.argtypes = [POINTER(FaB200BwdParams)]
    lib.fa_b200_backward.restype = c_int
    lib.fa_b200_merge_partial.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]
    lib.fa_b200_merge_partial.restype = c_int
    lib.fa_b200_combine_partials.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]
    lib.fa_b200_combine_partials.restype = c_int
    lib.fa_b200_cast_output.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]
    lib.fa_b200_cast_output.restype = c_int
    lib.fa_b200_host_ctx_create.argtypes = [c_int] * 7 + [POINTER(c_void_p)]
    lib.fa_b200_host_ctx_create.restype = c_int
    lib.fa_b200_forward_host.argtypes = [c_void_p] * 6
    lib.fa_b200_forward_host.restype = c_int
    lib.fa_b200_host_ctx_sync.argtypes = [c_void_p]
    lib.fa_b200_host_ctx_sync.restype = c_int
    lib.fa_b200_host_ctx_elapsed_ms.argtypes = [c_void_p, POINTER(c_float)]
    lib.fa_b200_host_ctx_elapsed_ms.restype = c_int
    lib.fa_b200_host_ctx_destroy.argtypes = [c_void_p]
    lib.fa_b200_host_ctx_destroy.restype = None
    lib.fa_b200_peer_alloc.argtypes = [ctypes.c_size_t, POINTER(c_void_p), ctypes.c_char_p]
    lib.fa_b200_peer_alloc.restype = c_int
    lib.fa_b200_peer_free.argtypes = [c_void_p]
    lib.fa_b200_peer_free.restype = c_int
    lib.fa_b200_peer_open.argtypes = [ctypes.c_char_p, POINTER(c_void_p)]
    lib.fa_b200_peer_open.restype = c_int
    lib.fa_b200_peer_close.argtypes = [c_void_p]
    lib.fa_b200_peer_close.restype = c_int
    lib.fa_b200_copy_async.argtypes = [c_void_p, c_void_p, ctypes.c_size_t, c_void_p]
    lib.fa_b200_copy_async.restype = c_int
    lib.fa_b200_work_item.argtypes = [c_int] * 7 + [POINTER(c_int)] * 4
    lib.fa_b200_work_item.restype = c_int
    lib.fa_b200_launch_count.argtypes = []
    lib.fa_b200_launch_count.restype = c_uint64
    lib.fa_b200_last_error.argtypes = []
    lib.fa_b200_last_error.restype = c_char_p
    lib.fa_b200_status_string.argtypes = [c_int]
    lib.fa_b200_status_string.restype = c_char_p
    lib.fa_b200_version.argtypes = []
    lib.fa_b200_version.restype = c_int
    lib.fa_b200_set_group_heads.argtypes = [c_int]
    lib.fa_b200_set_group_heads.restype = None
    lib.fa_b200_tmap_cache_stats.argtypes = [POINTER(c_uint64), POINTER(c_uint64)]
    lib.fa_b200_tmap_cache_stats.restype = None
    lib.fa_b200_ring_create.argtypes = [c_int] * 7 + [POINTER(c_void_p)]
    lib.fa_b200_ring_create.restype = c_int
    lib.fa_b200_ring_export.argtypes = [c_void_p, ctypes.c_char_p]
    lib.fa_b200_ring_export.restype = c_int
    lib.fa_b200_ring_connect.argtypes = [c_void_p, ctypes.c_char_p]
    lib.fa_b200_ring_connect.restype = c_int
    lib.fa_b200_ring_forward.argtypes = [c_void_p] * 6 + [c_int, c_float, c_void_p]
    lib.fa_b200_ring_forward.restype = c_int
    lib.fa_b200_ring_kv_buffers.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p)]
    lib.fa_b200_ring_kv_buffers.restype = c_int
    lib.fa_b200_ring_wait_consumed.argtypes = [c_void_p, c_void_p]
    lib.fa_b200_ring_wait_consumed.restype = c_int
    lib.fa_b200_ring_device_bytes.argtypes = [c_void_p]
    lib.fa_b200_ring_device_bytes.restype = ctypes.c_size_t
    lib.fa_b200_ring_set_profile.argtypes = [c_void_p, c_int]
    lib.fa_b200_ring_set_profile.restype = c_int
    lib.fa_b200_ring_timeline.argtypes = [c_void_p, POINTER(c_float), c_int]
    lib.fa_b200_ring_timeline.restype = c_int
    lib.fa_b200_ring_destroy.argtypes = [c_void_p]
    lib.fa_b200_ring_destroy.restype = c_int
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        raise FaB200Error(status, load().fa_b200_last_error().decode())


def work_item(B: int, H: int, N: int, d: int, causal: bool, index: int, N_kv: int = 0):
    """(num_items, bh, q0, tiles0, tiles1) of work item `index` of the launch for this shape (host-only)."""
    out = [c_int(0) for _ in range(4)]
    n = load().fa_b200_work_item(B, H, N, N_kv, d, 1 if causal else 0, index, *[ctypes.byref(x) for x in out])
    if n == 0:
        raise FaB200Error(2, load().fa_b200_last_error().decode())
    return (n,) + tuple(x.value for x in out)


def launch_count() -> int:
    return int(load().fa_b200_launch_count())


def tmap_cache_stats():
    """(hits, misses) of the library's TMA tensor-map cache."""
    h, m = c_uint64(0), c_uint64(0)
    load().fa_b200_tmap_cache_stats(ctypes.byref(h), ctypes.byref(m))
    return int(h.value), int(m.value)
