"""CPU tests of the drop-in boundary: libfa_b200.so loads, exports every symbol include/fa_b200.h
declares, and validates its arguments without touching a GPU (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

from flash_attention_impls_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fa_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fa_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_functions() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared_functions():
        assert getattr(lib, name) is not None
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    for name in _declared_functions():
        assert re.search(rf"\bT {name}\b", out), f"{name} is not an exported text symbol"


def test_reference_named_cxx_shims_are_exported():
    """flash_attention_cutlass_dispatch & co keep the reference's C++ linkage (flash_attn_cutlass.cu:519-529)."""
    out = subprocess.check_output(["nm", "-D", "-C", "--defined-only", _lib.LIB_PATH], text=True)
    for fn in ("flash_attention_cutlass_dispatch", "flash_attention_forward_dispatch",
               "flash_attention_small_tile_dispatch", "attention_reference_dispatch"):
        assert re.search(rf"{fn}\(cutlass::half_t const\*, cutlass::half_t const\*, cutlass::half_t const\*, "
                         rf"cutlass::half_t\*, int, int, int, int, CUstream_st\*\)", out), fn


def test_struct_layout_matches_header():
    # 7 pointers, 7 ints + float, 11 int64 strides, stream, workspace pointer + size  (x86-64 LP64)
    assert ctypes.sizeof(_lib.FaB200Params) == 7 * 8 + 8 * 4 + 11 * 8 + 8 + 16 + 8   # ... + precise (int, padded)


def test_version_and_status_strings():
    lib = _lib.load()
    assert lib.fa_b200_version() == (0 << 16) | 6
    assert lib.fa_b200_status_string(0) == b"ok"
    assert lib.fa_b200_status_string(3) == b"unsupported head_dim"
    assert lib.fa_b200_status_string(99) == b"unknown status"


def _params(**kw):
    p = _lib.FaB200Params()
    p.Q = p.K = p.V = p.O = 0x1000          # never dereferenced: validation fails first
    p.B, p.H, p.N, p.d = 1, 1, 128, 64
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.mark.parametrize("kw,status", [
    (dict(Q=None), 1), (dict(O=None), 1), (dict(B=0), 2), (dict(N=-5), 2), (dict(N_kv=-1), 2),
    (dict(d=44), 3), (dict(d=0), 3), (dict(d=136), 3), (dict(d=256), 3), (dict(dtype=7), 4),
    (dict(Q=0x1008), 5), (dict(q_stride_h=8 * 1024 + 4), 5), (dict(o_stride_n=68), 5), (dict(q_stride_n=32), 2),
    (dict(stat_stride_h=64), 2),
])
def test_argument_validation_returns_codes_without_a_gpu(kw, status):
    """Error convention of SURVEY.md section 8b: an int status instead of the reference's stderr +
    silent no-launch on unsupported head_dim (flash_attn_cutlass.cu:540-542)."""
    lib = _lib.load()
    rc = lib.fa_b200_forward(ctypes.byref(_params(**kw)))
    assert rc == status
    assert lib.fa_b200_last_error() != b""
    assert lib.fa_b200_forward(None) == 1


def test_merge_and_cast_validation():
    lib = _lib.load()
    assert lib.fa_b200_merge_partial(None, None, None, None, 4, 64, 1, None) == 1
    assert lib.fa_b200_merge_partial(0x1000, 0x1000, 0x1000, 0x1000, 4, 60, 1, None) == 2
    assert lib.fa_b200_cast_output(0x1000, 0x1000, 4, 64, 5, None) == 4


def test_combine_and_peer_validation():
    lib = _lib.load()
    assert lib.fa_b200_combine_partials(None, None, 2, None, None, 4, 64, 1, None) == 1
    assert lib.fa_b200_combine_partials(0x1000, 0x1000, 0, 0x1000, None, 4, 64, 1, None) == 2       # nparts
    assert lib.fa_b200_combine_partials(0x1000, 0x1000, 2, 0x1000, None, 4, 60, 1, None) == 2       # d % 8
    assert lib.fa_b200_combine_partials(0x1000, 0x1000, 2, 0x1000, None, 1 << 31, 64, 1, None) == 2  # rows < 2^31
    assert lib.fa_b200_combine_partials(0x1000, 0x1000, 2, 0x1000, None, 4, 64, 9, None) == 4       # dtype
    assert lib.fa_b200_combine_partials(0x1008, 0x1000, 2, 0x1000, None, 4, 64, 1, None) == 5       # alignment
    assert lib.fa_b200_peer_alloc(0, None, None) == 1
    assert lib.fa_b200_peer_open(None, None) == 1
    assert lib.fa_b200_copy_async(None, None, 16, None) == 1
    assert lib.fa_b200_peer_free(None) == 0 and lib.fa_b200_peer_close(None) == 0      # NULL is a no-op, like cudaFree


def test_backward_validation_returns_codes_without_a_gpu():
    lib = _lib.load()
    assert lib.fa_b200_backward(None) == 1
    p = _lib.FaB200BwdParams()
    for name in ("Q", "K", "V", "O", "dO", "lse", "dQ", "dK", "dV", "delta"):
        setattr(p, name, 0x1000)
    p.B, p.H, p.N, p.d, p.dtype = 1, 2, 128, 64, 1
    p.dK = None
    assert lib.fa_b200_backward(ctypes.byref(p)) == 1          # missing output
    p.dK = 0x1000
    p.N = 0
    assert lib.fa_b200_backward(ctypes.byref(p)) == 2
    p.N, p.d = 128, 96
    assert lib.fa_b200_backward(ctypes.byref(p)) == 3
    p.d, p.dtype = 64, 3
    assert lib.fa_b200_backward(ctypes.byref(p)) == 4
    p.dtype, p.dO = 0, 0x1004
    assert lib.fa_b200_backward(ctypes.byref(p)) == 5
    p.dO, p.N_kv = 0x1000, -3
    assert lib.fa_b200_backward(ctypes.byref(p)) == 2
    p.N_kv = 0
    p.do_stride[2] = 68
    assert lib.fa_b200_backward(ctypes.byref(p)) == 5          # stride not a multiple of 8
    p.do_stride[2] = 32
    assert lib.fa_b200_backward(ctypes.byref(p)) == 2          # row stride smaller than d
    # 10 pointers, 6 ints + float (+ padding), stream, N_kv (+ padding), 6 x 3 strides
    assert ctypes.sizeof(_lib.FaB200BwdParams) == 10 * 8 + 8 * 4 + 8 + 8 + 18 * 8


def test_precise_flag_is_part_of_the_parameter_block():
    """`precise` is the last field; a zeroed block means the default single 16-bit P."""
    p = _params()
    assert p.precise == 0
    assert _lib.FaB200Params.precise.offset == ctypes.sizeof(_lib.FaB200Params) - 8


def test_no_cpu_fallback_in_python_surface():
    import torch
    import flash_attention_impls_b200 as fa
    q = torch.zeros(1, 1, 128, 64, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fa.flash_attention(q, q, q)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fa.attention_forward(q, q, q)


def test_product_package_does_not_import_oracle():
    """The oracle is the checker only: nothing under flash_attention_impls_b200/ may reference it."""
    pkg = os.path.join(ROOT, "flash_attention_impls_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle." not in text and "import oracle" not in text and "liboracle" not in text, f


def test_workspace_query_policy():
    """Split-KV is only proposed for launches that would leave most SMs idle and are long enough to cut."""
    lib = _lib.load()
    assert lib.fa_b200_workspace_bytes(4, 32, 8192, 0, 128) == 0          # c3: 4096 items, no split
    assert lib.fa_b200_workspace_bytes(8, 16, 1024, 0, 64) == 0           # c2: 512 items of 8 K/V tiles: too short for a tail split
    # c3 sharded over 8 GPUs: 512 equal items on 148 SMs = 3.46 waves: the 68 items of the last wave are split in two
    assert lib.fa_b200_workspace_bytes(1, 16, 8192, 0, 128) == 2 * 68 * 256 * (128 * 2 + 8)
    assert lib.fa_b200_workspace_bytes(1, 37, 1024, 0, 64) == 0           # 148 items: exactly one wave
    assert lib.fa_b200_workspace_bytes(1, 1, 512, 0, 64) == 0             # too short to split
    n = lib.fa_b200_workspace_bytes(1, 1, 8192, 0, 64)                    # the reference's (1,1,8192,64) sweep point
    assert n > 0 and n % (8192 * (64 * 2 + 8)) == 0
    assert 2 <= n // (8192 * (64 * 2 + 8)) <= 32
    assert lib.fa_b200_workspace_bytes(1, 1, 8192, 0, 44) == 0            # invalid shape -> 0


def test_ring_handle_validation_without_a_gpu():
    """fa_b200_ring_* (SURVEY.md section 8b ownership row): argument errors are reported before any device call."""
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.fa_b200_ring_create(2, 0, 1, 1, 128, 64, 0, None) == 1
    assert lib.fa_b200_ring_create(0, 0, 1, 1, 128, 64, 0, ctypes.byref(h)) == 2          # world
    assert lib.fa_b200_ring_create(65, 0, 1, 1, 128, 64, 0, ctypes.byref(h)) == 2
    assert lib.fa_b200_ring_create(2, 2, 1, 1, 128, 64, 0, ctypes.byref(h)) == 2          # rank
    assert lib.fa_b200_ring_create(2, 0, 1, 1, 0, 64, 0, ctypes.byref(h)) == 2            # n_local
    assert lib.fa_b200_ring_create(2, 0, 1, 1, 128, 44, 0, ctypes.byref(h)) == 3          # head_dim
    assert lib.fa_b200_ring_create(2, 0, 1, 1, 128, 64, 5, ctypes.byref(h)) == 4          # dtype
    assert not h.value
    assert lib.fa_b200_ring_forward(None, None, None, None, None, None, 0, 0.0, None) == 1
    assert lib.fa_b200_ring_export(None, None) == 1 and lib.fa_b200_ring_connect(None, None) == 1
    assert lib.fa_b200_ring_device_bytes(None) == 0 and lib.fa_b200_ring_destroy(None) == 0
    assert _lib.FA_B200_RING_EXPORT_BYTES == int(re.search(r"#define FA_B200_RING_EXPORT_BYTES (\d+)", open(HEADER).read()).group(1))


def test_debug_overrides_are_compiled_out_of_the_release_library():
    """The UMMA-descriptor / split-count environment overrides exist only in -DFA_B200_DEBUG builds."""
    blob = open(_lib.LIB_PATH, "rb").read()
    for name in (b"FA_B200_DESC_HI_QK", b"FA_B200_DESC_HI_V", b"FA_B200_IDESC_QK", b"FA_B200_IDESC_PV", b"FA_B200_NSPLIT"):
        assert name not in blob, name
    assert b"FA_B200_GROUP_HEADS" in blob      # the one tuning knob left, read once (and settable through the ABI)
