"""Host-side mirror of the reference's operator surface for the attention-forward path.

Reference interfaces mirrored (same names, argument meaning and error behaviour):
  * `flash_attention(q, k, v, causal=False)`           code/triton_fa2/FA2-triton.py:240-244
  * `_FlashAttnFn.forward` -> (o, m, l)                code/triton_fa2/FA2-triton.py:173-205
  * `flash_attention_forward(Q,K,V,O,l,m,B,H,N,d,M)`   code/cuda_fa1/flashAttention.h:8-11 (+ main.cu:297-303)
  * `flash_attention_cutlass_dispatch(Q,K,V,O,B,H,N,d,stream)`  code/cutlass_cuda_fa1/run/flash_attn_cutlass.cu:519-529

PyTorch is used for device memory and streams only; every FLOP runs in libfa_b200.so.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import FA_B200_BF16, FA_B200_FP16, FaB200Params


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return FA_B200_FP16
    if t.dtype == torch.bfloat16:
        return FA_B200_BF16
    raise TypeError(f"unsupported dtype {t.dtype}: fp16 or bf16 required")


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if not t.is_cuda:
            # same contract as the reference: FA2-triton.py:176 asserts q.is_cuda
            raise RuntimeError("flash_attention_impls_b200 runs on CUDA (sm_100) tensors only; there is no CPU fallback")


def attention_forward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, causal: bool = False,
                      softmax_scale: Optional[float] = None, out: Optional[torch.Tensor] = None,
                      lse: Optional[torch.Tensor] = None, l: Optional[torch.Tensor] = None,
                      m: Optional[torch.Tensor] = None, return_lse: bool = True, allow_split: bool = True,
                      precise: bool = False):
    """O = softmax(Q K^T * scale [+ causal mask]) V on [B,H,N,d] tensors; returns (O, lse).

    q: [B,H,N,d]; k, v: [B,H,N_kv,d]; fp16 or bf16.  Only the last dim has to be contiguous: batch, head and row
    strides are passed to the kernel's TMA descriptors as they are (multiples of 8 elements), so row sub-ranges of
    a longer sequence (what the ring driver passes) and `x.transpose(1, 2)` views of [B,N,H,d] tensors need no copy.
    k and v share their strides (else both are copied); views a TMA descriptor cannot address - a stride-0 broadcast such
    as `k.expand(...)` for GQA, strides that are not multiples of 8 - are copied first.  precise=True feeds P to the tensor cores as a hi+lo pair of 16-bit operands
    (fp32-like P, the accuracy of the reference's CUDA-core FA1 kernel; about 1.4x the time) - see
    `fa_b200_params.precise` in include/fa_b200.h.
    """
    _require_cuda(q, k, v)
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise ValueError("q, k, v must be [B, H, N, d]")
    B, H, N, d = q.shape
    Bk, Hk, Nkv, dk = k.shape
    if (Bk, Hk, dk) != (B, H, d) or v.shape != k.shape:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)}")
    if k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError("q, k, v must share a dtype")
    dtype = _dtype_code(q)

    # Stride 0 (a broadcast view such as k.expand(B, H, N, d) for GQA/MQA) is NOT expressible to TMA, and the C ABI
    # reads 0 as "dense default", so such a view must never reach it as is (see _tma_ok).
    # inputs the descriptors cannot address are copied (broadcast heads, odd strides); outputs must be addressable
    q, k, v = (t if _tma_ok(t) else t.contiguous() for t in (q, k, v))

    def strides(t: torch.Tensor, what: str):
        if not _tma_ok(t):
            raise ValueError(f"{what}: needs a contiguous last dimension and batch/head/row strides that are positive "
                             "multiples of 8 elements")
        return _strides3(t)

    qs, ks, vs = strides(q, "q"), strides(k, "k"), strides(v, "v")
    if ks != vs:      # K and V share one set of strides in the ABI
        k, v = k.contiguous(), v.contiguous()
        ks = vs = strides(k, "k")
    if out is None:
        out = torch.empty((B, H, N, d), dtype=q.dtype, device=q.device)
    if out.shape != (B, H, N, d) or out.dtype != q.dtype:
        raise ValueError("out must be [B,H,N,d] of q's dtype")
    os_ = strides(out, "out")
    if lse is None and return_lse:
        lse = torch.empty((B, H, N), dtype=torch.float32, device=q.device)
    ss = None
    for s_ in (lse, l, m):
        if s_ is not None:
            if s_.dtype != torch.float32 or s_.shape != (B, H, N) or (N > 1 and s_.stride(2) != 1) or \
                    any(s_.shape[i] > 1 and s_.stride(i) <= 0 for i in range(2)):
                raise ValueError("lse / l / m must be fp32 [B,H,N] with contiguous rows and positive strides")
            st = tuple(s_.stride(i) if s_.shape[i] > 1 else 0 for i in range(2))
            if ss is not None and st != ss:
                raise ValueError("lse, l and m must share their strides")
            ss = st
    ss = ss or (0, 0)

    p = FaB200Params()
    p.Q, p.K, p.V, p.O = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    p.lse = lse.data_ptr() if lse is not None else None
    p.l = l.data_ptr() if l is not None else None
    p.m = m.data_ptr() if m is not None else None
    p.B, p.H, p.N, p.d = B, H, N, d
    p.N_kv = 0 if Nkv == N else Nkv
    p.dtype = dtype
    p.causal = 1 if causal else 0
    p.softmax_scale = float(softmax_scale) if softmax_scale else 0.0
    p.q_stride_b, p.q_stride_h, p.q_stride_n = qs
    p.kv_stride_b, p.kv_stride_h, p.kv_stride_n = ks
    p.o_stride_b, p.o_stride_h, p.o_stride_n = os_
    p.stat_stride_b, p.stat_stride_h = ss
    p.stream = _stream_ptr(q)
    p.precise = 1 if precise else 0
    with torch.cuda.device(q.device):
        # split-KV scratch for launches that would leave most SMs idle (torch's caching allocator owns it;
        # the library itself never allocates)
        ws_bytes = int(_lib.load().fa_b200_workspace_bytes(B, H, N, p.N_kv, d)) if allow_split else 0
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device) if ws_bytes else None
        if ws is not None:
            p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
        _lib.check(_lib.load().fa_b200_forward(ctypes.byref(p)))
    return out, lse


def _tma_ok(t: torch.Tensor) -> bool:
    """What a TMA descriptor can address: d contiguous, every other stride a positive multiple of 8 elements."""
    return t.stride(3) == 1 and all(t.shape[i] == 1 or (t.stride(i) > 0 and t.stride(i) % 8 == 0) for i in range(3))


def _strides3(t: torch.Tensor):
    # size-1 axes report arbitrary strides in torch: give the kernel the dense default (0) for those
    return tuple(t.stride(i) if t.shape[i] > 1 else 0 for i in range(3))


def attention_backward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, lse: torch.Tensor,
                       do: torch.Tensor, *, causal: bool = False, softmax_scale: Optional[float] = None):
    """(dQ, dK, dV) of O = softmax(Q K^T scale [+ causal mask]) V for the upstream gradient `do`, from the forward's
    output `o` and logsumexp `lse` (reference: FA2-triton.py:98-170, :207-237).

    q, o, do: [B,H,N,d]; k, v: [B,H,N_kv,d]; fp16/bf16 CUDA tensors, d in {32, 64, 128}.  As in the reference's backward
    (which hands every tensor's strides to its kernel) batch / head / row strides are free: views such as
    `x.transpose(1, 2)` of [B,N,H,d] storage go to the kernels' TMA descriptors as they are; only views a descriptor
    cannot address (stride 0, strides that are not multiples of 8) are copied first.  The gradients come back dense.
    Three launches of libfa_b200.so (fa_b200_backward), no atomics."""
    _require_cuda(q, k, v, o, lse, do)
    B, H, N, d = q.shape
    Nkv = k.shape[2]
    if k.shape != (B, H, Nkv, d) or v.shape != k.shape:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)}")
    for t_, name in ((o, "o"), (do, "do")):
        if t_.shape != q.shape:
            raise ValueError(f"{name} must have q's shape")
    for t_, name in ((k, "k"), (v, "v"), (o, "o"), (do, "do")):
        if t_.dtype != q.dtype:
            raise TypeError(f"{name} must have q's dtype")
    if lse.shape != (B, H, N) or lse.dtype != torch.float32:
        raise ValueError("lse must be fp32 [B,H,N]")
    q, k, v, o, do = (t_ if _tma_ok(t_) else t_.contiguous() for t_ in (q, k, v, o, do))
    if _strides3(k) != _strides3(v):      # K and V share one set of strides in the ABI
        k, v = k.contiguous(), v.contiguous()
    lse = lse.contiguous()
    dq, dk, dv = torch.empty_like(q, memory_format=torch.contiguous_format), \
        torch.empty_like(k, memory_format=torch.contiguous_format), torch.empty_like(v, memory_format=torch.contiguous_format)
    delta = torch.empty((B, H, N), dtype=torch.float32, device=q.device)
    p = _lib.FaB200BwdParams()
    p.Q, p.K, p.V, p.O, p.dO, p.lse = (t_.data_ptr() for t_ in (q, k, v, o, do, lse))
    p.dQ, p.dK, p.dV, p.delta = dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), delta.data_ptr()
    p.B, p.H, p.N, p.d = B, H, N, d
    p.N_kv = 0 if Nkv == N else Nkv
    p.dtype = _dtype_code(q)
    p.causal = 1 if causal else 0
    p.softmax_scale = float(softmax_scale) if softmax_scale else 0.0
    for name, t_ in (("q_stride", q), ("kv_stride", k), ("o_stride", o), ("do_stride", do)):
        for i, s_ in enumerate(_strides3(t_)):
            getattr(p, name)[i] = s_
    p.stream = _stream_ptr(q)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().fa_b200_backward(ctypes.byref(p)))
    return dq, dk, dv


class FlashAttnFunction(torch.autograd.Function):
    """Mirror of the reference's `_FlashAttnFn` (FA2-triton.py:173-237): forward saves (q, k, v) and the row
    statistics, backward returns (dQ, dK, dV, None).  The statistics are kept as one logsumexp instead of (m, l)."""

    @staticmethod
    def forward(ctx, q, k, v, causal: bool):
        o, lse = attention_forward(q, k, v, causal=causal)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.causal = causal
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        dq, dk, dv = attention_backward(q, k, v, o, lse, do, causal=ctx.causal)
        return dq, dk, dv, None


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False) -> torch.Tensor:
    """Drop-in for the reference's `flash_attention(q, k, v, causal)` (FA2-triton.py:240-244).

    Like the reference, fp32 inputs are down-cast to fp16 for the kernel and the result is cast back; when an input
    requires grad the call goes through `FlashAttnFunction`, as the reference's goes through `_FlashAttnFn`.
    """
    _require_cuda(q, k, v)
    orig = q.dtype
    if orig == torch.float32:
        q, k, v = q.half(), k.half(), v.half()
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
        o = FlashAttnFunction.apply(q, k, v, causal)
    else:
        o, _ = attention_forward(q.contiguous(), k.contiguous(), v.contiguous(), causal=causal, return_lse=False)
    return o.to(orig) if orig == torch.float32 else o


def flash_attention_with_stats(q, k, v, causal: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(o, m, l) exactly as `_FlashAttnFn.forward` keeps them (FA2-triton.py:179-182, 203):
    m = row max of the scaled scores, l = sum exp(s - m); logsumexp = m + log(l)."""
    B, H, N, _ = q.shape
    m = torch.empty((B, H, N), dtype=torch.float32, device=q.device)
    l = torch.empty_like(m)
    o, _ = attention_forward(q.contiguous(), k.contiguous(), v.contiguous(), causal=causal, l=l, m=m,
                             return_lse=False)
    return o, m, l


def flash_attention_forward(Q, K, V, O, l, m, B: int, H: int, N: int, d: int, M: int = 0, stream=None) -> None:
    """Host call with the argument list of the reference kernel `flash_attention_forward`
    (flashAttention.h:8-11): fp16 [B,H,N,d] in, O fp16, l and m fp32 [B,H,N] out.  M is ignored."""
    _require_cuda(Q, K, V, O, l, m)
    s = stream.cuda_stream if stream is not None else _stream_ptr(Q)
    with torch.cuda.device(Q.device):
        _lib.check(_lib.load().fa_b200_forward_legacy(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(),
                                                      l.data_ptr(), m.data_ptr(), B, H, N, d, M, s))


def flash_attention_cutlass_dispatch(Q, K, V, O, batch_size: int, num_heads: int, seq_len: int, head_dim: int,
                                     stream=None) -> None:
    """Same argument list as the reference dispatcher (flash_attn_cutlass.cu:519-529); fp16, O only."""
    _require_cuda(Q, K, V, O)
    s = stream.cuda_stream if stream is not None else _stream_ptr(Q)
    with torch.cuda.device(Q.device):
        _lib.check(_lib.load().fa_b200_forward_fp16(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(),
                                                    batch_size, num_heads, seq_len, head_dim, s))


# the three CUDA-core dispatchers of flash_attn_unified.cu:545-617 share the signature
flash_attention_forward_dispatch = flash_attention_cutlass_dispatch
flash_attention_small_tile_dispatch = flash_attention_cutlass_dispatch
attention_reference_dispatch = flash_attention_cutlass_dispatch


def merge_partial(o_acc: torch.Tensor, lse_acc: torch.Tensor, o_part: torch.Tensor, lse_part: torch.Tensor) -> None:
    """In-place logsumexp merge of an attention partial into the fp32 accumulator (ring attention)."""
    _require_cuda(o_acc, lse_acc, o_part, lse_part)
    d = o_acc.shape[-1]
    rows = o_acc.numel() // d
    assert o_acc.dtype == torch.float32 and lse_acc.dtype == torch.float32 and lse_part.dtype == torch.float32
    assert o_acc.is_contiguous() and o_part.is_contiguous() and lse_acc.is_contiguous() and lse_part.is_contiguous()
    assert o_part.shape == o_acc.shape and lse_acc.numel() == rows and lse_part.numel() == rows
    with torch.cuda.device(o_acc.device):
        _lib.check(_lib.load().fa_b200_merge_partial(o_acc.data_ptr(), lse_acc.data_ptr(), o_part.data_ptr(),
                                                     lse_part.data_ptr(), rows, d, _dtype_code(o_part),
                                                     _stream_ptr(o_acc)))


def combine_partials(o_parts: torch.Tensor, lse_parts: torch.Tensor, out: Optional[torch.Tensor] = None,
                     lse: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """One-pass logsumexp merge of `o_parts` [P, ..., d] (fp16/bf16) with `lse_parts` [P, ...] (fp32) over the
    leading axis -> (O [..., d] in o_parts.dtype, lse [...] fp32).  Rows of a partial with lse = -inf are skipped."""
    _require_cuda(o_parts, lse_parts)
    assert o_parts.is_contiguous() and lse_parts.is_contiguous() and lse_parts.dtype == torch.float32
    P, d = o_parts.shape[0], o_parts.shape[-1]
    rows = o_parts[0].numel() // d
    assert lse_parts.shape[0] == P and lse_parts[0].numel() == rows
    if out is None:
        out = torch.empty(o_parts.shape[1:], dtype=o_parts.dtype, device=o_parts.device)
    if lse is None:
        lse = torch.empty(lse_parts.shape[1:], dtype=torch.float32, device=o_parts.device)
    assert out.is_contiguous() and lse.is_contiguous() and out.dtype == o_parts.dtype
    with torch.cuda.device(o_parts.device):
        _lib.check(_lib.load().fa_b200_combine_partials(o_parts.data_ptr(), lse_parts.data_ptr(), P, out.data_ptr(),
                                                        lse.data_ptr(), rows, d, _dtype_code(o_parts),
                                                        _stream_ptr(o_parts)))
    return out, lse


def cast_output(o_acc: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    _require_cuda(o_acc)
    out = torch.empty(o_acc.shape, dtype=dtype, device=o_acc.device)
    d = o_acc.shape[-1]
    with torch.cuda.device(o_acc.device):
        _lib.check(_lib.load().fa_b200_cast_output(out.data_ptr(), o_acc.data_ptr(), o_acc.numel() // d, d,
                                                   _dtype_code(out), _stream_ptr(o_acc)))
    return out


class HostPipeline:
    """End-to-end path for HOST buffers: pinned host Q,K,V -> device -> kernel -> pinned host O, lse.

    Thin wrapper over the native C-ABI pipeline (`fa_b200_host_ctx_create` / `fa_b200_forward_host`,
    csrc/fa_host.cu): the (b,h) slices are independent, so one forward is cut into `chunks` groups of slices and
    pipelined on three streams (H2D, compute, D2H) through double-buffered device staging that the context owns:
    while chunk c runs on the tensor cores, chunk c+1 is crossing PCIe and chunk c-1's output is going back.
    This is the call bench.py's `e2e` figure times.
    """

    def __init__(self, B: int, H: int, N: int, d: int, dtype: torch.dtype, causal: bool = False,
                 chunks: int = 8, device=None):
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.B, self.H, self.N, self.d, self.dtype, self.causal = B, H, N, d, dtype, causal
        code = FA_B200_FP16 if dtype == torch.float16 else FA_B200_BF16 if dtype == torch.bfloat16 else None
        if code is None:
            raise TypeError(f"unsupported dtype {dtype}: fp16 or bf16 required")
        self._ctx = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().fa_b200_host_ctx_create(B, H, N, d, code, 1 if causal else 0, chunks,
                                                           ctypes.byref(self._ctx)))

    def bytes_per_call(self) -> Tuple[int, int]:
        e = torch.empty((), dtype=self.dtype).element_size()
        n = self.B * self.H * self.N * self.d
        return 3 * n * e, n * e + self.B * self.H * self.N * 4

    def __call__(self, q_host, k_host, v_host, o_host, lse_host=None):
        """q,k,v,o: pinned CPU tensors `[B,H,N,d]` of the context's dtype; lse: pinned fp32 `[B,H,N]` or None.
        Enqueues one forward and returns; call synchronize() before reading o_host / lse_host."""
        shape = (self.B, self.H, self.N, self.d)
        for t in (q_host, k_host, v_host, o_host):
            if t.is_cuda or not t.is_pinned() or not t.is_contiguous() or tuple(t.shape) != shape or t.dtype != self.dtype:
                raise ValueError("HostPipeline expects contiguous pinned host tensors [B,H,N,d] of the context's dtype")
        if lse_host is not None and (lse_host.is_cuda or not lse_host.is_pinned() or not lse_host.is_contiguous()
                                     or lse_host.dtype != torch.float32 or tuple(lse_host.shape) != shape[:3]):
            raise ValueError("lse_host must be a contiguous pinned fp32 [B,H,N] tensor")
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().fa_b200_forward_host(self._ctx, q_host.data_ptr(), k_host.data_ptr(),
                                                        v_host.data_ptr(), o_host.data_ptr(),
                                                        lse_host.data_ptr() if lse_host is not None else None))
        return o_host, lse_host

    def synchronize(self):
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().fa_b200_host_ctx_sync(self._ctx))

    def elapsed_ms_last_call(self) -> float:
        """Device time between the first H2D copy and the end of the last D2H copy of the most recent call
        (CUDA events recorded inside fa_b200_forward_host); call after synchronize()."""
        ms = ctypes.c_float(0.0)
        _lib.check(_lib.load().fa_b200_host_ctx_elapsed_ms(self._ctx, ctypes.byref(ms)))
        return float(ms.value)

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            _lib.load().fa_b200_host_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
