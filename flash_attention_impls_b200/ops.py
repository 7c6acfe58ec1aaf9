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
                      m: Optional[torch.Tensor] = None, return_lse: bool = True):
    """O = softmax(Q K^T * scale [+ causal mask]) V on [B,H,N,d] tensors; returns (O, lse).

    q: [B,H,N,d]; k, v: [B,H,N_kv,d]; fp16 or bf16, last dim contiguous, rows dense (stride d); the
    (b,h) slices may be strided views of a longer sequence (what the ring driver passes).
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

    def bh_stride(t: torch.Tensor, rows: int, what: str) -> int:
        if t.stride(3) != 1 or t.stride(2) != t.shape[3]:
            raise ValueError(f"{what}: rows must be dense (stride(-1)=1, stride(-2)=d); call .contiguous()")
        if H > 1 and B > 1 and t.stride(0) != t.stride(1) * H:
            raise ValueError(f"{what}: batch and head strides must collapse to one (b*H+h) stride")
        return t.stride(1) if H > 1 else (t.stride(0) if B > 1 else rows * t.shape[3])

    qs, ks, vs = bh_stride(q, N, "q"), bh_stride(k, Nkv, "k"), bh_stride(v, Nkv, "v")
    if ks != vs:
        raise ValueError("k and v must share their (b,h) stride")
    if out is None:
        out = torch.empty((B, H, N, d), dtype=q.dtype, device=q.device)
    os_ = bh_stride(out, N, "out")
    if lse is None and return_lse:
        lse = torch.empty((B, H, N), dtype=torch.float32, device=q.device)
    ss = 0
    for s in (lse, l, m):
        if s is not None:
            if s.dtype != torch.float32 or s.shape != (B, H, N) or s.stride(2) != 1:
                raise ValueError("lse / l / m must be fp32 [B,H,N] with dense rows")
            st = s.stride(1) if H > 1 else (s.stride(0) if B > 1 else N)
            if ss and st != ss:
                raise ValueError("lse, l and m must share their (b,h) stride")
            ss = st

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
    p.q_stride_bh, p.kv_stride_bh, p.o_stride_bh, p.stat_stride_bh = qs, ks, os_, ss
    p.stream = _stream_ptr(q)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().fa_b200_forward(ctypes.byref(p)))
    return out, lse


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False) -> torch.Tensor:
    """Drop-in for the reference's `flash_attention(q, k, v, causal)` (FA2-triton.py:240-244).

    Like the reference, fp32 inputs are down-cast to fp16 for the kernel and the result is cast back.
    """
    _require_cuda(q, k, v)
    orig = q.dtype
    if orig == torch.float32:
        q, k, v = q.half(), k.half(), v.half()
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

    The (b,h) slices are independent, so the call is pipelined over `chunks` groups of slices on three
    streams (H2D, compute, D2H) with double-buffered device staging: while chunk c runs on the tensor
    cores, chunk c+1 is crossing PCIe and chunk c-1's output is going back.  This is the call bench.py's
    `e2e` figure times.  Device staging is allocated once here, never per call.
    """

    def __init__(self, B: int, H: int, N: int, d: int, dtype: torch.dtype, causal: bool = False,
                 chunks: int = 8, device=None):
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.BH, self.N, self.d, self.dtype, self.causal = B * H, N, d, dtype, causal
        chunks = max(1, min(chunks, self.BH))
        base, rem = divmod(self.BH, chunks)
        self.ranges, s = [], 0
        for c in range(chunks):
            n = base + (1 if c < rem else 0)
            self.ranges.append((s, s + n))
            s += n
        cmax = base + (1 if rem else 0)
        with torch.cuda.device(self.device):
            mk = lambda *shape, dt=dtype: torch.empty(shape, dtype=dt, device=self.device)
            self.slots = [dict(q=mk(1, cmax, N, d), k=mk(1, cmax, N, d), v=mk(1, cmax, N, d), o=mk(1, cmax, N, d),
                               lse=mk(1, cmax, N, dt=torch.float32)) for _ in range(2)]
            self.s_h2d, self.s_comp, self.s_d2h = (torch.cuda.Stream(self.device) for _ in range(3))
            self.ev_in_free = [None, None]    # compute of the chunk that last used the slot's inputs
            self.ev_out_free = [None, None]   # D2H of the chunk that last used the slot's outputs

    def bytes_per_call(self) -> Tuple[int, int]:
        e = torch.empty((), dtype=self.dtype).element_size()
        n = self.BH * self.N * self.d
        return 3 * n * e, n * e + self.BH * self.N * 4

    def __call__(self, q_host, k_host, v_host, o_host, lse_host):
        """All five are pinned CPU tensors: q,k,v,o `[B,H,N,d]` dtype, lse `[B,H,N]` fp32."""
        for t in (q_host, k_host, v_host, o_host, lse_host):
            if t.is_cuda or not t.is_pinned():
                raise ValueError("HostPipeline expects pinned host tensors")
        N, d = self.N, self.d
        qh, kh, vh, oh = (t.view(1, self.BH, N, d) for t in (q_host, k_host, v_host, o_host))
        lh = lse_host.view(1, self.BH, N)
        with torch.cuda.device(self.device):
            for c, (b, e) in enumerate(self.ranges):
                n, slot = e - b, self.slots[c % 2]
                with torch.cuda.stream(self.s_h2d):
                    if self.ev_in_free[c % 2] is not None:
                        self.s_h2d.wait_event(self.ev_in_free[c % 2])
                    slot["q"][:, :n].copy_(qh[:, b:e], non_blocking=True)
                    slot["k"][:, :n].copy_(kh[:, b:e], non_blocking=True)
                    slot["v"][:, :n].copy_(vh[:, b:e], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(self.s_h2d)
                with torch.cuda.stream(self.s_comp):
                    self.s_comp.wait_event(ev_in)
                    if self.ev_out_free[c % 2] is not None:
                        self.s_comp.wait_event(self.ev_out_free[c % 2])
                    attention_forward(slot["q"][:, :n], slot["k"][:, :n], slot["v"][:, :n], causal=self.causal,
                                      out=slot["o"][:, :n], lse=slot["lse"][:, :n])
                    ev_c = torch.cuda.Event()
                    ev_c.record(self.s_comp)
                    self.ev_in_free[c % 2] = ev_c
                with torch.cuda.stream(self.s_d2h):
                    self.s_d2h.wait_event(ev_c)
                    oh[:, b:e].copy_(slot["o"][:, :n], non_blocking=True)
                    lh[:, b:e].copy_(slot["lse"][:, :n], non_blocking=True)
                    ev_o = torch.cuda.Event()
                    ev_o.record(self.s_d2h)
                    self.ev_out_free[c % 2] = ev_o
        return o_host, lse_host

    def synchronize(self):
        self.s_d2h.synchronize()
