"""oracle/oracle.py — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liboracle.so (the C
restatement of the reference's naive attention, oracle/naive_attention.c) plus an independent numpy
float64 restatement used to cross-check the C code on small cases.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_size_t, c_uint32

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None
_fp = POINTER(c_float)


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        lib.oracle_attention_f32.argtypes = [_fp, _fp, _fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int,
                                             c_float, c_int, c_int, c_int]
        lib.oracle_attention_f32.restype = c_int
        lib.oracle_max_symmetric_rel_err.argtypes = [_fp, _fp, c_size_t]
        lib.oracle_max_symmetric_rel_err.restype = c_float
        lib.oracle_merge_partial.argtypes = [_fp, _fp, _fp, _fp, c_size_t, c_int]
        lib.oracle_merge_partial.restype = None
        lib.fixture_reference_stream.argtypes = [_fp, c_size_t, c_float, c_float]
        lib.fixture_normal_bf16.argtypes = [_fp, c_size_t, c_uint32, c_float]
        lib.fixture_uniform_bf16.argtypes = [_fp, c_size_t, c_uint32, c_float, c_float]
        _lib = lib
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(_fp)


def attention(q: np.ndarray, k: np.ndarray, v: np.ndarray, causal: bool = False, scale: float = 0.0,
              nthreads: int = 0, row_begin: int = 0, row_end: int = 0):
    """q [B,H,N,d], k/v [B,H,Nkv,d] float32 -> (O [B,H,N,d], lse, l, m [B,H,N]) float32.
    Rows outside [row_begin,row_end) are left zero."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(k, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    B, H, N, d = q.shape
    Nkv = k.shape[2]
    o = np.zeros((B, H, N, d), np.float32)
    lse = np.zeros((B, H, N), np.float32)
    l = np.zeros((B, H, N), np.float32)
    m = np.zeros((B, H, N), np.float32)
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    rc = load().oracle_attention_f32(_p(q), _p(k), _p(v), _p(o), _p(lse), _p(l), _p(m), B * H, N, Nkv, d,
                                     1 if causal else 0, scale, nthreads, row_begin, row_end)
    if rc:
        raise RuntimeError(f"oracle_attention_f32 failed: {rc}")
    return o, lse, l, m


def attention_f64(q, k, v, causal=False, scale=None):
    """Independent numpy float64 restatement of main.cu:165-202 (+ causal rule FA2-triton.py:70-73)."""
    q = np.asarray(q, np.float64); k = np.asarray(k, np.float64); v = np.asarray(v, np.float64)
    N, Nkv, d = q.shape[2], k.shape[2], q.shape[3]
    s = np.einsum("bhid,bhjd->bhij", q, k) * (scale if scale else 1.0 / math.sqrt(d))
    if causal:
        i = np.arange(N)[:, None]; j = np.arange(Nkv)[None, :]
        s = np.where(j > i + (Nkv - N), -np.inf, s)
    m = s.max(axis=-1)
    msafe = np.where(np.isfinite(m), m, 0.0)
    e = np.exp(s - msafe[..., None])
    l = e.sum(axis=-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        o = np.einsum("bhij,bhjd->bhid", e, v) / np.where(l > 0, l, 1.0)[..., None]
        lse = np.where(l > 0, msafe + np.log(np.where(l > 0, l, 1.0)), -np.inf)
    return o, lse, l, m


def attention_backward_f64(q, k, v, do, causal=False, scale=None, row_block=0):
    """Gradients of O = softmax(Q K^T scale [+ causal mask]) V w.r.t. Q, K, V for an upstream gradient dO, in numpy
    float64, per (b,h) slice to bound memory.  This is the mathematics the reference's Triton backward recomputes
    (FA2-triton.py:98-170: P from the saved row statistics, dV = P^T dO, dP = dO V^T, dS, dQ = dS K, dK = dS^T Q) with
    the standard softmax Jacobian
        dS_ij = scale * P_ij * (dP_ij - sum_k P_ik dP_ik),     sum_k P_ik dP_ik = dO_i . O_i
    The reference kernel itself forms `(dP - sum_block(dP*P) * P) * scale` with the sum taken over one 128-key block
    (FA2-triton.py:158-159), which is not this Jacobian, and the reference never checks its backward;
    the pin for this function is therefore torch.autograd through the reference's own `sdpa_reference`
    (tests/golden/make_golden_bwd.py -> tests/golden/sdpa_bwd_golden.npz).
    `row_block` > 0 processes the query rows in blocks of that many rows (same result; bounds memory at long N).
    Returns (dQ, dK, dV, delta) as float64, delta[b,h,i] = dO_i . O_i."""
    q = np.asarray(q, np.float64); k = np.asarray(k, np.float64); v = np.asarray(v, np.float64)
    do = np.asarray(do, np.float64)
    B, H, N, d = q.shape
    Nkv = k.shape[2]
    sc = scale if scale else 1.0 / math.sqrt(d)
    dq = np.zeros_like(q); dk = np.zeros_like(k); dv = np.zeros_like(v)
    delta = np.zeros((B, H, N))
    rb = row_block if row_block else N          # query rows are independent: blocks of rows bound the N x Nkv temporaries
    j = np.arange(Nkv)[None, :]
    for b in range(B):
        for h in range(H):
            for r0 in range(0, N, rb):
                r1 = min(N, r0 + rb)
                s = (q[b, h, r0:r1] @ k[b, h].T) * sc
                if causal:
                    s = np.where(j > np.arange(r0, r1)[:, None] + (Nkv - N), -np.inf, s)
                m = s.max(axis=-1, keepdims=True)
                msafe = np.where(np.isfinite(m), m, 0.0)
                e = np.exp(s - msafe)
                l = e.sum(axis=-1, keepdims=True)
                p = e / np.where(l > 0, l, 1.0)
                o = p @ v[b, h]
                dp = do[b, h, r0:r1] @ v[b, h].T
                dl = (do[b, h, r0:r1] * o).sum(axis=-1, keepdims=True)
                ds = p * (dp - dl) * sc
                dq[b, h, r0:r1] = ds @ k[b, h]
                dk[b, h] += ds.T @ q[b, h, r0:r1]
                dv[b, h] += p.T @ do[b, h, r0:r1]
                delta[b, h, r0:r1] = dl[:, 0]
    return dq, dk, dv, delta


def merge_partial(o_a, lse_a, o_b, lse_b):
    o_a = np.ascontiguousarray(o_a, np.float32).copy(); lse_a = np.ascontiguousarray(lse_a, np.float32).copy()
    o_b = np.ascontiguousarray(o_b, np.float32); lse_b = np.ascontiguousarray(lse_b, np.float32)
    d = o_a.shape[-1]
    load().oracle_merge_partial(_p(o_a), _p(lse_a), _p(o_b), _p(lse_b), o_a.size // d, d)
    return o_a, lse_a


def symmetric_rel_err(a, b) -> float:
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return float(load().oracle_max_symmetric_rel_err(_p(a), _p(b), a.size))


def reference_stream(n: int, mean: float = 0.0, stddev: float = 0.02) -> np.ndarray:
    """Set R: the reference's mt19937(42) / N(0,0.02) stream (main.cu:43-61)."""
    out = np.empty(n, np.float32)
    load().fixture_reference_stream(_p(out), n, mean, stddev)
    return out


def set_s(shape_q, shape_kv, seeds=(1, 2, 3)):
    """Set S: Q,K ~ N(0,1), V ~ U(-0.5,0.5), bf16-representable."""
    nq, nk = int(np.prod(shape_q)), int(np.prod(shape_kv))
    q = np.empty(nq, np.float32); k = np.empty(nk, np.float32); v = np.empty(nk, np.float32)
    lib = load()
    lib.fixture_normal_bf16(_p(q), nq, seeds[0], 1.0)
    lib.fixture_normal_bf16(_p(k), nk, seeds[1], 1.0)
    lib.fixture_uniform_bf16(_p(v), nk, seeds[2], -0.5, 0.5)
    return q.reshape(shape_q), k.reshape(shape_kv), v.reshape(shape_kv)


def set_r(shape):
    """Set R tensors: Q == K == V (fresh generator per tensor in the reference)."""
    x = reference_stream(int(np.prod(shape))).reshape(shape)
    return x, x.copy(), x.copy()
