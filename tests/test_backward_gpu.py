"""GPU parity tests of the backward pass: libfa_b200.so's fa_b200_backward, through the
ctypes binding, against the CPU oracle (numpy float64, pinned against torch.autograd through the reference's own
`sdpa_reference`), against the committed golden gradients, and through the autograd mirror of the reference's
`_FlashAttnFn` (FA2-triton.py:173-237).

Tolerance: the kernels feed P and dS to the tensor cores as 16-bit operands and emit 16-bit gradients, so every
gradient is compared relative to the largest magnitude of its reference tensor: <= 3e-3 for fp16 (11-bit
significand), <= 1.5e-2 for bf16 (8-bit significand).  Measured: 5e-4 and 5e-3.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {torch.float16: 3e-3, torch.bfloat16: 1.5e-2}


@pytest.fixture(scope="module")
def fa():
    assert torch.cuda.is_available()
    import flash_attention_impls_b200 as fa
    fa.load()
    return fa


def _rel(got, ref):
    # a reference that is identically zero (dQ, dK of a one-row causal problem) is compared on an absolute 1e-4 scale
    return float(np.abs(got.float().cpu().numpy() - ref).max() / max(1e-4, np.abs(ref).max()))


SHAPES = [
    # B, H, N, d, dtype, causal
    (1, 1, 128, 64, torch.float16, False),
    (1, 1, 1, 64, torch.float16, True),           # one row: dQ = 0, dV = dO, dK = 0
    (1, 2, 127, 128, torch.bfloat16, True),       # one short of a tile
    (1, 2, 129, 128, torch.float16, False),       # one over a tile
    (2, 3, 333, 32, torch.bfloat16, False),
    (1, 2, 200, 64, torch.float16, True),
    (1, 2, 640, 128, torch.bfloat16, True),       # 5 tiles: dQ kernel visits 1..5 K/V tiles, dK/dV kernel 5..1 Q tiles
    (3, 50, 300, 64, torch.bfloat16, False),      # 450 CTAs per kernel: more than one wave
    (1, 2, 1024, 128, torch.float16, False),
]


@pytest.mark.parametrize("B,H,N,d,dtype,causal", SHAPES)
def test_backward_matches_oracle(fa, B, H, N, d, dtype, causal):
    from oracle import oracle
    q, k, v = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(51, 52, 53))
    do, _, _ = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(54, 55, 56))
    dev = torch.device("cuda:0")
    tq, tk, tv, tdo = (torch.from_numpy(x).to(dev, dtype) for x in (q, k, v, do))
    o, lse = fa.attention_forward(tq, tk, tv, causal=causal)
    before = fa.launch_count()
    dq, dk, dv = fa.attention_backward(tq, tk, tv, o, lse, tdo, causal=causal)
    torch.cuda.synchronize()
    assert fa.launch_count() - before == 3          # delta pre-pass, dQ kernel, dK/dV kernel
    rq, rk, rv, _ = oracle.attention_backward_f64(q, k, v, do, causal=causal)
    for got, ref in ((dq, rq), (dk, rk), (dv, rv)):
        assert torch.isfinite(got.float()).all()
        assert _rel(got, ref) <= TOL[dtype]


@pytest.mark.parametrize("name", ["bwd_noncausal_d64", "bwd_causal_d64", "bwd_causal_ragged_d128", "bwd_noncausal_ragged_d32"])
def test_backward_matches_golden_gradients_of_reference_sdpa(fa, name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "sdpa_bwd_golden.npz"))
    dev = torch.device("cuda:0")
    tq, tk, tv, tdo = (torch.from_numpy(g[f"{name}/{x}"]).to(dev) for x in ("q", "k", "v", "do"))   # fp16
    causal = bool(g[f"{name}/causal"])
    o, lse = fa.attention_forward(tq, tk, tv, causal=causal)
    dq, dk, dv = fa.attention_backward(tq, tk, tv, o, lse, tdo, causal=causal)
    torch.cuda.synchronize()
    for got, key in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
        assert _rel(got, g[f"{name}/{key}"]) <= TOL[torch.float16]


def test_autograd_function_mirrors_the_references_flash_attn_fn(fa):
    """`flash_attention(q, k, v, causal)` with inputs that require grad goes through FlashAttnFunction, as the
    reference's goes through _FlashAttnFn (FA2-triton.py:240-244); fp32 inputs are cast to fp16 and back."""
    from oracle import oracle
    B, H, N, d = 1, 2, 384, 64
    q, k, v = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(61, 62, 63))
    do, _, _ = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(64, 65, 66))
    dev = torch.device("cuda:0")
    tq, tk, tv = (torch.from_numpy(x).to(dev).requires_grad_(True) for x in (q, k, v))     # fp32 leaves
    o = fa.flash_attention(tq, tk, tv, causal=True)
    assert o.dtype == torch.float32
    o.backward(torch.from_numpy(do).to(dev))
    rq, rk, rv, _ = oracle.attention_backward_f64(q, k, v, do, causal=True)
    for got, ref in ((tq.grad, rq), (tk.grad, rk), (tv.grad, rv)):
        assert got is not None and got.dtype == torch.float32
        assert _rel(got, ref) <= TOL[torch.float16]
    # no grad required -> plain forward, no graph
    with torch.no_grad():
        assert not fa.flash_attention(tq, tk, tv).requires_grad


def test_backward_is_deterministic_and_overwrites_its_outputs(fa):
    """One writer per output tile: two runs are bit-identical (the reference accumulates dK/dV with fp16 atomics,
    FA2-triton.py:164-167), and stale output contents do not leak into the result."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(5)
    B, H, N, d = 2, 4, 1000, 128
    q, k, v, do = (torch.randn((B, H, N, d), generator=g, device=dev).bfloat16() for _ in range(4))
    o, lse = fa.attention_forward(q, k, v, causal=True)
    a = fa.attention_backward(q, k, v, o, lse, do, causal=True)
    b = fa.attention_backward(q, k, v, o, lse, do, causal=True)
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    # causal: the last key is seen only by the last query -> dV[N-1] = P[N-1,N-1] * dO[N-1]
    assert torch.isfinite(a[2].float()).all()


def _elementwise(got, ref, rtol, atol_row):
    """max over elements of |got - ref| / (rtol*|ref| + atol_row*rms(row of ref)): <= 1 passes.  Element-wise, unlike
    `_rel` (relative to the largest magnitude of the whole tensor).  The absolute term scales with the element's own ROW:
    an output element is a sum over 8192 products whose 16-bit rounding noise is set by the magnitude of the row's terms,
    not by the (possibly cancelling) sum, and causal gradients span a 50x range of row magnitudes inside one tensor."""
    got = got.float().cpu().numpy().astype(np.float64)
    row = np.sqrt(np.mean(ref * ref, axis=-1, keepdims=True))
    floor = 1e-4 * float(np.sqrt(np.mean(ref * ref)))       # rows that are exactly zero (causal row 0 of dQ: got ~1e-7)
    return float((np.abs(got - ref) / (rtol * np.abs(ref) + atol_row * row + floor)).max())


def test_full_size_c4_backward_against_oracle(fa):
    """B=4 H=32 N=8192 d=128 bf16 causal (BASELINE c4's shape).  One whole (b,h) slice - all 8192 rows of dQ, dK AND
    dV - is checked against the float64 oracle (row-blocked, so it fits in memory) with an ELEMENT-WISE tolerance:
    |err| <= 2^-7*|ref| + 2e-2*rms(row of ref)  (one bf16 ulp of the element plus the rounding noise of the row's 8192
    bf16 products; a CPU emulation of the kernel's roundings - P, dS and the outputs in bf16 - reaches 0.42 of this
    bound, the kernel measured 0.6; the previous relative-to-max figure is printed beside it).  A second slice gets the prefix check: dQ of the first 512 queries of a causal problem
    only involves the first 512 keys.  Plus finiteness of the full result and bit-exact (b,h)-shard equivalence."""
    from oracle import oracle
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(9)
    B, H, N, d = 4, 32, 8192, 128
    q = torch.randn((B, H, N, d), generator=g, device=dev).bfloat16()
    k = torch.randn((B, H, N, d), generator=g, device=dev).bfloat16()
    v = (torch.rand((B, H, N, d), generator=g, device=dev) - 0.5).bfloat16()
    do = torch.randn((B, H, N, d), generator=g, device=dev).bfloat16()
    o, lse = fa.attention_forward(q, k, v, causal=True)
    dq, dk, dv = fa.attention_backward(q, k, v, o, lse, do, causal=True)
    torch.cuda.synchronize()
    for t in (dq, dk, dv):
        assert torch.isfinite(t.float()).all()
    # (1) one full slice, every row, all three gradients
    b, h = 3, 31
    qs, ks, vs, dos = (t[b:b + 1, h:h + 1].float().cpu().numpy() for t in (q, k, v, do))
    rq, rk, rv, _ = oracle.attention_backward_f64(qs, ks, vs, dos, causal=True, row_block=1024)
    # dQ rows near the start of a causal sequence are sums of nearly cancelling terms p_ij (dP_ij - delta_i), and delta is
    # formed from the forward's 16-bit O (the standard FlashAttention-2 pre-pass: rowsum(dO o O)), which puts
    # 2^-9 |dO.O| into every term: their row scale factor is 6e-2 (emulated 2e-2 x 0.67 x 3, measured 3.9e-2)
    worst = {n_: _elementwise(got[b:b + 1, h:h + 1], ref, 2.0 ** -7, at) for n_, got, ref, at in
             (("dq", dq, rq, 6e-2), ("dk", dk, rk, 2e-2), ("dv", dv, rv, 2e-2))}
    rel = {n_: _rel(got[b:b + 1, h:h + 1], ref) for n_, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv))}
    print("c4 backward vs oracle: element-wise ratio", worst, "relative-to-max", rel)
    assert max(worst.values()) <= 1.0, (worst, rel)
    assert max(rel.values()) <= TOL[torch.bfloat16]
    # (2) prefix check on another slice
    qs, ks, vs, dos = (t[0:1, 0:1, :512].float().cpu().numpy() for t in (q, k, v, do))
    rq, _, _, _ = oracle.attention_backward_f64(qs, ks, vs, dos, causal=True)
    assert _rel(dq[0:1, 0:1, :512], rq) <= TOL[torch.bfloat16]
    # (3) (b,h) shard equivalence: a slice computed alone is bit-identical
    qs, ks, vs, dos, os_, ls_ = (t[1:2, 5:9].contiguous() for t in (q, k, v, do, o, lse))
    dq2, dk2, dv2 = fa.attention_backward(qs, ks, vs, os_, ls_, dos, causal=True)
    assert torch.equal(dq2, dq[1:2, 5:9]) and torch.equal(dk2, dk[1:2, 5:9]) and torch.equal(dv2, dv[1:2, 5:9])


@pytest.mark.parametrize("B,H,N,Nkv,d,dtype,causal", [
    (1, 2, 300, 500, 64, torch.float16, False),      # N_kv > N
    (2, 2, 500, 300, 128, torch.bfloat16, False),    # N_kv < N
    (1, 3, 300, 500, 128, torch.bfloat16, True),     # causal, bottom-right aligned: every query sees >= 201 keys
    (1, 2, 500, 300, 64, torch.float16, True),       # causal with N_kv < N: the first 200 queries see NO key (lse = -inf)
    (1, 2, 128, 700, 32, torch.bfloat16, True),      # one query tile against many key tiles
    (1, 2, 700, 128, 128, torch.float16, True),      # many query tiles, one key tile; query tiles 0..3 see nothing
])
def test_backward_with_a_different_number_of_keys(fa, B, H, N, Nkv, d, dtype, causal):
    """N_kv != N with the forward's mask rule col > row + (N_kv - N): whole tiles that see nothing get zero gradients,
    rows with lse = -inf contribute nothing."""
    from oracle import oracle
    q, k, v = oracle.set_s((B, H, N, d), (B, H, Nkv, d), seeds=(91, 92, 93))
    do, _, _ = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(94, 95, 96))
    dev = torch.device("cuda:0")
    tq, tk, tv, tdo = (torch.from_numpy(x).to(dev, dtype) for x in (q, k, v, do))
    o, lse = fa.attention_forward(tq, tk, tv, causal=causal)
    dq, dk, dv = fa.attention_backward(tq, tk, tv, o, lse, tdo, causal=causal)
    torch.cuda.synchronize()
    rq, rk, rv, _ = oracle.attention_backward_f64(q, k, v, do, causal=causal)
    for got, ref in ((dq, rq), (dk, rk), (dv, rv)):
        assert torch.isfinite(got.float()).all()
        assert _rel(got, ref) <= TOL[dtype]
    if causal and Nkv < N:      # queries that see no key: exactly zero dQ rows
        assert torch.all(dq[:, :, :N - Nkv] == 0)


@pytest.mark.parametrize("d,dtype,causal", [(128, torch.bfloat16, True), (64, torch.float16, False)])
def test_backward_takes_strided_views_without_copies(fa, d, dtype, causal):
    """[B,N,H,d] storage viewed as [B,H,N,d] (row stride H*d, head stride d) for q, k, v, o and do - what the
    reference's backward accepts through `x.stride(i)` (FA2-triton.py:219-227) - gives bit-identical gradients to
    the dense layout, and goes through the autograd mirror as well."""
    from oracle import oracle
    B, H, N = 2, 3, 400
    q, k, v = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(101, 102, 103))
    do, _, _ = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(104, 105, 106))
    dev = torch.device("cuda:0")
    dense = [torch.from_numpy(x).to(dev, dtype) for x in (q, k, v, do)]
    views = [t.transpose(1, 2).contiguous().transpose(1, 2) for t in dense]       # same values, [B,N,H,d] storage
    assert not views[0].is_contiguous()
    o_d, lse_d = fa.attention_forward(dense[0], dense[1], dense[2], causal=causal)
    o_v = torch.empty((B, N, H, d), dtype=dtype, device=dev).transpose(1, 2)
    fa.attention_forward(views[0], views[1], views[2], causal=causal, out=o_v, lse=lse_d.clone())
    g_d = fa.attention_backward(dense[0], dense[1], dense[2], o_d, lse_d, dense[3], causal=causal)
    g_v = fa.attention_backward(views[0], views[1], views[2], o_v, lse_d, views[3], causal=causal)
    torch.cuda.synchronize()
    for a_, b_ in zip(g_d, g_v):
        assert torch.equal(a_, b_)
    rq, rk, rv, _ = oracle.attention_backward_f64(q, k, v, do, causal=causal)
    for got, ref in zip(g_v, (rq, rk, rv)):
        assert _rel(got, ref) <= TOL[dtype]
    # autograd on strided leaves
    leaves = [t.clone().requires_grad_(True) for t in views[:3]]
    out = fa.flash_attention(leaves[0], leaves[1], leaves[2], causal=causal)
    out.backward(views[3])
    for leaf, ref in zip(leaves, g_d):
        assert torch.equal(leaf.grad, ref)
