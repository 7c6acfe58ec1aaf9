#!/usr/bin/env python
"""Times fa_b200_backward (delta pre-pass + dQ kernel + dK/dV kernel) with CUDA events at BASELINE's c3/c4/c2 shapes.
FLOPs: 7 tile products as executed would be 14*B*H*N^2*d; the usual convention counts the 5 a fused backward needs,
10*B*H*N^2*d (2.5x the forward), halved for causal - that is what is printed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flash_attention_impls_b200 as fa  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [(4, 32, 8192, 128, torch.bfloat16, False), (4, 32, 8192, 128, torch.bfloat16, True),
          (8, 16, 1024, 64, torch.float16, False)]
if "--one" in sys.argv:          # short run for an ncu capture: the c3 shape on a quarter of the heads
    SHAPES = [(1, 32, 8192, 128, torch.bfloat16, False)]
for (B, H, N, d, dtype, causal) in SHAPES:
    g = torch.Generator(device=dev); g.manual_seed(1)
    q, k, v, do = (torch.randn((B, H, N, d), generator=g, device=dev).to(dtype) for _ in range(4))
    if "--zeros" in sys.argv:    # cycle-domain timing: zero operands keep the GPU under its power limit, clock at max
        q, k, v, do = (torch.zeros_like(t) for t in (q, k, v, do))
    o, lse = fa.attention_forward(q, k, v, causal=causal)
    for _ in range(1 if "--one" in sys.argv else 3):
        fa.attention_backward(q, k, v, o, lse, do, causal=causal)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 2 if "--one" in sys.argv else 10
    e0.record()
    for _ in range(n):
        fa.attention_backward(q, k, v, o, lse, do, causal=causal)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 10.0 * B * H * N * N * d / (2 if causal else 1)
    print(f"BWD_TIMING B={B} H={H} N={N} d={d} {str(dtype).split('.')[-1]} causal={int(causal)}: {ms:.3f} ms "
          f"-> {fl / ms * 1e-9:.1f} TFLOP/s (5-product convention){' [zero inputs]' if '--zeros' in sys.argv else ''}", flush=True)
