#!/usr/bin/env python
"""cuDNN SDPA (torch) and this repo's forward on zero inputs vs random inputs: separates the cycle count of a kernel
(zero inputs: the SM clock stays at its maximum) from the power-management equilibrium it reaches on real data."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel
import flash_attention_impls_b200 as fa

def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    per = sorted(a.elapsed_time(b) for a, b in ev)
    return per[len(per) // 2], per[0]

dev = torch.device("cuda:0")
for (B, H, N, d, causal) in ((4, 32, 8192, 128, False), (4, 32, 8192, 128, True), (4, 32, 8192, 64, False)):
    fl = 4.0 * B * H * N * N * d / (2 if causal else 1)
    for kind in ("zeros", "randn"):
        mk = (lambda: torch.zeros(B, H, N, d, device=dev, dtype=torch.bfloat16)) if kind == "zeros" else \
             (lambda: torch.randn(B, H, N, d, device=dev, dtype=torch.bfloat16))
        q, k, v = mk(), mk(), mk()
        with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
            med, mn = t(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=causal))
        med2, mn2 = t(lambda: fa.attention_forward(q, k, v, causal=causal))
        print(f"ZERO_VS_RANDOM B={B} H={H} N={N} d={d} causal={int(causal)} {kind}: cuDNN {fl/med*1e-9:7.1f} TFLOP/s median "
              f"({fl/mn*1e-9:7.1f} best) | this repo {fl/med2*1e-9:7.1f} ({fl/mn2*1e-9:7.1f} best)", flush=True)
