#!/usr/bin/env python
"""Bring-up check of the backward kernels against the CPU oracle (numpy float64): prints the error of dQ, dK, dV
relative to max|ref| for a list of shapes; non-fatal, so that one run shows which product is wrong.  Test tool."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flash_attention_impls_b200 as fa  # noqa: E402
from oracle import oracle  # noqa: E402

SHAPES = [  # B, H, N, d, dtype, causal
    (1, 1, 128, 64, torch.float16, False),
    (1, 1, 128, 128, torch.bfloat16, False),
    (1, 2, 256, 128, torch.bfloat16, False),
    (1, 2, 256, 128, torch.bfloat16, True),
    (1, 2, 200, 64, torch.float16, True),
    (2, 3, 333, 32, torch.bfloat16, False),
    (1, 2, 1024, 128, torch.bfloat16, True),
]
if len(sys.argv) > 1:
    SHAPES = SHAPES[:int(sys.argv[1])]
dev = torch.device("cuda:0")
bad = 0
for (B, H, N, d, dtype, causal) in SHAPES:
    q, k, v = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(51, 52, 53))
    do, _, _ = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(54, 55, 56))
    tq, tk, tv, tdo = (torch.from_numpy(x).to(dev, dtype) for x in (q, k, v, do))
    o, lse = fa.attention_forward(tq, tk, tv, causal=causal)
    try:
        dq, dk, dv = fa.attention_backward(tq, tk, tv, o, lse, tdo, causal=causal)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"SHAPE {(B, H, N, d, str(dtype), causal)}: EXCEPTION {e}", flush=True)
        bad += 1
        break
    rq, rk, rv, _ = oracle.attention_backward_f64(q, k, v, do, causal=causal)
    errs = []
    for got, ref in ((dq, rq), (dk, rk), (dv, rv)):
        g = got.float().cpu().numpy()
        errs.append(float(np.abs(g - ref).max() / max(1e-9, np.abs(ref).max())))
    ok = all(e <= 2e-2 for e in errs) and all(torch.isfinite(t.float()).all().item() for t in (dq, dk, dv))
    bad += 0 if ok else 1
    print(f"SHAPE {(B, H, N, d, str(dtype).split('.')[-1], causal)}: rel err dQ {errs[0]:.3e} dK {errs[1]:.3e} dV {errs[2]:.3e} "
          f"{'PASS' if ok else 'FAIL'}", flush=True)
print("BWD_CHECK", "OK" if bad == 0 else f"{bad} FAILED")
