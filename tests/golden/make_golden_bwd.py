"""Generates tests/golden/sdpa_bwd_golden.npz: gradients of the REFERENCE's own Python oracle `sdpa_reference`
(/root/reference/code/triton_fa2/FA2-triton.py:311-323) obtained with torch.autograd on CPU in the build container,
for a seeded upstream gradient dO.  These pin oracle.attention_backward_f64 (the reference's Triton backward kernel is
never checked by the reference and does not implement the softmax Jacobian, see the oracle's docstring).
Re-run only where /root/reference exists:   python tests/golden/make_golden_bwd.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = "/root/reference/code/triton_fa2/FA2-triton.py"
spec = importlib.util.spec_from_file_location("fa2_triton_ref", REF)   # hyphen in the file name
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

CASES = [
    # name, B, H, N, d, causal
    ("bwd_noncausal_d64", 1, 2, 128, 64, False),
    ("bwd_causal_d64", 1, 2, 192, 64, True),
    ("bwd_causal_ragged_d128", 1, 1, 160, 128, True),
    ("bwd_noncausal_ragged_d32", 2, 1, 100, 32, False),
]
out = {}
for name, B, H, N, d, causal in CASES:
    q, k, v = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(31, 32, 33))
    do, _, _ = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(34, 35, 36))
    q, k, v, do = (x.astype(np.float16) for x in (q, k, v, do))
    tq, tk, tv = (torch.from_numpy(x.astype(np.float32)).requires_grad_(True) for x in (q, k, v))
    o = mod.sdpa_reference(tq, tk, tv, causal=causal)
    o.backward(torch.from_numpy(do.astype(np.float32)))
    for key, val in (("q", q), ("k", k), ("v", v), ("do", do)):
        out[f"{name}/{key}"] = val
    out[f"{name}/dq"] = tq.grad.numpy().astype(np.float32)
    out[f"{name}/dk"] = tk.grad.numpy().astype(np.float32)
    out[f"{name}/dv"] = tv.grad.numpy().astype(np.float32)
    out[f"{name}/causal"] = np.array(int(causal))
path = os.path.join(ROOT, "tests", "golden", "sdpa_bwd_golden.npz")
np.savez_compressed(path, **out)
print("wrote", len(CASES), "cases;", os.path.getsize(path), "bytes")
