"""Generates tests/golden/sdpa_golden.npz by running the REFERENCE's own Python oracle
`sdpa_reference` (/root/reference/code/triton_fa2/FA2-triton.py:311-323) on CPU in the build container.
The reference tree does not travel to the GPU box, so the vectors are committed; re-run this script only
where /root/reference exists:   python tests/golden/make_golden.py
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
    # name, B, H, N, d, causal, set
    ("c1_fp32_noncausal", 1, 1, 128, 64, False, "S"),     # BASELINE config 1
    ("causal_d64", 1, 2, 192, 64, True, "S"),
    ("causal_ragged_d128", 1, 1, 160, 128, True, "S"),
    ("noncausal_ragged_d128", 2, 1, 100, 128, False, "S"),
    ("setR_d64", 1, 2, 128, 64, False, "R"),              # the reference's own input distribution
]
out = {}
for name, B, H, N, d, causal, which in CASES:
    if which == "S":
        q, k, v = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(11, 12, 13))
    else:
        q, k, v = oracle.set_r((B, H, N, d))
    # inputs are stored as fp16 (exact for Set S, which is bf16-representable inside fp16's range;
    # Set R is rounded to fp16 first, which is what the reference feeds its kernels, main.cu:53-56)
    q, k, v = (x.astype(np.float16) for x in (q, k, v))
    assert all(np.array_equal(x.astype(np.float32).astype(np.float16), x) for x in (q, k, v))
    q32, k32, v32 = (x.astype(np.float32) for x in (q, k, v))
    o = mod.sdpa_reference(torch.from_numpy(q32), torch.from_numpy(k32), torch.from_numpy(v32), causal=causal)
    out[name + "/q"], out[name + "/k"], out[name + "/v"] = q, k, v
    out[name + "/o"] = o.numpy().astype(np.float32)
    out[name + "/causal"] = np.array(int(causal))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sdpa_golden.npz"), **out)
print("wrote", len(CASES), "cases;", os.path.getsize(os.path.join(ROOT, "tests", "golden", "sdpa_golden.npz")), "bytes")
