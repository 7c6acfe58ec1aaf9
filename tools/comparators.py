#!/usr/bin/env python
"""Same-box comparison of the forward kernel with the Blackwell attention kernels already on the box
(SURVEY.md section 2.4 "the honest bar").

  python tools/comparators.py [--iters 20] [--out gpurun_out/comparators]

Arms, all on the same synthetic Set-S inputs, same shapes (BASELINE c2 / c3 / c4 plus fp16 variants), CUDA events
on the launching stream, 3 warm-ups, L2 flushed between iterations when the inputs fit in it:
  * this repo            fa.attention_forward -> libfa_b200.so
  * torch SDPA, cuDNN    torch.nn.functional.scaled_dot_product_attention under SDPBackend.CUDNN_ATTENTION
  * torch SDPA, flash    ... under SDPBackend.FLASH_ATTENTION (torch's bundled FlashAttention-2, mma.sync)
  * flash_attn           flash_attn.flash_attn_func (the pip package in the image; layout [B,N,H,d])
  * CUTLASS example 77   examples/77_blackwell_fmha from the reference's vendored CUTLASS 4.3.0 tree, built as an
                         EXTERNAL binary by `make -C oracle ref_cutlass77` (oracle/_ref/cutlass77_fmha_fp16; fp16
                         only; timed by its own harness, TFLOPS/s parsed from its output)
Library kernels are comparators, never part of the product path.  An arm that cannot run here says why.
Writes <out>.json and <out>.md.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

SHAPES = [
    # name, B, H, N, d, dtype, causal
    ("c2", 8, 16, 1024, 64, "fp16", False),
    ("c3", 4, 32, 8192, 128, "bf16", False),
    ("c3-fp16", 4, 32, 8192, 128, "fp16", False),
    ("c4", 4, 32, 8192, 128, "bf16", True),
    ("c4-fp16", 4, 32, 8192, 128, "fp16", True),
    ("d64-8k", 4, 32, 8192, 64, "bf16", False),
]


def flops(B, H, N, d, causal):
    f = 4.0 * B * H * N * N * d
    return f / 2 if causal else f


def time_fn(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        if flush is not None:
            flush.zero_()
        e0[i].record()
        fn()
        e1[i].record()
    torch.cuda.synchronize()
    per = sorted(a.elapsed_time(b) for a, b in zip(e0, e1))
    total = e0[0].elapsed_time(e1[-1]) / iters if flush is None else sum(per) / iters
    return total, per[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "comparators"))
    ap.add_argument("--shapes", default="")
    args = ap.parse_args()
    import flash_attention_impls_b200 as fa
    from torch.nn.attention import SDPBackend, sdpa_kernel
    dev = torch.device("cuda", 0)
    results = []
    cutlass_bin = os.path.join(ROOT, "oracle", "_ref", "cutlass77_fmha_fp16")
    want = set(filter(None, args.shapes.split(",")))
    for name, B, H, N, d, dt, causal in SHAPES:
        if want and name not in want:
            continue
        dtype = torch.bfloat16 if dt == "bf16" else torch.float16
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        q = torch.randn((B, H, N, d), generator=g, device=dev).to(dtype)
        k = torch.randn((B, H, N, d), generator=g, device=dev).to(dtype)
        v = (torch.rand((B, H, N, d), generator=g, device=dev) - 0.5).to(dtype)
        o = torch.empty_like(q)
        lse = torch.empty((B, H, N), dtype=torch.float32, device=dev)
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if 3 * q.numel() * 2 < (256 << 20) else None
        fl = flops(B, H, N, d, causal)
        row = {"shape": name, "B": B, "H": H, "N": N, "d": d, "dtype": dt, "causal": causal, "arms": {}}

        def record(arm, fn, ref=None):
            try:
                ms, ms_min = time_fn(fn, args.iters, flush)
                rec = {"ms": ms, "ms_min": ms_min, "tflops": fl / (ms * 1e-3) * 1e-12}
                if ref is not None:
                    rec["max_abs_vs_this_repo"] = float((ref().float() - o.float()).abs().max())
                row["arms"][arm] = rec
            except Exception as e:  # an arm that cannot run on this box is reported, not hidden
                row["arms"][arm] = {"unavailable": (type(e).__name__ + ": " + str(e).splitlines()[0])[:200]}
            torch.cuda.synchronize()

        record("this repo (libfa_b200.so)", lambda: fa.attention_forward(q, k, v, causal=causal, out=o, lse=lse))
        for arm, backend in (("torch SDPA cuDNN", SDPBackend.CUDNN_ATTENTION), ("torch SDPA flash (FA2)", SDPBackend.FLASH_ATTENTION)):
            def sdpa(backend=backend):
                with sdpa_kernel(backend):
                    return F.scaled_dot_product_attention(q, k, v, is_causal=causal)
            record(arm, sdpa, ref=sdpa)
        try:
            from flash_attn import flash_attn_func
            qn, kn, vn = (t.transpose(1, 2).contiguous() for t in (q, k, v))     # [B,N,H,d]
            record("flash_attn %s" % __import__("flash_attn").__version__, lambda: flash_attn_func(qn, kn, vn, causal=causal),
                   ref=lambda: flash_attn_func(qn, kn, vn, causal=causal).transpose(1, 2))
        except Exception as e:
            row["arms"]["flash_attn"] = {"unavailable": (type(e).__name__ + ": " + str(e).splitlines()[0])[:200]}
        if dt == "fp16":
            if os.path.exists(cutlass_bin):
                try:
                    cmd = [cutlass_bin, f"--b={B}", f"--h={H}", f"--q={N}", f"--k={N}", f"--d={d}",
                           f"--mask={'causal' if causal else 'no'}", f"--iterations={args.iters}", "--warmup_iterations=3"]
                    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300).stdout
                    best = None
                    for ln in out.splitlines():
                        m = re.search(r"^\s*\[(OK|--)\]\s+(.*?)\s*:\s*([0-9.]+) TFLOPS/s", ln)
                        if m and (best is None or float(m.group(3)) > best[1]):
                            best = (m.group(2).strip(), float(m.group(3)))
                    if best:
                        row["arms"]["CUTLASS ex77 (external binary)"] = {"tflops": best[1], "ms": fl / best[1] * 1e-9, "variant": best[0],
                                                                          "timer": "the example's own harness"}
                    else:
                        row["arms"]["CUTLASS ex77 (external binary)"] = {"unavailable": "no result line: " + out[-200:]}
                except Exception as e:
                    row["arms"]["CUTLASS ex77 (external binary)"] = {"unavailable": str(e)[:200]}
            else:
                row["arms"]["CUTLASS ex77 (external binary)"] = {"unavailable": "oracle/_ref/cutlass77_fmha_fp16 not built"}
        results.append(row)
        del q, k, v, o, lse
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out + ".json", "w") as f:
        json.dump({"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
                   "iters": args.iters, "results": results}, f, indent=1)
    lines = ["| shape | arm | ms | TFLOP/s | vs this repo | max-abs diff vs this repo |", "|---|---|---|---|---|---|"]
    for row in results:
        mine = row["arms"].get("this repo (libfa_b200.so)", {}).get("tflops")
        tag = f"{row['shape']} `{row['B']},{row['H']},{row['N']},{row['d']}` {row['dtype']}{' causal' if row['causal'] else ''}"
        for arm, r in row["arms"].items():
            if "unavailable" in r:
                lines.append(f"| {tag} | {arm} | - | - | - | unavailable: {r['unavailable']} |")
            else:
                rel = f"{r['tflops'] / mine:.2f}x" if mine else "-"
                diff = f"{r['max_abs_vs_this_repo']:.1e}" if "max_abs_vs_this_repo" in r else "-"
                lines.append(f"| {tag} | {arm} | {r['ms']:.3f} | {r['tflops']:.0f} | {rel} | {diff} |")
    with open(args.out + ".md", "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
