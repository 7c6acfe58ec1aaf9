#!/usr/bin/env python
"""Ring-attention evaluation under ONE torchrun process group (saves GPU-box time over several launches):

  1. parity of both transports ("peer" = CUDA-IPC publish buffers + copy-engine pulls, "p2p" = NCCL send/recv)
     against the CPU oracle on a small sequence, causal and non-causal, three calls each (the peer transport
     double-buffers its publish slots across calls);
  2. timing of the c5 shape (B=1 H=32 N=131072 d=128 bf16 causal, or --n-local rows per rank) per transport:
     CUDA events on the launching stream, barrier on both sides, max over ranks;
  3. a per-step CUDA-event timeline of one call per transport (FA_B200_RING_PROFILE).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/ring_eval.py

Test/diagnostic tool: it may use the oracle (as the checker), the package never does.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flash_attention_impls_b200 as fa  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-total", type=int, default=131072)
    ap.add_argument("--n-local", type=int, default=0, help="rows per rank (overrides --n-total)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--transports", default="peer,p2p")
    ap.add_argument("--skip-parity", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    transports = [t for t in args.transports.split(",") if t]
    out = {"world": world, "parity": {}, "timing": {}}

    # ---- 1. parity vs the oracle (rank 0 computes the oracle, every rank checks its own shard)
    if not args.skip_parity:
        from oracle import oracle
        N, d, Hh = 256 * 2 * world, 128, 4
        q, k, v = oracle.set_s((1, Hh, N, d), (1, Hh, N, d), seeds=(41, 42, 43))
        tq, tk, tv = (torch.from_numpy(x).to(dev, torch.bfloat16) for x in (q, k, v))
        for causal in (False, True):
            o_ref, lse_ref, _, _ = oracle.attention(q, k, v, causal=causal)
            if causal:
                ql, kl, vl = (fa.zigzag_split(t, world, rank) for t in (tq, tk, tv))
                o_ref_l = fa.zigzag_split(torch.from_numpy(o_ref), world, rank).numpy()
                lse_ref_l = fa.zigzag_split(torch.from_numpy(lse_ref), world, rank).numpy()
            else:
                c = N // world
                ql, kl, vl = (t[:, :, rank * c:(rank + 1) * c].contiguous() for t in (tq, tk, tv))
                o_ref_l, lse_ref_l = o_ref[:, :, rank * c:(rank + 1) * c], lse_ref[:, :, rank * c:(rank + 1) * c]
            for tr in transports:
                worst_o, worst_l = 0.0, 0.0
                for _ in range(3):
                    o, lse = fa.ring_attention(ql, kl, vl, causal=causal, transport=tr)
                    torch.cuda.synchronize()
                    worst_o = max(worst_o, float(np.abs(o.float().cpu().numpy() - o_ref_l).max()))
                    worst_l = max(worst_l, float((np.abs(lse.cpu().numpy() - lse_ref_l) /
                                                  np.maximum(1.0, np.abs(lse_ref_l))).max()))
                t = torch.tensor([worst_o, worst_l], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ok = bool(t[0].item() <= 2e-3 and t[1].item() <= 1e-4)
                out["parity"][f"{tr}_causal{int(causal)}"] = {"o_maxabs": t[0].item(), "lse_rel": t[1].item(), "pass": ok}
                if rank == 0:
                    print(f"PARITY {tr} causal={int(causal)} N={N} world={world}: O max-abs {t[0].item():.3e} "
                          f"lse rel {t[1].item():.3e} {'PASS' if ok else 'FAIL'}", flush=True)

    # ---- 2./3. timing at the c5 shape
    n_local = args.n_local or args.n_total // world
    n_total = n_local * world
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    shape = (1, 32, n_local, 128)
    q = torch.randn(shape, generator=g, device=dev, dtype=torch.float32).bfloat16()
    k = torch.randn(shape, generator=g, device=dev, dtype=torch.float32).bfloat16()
    v = (torch.rand(shape, generator=g, device=dev, dtype=torch.float32) - 0.5).bfloat16()
    flops = 4.0 * 32 * float(n_total) ** 2 * 128 / 2
    for tr in transports:
        for _ in range(args.warmup):
            fa.ring_attention(q, k, v, causal=True, transport=tr)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fa.ring_attention(q, k, v, causal=True, transport=tr)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        out["timing"][tr] = {"n_total": n_total, "n_local": n_local, "ms_per_step": ms,
                             "tflops_total": flops / (ms * 1e-3) * 1e-12,
                             "tflops_per_gpu": flops / (ms * 1e-3) * 1e-12 / world}
        if rank == 0:
            print(f"TIMING {tr} world={world} N={n_total}: {ms:.3f} ms/step, "
                  f"{out['timing'][tr]['tflops_per_gpu']:.1f} TFLOP/s per GPU", flush=True)
        os.environ["FA_B200_RING_PROFILE"] = "1"
        dist.barrier()
        fa.ring_attention(q, k, v, causal=True, transport=tr)
        torch.cuda.synchronize()
        del os.environ["FA_B200_RING_PROFILE"]
        dist.barrier()
    if rank == 0:
        print("RING_EVAL " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
