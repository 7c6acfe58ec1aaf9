#!/usr/bin/env python
"""bench.py — the hot path's benchmark (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c4|c2|c5] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one attention forward over the whole workload.  Default workload is BASELINE.json
configs[2] ("c3"): B=4 H=32 N=8192 d=128 bf16 non-causal, the configuration the metric is quoted on.
Each rank runs that workload on its own GPU ((batch,head) units are independent: no collective), so
N-GPU runs are weak scaling and `value` is the aggregate TFLOP/s over all ranks.  With --workload c5
the ranks cooperate on ONE sequence with ring attention (K/V blocks pulled from peer memory over NVLink,
or NCCL send/recv with --ring-transport p2p), which is strong scaling.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, H, N, d, dtype, causal)   -- BASELINE.json configs[1..4]
    "c2": (8, 16, 1024, 64, "fp16", False),
    "c3": (4, 32, 8192, 128, "bf16", False),
    "c4": (4, 32, 8192, 128, "bf16", True),
    "c5": (1, 32, 131072, 128, "bf16", True),
}
METRIC = "attention_fwd_tflops"
UNIT = "TFLOP/s"


def flops_of(B, H, N, d, causal, Nkv=None):
    f = 4.0 * B * H * N * (Nkv if Nkv else N) * d      # main.cu:456, test_flash_attn.cu:308
    return f / 2 if causal else f


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops"]), tflops_sustained=float(p.get("bf16_tflops_sustained", 0)) or None,
                    hbm=float(p["hbm_gbs"]), source="MEASURED_PEAKS.json (measured)")
    except Exception:
        return dict(tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0, source="B200_PROFILING.md fallback")


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, torch_device):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop_evt = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_device).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_attention_sample(B, H, N, d, causal, target_s, threads):
    """Times the oracle port (naive fp32 softmax(QK^T)V restating main.cu:165-202) on the host cores on a
    bounded sample of the workload: the last `rows` query rows of `nbh` (b,h) slices against all N keys,
    sized from a short calibration run to take about `target_s` seconds."""
    from oracle import oracle
    max_bh = min(B * H, 8)
    q, k, v = oracle.set_s((1, max_bh, N, d), (1, max_bh, N, d))

    def run(sample):
        nbh, rows = sample
        r0 = N - rows            # the last rows: for causal runs these see (almost) all keys
        t0 = time.perf_counter()
        oracle.attention(q[:, :nbh], k[:, :nbh], v[:, :nbh], causal=causal, nthreads=threads, row_begin=r0, row_end=N)
        dt = time.perf_counter() - t0
        fl = 4.0 * rows * N * d * nbh
        if causal:               # rows r0..N-1 see r+1 keys each
            fl = 4.0 * d * (rows * (r0 + N + 1) / 2.0) * nbh
        return dt, fl
    cal_rows = min(N, 128)
    dt, fl = run((1, cal_rows))
    want_rows = cal_rows * target_s / max(dt, 1e-4)
    if want_rows <= N:
        sample = (1, int(max(16, want_rows)))
    else:
        sample = (int(min(max_bh, max(1, round(want_rows / N)))), N)
    return sample, run


def sample_text(sample, N, d, causal):
    return (f"last {sample[1]} query rows x {N} keys of {sample[0]} (b,h) slice(s), d={d}, fp32 naive "
            f"softmax(QK^T)V (oracle port of main.cu:165-202), causal={int(causal)}")


def reference_arm(args, wl):
    """`--impl reference`: the reference has no CPU implementation of this path and its CUDA kernels are not
    host code, so the CPU arm is the oracle port with every host thread (cpu_baseline.kind = "port")."""
    B, H, N, d, dtype, causal = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    target = max(0.05, min(1.0, 150.0 / max(1, steps + warmup)))
    rows, run = cpu_attention_sample(B, H, N, d, causal, target, threads)
    for _ in range(warmup):
        run(rows)
    t_tot, f_tot = 0.0, 0.0
    for _ in range(steps):
        dt, fl = run(rows)
        t_tot += dt
        f_tot += fl
    tf = f_tot / t_tot * 1e-12
    sample = sample_text(rows, N, d, causal) + " per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": tf, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": t_tot / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "B": B, "H": H, "N": N, "d": d, "causal": causal,
                   "note": "reference has no CPU path; oracle port of main.cu:165-202 on host cores"},
        "cpu_baseline": {"value": tf, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": tf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer arm (0 = min(steps, 5))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ring-transport", default="auto", choices=["auto", "peer", "p2p"],
                    help="c5 only: K/V block transport (auto = CUDA-IPC peer buffers + copy-engine pulls)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, wl)
        return
    args.warmup = max(3, args.warmup)

    import torch
    import torch.distributed as dist

    import flash_attention_impls_b200 as fa

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference "
                         "for the CPU baseline arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fa.load()

    B, H, N, d, dtype_name, causal = wl
    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float16
    ring = args.workload == "c5" and world > 1
    peaks = load_peaks()

    # ---- synthetic inputs (Set S distribution, generated on the device; resident before the timed region)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    if ring:
        n_local = N // world
        shape = (B, H, n_local, d)
    else:
        shape = (B, H, N, d)
    q = torch.randn(shape, generator=g, device=dev, dtype=torch.float32).to(dtype)
    k = torch.randn(shape, generator=g, device=dev, dtype=torch.float32).to(dtype)
    v = (torch.rand(shape, generator=g, device=dev, dtype=torch.float32) - 0.5).to(dtype)
    o = torch.empty_like(q)
    lse = torch.empty(shape[:3], dtype=torch.float32, device=dev)

    if ring:
        def step():
            return fa.ring_attention(q, k, v, causal=causal, transport=args.ring_transport)
        step_flops_total = flops_of(B, H, N, d, causal)          # one sequence shared by all ranks
        scaling = "strong"
    else:
        def step():
            fa.attention_forward(q, k, v, causal=causal, out=o, lse=lse)
        step_flops_total = flops_of(B, H, N, d, causal) * world  # every rank runs the full workload
        scaling = "weak"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # L2 hygiene: c3/c4/c5 stream > 126 MB per step; smaller workloads get an L2 flush (a 512 MB write)
    # between timed iterations, outside the per-step event pairs.
    in_bytes = 3 * q.numel() * q.element_size()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if in_bytes < (256 << 20) else None
    sampler = ClockSampler(dev)
    sampler.start()
    launches0 = fa.launch_count()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        if flush is not None:
            flush.zero_()
        ev0[i].record()
        step()
        ev1[i].record()
    torch.cuda.synchronize()
    # without a flush the steps run back to back and the span first-start -> last-end is the step time
    total_ms = ev0[0].elapsed_time(ev1[-1]) if flush is None else sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    barrier()
    launches = fa.launch_count() - launches0
    clocks = sampler.stop()
    per_step = sorted(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = step_flops_total / (ms_per_step * 1e-3) * 1e-12

    # ---- roofline of the dominant kernel (fa_fwd_sm100_kernel): algorithmic FLOPs / mean launch duration
    kern_flops = flops_of(B, H, N, d, causal) if not ring else None
    roofline = None
    if kern_flops is not None:
        mean_launch_ms = sum(per_step) / len(per_step)
        achieved = kern_flops / (mean_launch_ms * 1e-3) * 1e-12
        e = 2
        alg_bytes = 4 * B * H * N * d * e + B * H * N * 4
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tflops"], "traffic": None,
                    "peak_source": peaks["source"] + ", burst bf16 matmul",
                    "frac_of_sustained": (achieved / peaks["tflops_sustained"]) if peaks["tflops_sustained"] else None,
                    "frac_of_datasheet_2250": achieved / 2250.0,
                    "kernel": "fa_fwd_sm100_kernel", "launch_ms_mean": mean_launch_ms, "launch_ms_min": per_step[0],
                    "algorithmic_flops_per_launch": kern_flops, "algorithmic_bytes_per_launch": alg_bytes,
                    "hbm_gbs_algorithmic": alg_bytes / (mean_launch_ms * 1e-3) * 1e-9, "hbm_peak_gbs": peaks["hbm"]}
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                roofline["traffic"] = json.load(f).get(args.workload)
        except Exception:
            pass

    # ---- e2e: pinned host buffers -> H2D -> kernel -> D2H, through the package's public host API
    e2e = None
    if not args.no_e2e and not ring:
        pipe = fa.HostPipeline(B, H, N, d, dtype, causal=causal, chunks=8, device=dev)
        hq = torch.empty((B, H, N, d), dtype=dtype).pin_memory()
        hk = torch.empty_like(hq).pin_memory()
        hv = torch.empty_like(hq).pin_memory()
        ho = torch.empty_like(hq).pin_memory()
        hl = torch.empty((B, H, N), dtype=torch.float32).pin_memory()
        hq.copy_(q.cpu()); hk.copy_(k.cpu()); hv.copy_(v.cpu())
        n_e2e = args.e2e_steps or min(args.steps, 5)
        for _ in range(2):
            pipe(hq, hk, hv, ho, hl)
            pipe.synchronize()
        barrier()
        # every step's result is read back on the host, so each step ends with a sync; the step time is the
        # device-side span first-H2D -> last-D2H recorded by the library around its own copies and kernels
        e2e_ms = 0.0
        for _ in range(n_e2e):
            pipe(hq, hk, hv, ho, hl)
            pipe.synchronize()
            e2e_ms += pipe.elapsed_ms_last_call()
        torch.cuda.synchronize()
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        bi, bo = pipe.bytes_per_call()
        e2e = {"value": step_flops_total / (e2e_ms / n_e2e * 1e-3) * 1e-12, "unit": UNIT,
               "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo, "ms_per_step": e2e_ms / n_e2e, "steps": n_e2e,
               "api": "fa_b200_forward_host (C ABI; pinned host -> 8-chunk H2D/compute/D2H pipeline on 3 streams)",
               "result_check": float(ho.view(-1)[:1024].float().abs().sum().item())}
        barrier()

    # ---- CPU baseline beside it (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample, run = cpu_attention_sample(B, H, N, d, causal, 12.0, threads)
        dt, fl = run(sample)
        cpu = {"value": fl / dt * 1e-12, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": sample_text(sample, N, d, causal) + f", {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": dtype_name, "data": "synthetic",
            "config": {"workload": args.workload, "B": B, "H": H, "N": N, "d": d, "causal": causal,
                       "layout": "[B,H,N,d] contiguous", "per_gpu_tflops": value / world,
                       "parallelism": ("ring%d (zig-zag; K/V blocks %s; lse merge)" % (
                           world, "over NCCL send/recv" if args.ring_transport == "p2p" else
                           "pulled from peer memory by the copy engines")) if ring else
                                      ("independent (b,h) units, %d rank(s), no collective" % world),
                       "l2": ("inputs per step (%.0f MB) exceed the 126 MB L2; no flush needed" % (in_bytes / 1e6))
                             if flush is None else "L2 flushed (512 MB write) between timed iterations"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
