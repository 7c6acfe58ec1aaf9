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
def bind_to_gpu_numa_node(torch, dev):
    """Binds the calling thread to the CPUs NVML reports as local to this GPU, so that the pinned host buffers of
    the e2e arm are first-touched on the GPU's own NUMA node.  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
        if not uuid.startswith("GPU-"):
            uuid = "GPU-" + uuid
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        node = None
        try:
            for n in sorted(os.listdir("/sys/devices/system/node")):
                if n.startswith("node") and n[4:].isdigit():
                    with open(f"/sys/devices/system/node/{n}/cpulist") as f:
                        first = f.read().split(",")[0].split("-")
                    lo, hi = int(first[0]), int(first[-1])
                    if lo <= cpus[0] <= hi:
                        node = int(n[4:])
        except Exception:
            pass
        return {"cpus": f"{cpus[0]}-{cpus[-1]} ({len(cpus)})", "numa_node": node}
    except Exception as e:                      # affinity is an optimisation, never a requirement
        return {"error": str(e)[:80]}


def timed_steps(torch, step, steps, flush=None):
    """K steps bracketed by per-step CUDA events on the current stream.  Returns (total_ms, sorted per-step ms):
    without a flush the steps run back to back and the total is first-start -> last-end."""
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        if flush is not None:
            flush.zero_()
        ev0[i].record()
        step()
        ev1[i].record()
    torch.cuda.synchronize()
    per = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    total = ev0[0].elapsed_time(ev1[-1]) if flush is None else sum(per)
    return total, sorted(per)


def synth(torch, shape, dtype, dev, seed):
    """Set S distribution: Q, K ~ N(0,1), V ~ U(-0.5, 0.5), generated on the device."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    q = torch.randn(shape, generator=g, device=dev, dtype=torch.float32).to(dtype)
    k = torch.randn(shape, generator=g, device=dev, dtype=torch.float32).to(dtype)
    v = (torch.rand(shape, generator=g, device=dev, dtype=torch.float32) - 0.5).to(dtype)
    return q, k, v


def kernel_record(torch, fa, name, dev, peaks, steps):
    """Device-timed figure of another BASELINE config on this GPU (sub-record of the line; not the headline)."""
    B, H, N, d, dtype_name, causal = WORKLOADS[name]
    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float16
    q, k, v = synth(torch, (B, H, N, d), dtype, dev, 4321)
    o = torch.empty_like(q)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=dev)
    in_bytes = 3 * q.numel() * q.element_size()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if in_bytes < (256 << 20) else None

    def step():
        fa.attention_forward(q, k, v, causal=causal, out=o, lse=lse)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    total, per = timed_steps(torch, step, steps, flush)
    ms = sum(per) / len(per)
    fl = flops_of(B, H, N, d, causal)
    alg_bytes = 4 * B * H * N * d * 2 + B * H * N * 4
    tf = fl / (ms * 1e-3) * 1e-12
    gbs = alg_bytes / (ms * 1e-3) * 1e-9
    return {"workload": name, "B": B, "H": H, "N": N, "d": d, "dtype": dtype_name, "causal": causal, "steps": steps,
            "launch_ms_mean": ms, "launch_ms_min": per[0], "tflops": tf, "frac_of_measured_tensor_peak": tf / peaks["tflops"],
            "hbm_gbs_algorithmic": gbs, "frac_of_measured_hbm_peak": gbs / peaks["hbm"],
            "l2": "L2 flushed (512 MB write) between timed iterations" if flush is not None else "inputs exceed L2"}


def zero_input_record(torch, fa, name, dev, peaks, steps, real_tflops):
    """The same launch on all-zero inputs.  Zero operands do not toggle the tensor pipe's datapath, the GPU stays far
    below its power limit and the SM clock at its maximum, so this figure is the kernel's CYCLE count; the headline (random
    inputs) is what power management leaves of it.  Explains the roofline fraction, is not a throughput claim."""
    B, H, N, d, dtype_name, causal = WORKLOADS[name]
    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float16
    q = torch.zeros((B, H, N, d), dtype=dtype, device=dev)
    k, v, o = torch.zeros_like(q), torch.zeros_like(q), torch.empty_like(q)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=dev)

    def step():
        fa.attention_forward(q, k, v, causal=causal, out=o, lse=lse)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    total, per = timed_steps(torch, step, steps)
    ms = sum(per) / len(per)
    tf = flops_of(B, H, N, d, causal) / (ms * 1e-3) * 1e-12
    return {"workload": name + " shape, all-zero inputs (SM clock stays at max: cycle-domain rate)", "steps": steps,
            "launch_ms_mean": ms, "launch_ms_min": per[0], "tflops": tf, "frac_of_measured_tensor_peak": tf / peaks["tflops"],
            "random_over_zero": real_tflops / tf,
            "note": "random_over_zero = headline kernel rate / this rate: what the power limit costs on real data"}


def backward_record(torch, fa, dev, peaks, steps):
    """SURVEY.md section 8f.4 as a sub-record: fa_b200_backward (delta pre-pass + dQ kernel + dK/dV kernel) at the c3 shape.
    FLOPs in the usual 5-product convention, 10*B*H*N^2*d (the two kernels execute 7 products: S and dP are recomputed)."""
    B, H, N, d, dtype_name, causal = WORKLOADS["c3"]
    q, k, v = synth(torch, (B, H, N, d), torch.bfloat16, dev, 99)
    do = torch.randn_like(q)
    o, lse = fa.attention_forward(q, k, v, causal=causal)

    def step():
        fa.attention_backward(q, k, v, o, lse, do, causal=causal)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    launches0 = fa.launch_count()
    total, per = timed_steps(torch, step, steps)
    ms = total / steps
    tf = 10.0 * B * H * N * N * d / (ms * 1e-3) * 1e-12
    return {"workload": "backward at the c3 shape (B=4 H=32 N=8192 d=128 bf16 non-causal)", "steps": steps, "ms_per_step": ms,
            "tflops_5_product_convention": tf, "frac_of_measured_tensor_peak": tf / peaks["tflops"],
            "tflops_as_executed_7_products": tf * 1.4, "gpu_launches": int(fa.launch_count() - launches0),
            "kernels": "bwd_delta_kernel, fa_bwd_dq_sm100_kernel, fa_bwd_dkdv_sm100_kernel"}


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
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (c2/c4 at N=1; strong c3 and ring c5 at N>1)")
    ap.add_argument("--ring-transport", default="auto", choices=["auto", "peer", "p2p"],
                    help="ring attention: K/V block transport (auto = the C-ABI ring: copy-engine pulls from peer memory)")
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
    affinity = bind_to_gpu_numa_node(torch, dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fa.load()

    B, H, N, d, dtype_name, causal = wl
    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float16
    ring = args.workload == "c5" and world > 1
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- synthetic inputs (Set S distribution, generated on the device; resident before the timed region)
    shape = (B, H, N // world, d) if ring else (B, H, N, d)
    q, k, v = synth(torch, shape, dtype, dev, 1234 + rank)
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
    barrier()
    total_ms, per_step = timed_steps(torch, step, args.steps, flush)
    barrier()
    launches = fa.launch_count() - launches0
    clocks = sampler.stop()
    total_ms = max_over_ranks(total_ms)
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
                tr = json.load(f)
            roofline["traffic"] = tr.get(args.workload)
            roofline["traffic_source"] = ("not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                          "launch from the committed ncu --set full capture, " + str(tr.get("_source")))
        except Exception:
            pass

    # ---- e2e: pinned host buffers -> H2D -> kernel -> D2H, through the package's public host API
    def e2e_arm(nbh0, nbh1, total_flops):
        """Times fa_b200_forward_host over the (b,h) slices [nbh0, nbh1) of the workload (all of them at N=1; this
        rank's shard in the strong-scaling sub-record).  The step time is host wall clock around call + sync (what a
        caller sees); the library's own device-event span first-H2D -> last-D2H is reported beside it."""
        n = nbh1 - nbh0
        pipe = fa.HostPipeline(1, n, N, d, dtype, causal=causal, chunks=min(8, n), device=dev)
        hq = torch.empty((1, n, N, d), dtype=dtype).pin_memory()
        hk, hv, ho = (torch.empty_like(hq).pin_memory() for _ in range(3))
        hl = torch.empty((1, n, N), dtype=torch.float32).pin_memory()
        src = [t.reshape(1, B * H, N, d)[:, nbh0:nbh1] for t in (q, k, v)]
        hq.copy_(src[0].cpu()); hk.copy_(src[1].cpu()); hv.copy_(src[2].cpu())
        n_e2e = args.e2e_steps or min(args.steps, 5)
        for _ in range(2):
            pipe(hq, hk, hv, ho, hl)
            pipe.synchronize()
        barrier()
        dev_ms, t0 = 0.0, time.perf_counter()
        for _ in range(n_e2e):
            pipe(hq, hk, hv, ho, hl)
            pipe.synchronize()          # the step's result is read on the host: every step ends with a sync
            dev_ms += pipe.elapsed_ms_last_call()
        host_ms = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        host_ms, dev_ms = max_over_ranks(host_ms), max_over_ranks(dev_ms)
        bi, bo = pipe.bytes_per_call()
        rec = {"value": total_flops / (host_ms / n_e2e * 1e-3) * 1e-12, "unit": UNIT,
               "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo, "ms_per_step": host_ms / n_e2e,
               "device_span_ms_per_step": dev_ms / n_e2e, "steps": n_e2e,
               "timer": "host wall clock around fa_b200_forward_host + fa_b200_host_ctx_sync, max over ranks; "
                        "device_span = the library's CUDA-event span first H2D -> last D2H",
               "api": "fa_b200_forward_host (C ABI; pinned host -> 8-chunk H2D/compute/D2H pipeline on 3 streams)",
               "host_affinity": affinity,
               "result_check": float(ho.view(-1)[:1024].float().abs().sum().item())}
        pipe.close()
        barrier()
        return rec

    e2e = None
    if not args.no_e2e and not ring:
        e2e = e2e_arm(0, B * H, step_flops_total)

    # ---- sub-records: the other BASELINE configs, driver-visible in the same line
    sub = {}
    if not args.no_sub and args.workload == "c3":
        sub_steps = max(3, min(args.steps, 10))
        if world == 1:
            # first, after a pause (it is meant to show the kernel with the clock at its maximum, not after a hot run)
            time.sleep(1.0)
            sub["c3_zero_inputs"] = zero_input_record(torch, fa, "c3", dev, peaks, sub_steps, roofline["achieved"])
            for name in ("c2", "c4"):
                sub[name] = kernel_record(torch, fa, name, dev, peaks, max(sub_steps, 20))
            sub["backward_c3"] = backward_record(torch, fa, dev, peaks, sub_steps)
        else:
            # (i) BASELINE configs[2]: c3 (batch,head)-SHARDED over the ranks - strong scaling, no collective
            b0, b1 = fa.bh_shard_range(B * H, world, rank)
            qs, ks, vs = (t.reshape(1, B * H, N, d)[:, b0:b1] for t in (q, k, v))
            os_ = o.reshape(1, B * H, N, d)[:, b0:b1]
            ls_ = lse.reshape(1, B * H, N)[:, b0:b1]

            def shard_step():
                fa.attention_forward(qs, ks, vs, causal=causal, out=os_, lse=ls_)
            for _ in range(3):
                shard_step()
            barrier()
            smp = ClockSampler(dev); smp.start()
            t_ms, per = timed_steps(torch, shard_step, sub_steps)
            barrier()
            clk = smp.stop()
            t_ms = max_over_ranks(t_ms) / sub_steps
            full_ms = sum(per_step) / len(per_step)               # this rank's full-c3 launch time from the headline run
            tf_total = flops_of(B, H, N, d, causal) / (t_ms * 1e-3) * 1e-12
            sub["strong_c3"] = {
                "config": "BASELINE configs[2]: B=4 H=32 N=8192 d=128 bf16 non-causal, (b,h) slices split over the ranks "
                          "(bh_shard_range), no collective", "scaling": "strong", "n_gpus": world,
                "bh_per_gpu": b1 - b0, "steps": sub_steps, "ms_per_step": t_ms, "value": tf_total, "unit": UNIT,
                "per_gpu_tflops": tf_total / world,
                "efficiency_vs_single_gpu_kernel": (full_ms / world) / t_ms,
                "single_gpu_launch_ms": full_ms, "clocks": clk,
                "note": "%d work items of 256 rows per GPU = %.2f waves of 148 SMs" % ((b1 - b0) * N // 256, (b1 - b0) * N / 256 / 148)}
            if not args.no_e2e:
                sub["strong_c3"]["e2e"] = e2e_arm(b0, b1, flops_of(B, H, N, d, causal))
            del qs, ks, vs, os_, ls_
        if world > 1:
            # (ii) BASELINE configs[4]: c5 causal ring attention, one sequence over all ranks
            del q, k, v, o, lse
            torch.cuda.empty_cache()
            sub["ring_c5"] = ring_record(torch, dist, fa, dev, rank, world, peaks, sub_steps, args.ring_transport, barrier,
                                         max_over_ranks)

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
            "tmap_cache": dict(zip(("hits", "misses"), fa.tmap_cache_stats())),
        }
        line.update(sub)
        print(json.dumps(line), flush=True)
    if world > 1:
        fa.release_peer_buffers()
        dist.destroy_process_group()


def ring_record(torch, dist, fa, dev, rank, world, peaks, steps, transport, barrier, max_over_ranks):
    """c5 (BASELINE configs[4]) as a sub-record: causal ring attention of ONE 131072-token sequence over all ranks
    (zig-zag partition), timed like the headline (CUDA events, barrier + sync both sides, max over ranks), plus -
    outside the timed region - a parity check of the ring against a single-GPU attention_forward over the whole
    sequence on one head, and a per-step timeline of one call."""
    B, H, N, d, dtype_name, causal = WORKLOADS["c5"]
    dtype = torch.bfloat16
    n_local = N // world
    q, k, v = synth(torch, (B, H, n_local, d), dtype, dev, 777 + rank)   # rank r's zig-zag rows [chunk r ; chunk 2P-1-r]

    def step():
        return fa.ring_attention(q, k, v, causal=causal, transport=transport)
    for _ in range(3):
        out, lse = step()
    barrier()
    smp = ClockSampler(dev); smp.start()
    launches0 = fa.launch_count()
    t_ms, per = timed_steps(torch, step, steps)
    barrier()
    launches = fa.launch_count() - launches0
    clk = smp.stop()
    t_ms = max_over_ranks(t_ms) / steps
    fl = flops_of(B, H, N, d, causal)
    per_gpu = fl / world / (t_ms * 1e-3) * 1e-12
    rec = {"config": "BASELINE configs[4]: B=1 H=32 N=131072 d=128 bf16 causal, ring attention over %d GPUs (zig-zag)" % world,
           "scaling": "strong", "n_gpus": world, "steps": steps, "ms_per_step": t_ms, "value": fl / (t_ms * 1e-3) * 1e-12,
           "unit": UNIT, "per_gpu_tflops": per_gpu, "frac_of_measured_tensor_peak": per_gpu / peaks["tflops"],
           "gpu_launches": int(launches), "clocks": clk}
    ring = fa.c_ring_for(q, None) if transport != "p2p" else None
    if ring is not None and ring.ok:
        block = 2 * q.numel() * q.element_size()
        rec["transport"] = ("C-ABI ring (fa_b200_ring_*): copy-engine pulls from the owner's CUDA-IPC buffer, two receive "
                            "slots, interprocess ready/pulled events + host sequence counters, no collective in the step loop")
        rec["handle_device_bytes"] = ring.device_bytes()
        rec["handle_bytes_in_kv_blocks"] = ring.device_bytes() / block
        rec["ring_traffic_bytes_per_rank_per_step"] = block
        # one profiled call: where does a step's time go?
        barrier()
        ring.set_profile(True)
        step()
        torch.cuda.synchronize()
        steps_tl, combine_done = ring.timeline()
        ring.set_profile(False)
        if rank == 0 and steps_tl:
            wait = sum(r_ - b_ for (_, b_, r_, _) in steps_tl)
            kern = sum(d_ - r_ for (_, _, r_, d_) in steps_tl)
            rec["timeline_rank0_ms"] = {"steps": [{"step": s_, "begin": round(b_, 3), "kv_ready": round(r_, 3), "attn_done": round(d_, 3)}
                                                  for (s_, b_, r_, d_) in steps_tl], "combine_done": combine_done,
                                        "sum_wait_for_kv": wait, "sum_attention_kernels": kern,
                                        "publish_and_flags_before_step0": steps_tl[0][1],
                                        "combine": (combine_done - steps_tl[-1][3]) if combine_done else None}
            parts = {"attention kernels": kern, "waiting for K/V blocks": wait, "publish": steps_tl[0][1],
                     "combine": (combine_done - steps_tl[-1][3]) if combine_done else 0.0}
            rec["limiter"] = max(parts, key=parts.get) + " (%.0f %% of the call)" % (100 * max(parts.values()) / max(combine_done or 1e-9, 1e-9))
    else:
        rec["transport"] = "NCCL send/recv (torch.distributed.batch_isend_irecv), all exchanges posted up front"
    barrier()

    # ---- parity, outside any timed region: head 0 of the ring result vs ONE attention_forward over the whole sequence
    out, lse = step()
    torch.cuda.synchronize()
    gath = []
    for t in (q[:, :1], k[:, :1], v[:, :1]):
        parts = [torch.empty_like(t.contiguous()) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        gath.append(fa.zigzag_gather(parts))
    o_one, lse_one = fa.attention_forward(gath[0], gath[1], gath[2], causal=True)
    o_loc, lse_loc = fa.zigzag_split(o_one, world, rank), fa.zigzag_split(lse_one.unsqueeze(-1), world, rank).squeeze(-1)
    err_o = float((out[:, :1].float() - o_loc.float()).abs().max())
    err_l = float(((lse[:, :1] - lse_loc).abs() / lse_loc.abs().clamp(min=1.0)).max())
    rec["parity_vs_single_gpu_head0"] = {"o_max_abs": max_over_ranks(err_o), "lse_max_rel": max_over_ranks(err_l),
                                         "gate": "O 4e-3 (two bf16 roundings), lse 1e-4 relative",
                                         "pass": bool(max_over_ranks(err_o) <= 4e-3 and max_over_ranks(err_l) <= 1e-4)}
    barrier()
    return rec


if __name__ == "__main__":
    main()
