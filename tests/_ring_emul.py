"""Single-process, single-GPU run of the C-ABI ring (`fa_b200_ring_*`) with world_size > 1.

Ranks that live in the same process connect without CUDA IPC (csrc/fa_ring.cu), so the complete protocol -
publish, ready / pulled events with their host sequence counters, copy-engine pulls through the two-slot window,
zig-zag step schedule, one-pass combine - runs with real world_size-2/3/4 semantics on the ONE GPU of the driver's
test box.  Every rank gets its own host thread and its own stream, as the C ABI prescribes (a forward holds its host
thread until every peer has entered the same call; ctypes releases the GIL during the call).

Run by tests/test_ring_gpu.py in a subprocess under a timeout, and every wait in here is bounded, so a protocol bug
is reported instead of hanging the session.  Prints one line starting with PASS or FAIL.
"""
import ctypes
import os
import sys
import threading
import time

os.environ.setdefault("FA_B200_RING_TIMEOUT_S", "20")        # a rank that never shows up is an error after 20 s, not a hang
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    world, causal, calls = int(sys.argv[1]), bool(int(sys.argv[2])), int(sys.argv[3])
    B, H, d = 1, 3, 128
    Nl = 384                      # 1.5 work items per rank; the causal halves (192 rows) are not tile-aligned
    N = world * Nl
    import flash_attention_impls_b200 as fa
    from flash_attention_impls_b200 import _lib
    from flash_attention_impls_b200.parallel import zigzag_gather, zigzag_split
    from oracle import oracle

    lib = fa.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    rings = []
    for r in range(world):
        h = ctypes.c_void_p()
        _lib.check(lib.fa_b200_ring_create(world, r, B, H, Nl, d, _lib.FA_B200_BF16, ctypes.byref(h)))
        rings.append(h)
    blobs = []
    for h in rings:
        b = ctypes.create_string_buffer(_lib.FA_B200_RING_EXPORT_BYTES)
        _lib.check(lib.fa_b200_ring_export(h, b))
        blobs.append(b.raw)
    for h in rings:
        _lib.check(lib.fa_b200_ring_connect(h, b"".join(blobs)))
    block = B * H * Nl * d * 2
    owned = int(lib.fa_b200_ring_device_bytes(rings[0]))
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    worst_o = worst_l = worst_single = 0.0
    for call in range(calls):
        q, k, v = oracle.set_s((B, H, N, d), (B, H, N, d), seeds=(100 + call, 200 + call, 300 + call))
        tq, tk, tv = (torch.from_numpy(x).to(dev, torch.bfloat16) for x in (q, k, v))
        if causal:
            shards = [[zigzag_split(t, world, r) for t in (tq, tk, tv)] for r in range(world)]
        else:
            shards = [[t[:, :, r * Nl:(r + 1) * Nl].contiguous() for t in (tq, tk, tv)] for r in range(world)]
        outs = [torch.empty((B, H, Nl, d), dtype=torch.bfloat16, device=dev) for _ in range(world)]
        lses = [torch.empty((B, H, Nl), dtype=torch.float32, device=dev) for _ in range(world)]
        torch.cuda.synchronize()
        status = [None] * world

        def run_rank(r):
            ql, kl, vl = shards[r]
            torch.cuda.set_device(dev)
            status[r] = lib.fa_b200_ring_forward(rings[r], ql.data_ptr(), kl.data_ptr(), vl.data_ptr(), outs[r].data_ptr(),
                                                 lses[r].data_ptr(), 1 if causal else 0, 0.0, streams[r].cuda_stream)
            if status[r]:
                status[r] = (status[r], lib.fa_b200_last_error().decode())

        threads = [threading.Thread(target=run_rank, args=(r,), daemon=True) for r in range(world)]
        for t_ in threads:
            t_.start()
        for t_ in threads:
            t_.join(timeout=40.0)
        if any(t_.is_alive() for t_ in threads) or any(status):
            print(f"FAIL world={world} causal={int(causal)} call={call}: host side, alive={[t_.is_alive() for t_ in threads]} "
                  f"status={status}", flush=True)
            os._exit(1)
        deadline = time.time() + 30.0
        while not all(s_.query() for s_ in streams) and time.time() < deadline:
            time.sleep(0.01)
        if not all(s_.query() for s_ in streams):
            print(f"FAIL world={world} causal={int(causal)} call={call}: streams did not drain within 30 s: "
                  f"{[s_.query() for s_ in streams]}", flush=True)
            os._exit(1)
        torch.cuda.synchronize()
        if causal:
            o, lse = zigzag_gather(outs), zigzag_gather(lses)
        else:
            o, lse = torch.cat(outs, dim=2), torch.cat(lses, dim=2)
        o_one, lse_one = fa.attention_forward(tq, tk, tv, causal=causal)       # the same sequence on one GPU, one call
        torch.cuda.synchronize()
        o_ref, lse_ref, _, _ = oracle.attention(q, k, v, causal=causal)
        worst_o = max(worst_o, float(np.abs(o.float().cpu().numpy() - o_ref).max()))
        worst_l = max(worst_l, float((np.abs(lse.cpu().numpy() - lse_ref) / np.maximum(1.0, np.abs(lse_ref))).max()))
        worst_single = max(worst_single, float((o.float() - o_one.float()).abs().max()))
    for h in rings:
        lib.fa_b200_ring_destroy(h)
    ok = worst_o <= 2e-3 and worst_l <= 1e-4 and worst_single <= 4e-3
    # footprint: published block (K|V) + two receive slots (K|V each) + the partial stack (world outputs + lse)
    want = 3 * 2 * block + world * (block + B * H * Nl * 4)
    ok = ok and owned == want
    print(f"{'PASS' if ok else 'FAIL'} world={world} causal={int(causal)} calls={calls} o_err={worst_o:.3e} "
          f"lse_rel={worst_l:.3e} vs_single_gpu={worst_single:.3e} handle_bytes={owned} (= {owned / (2 * block):.2f} K|V blocks)")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
