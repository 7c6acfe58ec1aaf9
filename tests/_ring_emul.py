"""Single-process, single-GPU run of the C-ABI ring (`fa_b200_ring_*`) with world_size > 1.

Ranks that live in the same process connect without CUDA IPC (csrc/fa_ring.cu), so the complete protocol -
publish, ready/ack sequence flags, copy-engine pulls through the two-slot window, zig-zag step schedule,
one-pass combine - runs with real world_size-2/4 semantics on the ONE GPU of the driver's test box.  Every rank
gets its own stream; a forward is enqueue-only, so the host enqueues rank 0's whole call (which waits, on the
device, for flags that rank 1 has not even enqueued yet) and then the others.

Run by tests/test_ring_gpu.py in a subprocess under a timeout (a protocol bug would leave streams waiting on a
flag forever; the subprocess dies and takes its waits with it).  Prints one line starting with PASS or FAIL.
"""
import ctypes
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per stream: no false dependencies
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
        for r in range(world):
            ql, kl, vl = shards[r]
            _lib.check(lib.fa_b200_ring_forward(rings[r], ql.data_ptr(), kl.data_ptr(), vl.data_ptr(), outs[r].data_ptr(),
                                                lses[r].data_ptr(), 1 if causal else 0, 0.0, streams[r].cuda_stream))
        # never block on a device-side wait that might not be satisfied: poll the streams with a deadline and, if they
        # do not drain, dump the sequence flags (which rank is waiting for whom) before giving up
        import time
        deadline = time.time() + 30.0
        while not all(s_.query() for s_ in streams) and time.time() < deadline:
            time.sleep(0.01)
        if not all(s_.query() for s_ in streams):
            side = torch.cuda.Stream(dev)
            dump = torch.zeros(128, dtype=torch.int32, device=dev)
            for r in range(world):
                kb, vb = ctypes.c_void_p(), ctypes.c_void_p()
                lib.fa_b200_ring_kv_buffers(rings[r], ctypes.byref(kb), ctypes.byref(vb))
                lib.fa_b200_copy_async(dump.data_ptr(), kb.value + 2 * block, 512, side.cuda_stream)
                side.synchronize()
                fl = dump.cpu().tolist()
                print(f"STUCK call={call} rank={r} stream_done={streams[r].query()} ready={fl[:world]} ack={fl[64:64 + world]}", flush=True)
            print(f"FAIL world={world} causal={int(causal)}: streams did not drain within 30 s", flush=True)
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
    # footprint: published block (K|V) + two receive slots (K|V each) + the partial stack (world outputs + lse) + flags
    want = 3 * 2 * block + world * (block + B * H * Nl * 4) + 2 * 64 * 4 + 8
    ok = ok and owned == want
    print(f"{'PASS' if ok else 'FAIL'} world={world} causal={int(causal)} calls={calls} o_err={worst_o:.3e} "
          f"lse_rel={worst_l:.3e} vs_single_gpu={worst_single:.3e} handle_bytes={owned} (= {owned / (2 * block):.2f} K|V blocks)")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
