"""GPU tests of the ring-attention driver.  With one visible GPU the ring degenerates to world_size 1
(still exercises accumulate/merge/cast on the device); with >= 2 GPUs a real NCCL ring is spawned."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, causal, N, d, q_out, transport="auto"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import flash_attention_impls_b200 as fa
        from oracle import oracle
        q, k, v = oracle.set_s((1, 4, N, d), (1, 4, N, d), seeds=(41, 42, 43))
        q, k, v = (torch.from_numpy(x).to(dev, torch.bfloat16) for x in (q, k, v))
        if causal:
            ql, kl, vl = (fa.zigzag_split(t, world, rank) for t in (q, k, v))
        else:
            c = N // world
            ql, kl, vl = (t[:, :, rank * c:(rank + 1) * c].contiguous() for t in (q, k, v))
        for _ in range(3):      # repeated calls exercise the double-buffered publish slots of the peer transport
            o, lse = fa.ring_attention(ql, kl, vl, causal=causal, transport=transport)
        torch.cuda.synchronize()
        q_out.put((rank, o.float().cpu().numpy(), lse.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("causal,transport", [(False, "peer"), (True, "peer"), (True, "p2p")])
def test_ring_attention_nccl(causal, transport):
    from flash_attention_impls_b200.parallel import zigzag_gather
    from oracle import oracle
    world = min(torch.cuda.device_count(), 4)
    if world < 1:
        pytest.skip("no GPU")
    N, d = 512 * world, 128
    port = 29800 + (os.getpid() % 1000) + (5 if causal else 0) + (11 if transport == "p2p" else 0)
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, causal, N, d, q_out, transport)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q_out.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    q, k, v = oracle.set_s((1, 4, N, d), (1, 4, N, d), seeds=(41, 42, 43))
    o_ref, lse_ref, _, _ = oracle.attention(q, k, v, causal=causal)
    if causal:
        o = zigzag_gather([torch.from_numpy(r[1]) for r in results]).numpy()
        lse = zigzag_gather([torch.from_numpy(r[2]) for r in results]).numpy()
    else:
        o = np.concatenate([r[1] for r in results], axis=2)
        lse = np.concatenate([r[2] for r in results], axis=2)
    assert np.abs(o - o_ref).max() <= 2e-3
    assert (np.abs(lse - lse_ref) / np.maximum(1.0, np.abs(lse_ref))).max() <= 1e-4


# ------------------------------------------------------------------------------ the C-ABI ring (fa_b200_ring_*)
@pytest.mark.parametrize("world,causal", [(2, True), (2, False), (4, True), (3, True)])
def test_c_abi_ring_protocol_on_one_gpu(world, causal):
    """world_size 2/3/4 of the C-ABI ring on ONE GPU (ranks in one process, one host thread each, connect without IPC):
    ready / pulled events and their host counters, two-slot pull window, zig-zag schedule and combine, three
    back-to-back calls, against the oracle and against a single attention_forward over the whole sequence.  Runs in a
    subprocess under a timeout, with every wait inside bounded, so a protocol bug cannot hang the test session."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ring_emul.py"), str(world), str(int(causal)), "3"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().startswith("PASS"), r.stdout + r.stderr


def test_c_abi_ring_world_one_and_argument_errors():
    import ctypes
    import flash_attention_impls_b200 as fa
    from flash_attention_impls_b200 import _lib
    from oracle import oracle
    lib = fa.load()
    dev = torch.device("cuda", 0)
    h = ctypes.c_void_p()
    _lib.check(lib.fa_b200_ring_create(1, 0, 1, 2, 300, 64, _lib.FA_B200_FP16, ctypes.byref(h)))
    q, k, v = oracle.set_s((1, 2, 300, 64), (1, 2, 300, 64), seeds=(7, 8, 9))
    tq, tk, tv = (torch.from_numpy(x).to(dev, torch.float16) for x in (q, k, v))
    o = torch.empty_like(tq)
    lse = torch.empty((1, 2, 300), dtype=torch.float32, device=dev)
    _lib.check(lib.fa_b200_ring_forward(h, tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), o.data_ptr(), lse.data_ptr(), 1, 0.0,
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    o_ref, lse_ref, _, _ = oracle.attention(q, k, v, causal=True)
    assert np.abs(o.float().cpu().numpy() - o_ref).max() <= 2e-3
    assert (np.abs(lse.cpu().numpy() - lse_ref) / np.maximum(1.0, np.abs(lse_ref))).max() <= 1e-4
    assert lib.fa_b200_ring_device_bytes(h) == 0          # world 1 owns nothing
    assert lib.fa_b200_ring_forward(h, None, tk.data_ptr(), tv.data_ptr(), o.data_ptr(), None, 0, 0.0, None) == 1   # ERR_NULL
    lib.fa_b200_ring_destroy(h)
    # a world-2 handle that has not been connected refuses to run
    _lib.check(lib.fa_b200_ring_create(2, 1, 1, 2, 256, 64, _lib.FA_B200_FP16, ctypes.byref(h)))
    assert lib.fa_b200_ring_forward(h, tq.data_ptr(), tk.data_ptr(), tv.data_ptr(), o.data_ptr(), None, 0, 0.0, None) == 2
    assert b"connect" in lib.fa_b200_last_error()
    lib.fa_b200_ring_destroy(h)
