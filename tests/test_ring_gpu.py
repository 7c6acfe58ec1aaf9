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
