"""world_size-2 (and 4) gloo tests of the ring-attention schedule on CPU.  The compute backend is the
oracle (test injection through ring_attention's `backend` argument); what is under test is the host logic
of the N>1 path: block rotation, zig-zag causal case analysis, strided sub-range calls, merge order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """CPU stand-in with the same two calls as the CUDA backend, built on the oracle."""

    def attention(self, q, k, v, causal, out, lse):
        from oracle import oracle
        o, s, _, _ = oracle.attention(q.float().numpy(), k.float().numpy(), v.float().numpy(), causal=causal, nthreads=1)
        out.copy_(torch.from_numpy(o).to(out.dtype))
        lse.copy_(torch.from_numpy(s))

    def combine(self, o_parts, lse_parts):
        from oracle import oracle
        o = np.zeros(o_parts.shape[1:], np.float32)
        s = np.full(lse_parts.shape[1:], -np.inf, np.float32)
        for i in range(o_parts.shape[0]):
            o_i = np.nan_to_num(o_parts[i].float().numpy())      # rows a step skipped are uninitialised (lse = -inf)
            o, s = oracle.merge_partial(o, s, o_i, lse_parts[i].numpy())
            o, s = o.reshape(o_parts.shape[1:]), s.reshape(lse_parts.shape[1:])
        return torch.from_numpy(o).to(o_parts.dtype), torch.from_numpy(s)


def _worker(rank, world, port, causal, N, d, q_out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flash_attention_impls_b200.parallel import ring_attention, zigzag_split
        from oracle import oracle
        q, k, v = oracle.set_s((1, 2, N, d), (1, 2, N, d), seeds=(21, 22, 23))
        q, k, v = map(torch.from_numpy, (q, k, v))
        if causal:
            ql, kl, vl = (zigzag_split(t, world, rank) for t in (q, k, v))
        else:
            c = N // world
            ql, kl, vl = (t[:, :, rank * c:(rank + 1) * c].contiguous() for t in (q, k, v))
        o, lse = ring_attention(ql, kl, vl, causal=causal, backend=OracleBackend())
        q_out.put((rank, o.numpy(), lse.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,causal", [(2, False), (2, True), (4, True)])
def test_ring_attention_schedule_matches_single_device_oracle(world, causal):
    from flash_attention_impls_b200.parallel import zigzag_gather
    from oracle import oracle
    N, d = 32 * world, 64
    port = 29500 + (os.getpid() % 2000) + world + (7 if causal else 0)
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, causal, N, d, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q_out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda t: t[0])
    q, k, v = oracle.set_s((1, 2, N, d), (1, 2, N, d), seeds=(21, 22, 23))
    o_ref, lse_ref, _, _ = oracle.attention(q, k, v, causal=causal)
    if causal:
        o = zigzag_gather([torch.from_numpy(r[1]) for r in results]).numpy()
        lse = zigzag_gather([torch.from_numpy(r[2]) for r in results]).numpy()
    else:
        o = np.concatenate([r[1] for r in results], axis=2)
        lse = np.concatenate([r[2] for r in results], axis=2)
    assert np.abs(o - o_ref).max() < 5e-6
    assert np.abs(lse - lse_ref).max() < 1e-5
