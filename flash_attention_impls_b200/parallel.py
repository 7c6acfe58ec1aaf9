"""Multi-GPU drivers for the attention-forward path.  The reference has no
multi-GPU code at all; both modes are new work named by BASELINE.json's north_star.

* (batch, head) sharding — `[B,H,N,d]` is contiguous in b*H+h (reference: flashAttention.cu:30), the
  (b,h) slices are independent, so rank g of P simply owns a contiguous range of them.  No collective.
* ring attention — the sequence is split over the ranks; K/V blocks travel with NCCL send/recv
  (torch.distributed P2P over NVLink/NVSwitch) while each rank runs the local kernel on the block it holds,
  in ring order (step s uses the block of rank r-s), and the per-block partials are merged with their
  logsumexp.  Causal runs use the zig-zag partition (rank r owns
  sequence chunks r and 2P-1-r) so that every rank does the same amount of work at every step.

One process per GPU; the compute calls go to libfa_b200.so through `ops`.  The `backend` argument
exists so the schedule/partition logic can be tested with gloo on CPU against the oracle
(tests/test_ring_cpu.py); the default backend is the CUDA library and refuses CPU tensors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


# ----------------------------------------------------------------------------- (b,h) sharding
def bh_shard_range(BH: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of the (b*H+h) slices owned by `rank`; remainders go to the first ranks."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, rem = divmod(BH, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


# ----------------------------------------------------------------------------- zig-zag partition
def zigzag_split(x: torch.Tensor, world_size: int, rank: int, dim: int = 2) -> torch.Tensor:
    """Rows of `x` along `dim` owned by `rank`: chunks `rank` and `2P-1-rank` of 2P equal chunks."""
    n = x.shape[dim]
    if n % (2 * world_size):
        raise ValueError(f"sequence length {n} must be a multiple of 2*world_size={2 * world_size}")
    c = n // (2 * world_size)
    lo = x.narrow(dim, rank * c, c)
    hi = x.narrow(dim, (2 * world_size - 1 - rank) * c, c)
    return torch.cat([lo, hi], dim=dim).contiguous()


def zigzag_gather(shards, dim: int = 2) -> torch.Tensor:
    """Inverse of zigzag_split given the list of all ranks' shards (rank order)."""
    P = len(shards)
    c = shards[0].shape[dim] // 2
    chunks = [None] * (2 * P)
    for r, s in enumerate(shards):
        chunks[r] = s.narrow(dim, 0, c)
        chunks[2 * P - 1 - r] = s.narrow(dim, c, c)
    return torch.cat(chunks, dim=dim)


# ----------------------------------------------------------------------------- backends
class _CudaBackend:
    """The product path: libfa_b200.so kernels on the current CUDA stream."""

    def attention(self, q, k, v, causal, out, lse):
        from . import ops
        ops.attention_forward(q, k, v, causal=causal, out=out, lse=lse)

    def merge(self, o_acc, lse_acc, o_part, lse_part):
        from . import ops
        ops.merge_partial(o_acc, lse_acc, o_part, lse_part)

    def finalize(self, o_acc, dtype):
        from . import ops
        return ops.cast_output(o_acc, dtype)


def _post_block_exchange(k, v, recv_k, recv_v, step, group, rank, world):
    """Post the transfers of ring step `step`: my own K/V block goes to the rank that needs it at that step,
    (rank + step) mod P, and the block I need, the one owned by (rank - step) mod P, comes in.

    Every GPU has full NVLink bandwidth to every peer through NVSwitch, so the block does not have to hop
    around the ring: each owner sends it directly.  That removes the step-to-step dependency of a forwarding
    ring: all P-1 exchanges are posted up front and run back to back on NCCL's stream while the kernels of
    the earlier steps compute."""
    dst, src = (rank + step) % world, (rank - step) % world
    g_dst = dist.get_global_rank(group, dst) if group is not None else dst
    g_src = dist.get_global_rank(group, src) if group is not None else src
    ops_ = [
        dist.P2POp(dist.isend, k, g_dst, group),
        dist.P2POp(dist.irecv, recv_k, g_src, group),
        dist.P2POp(dist.isend, v, g_dst, group),
        dist.P2POp(dist.irecv, recv_v, g_src, group),
    ]
    return dist.batch_isend_irecv(ops_)


def ring_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False,
                   group: Optional[dist.ProcessGroup] = None, backend=None):
    """Ring attention over the ranks of `group`.

    q, k, v: this rank's shard `[B, H, N_local, d]` of a sequence of length `P * N_local`.
      non-causal: any equal partition of the sequence (rank order is irrelevant to the result);
      causal:     the zig-zag partition (`zigzag_split`): local rows = [chunk r ; chunk 2P-1-r].
    Returns (O_local `[B,H,N_local,d]` in q.dtype, lse_local `[B,H,N_local]` fp32).

    Step s (s = 0..P-1) works on the block owned by rank (r - s) mod P; the P-1 remote blocks are fetched
    with NCCL send/recv posted up front (see _post_block_exchange), so step s never waits on step s-1's
    transfer.  With the zig-zag layout the causal structure per step is one of
      src == r : square causal on the local block
      src <  r : every local query row sees only the FIRST half of the visiting block (no mask)
      src >  r : only the SECOND half of the local query rows see the visiting block (no mask)
    so the kernel is only ever asked for plain or square-causal attention on (strided) row ranges.
    """
    be = backend if backend is not None else _CudaBackend()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B, H, Nl, d = q.shape
    if k.shape != q.shape or v.shape != q.shape:
        raise ValueError("ring_attention expects equally sized local q, k, v shards")
    if causal and Nl % 2:
        raise ValueError("causal ring attention needs an even local length (zig-zag halves)")
    half = Nl // 2

    o_acc = torch.zeros((B, H, Nl, d), dtype=torch.float32, device=q.device)
    lse_acc = torch.full((B, H, Nl), float("-inf"), dtype=torch.float32, device=q.device)
    o_part = torch.empty((B, H, Nl, d), dtype=q.dtype, device=q.device)
    lse_part = torch.empty((B, H, Nl), dtype=torch.float32, device=q.device)

    own_k, own_v = k.contiguous(), v.contiguous()
    blocks = [(own_k, own_v)] + [(torch.empty_like(own_k), torch.empty_like(own_v)) for _ in range(world - 1)]
    reqs = [None] + [_post_block_exchange(own_k, own_v, blocks[s][0], blocks[s][1], s, group, rank, world)
                     for s in range(1, world)]

    for step in range(world):
        src = (rank - step) % world
        if step > 0:
            for r_ in reqs[step]:
                r_.wait()
        cur_k, cur_v = blocks[step]

        if not causal or src == rank:
            be.attention(q, cur_k, cur_v, causal and src == rank, o_part, lse_part)
        elif src < rank:
            # all local queries vs the first half (chunk `src`) of the visiting block
            be.attention(q, cur_k[:, :, :half], cur_v[:, :, :half], False, o_part, lse_part)
        else:
            # only the second half of the local queries (chunk 2P-1-r) sees the visiting block
            lse_part[:, :, :half].fill_(float("-inf"))
            be.attention(q[:, :, half:], cur_k, cur_v, False, o_part[:, :, half:], lse_part[:, :, half:])
        be.merge(o_acc, lse_acc, o_part, lse_part)

    return be.finalize(o_acc, q.dtype), lse_acc
