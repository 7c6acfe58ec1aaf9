"""Multi-GPU drivers for the attention-forward path.  The reference has no
multi-GPU code at all; both modes are new work named by BASELINE.json's north_star.

* (batch, head) sharding — `[B,H,N,d]` is contiguous in b*H+h (reference: flashAttention.cu:30), the
  (b,h) slices are independent, so rank g of P simply owns a contiguous range of them.  No collective.
* ring attention — the sequence is split over the ranks; each rank runs the local kernel on one K/V block
  at a time, in ring order (step s uses the block of rank r-s); every step writes its partial (O_s, lse_s) into
  slot s of a stacked buffer and ONE pass at the end merges the P partials with their logsumexp (2 bytes read per
  element and partial, instead of a 10-byte read-modify-write of an fp32 accumulator after every step).  The blocks travel over NVLink/NVSwitch in one of two ways: `transport="peer"` (default on a
  single node): every rank publishes its block in a CUDA-IPC buffer and the others PULL it with the copy
  engines, which needs no SM; `transport="p2p"`: NCCL send/recv through torch.distributed (also what the CPU/gloo
  tests of the schedule use).  Causal runs use the zig-zag partition (rank r owns
  sequence chunks r and 2P-1-r) so that every rank does the same amount of work at every step.

One process per GPU; the compute calls go to libfa_b200.so through `ops`.  The `backend` argument
exists so the schedule/partition logic can be tested with gloo on CPU against the oracle
(tests/test_ring_cpu.py); the default backend is the CUDA library and refuses CPU tensors.
"""
from __future__ import annotations

import os
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

    def combine(self, o_parts, lse_parts):
        from . import ops
        return ops.combine_partials(o_parts, lse_parts)


class _RingProfile:
    """FA_B200_RING_PROFILE=1: CUDA-event timeline of one ring_attention call (diagnostics only)."""

    def __init__(self, device):
        self.marks = []
        self.device = device
        self.mark("start")

    def mark(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.marks.append((name, ev))

    def report(self, rank):
        torch.cuda.synchronize(self.device)
        t0 = self.marks[0][1]
        line = " ".join(f"{n}={t0.elapsed_time(e):.2f}" for n, e in self.marks[1:])
        print(f"[ring profile rank {rank}] ms since start: {line}", flush=True)


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


class _PeerRing:
    """K/V exchange over peer memory: CUDA-IPC buffers + copy-engine pulls (see include/fa_b200.h, peer section).

    One instance per (block bytes, group, device), cached in `_peer_rings`.  Each rank owns two publish buffers
    (K|V, double-buffered across calls) that every other rank of the node has mapped.  A call writes the local block
    into publish buffer `n % 2`, runs one tiny all-reduce as the cross-rank "everything is published" barrier, then
    pulls the P-1 remote blocks on a side stream, one event per block.  Because call n+1 uses the other buffer and
    its barrier is stream-ordered after call n's compute, a buffer is only overwritten two calls later, when every
    peer has finished pulling from it."""

    def __init__(self, block_bytes: int, group, device):
        import ctypes
        from . import _lib
        self.lib, self.ct = _lib, ctypes
        self.group, self.device, self.block_bytes = group, device, block_bytes
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.calls = 0
        self.mine, handles = [], []
        with torch.cuda.device(device):
            for _ in range(2):
                ptr = ctypes.c_void_p()
                h = ctypes.create_string_buffer(64)
                _lib.check(_lib.load().fa_b200_peer_alloc(2 * block_bytes, ctypes.byref(ptr), h))
                self.mine.append(ptr.value)
                handles.append(h.raw)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, handles, group=group)
            self.peers = []            # peers[r][buf] -> device pointer of rank r's publish buffer
            for r, hs in enumerate(gathered):
                if r == self.rank:
                    self.peers.append(list(self.mine))
                    continue
                ptrs = []
                for raw in hs:
                    p_ = ctypes.c_void_p()
                    _lib.check(_lib.load().fa_b200_peer_open(raw, ctypes.byref(p_)))
                    ptrs.append(p_.value)
                self.peers.append(ptrs)
            self.copy_stream = torch.cuda.Stream(device)
            self.flag = torch.zeros(1, dtype=torch.float32, device=device)

    def exchange(self, k, v, blocks):
        """Publishes (k, v), then starts pulling blocks[s] <- rank (r - s) mod P for s = 1..P-1.
        Returns one CUDA event per step (None for step 0)."""
        lib, buf = self.lib.load(), self.calls % 2
        self.calls += 1
        cur = torch.cuda.current_stream(self.device)
        nb = self.block_bytes
        with torch.cuda.device(self.device):
            self.lib.check(lib.fa_b200_copy_async(self.mine[buf], k.data_ptr(), nb, cur.cuda_stream))
            self.lib.check(lib.fa_b200_copy_async(self.mine[buf] + nb, v.data_ptr(), nb, cur.cuda_stream))
            dist.all_reduce(self.flag, group=self.group)          # every rank's block is published
            self.copy_stream.wait_stream(cur)
            events = [None]
            for s in range(1, self.world):
                src = self.peers[(self.rank - s) % self.world][buf]
                bk, bv = blocks[s]
                self.lib.check(lib.fa_b200_copy_async(bk.data_ptr(), src, nb, self.copy_stream.cuda_stream))
                self.lib.check(lib.fa_b200_copy_async(bv.data_ptr(), src + nb, nb, self.copy_stream.cuda_stream))
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                events.append(ev)
        return events


    def close(self):
        """Unmaps the peers' buffers and frees this rank's publish buffers (collective: every rank must call it,
        after a barrier, so that nobody is still pulling)."""
        lib = self.lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for r, ptrs in enumerate(self.peers):
                if r != self.rank:
                    for p_ in ptrs:
                        lib.fa_b200_peer_close(p_)
            for p_ in self.mine:
                lib.fa_b200_peer_free(p_)
        self.peers, self.mine = [], []


_peer_rings = {}


def release_peer_buffers(group: Optional[dist.ProcessGroup] = None) -> None:
    """Frees the CUDA-IPC publish buffers the "peer" transport caches per (block size, group, device).  Collective
    over `group`; call it before destroying the process group if the buffers should not live until process exit."""
    if not _peer_rings:
        return
    dist.barrier(group)
    for key in [k for k in _peer_rings if k[1] == id(group)]:
        _peer_rings.pop(key).close()


def _peer_ring_for(k, group):
    key = (k.numel() * k.element_size(), id(group), k.device.index)
    ring = _peer_rings.get(key)
    if ring is None:
        ring = _peer_rings[key] = _PeerRing(key[0], group, k.device)
    return ring


def ring_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False,
                   group: Optional[dist.ProcessGroup] = None, backend=None, transport: str = "auto"):
    """Ring attention over the ranks of `group`.

    q, k, v: this rank's shard `[B, H, N_local, d]` of a sequence of length `P * N_local`.
      non-causal: any equal partition of the sequence (rank order is irrelevant to the result);
      causal:     the zig-zag partition (`zigzag_split`): local rows = [chunk r ; chunk 2P-1-r].
    Returns (O_local `[B,H,N_local,d]` in q.dtype, lse_local `[B,H,N_local]` fp32).

    Step s (s = 0..P-1) works on the block owned by rank (r - s) mod P and leaves its partial in slot s; one
    `combine_partials` pass merges the slots at the end; all P-1 remote blocks are requested up
    front, so step s never waits on step s-1's transfer.  transport: "peer" = copy-engine pulls from CUDA-IPC
    buffers (single node, no SM used), "p2p" = torch.distributed send/recv (NCCL or gloo), "auto" = "peer" for
    CUDA tensors with the built-in backend, else "p2p".  With the zig-zag layout the causal structure per step
    is one of
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

    # slot s holds the partial of step s; a slot row that a step does not compute keeps lse = -inf and is skipped
    o_parts = torch.empty((world, B, H, Nl, d), dtype=q.dtype, device=q.device)
    lse_parts = torch.full((world, B, H, Nl), float("-inf"), dtype=torch.float32, device=q.device)

    own_k, own_v = k.contiguous(), v.contiguous()
    blocks = [(own_k, own_v)] + [(torch.empty_like(own_k), torch.empty_like(own_v)) for _ in range(world - 1)]
    if transport == "auto":
        transport = "peer" if (backend is None and q.is_cuda and world > 1) else "p2p"
    if transport not in ("peer", "p2p"):
        raise ValueError("transport must be 'auto', 'peer' or 'p2p'")
    reqs, events = None, None
    if world > 1 and transport == "peer":
        events = _peer_ring_for(own_k, group).exchange(own_k, own_v, blocks)
    elif world > 1:
        reqs = [None] + [_post_block_exchange(own_k, own_v, blocks[s][0], blocks[s][1], s, group, rank, world)
                         for s in range(1, world)]

    prof = _RingProfile(q.device) if os.environ.get("FA_B200_RING_PROFILE") else None
    for step in range(world):
        src = (rank - step) % world
        if prof:
            prof.mark(f"s{step}:begin")
        if step > 0:
            if events is not None:
                torch.cuda.current_stream(q.device).wait_event(events[step])
            else:
                for r_ in reqs[step]:
                    r_.wait()
        if prof:
            prof.mark(f"s{step}:kv_ready")
        cur_k, cur_v = blocks[step]
        o_part, lse_part = o_parts[step], lse_parts[step]

        if not causal or src == rank:
            be.attention(q, cur_k, cur_v, causal and src == rank, o_part, lse_part)
        elif src < rank:
            # all local queries vs the first half (chunk `src`) of the visiting block
            be.attention(q, cur_k[:, :, :half], cur_v[:, :, :half], False, o_part, lse_part)
        else:
            # only the second half of the local queries (chunk 2P-1-r) sees the visiting block
            be.attention(q[:, :, half:], cur_k, cur_v, False, o_part[:, :, half:], lse_part[:, :, half:])
        if prof:
            prof.mark(f"s{step}:attn_done")

    if world == 1:
        out, lse = o_parts[0], lse_parts[0]
    else:
        out, lse = be.combine(o_parts, lse_parts)
    if prof:
        prof.mark("combine_done")
        prof.report(rank)
    return out, lse
