"""Multi-GPU drivers for the attention-forward path.  The reference has no
multi-GPU code at all; both modes are new work named by BASELINE.json's north_star.

* (batch, head) sharding — `[B,H,N,d]` is contiguous in b*H+h (reference: flashAttention.cu:30), the
  (b,h) slices are independent, so rank g of P simply owns a contiguous range of them.  No collective.
* ring attention — the sequence is split over the ranks; each rank runs the local kernel on one K/V block
  at a time, in ring order (step s uses the block of rank r-s); every step writes its partial (O_s, lse_s) into
  slot s of a stacked buffer and ONE pass at the end merges the P partials with their logsumexp (2 bytes read per
  element and partial, instead of a 10-byte read-modify-write of an fp32 accumulator after every step).  The blocks
  travel over NVLink/NVSwitch in one of two ways: `transport="peer"` (default on a single node): the C-ABI ring
  (`fa_b200_ring_*`, csrc/fa_ring.cu) - every rank publishes its block in a CUDA-IPC buffer and the others PULL it
  with the copy engines through a window of two receive slots, ordered by interprocess events, which needs no SM
  and no collective; `transport="p2p"`: NCCL send/recv through torch.distributed (also what the CPU/gloo tests of the
  schedule use).  Causal runs use the zig-zag partition (rank r owns
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


class _CRing:
    """`transport="peer"`: the ring driver behind the C ABI (`fa_b200_ring_*`, csrc/fa_ring.cu).

    One handle per (shard shape, dtype, group, device), cached in `_c_rings`.  Creation is the only place where the
    ranks talk on the host: every rank creates its handle, and ONE all-gather carries (hostname, status, 128-byte
    export blob), so that all ranks take the same decision - the peer transport is used only when every rank sits
    on the same host and every create and every connect succeeded; otherwise every rank destroys its handle and the
    group falls back to NCCL send/recv.  After that a forward is enqueue-only on the device: copy-engine pulls through
    a window of two receive slots, ordered by interprocess events (no collective; the hosts only exchange sequence
    counters through shared memory)."""

    def __init__(self, shape, dtype, group, device):
        import ctypes
        import socket
        from . import _lib
        self.lib, self.ct = _lib, ctypes
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.shape, self.dtype = tuple(shape), dtype
        self.handle = ctypes.c_void_p()
        self.ok = False
        B, H, Nl, d = self.shape
        code = _lib.FA_B200_BF16 if dtype == torch.bfloat16 else _lib.FA_B200_FP16
        lib = _lib.load()
        blob = ctypes.create_string_buffer(_lib.FA_B200_RING_EXPORT_BYTES)
        with torch.cuda.device(device):
            st = lib.fa_b200_ring_create(self.world, self.rank, B, H, Nl, d, code, ctypes.byref(self.handle))
            err = lib.fa_b200_last_error().decode() if st else ""
            if st == 0:
                st = lib.fa_b200_ring_export(self.handle, blob)
            info = [None] * self.world
            dist.all_gather_object(info, (socket.gethostname(), int(st), blob.raw, err), group=group)
            good = all(i[1] == 0 for i in info) and len({i[0] for i in info}) == 1
            st2 = -1
            if good:
                st2 = lib.fa_b200_ring_connect(self.handle, b"".join(i[2] for i in info))
            flags = [None] * self.world
            dist.all_gather_object(flags, int(st2), group=group)
            self.ok = good and all(f == 0 for f in flags)
            if not self.ok:
                self.why = next((i[3] for i in info if i[1]), "") or ("ranks span several hosts" if len({i[0] for i in info}) > 1
                                                                      else "fa_b200_ring_connect failed on a rank")
                self.close()

    def forward(self, q, k, v, causal, softmax_scale=0.0):
        out = torch.empty_like(q)
        lse = torch.empty(q.shape[:3], dtype=torch.float32, device=q.device)
        with torch.cuda.device(self.device):
            self.lib.check(self.lib.load().fa_b200_ring_forward(
                self.handle, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(),
                1 if causal else 0, float(softmax_scale), torch.cuda.current_stream(self.device).cuda_stream))
        return out, lse

    def set_profile(self, on: bool):
        self.lib.check(self.lib.load().fa_b200_ring_set_profile(self.handle, 1 if on else 0))

    def timeline(self):
        """[(step, begin_ms, kv_ready_ms, attn_done_ms), ...], combine_done_ms of the last (synchronised) forward."""
        buf = (self.ct.c_float * (3 * self.world + 2))()
        n = self.lib.load().fa_b200_ring_timeline(self.handle, buf, len(buf))
        vals = [float(buf[i]) for i in range(max(n, 0))]
        steps = [(s, vals[3 * s], vals[3 * s + 1], vals[3 * s + 2]) for s in range(self.world) if 3 * s + 2 < len(vals)]
        return steps, (vals[-1] if vals else None)

    def device_bytes(self) -> int:
        return int(self.lib.load().fa_b200_ring_device_bytes(self.handle))

    def close(self):
        """Collective in effect: no rank may still be pulling (callers synchronise the ranks first)."""
        if self.handle:
            with torch.cuda.device(self.device):
                self.lib.load().fa_b200_ring_destroy(self.handle)
            self.handle = self.ct.c_void_p()


_c_rings = {}


def release_peer_buffers(group: Optional[dist.ProcessGroup] = None) -> None:
    """Destroys the ring handles the "peer" transport caches per (shard shape, dtype, group, device).  Collective
    over `group`; call it before destroying the process group if the buffers should not live until process exit."""
    if not _c_rings:
        return
    torch.cuda.synchronize()
    dist.barrier(group)
    for key in [k for k in _c_rings if k[2] == id(group)]:
        _c_rings.pop(key).close()


def c_ring_for(q: torch.Tensor, group) -> "_CRing":
    """The cached C-ABI ring handle for shards shaped like `q` (created collectively on first use)."""
    key = (tuple(q.shape), q.dtype, id(group), q.device.index)
    ring = _c_rings.get(key)
    if ring is None:
        ring = _c_rings[key] = _CRing(q.shape, q.dtype, group, q.device)
    return ring


def ring_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False,
                   group: Optional[dist.ProcessGroup] = None, backend=None, transport: str = "auto"):
    """Ring attention over the ranks of `group`.

    q, k, v: this rank's shard `[B, H, N_local, d]` of a sequence of length `P * N_local`.
      non-causal: any equal partition of the sequence (rank order is irrelevant to the result);
      causal:     the zig-zag partition (`zigzag_split`): local rows = [chunk r ; chunk 2P-1-r].
    Returns (O_local `[B,H,N_local,d]` in q.dtype, lse_local `[B,H,N_local]` fp32).

    Step s (s = 0..P-1) works on the block owned by rank (r - s) mod P and leaves its partial in slot s; one
    `combine_partials` pass merges the slots at the end.  transport: "peer" = the C-ABI ring (`fa_b200_ring_*`):
    copy-engine pulls from the owner's CUDA-IPC buffer through a window of two receive slots, flag-ordered, no SM
    used, everything preallocated in the handle (single node); "p2p" = torch.distributed send/recv (NCCL or gloo; all
    P-1 exchanges posted up front); "auto" = "peer" for CUDA tensors with the built-in backend when every rank of
    the group is on one host and can map its peers (decided collectively, once), else "p2p".  With the zig-zag
    layout the causal structure per step is one of
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
    if transport not in ("auto", "peer", "p2p"):
        raise ValueError("transport must be 'auto', 'peer' or 'p2p'")
    if transport in ("auto", "peer") and backend is None and q.is_cuda and world > 1:
        ring = c_ring_for(q, group)
        if ring.ok:
            return ring.forward(q.contiguous(), k.contiguous(), v.contiguous(), causal)
        if transport == "peer":
            raise RuntimeError(f"ring_attention(transport='peer') is not available for this group: {ring.why}")
    elif transport == "peer" and world > 1:
        raise ValueError("transport='peer' needs CUDA tensors and the built-in backend")

    # ---- "p2p": NCCL / gloo send-recv schedule (also the host logic the CPU tests drive with the oracle backend)
    # slot s holds the partial of step s; a slot row that a step does not compute keeps lse = -inf and is skipped
    o_parts = torch.empty((world, B, H, Nl, d), dtype=q.dtype, device=q.device)
    lse_parts = torch.full((world, B, H, Nl), float("-inf"), dtype=torch.float32, device=q.device)

    own_k, own_v = k.contiguous(), v.contiguous()
    blocks = [(own_k, own_v)] + [(torch.empty_like(own_k), torch.empty_like(own_v)) for _ in range(world - 1)]
    reqs = None
    if world > 1:
        reqs = [None] + [_post_block_exchange(own_k, own_v, blocks[s][0], blocks[s][1], s, group, rank, world)
                         for s in range(1, world)]

    prof = _RingProfile(q.device) if os.environ.get("FA_B200_RING_PROFILE") else None
    for step in range(world):
        src = (rank - step) % world
        if prof:
            prof.mark(f"s{step}:begin")
        if step > 0:
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
