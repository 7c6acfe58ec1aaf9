"""flash_attention_impls_b200 — B200 (sm_100a) FlashAttention forward behind the reference's launch surface.

The product is `lib/libfa_b200.so` (C ABI: include/fa_b200.h).  This package is the thin host-side
mirror of the reference's operator interface plus the multi-GPU drivers ((b,h) sharding, ring attention).
"""
from ._lib import FaB200Error, LIB_PATH, launch_count, load, tmap_cache_stats  # noqa: F401
from .ops import (  # noqa: F401
    FlashAttnFunction,
    HostPipeline,
    attention_backward,
    attention_forward,
    attention_reference_dispatch,
    cast_output,
    combine_partials,
    flash_attention,
    flash_attention_cutlass_dispatch,
    flash_attention_forward,
    flash_attention_forward_dispatch,
    flash_attention_small_tile_dispatch,
    flash_attention_with_stats,
    merge_partial,
)
from .parallel import (  # noqa: F401
    bh_shard_range,
    c_ring_for,
    release_peer_buffers,
    ring_attention,
    zigzag_gather,
    zigzag_split,
)

__all__ = [
    "attention_forward", "attention_backward", "FlashAttnFunction", "flash_attention", "flash_attention_with_stats", "flash_attention_forward",
    "flash_attention_cutlass_dispatch", "flash_attention_forward_dispatch",
    "flash_attention_small_tile_dispatch", "attention_reference_dispatch", "merge_partial", "combine_partials", "cast_output", "HostPipeline",
    "ring_attention", "release_peer_buffers", "zigzag_split", "zigzag_gather", "bh_shard_range", "launch_count", "load",
]
