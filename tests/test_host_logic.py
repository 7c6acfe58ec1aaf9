"""CPU tests of the host-side partition logic used by the multi-GPU drivers."""
import pytest
import torch

from flash_attention_impls_b200.parallel import bh_shard_range, zigzag_gather, zigzag_split


@pytest.mark.parametrize("BH,P", [(128, 1), (128, 2), (128, 8), (10, 4), (3, 8)])
def test_bh_shard_ranges_tile_the_heads(BH, P):
    ranges = [bh_shard_range(BH, P, r) for r in range(P)]
    assert ranges[0][0] == 0 and ranges[-1][1] == BH
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


def test_bh_shard_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        bh_shard_range(8, 2, 2)


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_zigzag_round_trip_and_balance(P):
    N = 16 * P
    x = torch.arange(N, dtype=torch.float32).view(1, 1, N, 1).expand(1, 2, N, 3).contiguous()
    shards = [zigzag_split(x, P, r) for r in range(P)]
    assert torch.equal(zigzag_gather(shards), x)
    # causal work per rank (number of visible (row, key) pairs) is identical across ranks
    work = []
    for s in shards:
        rows = s[0, 0, :, 0]
        work.append(int((rows + 1).sum().item()))
    assert len(set(work)) == 1


def test_zigzag_requires_divisibility():
    with pytest.raises(ValueError):
        zigzag_split(torch.zeros(1, 1, 10, 4), 4, 0)
