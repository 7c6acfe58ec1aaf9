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


# ----------------------------------------------------------------------------- tile scheduler order
@pytest.mark.parametrize("B,H,N,Nkv,d,causal,group", [
    (4, 32, 8192, 0, 128, False, None), (4, 32, 8192, 0, 128, True, None), (2, 24, 1100, 0, 128, True, "5"),
    (1, 7, 520, 0, 64, True, "3"), (3, 5, 300, 900, 64, True, None), (1, 9, 900, 300, 128, True, "4"),
    (1, 1, 1, 0, 64, True, None),
])
def test_work_item_list_is_a_permutation_with_the_right_trip_counts(monkeypatch, B, H, N, Nkv, d, causal, group):
    """fa_b200_work_item decodes the kernel's own item mapping (get_item): every (b*H+h, q-block) exactly once,
    trip counts that skip the K/V tiles above the causal diagonal, and - for causal launches - items ordered
    longest-first inside each group of heads."""
    from flash_attention_impls_b200 import _lib
    # the environment variable is only read once per process; the knob itself is a C-ABI call
    _lib.load().fa_b200_set_group_heads(int(group) if group is not None else 0)
    nkv = Nkv or N
    n_items = _lib.work_item(B, H, N, d, causal, 0, Nkv)[0]
    nqb = (N + 255) // 256
    assert n_items == B * H * nqb
    seen, per_group_cost = set(), {}
    items = [_lib.work_item(B, H, N, d, causal, i, Nkv)[1:] for i in range(n_items)]
    for idx, (bh, q0, t0, t1) in enumerate(items):
        assert 0 <= bh < B * H and q0 % 256 == 0 and 0 <= q0 < nqb * 256
        assert (bh, q0) not in seen
        seen.add((bh, q0))
        for tile, t in ((0, t0), (1, t1)):
            r0 = q0 + 128 * tile
            if r0 >= N:
                want = 0
            elif not causal:
                want = (nkv + 127) // 128
            else:
                last_col = min(r0 + 127, N - 1) + (nkv - N)
                want = 0 if last_col < 0 else min((nkv + 127) // 128, last_col // 128 + 1)
            assert t == want
    assert len(seen) == n_items
    if causal:
        # inside a group the q-blocks are visited in descending order (heaviest first)
        g = int(group) if group else max(1, min(B * H, (64 << 20) // (4 * nkv * d)))
        per = g * nqb
        for start in range(0, n_items, per):
            q0s = [q0 for (_, q0, _, _) in items[start:start + per]]
            assert q0s == sorted(q0s, reverse=True)
            heads = {bh for (bh, _, _, _) in items[start:start + per]}
            assert len(heads) <= g and max(heads) - min(heads) < g


def test_work_item_rejects_bad_arguments():
    from flash_attention_impls_b200 import _lib
    with pytest.raises(_lib.FaB200Error):
        _lib.work_item(1, 1, 128, 64, False, 5)
    with pytest.raises(_lib.FaB200Error):
        _lib.work_item(1, 1, 128, 44, False, 0)
