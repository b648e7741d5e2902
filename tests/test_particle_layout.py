"""Host-side check of the particle kernel's staging layout (cm3_b200/csrc/particle.cu, PtGeom):
the swizzle mode chosen per record size makes the per-env 16-byte stores of a warp conflict-free.
Pure arithmetic - no GPU, no oracle."""
import pytest


def sw_bits_for(chunks):
    """particle.cu: sw_bits_for - 16-byte chunks per env record S = odd * 2^k -> k clamped to 3."""
    return 0 if chunks % 2 else 1 if chunks % 4 else 2 if chunks % 8 else 3


def swizzle(off, bits):
    """TMA 32/64/128-byte swizzle: address bits [4, 4+bits) ^= bits [7, 7+bits)."""
    return off ^ ((off >> 3) & (((1 << bits) - 1) << 4))


@pytest.mark.parametrize("n_agents", [1, 2, 3, 4])
@pytest.mark.parametrize("real_size", [4, 8])
def test_staging_stores_are_bank_conflict_free(n_agents, real_size):
    n_others = max(n_agents - 1, 1)
    for record_bytes in (n_agents * 4 * real_size, n_agents * 4 * n_others * real_size):
        chunks = record_bytes // 16
        bits = sw_bits_for(chunks)
        tile = 32 * record_bytes
        width = (16 << bits) if bits else 128
        assert tile % width == 0 and tile // width <= 256  # TMA box rows
        seen = set()
        for c in range(chunks):          # one 16-byte store instruction per chunk of the record
            for quarter in range(4):     # a 128-bit shared store is processed per quarter-warp
                groups = [(swizzle((lane * chunks + c) * 16, bits) >> 4) & 7
                          for lane in range(quarter * 8, quarter * 8 + 8)]
                assert len(set(groups)) == 8, (n_agents, real_size, record_bytes, c, quarter)
            seen.update(swizzle((lane * chunks + c) * 16, bits) for lane in range(32))
        assert seen == set(range(0, tile, 16))  # a permutation of the tile's chunks
