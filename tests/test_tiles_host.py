"""CPU tests of the packed 2-bit tile decoders (cm3_b200/tiles.py; layout: include/cm3env.h, CM3_TILE_U2):
a test-side encoder written from the header's bit layout, decoded by the NumPy and the torch paths."""
import numpy as np
import pytest
import torch

from cm3_b200.tiles import unpack_grid_u2, unpack_window_u2


def _code(v):
    return np.where(v < 0, 3, v).astype(np.uint64)   # 0 -> 0, +1 -> 1, -1 -> 3 (two's complement in 2 bits)


def pack_window(win):
    """[..., W, W, 3] in {-1, 0, 1} -> [..., W, RW] uint32: cell (dc, ch) of a row at bits 2 (3 dc + ch)."""
    W = win.shape[-2]
    rw = (6 * W + 31) // 32
    flat = _code(win.reshape(win.shape[:-2] + (3 * W,)))              # [..., W, 3W]
    k = np.arange(3 * W, dtype=np.uint64)
    acc = (flat << (2 * k)).sum(axis=-1, dtype=np.uint64) if 6 * W <= 64 else None
    words = np.stack([(acc >> np.uint64(32 * i)) & np.uint64(0xFFFFFFFF) for i in range(rw)], axis=-1)
    return words.astype(np.uint32)


def pack_grid(grid):
    """[..., R, C + 1, 2] -> [..., R, GW] uint32: cell (j, ch) at bits 2 (2 j + ch), 8 cells per word."""
    nc = grid.shape[-2]
    gw = (nc + 7) // 8
    out = np.zeros(grid.shape[:-2] + (gw,), dtype=np.uint64)
    code = _code(grid)
    for j in range(nc):
        for ch in range(2):
            out[..., j // 8] |= code[..., j, ch] << np.uint64(4 * (j % 8) + 2 * ch)
    return out.astype(np.uint32)


@pytest.mark.parametrize("n_obs", [1, 2, 3])
def test_window_round_trip(n_obs):
    rng = np.random.default_rng(n_obs)
    W = 2 * n_obs + 1
    win = rng.integers(-1, 2, size=(7, 3, W, W, 3)).astype(np.int8)
    words = pack_window(win)
    assert words.shape == (7, 3, W, (6 * W + 31) // 32)
    got = unpack_window_u2(words, n_obs)
    assert got.dtype == np.int8
    np.testing.assert_array_equal(got, win)
    # the kernel hands the words out as int32: same bits
    got_t = unpack_window_u2(torch.from_numpy(words.view(np.int32)), n_obs)
    assert got_t.dtype == torch.int8
    np.testing.assert_array_equal(got_t.numpy(), win)


@pytest.mark.parametrize("n_columns", [2, 8, 16, 20, 28])
def test_grid_round_trip(n_columns):
    rng = np.random.default_rng(n_columns)
    grid = rng.integers(-1, 2, size=(5, 3, n_columns + 1, 2)).astype(np.int8)
    words = pack_grid(grid)
    assert words.shape == (5, 3, (n_columns + 1 + 7) // 8)
    np.testing.assert_array_equal(unpack_grid_u2(words, n_columns), grid)
    np.testing.assert_array_equal(unpack_grid_u2(torch.from_numpy(words.view(np.int32)), n_columns).numpy(), grid)


def test_known_word():
    """One window row of the stage-2 board written out by hand: cells (+1, 0, -1), (0, 0, 0), (-1, -1, +1), 0 ..."""
    row = np.zeros((5, 3), dtype=np.int8)
    row[0] = (1, 0, -1)
    row[2] = (-1, -1, 1)
    word = (1 << 0) | (3 << 4) | (3 << 12) | (3 << 14) | (1 << 16)
    win = np.zeros((5, 5, 3), dtype=np.int8)
    win[1] = row
    words = np.zeros((5, 1), dtype=np.uint32)
    words[1, 0] = word
    np.testing.assert_array_equal(unpack_window_u2(words, 2), win)
