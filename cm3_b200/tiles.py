"""Decoders of the packed 2-bit tile encoding (include/cm3env.h: CM3_TILE_U2; VecCheckers(tile_dtype="u2")).

Every cell of the two bulky Checkers outputs - the per-agent window `obs_self_t` (get_obs,
env/checkers.py:97-109) and the global grid (get_valid_grid, :66-76) - holds -1, 0 or +1; the packed
form stores it in 2 bits (0 -> 0, 1 -> +1, 3 -> -1, i.e. two's complement), least significant cell
first, one window row / 8 grid cells per 32-bit word.  These functions turn the words back into the
reference's arrays; they work on NumPy arrays and on torch tensors (any device) alike.
"""
import numpy as np


def _codes(words, n_cells, is_torch):
    """words [..., nw] (32-bit) -> values [..., n_cells] in {-1, 0, 1} (int8), cell k at bits 2 k."""
    if is_torch:
        import torch
        w = words.to(torch.int64) & 0xFFFFFFFF
        k = torch.arange(n_cells, device=words.device)
        word = w[..., (k // 16)]
        code = (word >> (2 * (k % 16))) & 3
        return (((code + 2) & 3) - 2).to(torch.int8)
    w = np.asarray(words).astype(np.int64) & 0xFFFFFFFF
    k = np.arange(n_cells)
    code = (w[..., k // 16] >> (2 * (k % 16))) & 3
    return (((code + 2) & 3) - 2).astype(np.int8)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def unpack_window_u2(words, n_obs):
    """obs_self_t words [..., W, RW] -> [..., W, W, 3] int8 (W = 2 n_obs + 1)."""
    W = 2 * n_obs + 1
    t = _is_torch(words)
    vals = _codes(words, 3 * W, t)            # [..., W, 3 W]: one row's cells in (dc, ch) order
    return vals.reshape(tuple(vals.shape[:-1]) + (W, 3))


def unpack_grid_u2(words, n_columns):
    """grid words [..., R, GW] -> [..., R, n_columns + 1, 2] int8."""
    t = _is_torch(words)
    nc = n_columns + 1
    w = words
    if t:
        import torch
        ww = w.to(torch.int64) & 0xFFFFFFFF
        j = torch.arange(nc, device=w.device)
        word = ww[..., (j // 8)]                                      # [..., R, nc]
        sh = (4 * (j % 8)).unsqueeze(-1) + 2 * torch.arange(2, device=w.device)   # [nc, 2]
        code = (word.unsqueeze(-1) >> sh) & 3
        return (((code + 2) & 3) - 2).to(torch.int8)
    ww = np.asarray(w).astype(np.int64) & 0xFFFFFFFF
    j = np.arange(nc)
    word = ww[..., j // 8]
    sh = (4 * (j % 8))[:, None] + 2 * np.arange(2)
    code = (word[..., None] >> sh) & 3
    return (((code + 2) & 3) - 2).astype(np.int8)
