"""Output buffer allocation shared by the batched facades.

The single-step output set of an env (and its pinned host mirror) is carved out of ONE
allocation with the fields back to back in the same order, so that cm3_*_step_host can bring
every field to the host with a single device-to-host copy instead of one per field (the call
is PCIe-bound; per-copy set-up time is pure overhead)."""
import numpy as np
import torch


def _to_int8_host(actions, shape):
    """Host actions as a C-contiguous int8 array of `shape`; values are saturated like the
    facades' device path.  int8 input is passed through without a conversion pass."""
    a = np.asarray(actions)
    if a.dtype != np.int8:
        a = np.clip(a, -128, 127).astype(np.int8)
    return a.reshape(shape)


class FieldDict(dict):
    """field -> tensor; `.block` is the byte tensor all fields are views of (packed) or None, and
    `.offsets` the byte offset of every field inside it."""
    block = None
    offsets = None


def alloc_fields(shapes, dtype_of, lead=(), device=None, pinned=False, packed=False):
    """dict field -> zeroed tensor of shape lead + shapes[field].

    packed: all fields are views of one byte block (`.block` of the result), laid out in order of
    decreasing element size, every field starting on a 16-byte boundary (the kernels' vector
    stores and TMA bulk copies need that; for batches that are a multiple of 32 envs every field
    size already is a multiple of 16 and the layout is gapless); the block is padded to a
    multiple of 256 bytes so that blocks can be stacked."""
    lead = tuple(lead)
    if not packed:
        out = FieldDict()
        for k, shp in shapes.items():
            t = torch.zeros(lead + tuple(shp), dtype=dtype_of(k), device=None if pinned else device)
            out[k] = t.pin_memory() if pinned else t
        return out
    order = sorted(shapes, key=lambda k: -torch.empty((), dtype=dtype_of(k)).element_size())
    sizes = {k: int(np.prod(lead + tuple(shapes[k]))) * torch.empty((), dtype=dtype_of(k)).element_size()
             for k in order}
    offsets, off = {}, 0
    for k in order:
        offsets[k] = off
        off += (sizes[k] + 15) // 16 * 16
    total = (off + 255) // 256 * 256
    block = torch.zeros(total, dtype=torch.uint8, device=None if pinned else device)
    if pinned:
        block = block.pin_memory()
    views = {k: block[offsets[k]:offsets[k] + sizes[k]].view(dtype_of(k)).view(lead + tuple(shapes[k])) for k in order}
    out = FieldDict((k, views[k]) for k in shapes)
    out.block = block
    out.offsets = offsets
    return out


class HostRolloutBuffers(object):
    """Everything cm3_*_rollout_host needs for T steps of one env facade: two packed device output
    sets, a pinned host area of T stacked output blocks with per-field [T, B, ...] views into it, a
    pinned [T, B, N] action array and its two device slots."""

    def __init__(self, env, T, outputs_struct):
        import ctypes as C
        self.T = T = int(T)
        self.dev = [env.alloc_outputs(), env.alloc_outputs()]
        self.block_bytes = self.dev[0].block.numel()
        self.host = torch.zeros(T, self.block_bytes, dtype=torch.uint8).pin_memory()
        shapes = env.field_shapes()
        self.views = {}
        for k, off in self.dev[0].offsets.items():
            dt = env.field_dtype(k)
            n = int(np.prod(shapes[k])) * dt.itemsize
            self.views[k] = self.host[:, off:off + n].view(dt).unflatten(1, tuple(shapes[k])).numpy()
        self.actions_host = torch.zeros(T, env.B, env.N, dtype=torch.int8).pin_memory()
        self.actions_dev = torch.zeros(2, env.B, env.N, dtype=torch.int8, device=env.device)
        self.outs_c = (type(outputs_struct(self.dev[0])) * 2)(outputs_struct(self.dev[0]), outputs_struct(self.dev[1]))
        self.blocks_c = (C.c_void_p * 2)(self.dev[0].block.data_ptr(), self.dev[1].block.data_ptr())
