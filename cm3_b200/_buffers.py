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
    """field -> tensor; `.block` is the byte tensor all fields are views of (packed) or None."""
    block = None


def alloc_fields(shapes, dtype_of, lead=(), device=None, pinned=False, packed=False):
    """dict field -> zeroed tensor of shape lead + shapes[field].

    packed: all fields are views of one byte block (`.block` of the result), laid out in order of
    decreasing element size (so every view is naturally aligned) with no gaps."""
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
    total = sum(sizes.values())
    block = torch.zeros(total, dtype=torch.uint8, device=None if pinned else device)
    if pinned:
        block = block.pin_memory()
    views, off = {}, 0
    for k in order:
        views[k] = block[off:off + sizes[k]].view(dtype_of(k)).view(lead + tuple(shapes[k]))
        off += sizes[k]
    out = FieldDict((k, views[k]) for k in shapes)
    out.block = block
    return out
