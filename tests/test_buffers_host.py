"""CPU: the packed output allocation of the batched facades (cm3_b200/_buffers.py) - the layout that
lets *_step_host_packed bring every field to the host with one copy."""
import numpy as np
import torch

from cm3_b200._buffers import FieldDict, _to_int8_host, alloc_fields

SHAPES = dict(grid=(64, 3, 9, 2), vec=(64, 2, 4), obs_self_t=(64, 2, 5, 5, 3), reward=(64,), done=(64,))


def dtype_of(k):
    return torch.uint8 if k == "done" else torch.int8 if k in ("grid", "obs_self_t") else torch.float32


def test_packed_fields_are_gapless_aligned_views_of_one_block():
    out = alloc_fields(SHAPES, dtype_of, packed=True)
    assert isinstance(out, FieldDict) and out.block is not None and list(out) == list(SHAPES)
    base = out.block.data_ptr()
    spans = sorted((v.data_ptr() - base, v.numel() * v.element_size(), k) for k, v in out.items())
    used = sum(n for _, n, _ in spans)
    # gapless fields, then padding to a multiple of 256 bytes (blocks are stacked by rollout_host)
    assert spans[0][0] == 0 and out.block.numel() == (used + 255) // 256 * 256
    assert out.offsets == {k: off for off, _, k in spans}
    for (o0, n0, _), (o1, _, _) in zip(spans, spans[1:]):
        assert o0 + n0 == o1                      # back to back: the block is exactly the fields
    for k, v in out.items():
        assert tuple(v.shape) == SHAPES[k] and v.dtype == dtype_of(k)
        assert (v.data_ptr() - base) % v.element_size() == 0
        assert (v.data_ptr() - base) % 16 == 0    # 64 envs: every field starts on a 16-byte boundary
    out["vec"].fill_(3.0)                         # views alias the block
    assert out.block.view(torch.float32)[:64 * 2 * 4].eq(3.0).all()


def test_unpacked_and_time_major_allocation():
    out = alloc_fields(SHAPES, dtype_of, lead=(5,), packed=False)
    assert out.block is None
    for k, v in out.items():
        assert tuple(v.shape) == (5,) + SHAPES[k] and v.dtype == dtype_of(k) and not v.any()


def test_host_actions_are_saturated_and_int8_passes_through():
    a = np.array([[0, 4], [300, -300]])
    np.testing.assert_array_equal(_to_int8_host(a, (2, 2)), np.array([[0, 4], [127, -128]], dtype=np.int8))
    b = np.arange(4, dtype=np.int8)
    assert np.shares_memory(_to_int8_host(b, (2, 2)), b)


def test_bind_host_to_gpu_is_a_no_op_without_nvml_or_when_disabled(monkeypatch):
    """cm3_b200.sharding.bind_host_to_gpu: an optimisation for multi-socket hosts, never a requirement - without a
    driver (this container) or with CM3_BIND_NUMA=0 it returns 0 and leaves the affinity mask alone."""
    import os
    from cm3_b200.sharding import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    monkeypatch.setenv("CM3_BIND_NUMA", "0")
    assert bind_host_to_gpu(0) == 0
    monkeypatch.delenv("CM3_BIND_NUMA")
    n = bind_host_to_gpu(0)
    assert n == 0 or n == len(os.sched_getaffinity(0))
    if n == 0:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
