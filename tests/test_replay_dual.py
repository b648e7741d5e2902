"""CPU: the dual (good / bad episode) replay buffer and the episode router against the reference's
semantics (alg/replay_buffer_dual.py:13-63, fed per finished episode by alg/train_onpolicy.py:329-356).
The checker is the reference's own class when the tree is present (build container), else the
restatement below, which the container run pins against it."""
import os
import random
import sys

import numpy as np
import pytest
import torch

from cm3_b200.replay import DeviceDualReplayBuffer, EpisodeRouter


class RestatedDual(object):
    """replay_buffer_dual.py restated (lists of transitions; same branches, same arithmetic)."""

    def __init__(self, size=5e4):
        self.memory_1, self.memory_2, self.maxsize, self.idx_1, self.idx_2 = [], [], int(size), 0, 0

    def add(self, episode, is_bad=False):
        for tr in episode:
            mem = self.memory_1 if is_bad else self.memory_2
            idx = self.idx_1 if is_bad else self.idx_2
            if idx >= len(mem):
                mem.append(tr)
            else:
                mem[idx] = tr
            if is_bad:
                self.idx_1 = (idx + 1) % self.maxsize
            else:
                self.idx_2 = (idx + 1) % self.maxsize

    def sample_counts(self, size):
        half = int(size / 2.0)
        n1, n2 = len(self.memory_1), len(self.memory_2)
        if half <= n1 and half > n2:
            return min(n1, size - n2), n2
        if half > n1 and half <= n2:
            return n1, min(n2, size - n1)
        if n1 < half and n2 < half:
            return n1, n2
        return half, half


def reference_dual():
    path = "/root/reference/alg"
    if not os.path.isdir(path):
        return None
    sys.path.insert(0, path)
    try:
        import importlib
        return importlib.import_module("replay_buffer_dual").Replay_Buffer
    finally:
        sys.path.remove(path)


def synthetic_rollout(T, B, seed, max_steps=9):
    """done / collisions-latch arrays with the kernel's semantics: collisions[t] is the running count
    of the episode that step t belongs to; it restarts after a done step."""
    rng = np.random.default_rng(seed)
    done = np.zeros((T, B), dtype=np.uint8)
    coll = np.zeros((T, B), dtype=np.int32)
    steps, run = np.zeros(B, dtype=int), np.zeros(B, dtype=int)
    for t in range(T):
        steps += 1
        run += 2 * (rng.random(B) < 0.08)
        coll[t] = run
        d = (steps == max_steps) | (rng.random(B) < 0.07)
        done[t] = d
        steps[d] = 0
        run[d] = 0
    return done, coll


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_router_files_every_transition_under_its_episodes_flag(seed):
    T, B, blocks = 17, 6, 5
    done, coll = synthetic_rollout(T * blocks, B, seed)
    uid = np.arange(T * blocks * B).reshape(T * blocks, B)          # a transition's identity
    Ref = reference_dual() or RestatedDual
    ref = Ref(size=10 ** 6)
    # the trainer's loop, one env at a time: buf_episode, then add(buf_episode, collisions != 0)
    for b in range(B):
        ep = []
        for t in range(T * blocks):
            ep.append(int(uid[t, b]))
            if done[t, b]:
                ref.add(ep, coll[t, b] != 0)
                ep = []
    router, buf = EpisodeRouter(), DeviceDualReplayBuffer(size=10 ** 6)
    for k in range(blocks):
        sl = slice(k * T, (k + 1) * T)
        tr = {"uid": torch.from_numpy(uid[sl].copy()), "x": torch.from_numpy(uid[sl].astype(np.float32) * 0.5)}
        ready, bad = router.push(tr, torch.from_numpy(done[sl].copy()), torch.from_numpy(coll[sl].copy()))
        if ready:
            buf.add(ready, bad)
            assert torch.equal(ready["x"], ready["uid"].float() * 0.5)    # fields stay aligned
    got1 = sorted(buf.memory_1.take()["uid"].tolist()) if len(buf.memory_1) else []
    got2 = sorted(buf.memory_2.take()["uid"].tolist()) if len(buf.memory_2) else []
    assert got1 == sorted(ref.memory_1) and got2 == sorted(ref.memory_2)
    assert got1 and got2
    # what is still held back is exactly the unfinished tail of every env
    held = sorted(router.pending["uid"].tolist()) if router.pending is not None else []
    want = sorted(int(uid[t, b]) for b in range(B) for t in range(T * blocks)
                  if not done[t:, b].any())
    assert held == want
    # inside one env the transitions of an episode arrive in time order
    for mem in (buf.memory_1.take()["uid"], buf.memory_2.take()["uid"]):
        for b in range(B):
            mine = [u for u in mem.tolist() if u % B == b]
            assert mine == sorted(mine)


def test_dual_rings_wrap_like_the_reference():
    Ref = reference_dual() or RestatedDual
    ref, buf = Ref(size=7), DeviceDualReplayBuffer(size=7)
    rng = random.Random(3)
    uid = 0
    for ep in range(12):
        n, bad = rng.randint(1, 5), rng.random() < 0.5
        ids = list(range(uid, uid + n))
        uid += n
        ref.add(ids, bad)
        buf.add({"uid": torch.tensor(ids)}, bad)
        assert buf.memory_1.take().get("uid", torch.zeros(0)).tolist() == list(ref.memory_1)
        assert buf.memory_2.take().get("uid", torch.zeros(0)).tolist() == list(ref.memory_2)
        assert (buf.memory_1.idx, buf.memory_2.idx) == (ref.idx_1, ref.idx_2)


@pytest.mark.parametrize("n1,n2,size", [(40, 3, 20), (3, 40, 20), (4, 5, 20), (40, 40, 20), (40, 40, 21), (10, 0, 20), (0, 0, 8)])
def test_dual_sample_batch_takes_the_references_shares(n1, n2, size):
    """replay_buffer_dual.py:38-63: half from each when both can give it, else the short memory whole
    and the remainder from the other."""
    Ref = reference_dual() or RestatedDual
    ref, buf = Ref(size=100), DeviceDualReplayBuffer(size=100)
    if n1:
        ref.add(list(range(n1)), True); buf.add({"uid": torch.arange(n1)}, True)
    if n2:
        ref.add(list(range(1000, 1000 + n2)), False); buf.add({"uid": torch.arange(1000, 1000 + n2)}, False)
    random.seed(0)
    want = np.array(ref.sample_batch(size))
    got = buf.sample_batch(size, generator=torch.Generator().manual_seed(0))
    ids = got["uid"].tolist() if got else []
    assert len(ids) == len(want) == sum(RestatedDual.sample_counts(ref, size))
    assert sum(1 for u in ids if u < 1000) == int((want < 1000).sum())
    assert len(set(ids)) == len(ids)
    k1 = sum(1 for u in ids if u < 1000)
    assert all(u < 1000 for u in ids[:k1]) and all(u >= 1000 for u in ids[k1:])   # memory_1 part first


def test_restatement_equals_the_reference_class():
    Ref = reference_dual()
    if Ref is None:
        pytest.skip("reference tree not present (GPU box)")
    a, b = Ref(size=5), RestatedDual(size=5)
    rng = random.Random(0)
    for ep in range(30):
        ids, bad = [rng.random() for _ in range(rng.randint(1, 4))], rng.random() < 0.4
        a.add(ids, bad); b.add(ids, bad)
        assert (a.memory_1, a.memory_2, a.idx_1, a.idx_2) == (b.memory_1, b.memory_2, b.idx_1, b.idx_2)
        for size in (2, 5, 9, 16):
            random.seed(1)
            got = a.sample_batch(size)
            assert len(got) == sum(b.sample_counts(size))
