"""Device replay buffers with the semantics of the reference's.

The reference keeps Python lists of object-array transitions on the host:
  * alg/replay_buffer.py:4-37   Replay_Buffer      - one ring of `size` transitions, `add(transition)`,
                                                     `sample_batch(n)` (everything, in memory order, while
                                                     len <= n; else n distinct uniform picks);
  * alg/replay_buffer_dual.py:4-63 Replay_Buffer   - two rings, memory_1 for transitions of "bad"
                                                     episodes and memory_2 for the others
                                                     (`add(episode, is_bad)`), and a `sample_batch` that
                                                     takes half of the batch from each when both can give it.
Here the rings are struct-of-arrays tensors in HBM - one `[size, ...]` tensor per transition field,
allocated on the first add - filled by whole blocks of transitions with one indexed copy per field,
and sampled with device index tensors, so a rollout never leaves the GPU between the step kernel
and the learner.  Slot order is the reference's: transition k of the add stream lands in slot
k mod size.

`EpisodeRouter` provides what the trainer's `buf_episode` list provides (train_onpolicy.py:302-356):
the dual buffer files a transition under good / bad by a property of its WHOLE episode
(`scenario.collisions != 0` when the episode has ended), so transitions are held back until their
episode terminates.  The per-step `collisions` output of the particle kernel (latched before the
in-kernel reset clears it, include/cm3env.h) carries that property to the terminal step.
"""
import torch


def _flatten(tr, lead):
    """field -> [n, ...] from field -> [*lead, ...]."""
    n = 1
    for d in lead:
        n *= int(d)
    return {k: v.reshape((n,) + tuple(v.shape[len(lead):])) for k, v in tr.items()}, n


class _Ring(object):
    """One ring of `maxsize` transitions (replay_buffer.py:6-16)."""

    def __init__(self, maxsize):
        self.maxsize, self.count = int(maxsize), 0   # count = transitions ever added
        self.store = None

    def __len__(self):
        return min(self.count, self.maxsize)

    @property
    def idx(self):
        """The reference's self.idx: the slot the next transition goes to."""
        return self.count % self.maxsize

    def add(self, flat, n):
        if n == 0:
            return
        if self.store is None:
            self.store = {k: torch.zeros((self.maxsize,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
                          for k, v in flat.items()}
        if set(flat) != set(self.store):
            raise ValueError("transition fields changed: %s vs %s" % (sorted(flat), sorted(self.store)))
        dev = next(iter(self.store.values())).device
        if n > self.maxsize:       # only the last `maxsize` survive a sequence of adds
            flat = {k: v[n - self.maxsize:] for k, v in flat.items()}
            self.count += n - self.maxsize
            n = self.maxsize
        pos = (self.count + torch.arange(n, device=dev)) % self.maxsize
        for k, v in flat.items():
            self.store[k].index_copy_(0, pos, v.to(dev))
        self.count += n

    def take(self, index=None):
        """All stored transitions in memory order (index None) or the rows `index`."""
        n = len(self)
        if self.store is None:
            return {}
        if index is None:
            return {k: v[:n] for k, v in self.store.items()}
        return {k: v.index_select(0, index) for k, v in self.store.items()}

    def sample_index(self, size, generator=None):
        """random.sample(memory, size): `size` distinct uniform slots."""
        dev = next(iter(self.store.values())).device
        return torch.randperm(len(self), device=dev, generator=generator)[:size]


class DeviceReplayBuffer(object):
    """alg/replay_buffer.py Replay_Buffer over device tensors."""

    def __init__(self, size=1e6):
        self.ring = _Ring(size)
        self.maxsize = self.ring.maxsize

    def __len__(self):
        return len(self.ring)

    @property
    def idx(self):
        return self.ring.idx

    def add(self, tr, lead=1):
        """`tr`: field -> tensor with `lead` leading batch axes ([n, ...], or [T, B, ...] with
        lead=2: transitions are taken in (t, b) order - time-major like the trainer's loop, the
        order inside a step is the env index).  replay_buffer.py:11-16 per transition."""
        first = next(iter(tr.values()))
        flat, n = _flatten(tr, first.shape[:lead])
        self.ring.add(flat, n)

    def sample_batch(self, size, generator=None):
        """replay_buffer.py:27-37: the whole memory (in slot order) while it holds <= size
        transitions, else `size` distinct uniformly chosen ones."""
        if len(self.ring) <= size:
            return self.ring.take()
        return self.ring.take(self.ring.sample_index(int(size), generator))


class DeviceDualReplayBuffer(object):
    """alg/replay_buffer_dual.py Replay_Buffer over device tensors: memory_1 holds the transitions
    of bad episodes, memory_2 the rest."""

    def __init__(self, size=5e4):
        self.memory_1, self.memory_2 = _Ring(size), _Ring(size)
        self.maxsize = self.memory_1.maxsize

    def add(self, episode, is_bad=False, lead=1):
        """replay_buffer_dual.py:13-24.  `episode`: field -> [n, ...] transitions; `is_bad` either one
        bool for all of them (the reference's call, one episode at a time) or a bool tensor [n] filing
        every transition by its own episode's flag (EpisodeRouter.push supplies it)."""
        first = next(iter(episode.values()))
        flat, n = _flatten(episode, first.shape[:lead])
        if torch.is_tensor(is_bad):
            bad = is_bad.reshape(n).bool()
            ib, ig = torch.nonzero(bad).squeeze(1), torch.nonzero(~bad).squeeze(1)
            self.memory_1.add({k: v.index_select(0, ib) for k, v in flat.items()}, int(ib.numel()))
            self.memory_2.add({k: v.index_select(0, ig) for k, v in flat.items()}, int(ig.numel()))
        elif is_bad:
            self.memory_1.add(flat, n)
        else:
            self.memory_2.add(flat, n)

    def sample_batch(self, size, generator=None):
        """replay_buffer_dual.py:38-63, case by case; where the reference concatenates two lists the
        result is the concatenation of the two selections in the same order (memory_1 part first)."""
        half = int(size / 2.0)
        n1, n2 = len(self.memory_1), len(self.memory_2)

        def cat(a, b):
            if not a:
                return b
            if not b:
                return a
            return {k: torch.cat([a[k], b[k]], dim=0) for k in a}
        if half <= n1 and half > n2:      # enough bad transitions but not enough good ones
            n_from_1 = min(n1, size - n2)
            return cat(self.memory_1.take(self.memory_1.sample_index(n_from_1, generator)), self.memory_2.take())
        if half > n1 and half <= n2:      # not enough bad transitions but enough good ones
            n_from_2 = min(n2, size - n1)
            return cat(self.memory_1.take(), self.memory_2.take(self.memory_2.sample_index(n_from_2, generator)))
        if n1 < half and n2 < half:       # neither
            return cat(self.memory_1.take(), self.memory_2.take())
        return cat(self.memory_1.take(self.memory_1.sample_index(half, generator)),
                   self.memory_2.take(self.memory_2.sample_index(half, generator)))


class EpisodeRouter(object):
    """Turns blocks of vectorised transitions ([T, B, ...], in-kernel episode reset on) into finished
    episodes' transitions with their episode-level bad flag, holding back the transitions of
    episodes still running at the end of a block - the device form of the trainer's
    `buf_episode` / `buf.add(buf_episode, scenario.collisions != 0)` (train_onpolicy.py:329-356)."""

    def __init__(self):
        self.pending = None   # field -> [P, ...] transitions of unfinished episodes, and "_env" [P]

    def push(self, tr, done, episode_flag_source):
        """tr: field -> [T, B, ...]; done [T, B] (uint8 / bool); episode_flag_source [T, B]: a value
        whose reading AT THE TERMINAL STEP of an episode classifies the whole episode as bad when
        non-zero (the particle kernel's `collisions` output).  Returns (transitions field -> [M, ...],
        bad [M] bool) of every transition whose episode has ended, oldest first."""
        T, B = done.shape
        dev = done.device
        d = done.bool()
        t_idx = torch.arange(T, device=dev).unsqueeze(1).expand(T, B)
        marks = torch.where(d, t_idx, torch.full_like(t_idx, T))
        term = torch.flip(torch.cummin(torch.flip(marks, [0]), dim=0).values, [0])   # first terminal step at or after t
        finished = term < T
        flag_tb = episode_flag_source.gather(0, term.clamp(max=T - 1)) != 0           # [T, B]
        first = term[0]                                                               # [B] first terminal step of the block
        flat, n = _flatten(tr, (T, B))
        env_of = torch.arange(B, device=dev).repeat(T)
        fin = finished.reshape(n)
        out_parts, bad_parts = [], []
        if self.pending is not None:
            penv = self.pending["_env"]
            ends = first[penv] < T
            sel = torch.nonzero(ends).squeeze(1)
            if sel.numel():
                out_parts.append({k: v.index_select(0, sel) for k, v in self.pending.items() if k != "_env"})
                bad_parts.append(flag_tb[first[penv[sel]], penv[sel]])
            keep = torch.nonzero(~ends).squeeze(1)
            self.pending = {k: v.index_select(0, keep) for k, v in self.pending.items()} if keep.numel() else None
        sel = torch.nonzero(fin).squeeze(1)
        if sel.numel():
            out_parts.append({k: v.index_select(0, sel) for k, v in flat.items()})
            bad_parts.append(flag_tb.reshape(n)[sel])
        rest = torch.nonzero(~fin).squeeze(1)
        if rest.numel():
            new = {k: v.index_select(0, rest) for k, v in flat.items()}
            new["_env"] = env_of.index_select(0, rest)
            self.pending = new if self.pending is None else {k: torch.cat([self.pending[k], new[k]], dim=0) for k in new}
        if not out_parts:
            return {}, torch.zeros(0, dtype=torch.bool, device=dev)
        out = {k: torch.cat([p[k] for p in out_parts], dim=0) for k in out_parts[0]}
        return out, torch.cat(bad_parts, dim=0)
