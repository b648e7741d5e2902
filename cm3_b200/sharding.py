"""Multi-GPU use of the env stepper: one process per GPU, the env batch cut into contiguous
global-env-id ranges, no collective on the stepping path, and an optional all-gather of the
rollout buffers for a learner that wants the whole batch on every GPU.

The reference has no counterpart: it runs one env instance per process and its only fan-out is one
process per seed without communication (alg/train_multiprocess.py:31-43).  What is here stands
where a data-parallel learner built on the reference's loop (alg/train_onpolicy.py:302-350) would
exchange its workers' rollouts.

torch.distributed is the plumbing (rendezvous, NCCL / gloo collectives, symmetric-memory handle
exchange); the stepping and, in "peer" mode, the gather itself are this repo's kernels.
"""
import os

import torch
import torch.distributed as dist


def split_envs(total_envs, world):
    """Contiguous ranges [(start, count)] * world; the first total % world ranks take one extra."""
    total_envs, world = int(total_envs), int(world)
    if world < 1 or total_envs < world:
        raise ValueError("need 1 <= world <= total_envs (got world=%d, total_envs=%d)" % (world, total_envs))
    base, extra = divmod(total_envs, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, n))
        start += n
    return out


def bind_host_to_gpu(local_rank):
    """Pins the calling process to the CPU cores next to GPU `local_rank` (NVML's ideal affinity mask), so that the
    pinned host buffers it allocates afterwards - and the thread that drives the GPU - sit on that GPU's NUMA node.
    One process per GPU with host-buffer I/O (cm3_*_rollout_host) otherwise lands on whichever socket the launcher
    picked and may copy across the socket link.  Returns the number of CPUs bound to, or 0 when NVML is unavailable
    (nothing is changed then).  CM3_BIND_NUMA=0 disables.  On the single-socket hosts of this pool the mask is
    every CPU and the call changes nothing (measured: e2e 51-52 GB/s either way); it is there for multi-socket hosts."""
    if os.environ.get("CM3_BIND_NUMA", "1") == "0":
        return 0
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = local_rank
        if vis:
            ids = vis.split(",")
            if local_rank < len(ids) and ids[local_rank].strip().isdigit():
                phys = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))   # never widen what the launcher / container allows
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001 - an optimisation, never a requirement
        return 0


class EnvShard(object):
    """The env-id range owned by this rank.  rank / world default to the initialised process
    group, else to torchrun's RANK / WORLD_SIZE, else to a single process."""

    def __init__(self, total_envs, rank=None, world=None, local_rank=None):
        if rank is None or world is None:
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
            else:
                rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        if not 0 <= rank < world:
            raise ValueError("rank %d outside world %d" % (rank, world))
        self.total_envs, self.rank, self.world = int(total_envs), int(rank), int(world)
        self.ranges = split_envs(total_envs, world)
        self.start, self.count = self.ranges[rank]
        self.local_rank = int(os.environ.get("LOCAL_RANK", rank)) if local_rank is None else int(local_rank)
        self.device = "cuda:%d" % self.local_rank

    @property
    def uniform(self):
        return self.total_envs % self.world == 0

    def owner_of(self, env_id):
        for r, (s, n) in enumerate(self.ranges):
            if s <= env_id < s + n:
                return r, env_id - s
        raise IndexError(env_id)

    def __repr__(self):
        return "EnvShard(rank %d/%d: envs [%d, %d) of %d)" % (self.rank, self.world, self.start,
                                                              self.start + self.count, self.total_envs)


def all_gather_rollout(local, shard, group=None, time_major=True):
    """Collective all-gather of rollout fields.  `local` maps field -> [T, B_local, ...] tensor
    (same T and trailing shape on every rank, B_local = shard.count which must be uniform).
    Returns field -> [T, B_total, ...] (time_major, one permuting copy after the collective) or
    field -> [world, T, B_local, ...] (the collective's native layout, no copy).  Works on any
    backend (NCCL on GPUs; gloo in the CPU tests)."""
    if not shard.uniform:
        raise ValueError("all_gather_rollout needs total_envs divisible by the world size")
    out = {}
    for k, x in local.items():
        x = x.contiguous()
        if x.shape[1] != shard.count:
            raise ValueError("%s: second dim %d != shard.count %d" % (k, x.shape[1], shard.count))
        g = torch.empty((shard.world,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        if shard.world == 1:
            g[0].copy_(x)
        else:
            # concatenation form [world * T, ...]: accepted by NCCL and gloo alike
            dist.all_gather_into_tensor(g.view((shard.world * x.shape[0],) + tuple(x.shape[1:])), x, group=group)
        if time_major:
            T = x.shape[0]
            perm = (1, 0) + tuple(range(2, g.dim()))
            g = g.permute(*perm).reshape((T, shard.total_envs) + tuple(x.shape[2:]))
        out[k] = g
    return out


class RolloutAllGather(object):
    """T fused env steps on this rank's shard with the outputs of ALL ranks delivered as
    field -> [T, total_envs, ...] on every GPU.

    mode "peer": the rollout buffers live in symmetric memory (every rank can address every
        rank's buffer over NVLink); cm3_*_rollout_gather stores each output element straight
        into all `world` buffers at this shard's env offset - compute and all-gather are ONE
        kernel, there is no local staging buffer and no collective launch.
        double_buffer=True: two symmetric buffers used in turn.  Rollout k + 1 only needs "every
        rank is done with buffer (k + 1) & 1" - the contents of rollout k - 1 - so it is enqueued
        right behind rollout k on the compute stream, while the barrier that publishes rollout k
        ("all ranks' stores have landed") runs on a second stream; the consumer's stream waits for
        that event.  With one buffer the two barriers and the kernel are strictly serial.
    mode "nccl": local rollout, then ncclAllGather per field (+ a permuting copy to time-major).
    mode "auto": "peer" when symmetric memory can be set up for the group, else "nccl".
    """

    def __init__(self, env, T, shard=None, group=None, mode="auto", fields=None, double_buffer=False):
        self.env, self.T = env, int(T)
        self.shard = shard or EnvShard(env.B if not dist.is_initialized() else env.B * dist.get_world_size())
        if self.shard.count != env.B:
            raise ValueError("env.B (%d) != shard.count (%d)" % (env.B, self.shard.count))
        if env.env_id_offset != self.shard.start:
            raise ValueError("env.env_id_offset (%d) != shard.start (%d)" % (env.env_id_offset, self.shard.start))
        self.group = group
        self.fields = tuple(fields) if fields is not None else tuple(env.field_shapes().keys())
        self.mode = mode
        self.nbuf = 2 if double_buffer else 1
        self._hdl = None
        self._k = 0
        if mode in ("auto", "peer"):
            try:
                self._setup_peer()
                self.mode = "peer"
            except Exception as e:  # noqa: BLE001 - symmetric memory is an optional capability
                if mode == "peer":
                    raise
                self.mode, self.peer_error = "nccl", repr(e)
        if self.mode == "nccl":
            self.local = {k: v for k, v in env.alloc_outputs(self.T, fields=self.fields).items()}

    # ------------------------------------------------------------------ peer mode
    def _field_layout(self):
        """Byte offsets of the [T, total_envs, ...] arrays inside one symmetric allocation."""
        shapes = self.env.field_shapes()
        layout, off = {}, 0
        for k in self.fields:
            dt = self.env.field_dtype(k)
            shp = (self.T, self.shard.total_envs) + tuple(shapes[k][1:])
            nbytes = int(torch.Size(shp).numel()) * dt.itemsize
            layout[k] = (off, shp, dt)
            off += (nbytes + 255) // 256 * 256
        return layout, off

    def _setup_peer(self):
        import torch.distributed._symmetric_memory as symm_mem
        if not self.shard.uniform:
            raise ValueError("peer mode needs total_envs divisible by the world size")
        if self.shard.world > 8:
            raise ValueError("peer mode addresses at most 8 GPUs (one NVSwitch node)")
        layout, total = self._field_layout()
        dev = self.env.device
        grp = self.group if self.group is not None else dist.group.WORLD
        self._bufs, self._hdls, self._gathered, self._dsts = [], [], [], []
        for i in range(self.nbuf):
            buf = symm_mem.empty(total, dtype=torch.uint8, device=dev)
            hdl = symm_mem.rendezvous(buf, grp)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            self._bufs.append(buf)
            self._hdls.append(hdl)
            self._gathered.append({k: buf[off:off + int(torch.Size(shp).numel()) * dt.itemsize].view(dt).view(shp)
                                   for k, (off, shp, dt) in layout.items()})
            self._dsts.append([{k: ptrs[r] + layout[k][0] for k in self.fields} for r in range(self.shard.world)])
        self._layout = layout
        self._hdl, self.gathered, self._dst_ptrs = self._hdls[0], self._gathered[0], self._dsts[0]
        if self.nbuf == 2:
            self._pub_stream = torch.cuda.Stream(device=dev)
            self._stepped = [torch.cuda.Event(), torch.cuda.Event()]
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]

    # ------------------------------------------------------------------ API
    def rollout(self, actions=None, seed=0, t0=0, auto_reset=False, time_major=True, wait=True):
        """Returns the gathered field dict of this rollout.  peer mode with double_buffer: the
        buffers alternate, and with wait=False the caller's stream is NOT made to wait for the
        publishing barrier - call wait_ready(result) (or rollout(..., wait=True)) before reading."""
        env = self.env
        if self.mode == "peer":
            i = self._k % self.nbuf
            self._k += 1
            hdl = self._hdls[i]
            hdl.barrier(channel=0)  # every rank is done reading the previous contents of buffer i
            env.rollout_gather(self.T, self._dsts[i], self.shard.total_envs, self.shard.start,
                               actions=actions, seed=seed, t0=t0, auto_reset=auto_reset)
            if self.nbuf == 1:
                hdl.barrier(channel=0)  # all ranks' stores have landed
                return self._gathered[0]
            cur = torch.cuda.current_stream(env.device)
            self._stepped[i].record(cur)
            with torch.cuda.stream(self._pub_stream):
                self._pub_stream.wait_event(self._stepped[i])
                hdl.barrier(channel=1)  # all ranks' stores into buffer i have landed
                self._ready[i].record(self._pub_stream)
            if wait and self._k >= 2:
                # steady state of a pipelined consumer: it reads rollout k - 1 while rollout k is in flight
                cur.wait_event(self._ready[(i + 1) % 2])
            self._last = i
            return self._gathered[i]
        env.rollout(self.T, actions=actions, seed=seed, t0=t0, auto_reset=auto_reset, out=self.local)
        return all_gather_rollout(self.local, self.shard, self.group, time_major=time_major)

    def wait_ready(self, gathered=None):
        """Makes the current stream wait until the given (default: the latest) rollout's data from
        every rank has landed in this GPU's buffer."""
        if self.mode != "peer" or self.nbuf == 1:
            return
        i = self._last if gathered is None else next(j for j, g in enumerate(self._gathered) if g is gathered)
        torch.cuda.current_stream(self.env.device).wait_event(self._ready[i])

    def bytes_moved_per_rollout(self):
        """Bytes this rank sends to peers (peer mode: stores over NVLink; nccl: all-gather send)."""
        shapes = self.env.field_shapes()
        per_env = 0
        for k in self.fields:
            n = 1
            for d in shapes[k][1:]:
                n *= d
            per_env += n * self.env.field_dtype(k).itemsize
        return per_env * self.env.B * self.T * (self.shard.world - 1)
