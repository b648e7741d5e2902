"""VecCheckers - B instances of the reference's Checkers env (env/checkers.py) stepped by one
CUDA kernel launch.  torch tensors only hold the HBM (and pinned host) buffers; all
computation happens in libcm3env.so through the C ABI of include/cm3env.h.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._buffers import HostRolloutBuffers, alloc_fields, _to_int8_host

FIELDS = L.CheckersOutputs.FIELDS
OBS_FIELDS = ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v")
# the fields of the reference's return tuple (env/checkers.py:262); goal_idx is this library's extra
REF_FIELDS = ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v", "reward", "local_rewards", "done")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class VecCheckers(object):
    """Constructor arguments after `num_envs` are those of Checkers.__init__
    (env/checkers.py:5-6), with the same defaults."""

    def __init__(self, num_envs, n_rows=3, n_columns=16, n_obs=2, agents_r=(0, 2),
                 agents_c=(16, 16), n_agents=1, max_steps=50, device="cuda:0",
                 dtype=torch.float32, env_id_offset=0, tile_dtype=None, random_goal=False):
        # the reference's own asserts (checkers.py:16-17)
        assert n_rows % 2 == 1
        assert n_columns % 2 == 0
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 or torch.float64")
        # grid / obs_self_t only hold {-1, 0, +1}: tile_dtype=torch.int8 writes them as bytes,
        # tile_dtype="u2" as 2 bits per cell packed in 32-bit words (cm3_b200/tiles.py decodes)
        self.tile_u2 = isinstance(tile_dtype, str) and tile_dtype == "u2"
        if self.tile_u2:
            tile_dtype = torch.int32
            if dtype != torch.float32:
                raise ValueError('tile_dtype="u2" needs dtype=torch.float32')
        tile_dtype = dtype if tile_dtype is None else tile_dtype
        if tile_dtype not in (dtype, torch.int8) and not self.tile_u2:
            raise ValueError('tile_dtype must be None (= dtype), torch.int8 or "u2"')
        self.tile_dtype = tile_dtype
        self.lib = L.load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.Cm3Error(-5, "VecCheckers needs a CUDA device (got %r); there is no CPU path" % (device,))
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self.B, self.N = int(num_envs), int(n_agents)
        self.n_rows, self.n_columns, self.n_obs = int(n_rows), int(n_columns), int(n_obs)
        self.max_steps = int(max_steps)
        self.dtype = dtype
        self.W = 2 * self.n_obs + 1
        self.L_others = 2 * max(self.N - 1, 1)
        self.total_rows = self.n_rows + 2 * self.n_obs          # checkers.py:24
        self.total_columns = self.n_columns + 2 * self.n_obs + 1  # checkers.py:25
        self.max_collectible = self.n_rows * self.n_columns      # checkers.py:28

        cfg = L.CheckersConfig()
        cfg.n_rows, cfg.n_columns, cfg.n_obs = self.n_rows, self.n_columns, self.n_obs
        cfg.n_agents, cfg.max_steps = self.N, self.max_steps
        if self.N > L.MAX_AGENTS:
            raise L.Cm3Error(-4, "n_agents=%d: at most %d agents are supported" % (self.N, L.MAX_AGENTS))
        for i in range(self.N):
            cfg.agents_r[i] = int(agents_r[i])
            cfg.agents_c[i] = int(agents_c[i])
        cfg.num_envs = self.B
        cfg.real = L.REAL_F64 if dtype == torch.float64 else L.REAL_F32
        cfg.tile = L.TILE_U2 if self.tile_u2 else L.TILE_I8 if tile_dtype == torch.int8 and dtype != torch.int8 else L.TILE_REAL
        cfg.device = dev_index
        cfg.env_id_offset = int(env_id_offset)
        self.env_id_offset = int(env_id_offset)
        # stage 1 draws a fresh goal before every episode (train_offpolicy.py:291-296): with
        # random_goal the in-kernel episode reset does the same from Philox
        self.random_goal = bool(random_goal)
        cfg.random_goal = 1 if self.random_goal else 0
        h = C.c_void_p()
        L.check(self.lib.cm3_checkers_create(C.byref(cfg), C.byref(h)))
        self._h = h
        tiles = C.c_int32(0)
        L.check(self.lib.cm3_checkers_tiles(h, C.byref(tiles)))

        dev = self.device
        self.state = dict(
            remaining=torch.zeros(self.B, dtype=torch.int64, device=dev),
            agents=torch.zeros(self.B, self.N, dtype=torch.int32, device=dev),
            meta=torch.zeros(self.B, dtype=torch.int32, device=dev))
        # per-tile launch-chaining words (include/cm3env.h: cm3_checkers_state.sync); not part of
        # the env state proper: never saved, never restored
        self._sync = torch.zeros(2, tiles.value, dtype=torch.int32, device=dev)
        self._st = L.CheckersState(_ptr(self.state["remaining"]), _ptr(self.state["agents"]),
                                   _ptr(self.state["meta"]), _ptr(self._sync))
        self.out = self.alloc_outputs()
        self._out_c = self._outputs_struct(self.out)
        self._actions_dev = torch.zeros(self.B, self.N, dtype=torch.int8, device=dev)
        self._host = None
        self._mapped = None
        self._is_reset = False

    # ------------------------------------------------------------------ buffers
    def field_shapes(self):
        B, N, W = self.B, self.N, self.W
        grid, win = (B, self.n_rows, self.n_columns + 1, 2), (B, N, W, W, 3)
        if self.tile_u2:   # packed rows (include/cm3env.h: CM3_TILE_U2)
            grid, win = (B, self.n_rows, (self.n_columns + 1 + 7) // 8), (B, N, W, (6 * W + 31) // 32)
        return dict(grid=grid, vec=(B, N, 4),
                    obs_others=(B, N, self.L_others), obs_self_t=win,
                    obs_self_v=(B, N, 4), reward=(B,), local_rewards=(B, N), done=(B,), goal_idx=(B, N))

    def bytes_per_env_step(self, fields=None):
        """Algorithmic bytes of one env-step: outputs + state read/write + actions."""
        state = 2 * (8 + 4 * self.N + 4)
        return self.out_bytes_per_env_step(fields) + state + self.N

    def field_dtype(self, k):
        if k in ("done", "goal_idx"):
            return torch.uint8
        return self.tile_dtype if k in ("grid", "obs_self_t") else self.dtype

    def out_bytes_per_env_step(self, fields=None):
        """Output bytes per env-step of the reference's return tuple (REF_FIELDS) or of `fields`."""
        fields = REF_FIELDS if fields is None else fields
        return sum(int(np.prod(s[1:])) * self.field_dtype(k).itemsize for k, s in self.field_shapes().items() if k in fields)

    def alloc_outputs(self, T=None, pinned_host=False, fields=None):
        """Zeroed output buffers: [B, ...] per field (T=None) or [T, B, ...] rollout buffers, for
        every field or for `fields`.  The single-step set is one packed allocation (every field
        16-byte aligned), so that step_host needs a single device-to-host copy."""
        lead = () if T is None else (int(T),)
        shapes = {k: v for k, v in self.field_shapes().items() if fields is None or k in fields}
        return alloc_fields(shapes, self.field_dtype, lead, device=self.device,
                            pinned=pinned_host, packed=T is None)

    @staticmethod
    def _outputs_struct(out):
        return L.CheckersOutputs(*[_ptr(out.get(f)) for f in FIELDS])

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self.lib.cm3_checkers_destroy(h)
            self._h = None

    # ------------------------------------------------------------------ API
    def goal_idx_from(self, goals):
        """The reference's one-hot goals, [N,2] (shared by all envs) or [B,N,2], to
        uint8 [B,N] = np.where(goals[idx]==1)[0][0] (checkers.py:235)."""
        g = np.asarray(goals.cpu() if torch.is_tensor(goals) else goals)
        if g.ndim not in (2, 3) or g.shape[-1] != 2 or g.shape[-2] != self.N:
            raise ValueError("goals must be one-hot with shape [n_agents,2] or [num_envs,n_agents,2]")
        if not np.all((g == 1).sum(axis=-1) >= 1):
            raise ValueError("every goals row needs an entry equal to 1 (checkers.py:235)")
        idx = np.argmax(g == 1, axis=-1).astype(np.uint8)
        return np.array(np.broadcast_to(idx, (self.B, self.N)), dtype=np.uint8, order="C")

    def reset(self, goals=None, mask=None, goal_idx=None):
        """Checkers.reset(goals) for every env (or those where mask != 0).  Returns the output
        dict (device tensors); reward / local_rewards are not meaningful after a reset.

        Without goals the selected envs KEEP the goals they have; on the very first reset they get
        the trainers' default, agent n -> goal n & 1 (np.eye(n_agents) for stage 2,
        train_offpolicy.py:298)."""
        gi = None
        if goals is None and goal_idx is None and not self._is_reset:
            goal_idx = np.arange(self.N, dtype=np.uint8) & 1
        if goal_idx is not None:
            gnp = np.array(np.broadcast_to(np.asarray(goal_idx, dtype=np.uint8), (self.B, self.N)), order="C")
            if gnp.max(initial=0) > 1:
                raise ValueError("goal index must be 0 or 1 (checkers.py:202-223)")
            gi = torch.from_numpy(gnp)
        elif goals is not None:
            gi = torch.from_numpy(self.goal_idx_from(goals))
        if gi is not None:
            gi = gi.to(self.device, non_blocking=False)
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
            if m.shape != (self.B,):
                raise ValueError("mask must have shape [num_envs]")
        L.check(self.lib.cm3_checkers_reset(self._h, C.byref(self._st), _ptr(gi), _ptr(m),
                                            C.byref(self._out_c), self._stream()))
        self._keep = (gi, m)
        self._is_reset = True
        return self.out

    def _actions_tensor(self, actions, lead=()):
        shape = tuple(lead) + (self.B, self.N)
        if torch.is_tensor(actions):
            a = actions
            if a.dtype != torch.int8:
                a = a.clamp(-128, 127).to(torch.int8)
            a = a.to(self.device).reshape(shape).contiguous()
        else:
            a = np.clip(np.asarray(actions, dtype=np.int64), -128, 127).astype(np.int8).reshape(shape)
            a = torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        return a

    def step(self, actions):
        """Checkers.step(actions) for all envs; actions [B,N] ints (any value, like the
        reference: values outside 0..4 earn -0.1).  Returns the output dict (device tensors,
        overwritten by the next call)."""
        a = self._actions_tensor(actions)
        L.check(self.lib.cm3_checkers_step(self._h, C.byref(self._st), _ptr(a),
                                           C.byref(self._out_c), self._stream()))
        self._keep = a
        return self.out

    def step_chained(self, actions, out, seed=0, t0=0, auto_reset=True):
        """One step that may overlap the previous launch on the device (cm3_checkers_step_chained):
        `actions` is an int8 device tensor [B,N] that was complete before the previous call was
        made (a slice of a pre-generated stream), `out` a dict of [B,...] (or [1,B,...]) buffers
        that is not one of the previous two calls' - a slot of a rollout ring of at least three."""
        oc = out if isinstance(out, L.CheckersOutputs) else self._outputs_struct(out)
        L.check(self.lib.cm3_checkers_step_chained(self._h, C.byref(self._st), _ptr(actions), int(seed) & (2**64 - 1),
                                                   int(t0), 1 if auto_reset else 0, C.byref(oc), self._stream()))
        self._keep = (actions, oc)
        return out

    def rollout(self, T, actions=None, seed=0, t0=0, auto_reset=False, out=None,
                record_actions=False):
        """T fused steps in one launch.  actions [T,B,N] or None (device Philox stream keyed
        by (seed, global env id, t0 + t)).  `out` = alloc_outputs(T) buffers to fill."""
        T = int(T)
        if out is None:
            out = self.alloc_outputs(T)
        a = None if actions is None else self._actions_tensor(actions, (T,))
        rec = torch.zeros(T, self.B, self.N, dtype=torch.int8, device=self.device) if record_actions else None
        oc = self._outputs_struct(out)
        L.check(self.lib.cm3_checkers_rollout(self._h, C.byref(self._st), _ptr(a), int(seed) & (2**64 - 1),
                                              int(t0), T, 1 if auto_reset else 0, _ptr(rec),
                                              C.byref(oc), self._stream()))
        self._keep = (a, oc)
        if record_actions:
            out = dict(out)
            out["actions"] = rec
        return out

    def plan_rollout(self, T, actions=None, out=None, auto_reset=True, seed=0):
        """A rollout whose ctypes arguments are built ONCE: the returned callable launches
        cm3_checkers_rollout(t0) with nothing but the foreign call on the host path (rollout()
        rebuilds the output struct and re-validates the action tensor on every call).  `actions`
        must already be an int8 device tensor [T,B,N] (or None for the Philox stream)."""
        T = int(T)
        out = self.alloc_outputs(T) if out is None else out
        if actions is not None and (actions.dtype != torch.int8 or actions.device != self.device
                                    or tuple(actions.shape) != (T, self.B, self.N) or not actions.is_contiguous()):
            raise ValueError("plan_rollout needs a contiguous int8 device tensor [T,B,N]")
        oc = self._outputs_struct(out)
        fn, h, st, a = self.lib.cm3_checkers_rollout, self._h, C.byref(self._st), _ptr(actions)
        seed, ar, ocr, dev = int(seed) & (2**64 - 1), 1 if auto_reset else 0, C.byref(oc), self.device

        def launch(t0=0):
            rc = fn(h, st, a, seed, int(t0), T, ar, None, ocr, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if rc != 0:
                L.check(rc)
        launch.out, launch.keep = out, (actions, oc)
        return launch

    def rollout_gather(self, T, dst_ptrs, dst_B, dst_env0, actions=None, seed=0, t0=0,
                       auto_reset=False):
        """Fused rollout + all-gather (cm3_checkers_rollout_gather): like rollout(), but every output element is stored
        to each destination in `dst_ptrs` - a list (one entry per GPU, at most 8) of dicts
        field -> raw device pointer of a [T, dst_B, ...] array, typically the symmetric-memory
        rollout buffers of all ranks (cm3_b200.sharding.RolloutAllGather).  This shard's envs land
        at rows [dst_env0, dst_env0 + B).  Fields missing from the dicts are not written."""
        T = int(T)
        if not 1 <= len(dst_ptrs) <= L.MAX_DST:
            raise ValueError("1..%d destinations" % L.MAX_DST)
        a = None if actions is None else self._actions_tensor(actions, (T,))
        arr = (L.CheckersOutputs * len(dst_ptrs))()
        for i, d in enumerate(dst_ptrs):
            arr[i] = L.CheckersOutputs(*[(C.c_void_p(int(d[f])) if d.get(f) else None) for f in FIELDS])
        L.check(self.lib.cm3_checkers_rollout_gather(self._h, C.byref(self._st), _ptr(a), int(seed) & (2**64 - 1), int(t0), T,
                            1 if auto_reset else 0, None, len(dst_ptrs), arr, int(dst_B),
                            int(dst_env0), self._stream()))
        self._keep = (a, arr)

    def step_host(self, actions, fields=FIELDS):
        """Host-buffer step: actions is a host int8 array [B,N]; the requested output fields are
        copied back into pinned host buffers and returned as NumPy views.  When every field is
        requested and the single-step buffers are packed (alloc_outputs), the outputs travel as one
        device-to-host copy (cm3_checkers_step_host_packed)."""
        if self._host is None:
            self._host = self.alloc_outputs(pinned_host=True)
            self._host_actions = torch.zeros(self.B, self.N, dtype=torch.int8).pin_memory()
        self._host_actions.numpy()[...] = _to_int8_host(actions, (self.B, self.N))
        dev_block, host_block = getattr(self.out, "block", None), getattr(self._host, "block", None)
        if dev_block is not None and host_block is not None and tuple(fields) == tuple(FIELDS):
            L.check(self.lib.cm3_checkers_step_host_packed(self._h, C.byref(self._st), _ptr(self._host_actions),
                                                       _ptr(self._actions_dev), C.byref(self._out_c),
                                                       _ptr(dev_block), _ptr(host_block), dev_block.numel(),
                                                       self._stream()))
        else:
            oh = L.CheckersOutputs(*[_ptr(self._host[f]) if f in fields else None for f in FIELDS])
            L.check(self.lib.cm3_checkers_step_host(self._h, C.byref(self._st), _ptr(self._host_actions),
                                                _ptr(self._actions_dev), C.byref(self._out_c),
                                                C.byref(oh), self._stream()))
        return {f: self._host[f].numpy() for f in fields}

    def step_mapped(self, actions, stream=None, copy=False):
        """Low-latency host step for small batches (the B = 1 drop-ins): the kernel reads the
        actions from, and writes every output field to, PINNED HOST memory directly (unified
        addressing), so the host path is one launch and one stream wait - no copy calls.  Returns
        field -> NumPy view of the pinned buffers (overwritten by the next call), or with copy=True
        views of ONE fresh host copy of the whole output block (what the B = 1 drop-ins hand out).
        `stream`: a cached ctypes stream handle (default: torch's current stream)."""
        m = self._mapped
        if m is None:
            host = self.alloc_outputs(pinned_host=True)
            acts = torch.zeros(self.B, self.N, dtype=torch.int8).pin_memory()
            m = self._mapped = dict(host=host, acts=acts, acts_np=acts.numpy(), oc=self._outputs_struct(host),
                                    views={f: host[f].numpy() for f in FIELDS}, step=self.lib.cm3_checkers_step,
                                    sync=self.lib.cm3_stream_synchronize, st=C.byref(self._st), a=_ptr(acts))
            m["ocr"] = C.byref(m["oc"])
            m["block_np"] = host.block.numpy()
            shapes = self.field_shapes()
            m["layout"] = [(f, host.offsets[f], int(np.prod(shapes[f])) * self.field_dtype(f).itemsize,
                            host[f].numpy().dtype, tuple(shapes[f])) for f in FIELDS]
        m["acts_np"][...] = _to_int8_host(actions, (self.B, self.N))
        s = stream if stream is not None else C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = m["step"](self._h, m["st"], m["a"], m["ocr"], s) or m["sync"](s)
        if rc != 0:
            L.check(rc)
        if not copy:
            return m["views"]
        blk = m["block_np"].copy()
        return {f: blk[off:off + n].view(dt).reshape(shape) for f, off, n, dt, shape in m["layout"]}

    def download(self):
        """The packed single-step outputs (whatever the last launch wrote to self.out) in ONE
        device-to-host copy; returns field -> NumPy view of the pinned host mirror."""
        if self._host is None:
            self._host = self.alloc_outputs(pinned_host=True)
            self._host_actions = torch.zeros(self.B, self.N, dtype=torch.int8).pin_memory()
        self._host.block.copy_(self.out.block, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return {f: self._host[f].numpy() for f in FIELDS}

    def rollout_host(self, actions, seed=0, t0=0, auto_reset=True):
        """T steps with host actions ([T,B,N] int8) and every output delivered to pinned host
        memory, double buffered (cm3_checkers_rollout_host): the device-to-host copy of step t
        overlaps the kernel of step t + 1.  Returns field -> NumPy view [T,B,...] (overwritten by
        the next call with the same T)."""
        a = np.asarray(actions)
        T = int(a.shape[0])
        hr = self._hr.get(T) if hasattr(self, "_hr") else None
        if hr is None:
            if not hasattr(self, "_hr"):
                self._hr = {}
            hr = self._hr[T] = HostRolloutBuffers(self, T, self._outputs_struct)
        hr.actions_host.numpy()[...] = _to_int8_host(a, (T, self.B, self.N))
        L.check(self.lib.cm3_checkers_rollout_host(self._h, C.byref(self._st), _ptr(hr.actions_host), _ptr(hr.actions_dev), T,
                                                  int(seed) & (2**64 - 1), int(t0), 1 if auto_reset else 0, hr.outs_c,
                                                  hr.blocks_c, _ptr(hr.host), hr.block_bytes, hr.block_bytes,
                                                  self._stream()))
        return hr.views

    # ------------------------------------------------------------------ state
    def get_state_host(self):
        """NumPy copies of the compact state through cm3_checkers_get_state (checkpointing without a
        tensor library; state_dict() is the device-side equivalent)."""
        host = {k: np.empty(tuple(v.shape), dtype=torch.empty(0, dtype=v.dtype).numpy().dtype) for k, v in self.state.items()}
        hs = L.CheckersState(*[C.c_void_p(host[f].ctypes.data) for f in self.state])
        L.check(self.lib.cm3_checkers_get_state(self._h, C.byref(self._st), C.byref(hs), self._stream()))
        return host

    def set_state_host(self, host):
        """Inverse of get_state_host (cm3_checkers_set_state); missing keys are left untouched."""
        keep = {k: np.ascontiguousarray(host[k], dtype=torch.empty(0, dtype=self.state[k].dtype).numpy().dtype).reshape(tuple(self.state[k].shape))
                for k in self.state if k in host}
        hs = L.CheckersState(*[(C.c_void_p(keep[f].ctypes.data) if f in keep else None) for f in self.state])
        L.check(self.lib.cm3_checkers_set_state(self._h, C.byref(self._st), C.byref(hs), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # pageable host memory: the arrays may go away

    def state_dict(self):
        return {k: v.clone() for k, v in self.state.items()}

    def load_state_dict(self, sd):
        for k in self.state:
            self.state[k].copy_(sd[k])

    def unpack_state(self):
        """Decoded compact state as NumPy arrays (debugging / tests)."""
        rem = self.state["remaining"].cpu().numpy().view(np.uint64)
        ag = self.state["agents"].cpu().numpy().view(np.uint32)
        meta = self.state["meta"].cpu().numpy().view(np.uint32)
        return dict(remaining=rem, r=ag & 0xFF, c=(ag >> 8) & 0xFF, n_green=(ag >> 16) & 0xFF,
                    n_orange=ag >> 24, steps=meta & 0xFFFFFF, goal_bits=meta >> 24)
