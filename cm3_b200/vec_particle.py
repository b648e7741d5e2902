"""VecParticle - B instances of MultiAgentEnv + the multi-goal_spread scenario
(multiagent/environment.py, core.py, scenarios/multi-goal_spread.py) stepped by one CUDA kernel
launch.  torch tensors only hold the buffers; the computation is in libcm3env.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._buffers import HostRolloutBuffers, alloc_fields, _to_int8_host

FIELDS = L.ParticleOutputs.FIELDS
OBS_FIELDS = ("global_state", "obs_others", "obs_self")
# the fields of the reference's return tuple (multiagent/environment.py:123); `collisions`
# (scenario.collisions latched per step) is this library's extra
REF_FIELDS = ("global_state", "obs_others", "obs_self", "reward", "reward_n", "done")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class VecParticle(object):
    """`config` is the dict of a config_particle_*.json (agents_x/y, landmarks_x/y, initial_std),
    `prob_random` the make_world argument (multi-goal_spread.py:19), `max_steps` the
    MultiAgentEnv argument (environment.py:16).  dtype float64 = free-running parity mode."""

    def __init__(self, num_envs, n_agents, config, prob_random=0.0, max_steps=50,
                 device="cuda:0", dtype=torch.float32, env_id_offset=0, **world_overrides):
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 or torch.float64")
        self.lib = L.load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.Cm3Error(-5, "VecParticle needs a CUDA device (got %r); there is no CPU path" % (device,))
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self.B, self.N = int(num_envs), int(n_agents)
        if self.N > L.MAX_AGENTS:
            raise L.Cm3Error(-4, "n_agents=%d: at most %d agents are supported" % (self.N, L.MAX_AGENTS))
        self.max_steps = int(max_steps)
        self.dtype = dtype
        self.L_others = 4 * max(self.N - 1, 1)
        self.config = dict(config)
        self.prob_random = float(prob_random)

        cfg = L.ParticleConfig()
        self.lib.cm3_particle_default_config(C.byref(cfg), self.N, self.max_steps)
        cfg.num_envs = self.B
        cfg.real = L.REAL_F64 if dtype == torch.float64 else L.REAL_F32
        cfg.device = dev_index
        cfg.env_id_offset = int(env_id_offset)
        self.env_id_offset = int(env_id_offset)
        for i in range(self.N):
            cfg.agents_x[i] = float(config["agents_x"][i])
            cfg.agents_y[i] = float(config["agents_y"][i])
            cfg.landmarks_x[i] = float(config["landmarks_x"][i])
            cfg.landmarks_y[i] = float(config["landmarks_y"][i])
        cfg.initial_std = float(config.get("initial_std", 0.0))
        cfg.prob_random = self.prob_random
        for k, v in world_overrides.items():  # dt, damping, contact_force, contact_margin, ...
            if not hasattr(cfg, k):
                raise TypeError("unknown world constant %r" % k)
            setattr(cfg, k, float(v))
        self.cfg = cfg
        h = C.c_void_p()
        L.check(self.lib.cm3_particle_create(C.byref(cfg), C.byref(h)))
        self._h = h
        tiles = C.c_int32(0)
        L.check(self.lib.cm3_particle_tiles(h, C.byref(tiles)))

        dev = self.device
        B, N = self.B, self.N
        self.state = dict(
            sv=torch.zeros(B, N, 4, dtype=dtype, device=dev),
            landmarks=torch.zeros(B, N, 2, dtype=dtype, device=dev),
            steps=torch.zeros(B, dtype=torch.int32, device=dev),
            collisions=torch.zeros(B, dtype=torch.int32, device=dev),
            reached=torch.zeros(B, dtype=torch.uint8, device=dev))
        # per-tile launch-chaining words (cm3_particle_state.sync); not env state: never saved
        self._sync = torch.zeros(2, tiles.value, dtype=torch.int32, device=dev)
        self._st = L.ParticleState(*([_ptr(self.state[k]) for k in
                                      ("sv", "landmarks", "steps", "collisions", "reached")] + [_ptr(self._sync)]))
        self.out = self.alloc_outputs()
        self._out_c = self._outputs_struct(self.out)
        self._actions_dev = torch.zeros(B, N, dtype=torch.int8, device=dev)
        self._host = None
        self._mapped = None
        self._reset_counter = 0

    # ------------------------------------------------------------------ buffers
    def field_shapes(self):
        B, N = self.B, self.N
        return dict(global_state=(B, N, 4), obs_others=(B, N, self.L_others), obs_self=(B, N, 4),
                    reward=(B,), reward_n=(B, N), done=(B,), collisions=(B,), reached=(B,))

    def bytes_per_env_step(self, fields=None):
        """Algorithmic bytes of one env-step."""
        el = 8 if self.dtype == torch.float64 else 4
        N = self.N
        state = 2 * (4 * N * el) + 2 * N * el + 2 * 4 + 2 * 4 + 2  # sv rw, landmarks r, steps rw, collisions rw, reached rw
        return self.out_bytes_per_env_step(fields) + state + N

    def field_dtype(self, k):
        return torch.uint8 if k in ("done", "reached") else torch.int32 if k == "collisions" else self.dtype

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
        return L.ParticleOutputs(*[_ptr(out.get(f)) for f in FIELDS])

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self.lib.cm3_particle_destroy(h)
            self._h = None

    # ------------------------------------------------------------------ API
    def reset(self, init_pos=None, init_landmarks=None, mask=None, seed=0, reset_counter=None):
        """MultiAgentEnv.reset().  With init_pos / init_landmarks ([B,N,2]) the state that
        reset_world() drew on the host is injected (parity protocol); otherwise the device draws
        it from Philox keyed by (seed, global env id, reset_counter)."""
        ip = il = None
        if init_pos is not None:
            if init_landmarks is None:
                raise ValueError("init_pos and init_landmarks must be given together")
            ip = torch.as_tensor(np.asarray(init_pos)).to(device=self.device, dtype=self.dtype).reshape(self.B, self.N, 2).contiguous()
            il = torch.as_tensor(np.asarray(init_landmarks)).to(device=self.device, dtype=self.dtype).reshape(self.B, self.N, 2).contiguous()
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
            if m.shape != (self.B,):
                raise ValueError("mask must have shape [num_envs]")
        if reset_counter is None:
            reset_counter = self._reset_counter
            self._reset_counter += 1
        L.check(self.lib.cm3_particle_reset(self._h, C.byref(self._st), _ptr(ip), _ptr(il), _ptr(m),
                                            int(seed) & (2**64 - 1), int(reset_counter),
                                            C.byref(self._out_c), self._stream()))
        self._keep = (ip, il, m)
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
        """MultiAgentEnv.step(action_n) for all envs; actions [B,N] ints (values outside 1..4
        apply no force, environment.py:197-200)."""
        a = self._actions_tensor(actions)
        L.check(self.lib.cm3_particle_step(self._h, C.byref(self._st), _ptr(a),
                                           C.byref(self._out_c), self._stream()))
        self._keep = a
        return self.out

    def step_chained(self, actions, out, seed=0, t0=0, auto_reset=True):
        """One step that may overlap the previous launch on the device (cm3_particle_step_chained);
        see VecCheckers.step_chained for the contract on `actions` and `out`."""
        oc = out if isinstance(out, L.ParticleOutputs) else self._outputs_struct(out)
        L.check(self.lib.cm3_particle_step_chained(self._h, C.byref(self._st), _ptr(actions), int(seed) & (2**64 - 1),
                                                   int(t0), 1 if auto_reset else 0, C.byref(oc), self._stream()))
        self._keep = (actions, oc)
        return out

    def plan_rollout(self, T, actions=None, out=None, auto_reset=True, seed=0):
        """See VecCheckers.plan_rollout: a rollout launch with its ctypes arguments built once."""
        T = int(T)
        out = self.alloc_outputs(T) if out is None else out
        if actions is not None and (actions.dtype != torch.int8 or actions.device != self.device
                                    or tuple(actions.shape) != (T, self.B, self.N) or not actions.is_contiguous()):
            raise ValueError("plan_rollout needs a contiguous int8 device tensor [T,B,N]")
        oc = self._outputs_struct(out)
        fn, h, st, a = self.lib.cm3_particle_rollout, self._h, C.byref(self._st), _ptr(actions)
        seed, ar, ocr, dev = int(seed) & (2**64 - 1), 1 if auto_reset else 0, C.byref(oc), self.device

        def launch(t0=0):
            rc = fn(h, st, a, seed, int(t0), T, ar, None, ocr, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if rc != 0:
                L.check(rc)
        launch.out, launch.keep = out, (actions, oc)
        return launch

    def rollout(self, T, actions=None, seed=0, t0=0, auto_reset=False, out=None,
                record_actions=False):
        T = int(T)
        if out is None:
            out = self.alloc_outputs(T)
        a = None if actions is None else self._actions_tensor(actions, (T,))
        rec = torch.zeros(T, self.B, self.N, dtype=torch.int8, device=self.device) if record_actions else None
        oc = self._outputs_struct(out)
        L.check(self.lib.cm3_particle_rollout(self._h, C.byref(self._st), _ptr(a), int(seed) & (2**64 - 1),
                                              int(t0), T, 1 if auto_reset else 0, _ptr(rec),
                                              C.byref(oc), self._stream()))
        self._keep = (a, oc)
        if record_actions:
            out = dict(out)
            out["actions"] = rec
        return out

    def rollout_gather(self, T, dst_ptrs, dst_B, dst_env0, actions=None, seed=0, t0=0,
                       auto_reset=False):
        """Fused rollout + all-gather (cm3_particle_rollout_gather): like rollout(), but every output element is stored
        to each destination in `dst_ptrs` - a list (one entry per GPU, at most 8) of dicts
        field -> raw device pointer of a [T, dst_B, ...] array, typically the symmetric-memory
        rollout buffers of all ranks (cm3_b200.sharding.RolloutAllGather).  This shard's envs land
        at rows [dst_env0, dst_env0 + B).  Fields missing from the dicts are not written."""
        T = int(T)
        if not 1 <= len(dst_ptrs) <= L.MAX_DST:
            raise ValueError("1..%d destinations" % L.MAX_DST)
        a = None if actions is None else self._actions_tensor(actions, (T,))
        arr = (L.ParticleOutputs * len(dst_ptrs))()
        for i, d in enumerate(dst_ptrs):
            arr[i] = L.ParticleOutputs(*[(C.c_void_p(int(d[f])) if d.get(f) else None) for f in FIELDS])
        L.check(self.lib.cm3_particle_rollout_gather(self._h, C.byref(self._st), _ptr(a), int(seed) & (2**64 - 1), int(t0), T,
                            1 if auto_reset else 0, None, len(dst_ptrs), arr, int(dst_B),
                            int(dst_env0), self._stream()))
        self._keep = (a, arr)

    def step_host(self, actions, fields=FIELDS):
        """Host-buffer step: actions is a host int8 array [B,N]; the requested output fields are
        copied back into pinned host buffers and returned as NumPy views.  When every field is
        requested and the single-step buffers are packed (alloc_outputs), the outputs travel as one
        device-to-host copy (cm3_particle_step_host_packed)."""
        if self._host is None:
            self._host = self.alloc_outputs(pinned_host=True)
            self._host_actions = torch.zeros(self.B, self.N, dtype=torch.int8).pin_memory()
        self._host_actions.numpy()[...] = _to_int8_host(actions, (self.B, self.N))
        dev_block, host_block = getattr(self.out, "block", None), getattr(self._host, "block", None)
        if dev_block is not None and host_block is not None and tuple(fields) == tuple(FIELDS):
            L.check(self.lib.cm3_particle_step_host_packed(self._h, C.byref(self._st), _ptr(self._host_actions),
                                                       _ptr(self._actions_dev), C.byref(self._out_c),
                                                       _ptr(dev_block), _ptr(host_block), dev_block.numel(),
                                                       self._stream()))
        else:
            oh = L.ParticleOutputs(*[_ptr(self._host[f]) if f in fields else None for f in FIELDS])
            L.check(self.lib.cm3_particle_step_host(self._h, C.byref(self._st), _ptr(self._host_actions),
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
                                    views={f: host[f].numpy() for f in FIELDS}, step=self.lib.cm3_particle_step,
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
        memory, double buffered (cm3_particle_rollout_host): the device-to-host copy of step t
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
        L.check(self.lib.cm3_particle_rollout_host(self._h, C.byref(self._st), _ptr(hr.actions_host), _ptr(hr.actions_dev), T,
                                                  int(seed) & (2**64 - 1), int(t0), 1 if auto_reset else 0, hr.outs_c,
                                                  hr.blocks_c, _ptr(hr.host), hr.block_bytes, hr.block_bytes,
                                                  self._stream()))
        return hr.views

    # ------------------------------------------------------------------ state
    def get_state_host(self):
        """NumPy copies of the compact state through cm3_particle_get_state (checkpointing without a
        tensor library; state_dict() is the device-side equivalent)."""
        host = {k: np.empty(tuple(v.shape), dtype=torch.empty(0, dtype=v.dtype).numpy().dtype) for k, v in self.state.items()}
        hs = L.ParticleState(*[C.c_void_p(host[f].ctypes.data) for f in self.state])
        L.check(self.lib.cm3_particle_get_state(self._h, C.byref(self._st), C.byref(hs), self._stream()))
        return host

    def set_state_host(self, host):
        """Inverse of get_state_host (cm3_particle_set_state); missing keys are left untouched."""
        keep = {k: np.ascontiguousarray(host[k], dtype=torch.empty(0, dtype=self.state[k].dtype).numpy().dtype).reshape(tuple(self.state[k].shape))
                for k in self.state if k in host}
        hs = L.ParticleState(*[(C.c_void_p(keep[f].ctypes.data) if f in keep else None) for f in self.state])
        L.check(self.lib.cm3_particle_set_state(self._h, C.byref(self._st), C.byref(hs), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # pageable host memory: the arrays may go away

    def state_dict(self):
        return {k: v.clone() for k, v in self.state.items()}

    def load_state_dict(self, sd):
        for k in self.state:
            self.state[k].copy_(torch.as_tensor(sd[k]).to(self.state[k].dtype))

    def set_state(self, pos=None, vel=None, landmarks=None, steps=None, collisions=None, reached=None):
        """State injection (parity protocol)."""
        def t(x, dt):
            return torch.as_tensor(np.asarray(x)).to(device=self.device, dtype=dt)
        if vel is not None:
            self.state["sv"][:, :, 0:2] = t(vel, self.dtype).reshape(self.B, self.N, 2)
        if pos is not None:
            self.state["sv"][:, :, 2:4] = t(pos, self.dtype).reshape(self.B, self.N, 2)
        if landmarks is not None:
            self.state["landmarks"].copy_(t(landmarks, self.dtype).reshape(self.B, self.N, 2))
        if steps is not None:
            self.state["steps"].copy_(t(steps, torch.int32))
        if collisions is not None:
            self.state["collisions"].copy_(t(collisions, torch.int32))
        if reached is not None:
            r = np.asarray(reached)
            if r.ndim == 2:  # [B,N] flags -> bitmask
                r = (r.astype(np.uint8) << np.arange(self.N, dtype=np.uint8)).sum(axis=1)
            self.state["reached"].copy_(t(r, torch.uint8))
