"""VecCheckers - B instances of the reference's Checkers env (env/checkers.py) stepped by one
CUDA kernel launch.  torch tensors only hold the HBM (and pinned host) buffers; all
computation happens in libcm3env.so through the C ABI of include/cm3env.h.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._buffers import alloc_fields, _to_int8_host

FIELDS = L.CheckersOutputs.FIELDS
OBS_FIELDS = ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class VecCheckers(object):
    """Constructor arguments after `num_envs` are those of Checkers.__init__
    (env/checkers.py:5-6), with the same defaults."""

    def __init__(self, num_envs, n_rows=3, n_columns=16, n_obs=2, agents_r=(0, 2),
                 agents_c=(16, 16), n_agents=1, max_steps=50, device="cuda:0",
                 dtype=torch.float32, env_id_offset=0, tile_dtype=None):
        # the reference's own asserts (checkers.py:16-17)
        assert n_rows % 2 == 1
        assert n_columns % 2 == 0
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 or torch.float64")
        # grid / obs_self_t only hold {-1, 0, +1}: tile_dtype=torch.int8 writes them as bytes
        tile_dtype = dtype if tile_dtype is None else tile_dtype
        if tile_dtype not in (dtype, torch.int8):
            raise ValueError("tile_dtype must be None (= dtype) or torch.int8")
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
        cfg.tile = L.TILE_I8 if tile_dtype == torch.int8 and dtype != torch.int8 else L.TILE_REAL
        cfg.device = dev_index
        cfg.env_id_offset = int(env_id_offset)
        self.env_id_offset = int(env_id_offset)
        h = C.c_void_p()
        L.check(self.lib.cm3_checkers_create(C.byref(cfg), C.byref(h)))
        self._h = h

        dev = self.device
        self.state = dict(
            remaining=torch.zeros(self.B, dtype=torch.int64, device=dev),
            agents=torch.zeros(self.B, self.N, dtype=torch.int32, device=dev),
            meta=torch.zeros(self.B, dtype=torch.int32, device=dev))
        self._st = L.CheckersState(_ptr(self.state["remaining"]), _ptr(self.state["agents"]),
                                   _ptr(self.state["meta"]))
        self.out = self.alloc_outputs()
        self._out_c = self._outputs_struct(self.out)
        self._actions_dev = torch.zeros(self.B, self.N, dtype=torch.int8, device=dev)
        self._host = None
        self._is_reset = False

    # ------------------------------------------------------------------ buffers
    def field_shapes(self):
        B, N, W = self.B, self.N, self.W
        return dict(grid=(B, self.n_rows, self.n_columns + 1, 2), vec=(B, N, 4),
                    obs_others=(B, N, self.L_others), obs_self_t=(B, N, W, W, 3),
                    obs_self_v=(B, N, 4), reward=(B,), local_rewards=(B, N), done=(B,))

    def bytes_per_env_step(self):
        """Algorithmic bytes of one env-step: outputs + state read/write + actions."""
        out = sum(int(np.prod(s[1:])) * self.field_dtype(k).itemsize for k, s in self.field_shapes().items())
        state = 2 * (8 + 4 * self.N + 4)
        return out + state + self.N

    def field_dtype(self, k):
        if k == "done":
            return torch.uint8
        return self.tile_dtype if k in ("grid", "obs_self_t") else self.dtype

    def out_bytes_per_env_step(self):
        return sum(int(np.prod(s[1:])) * self.field_dtype(k).itemsize for k, s in self.field_shapes().items())

    def alloc_outputs(self, T=None, pinned_host=False):
        """Zeroed output buffers: [B, ...] per field (T=None) or [T, B, ...] rollout buffers.  The
        single-step set of a batch that is a multiple of 32 envs is one packed allocation (every
        field stays 16-byte aligned), so that step_host needs a single device-to-host copy."""
        lead = () if T is None else (int(T),)
        return alloc_fields(self.field_shapes(), self.field_dtype, lead, device=self.device,
                            pinned=pinned_host, packed=T is None and self.B % 32 == 0)

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
        dict (device tensors); reward / local_rewards are not meaningful after a reset."""
        gi = None
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
