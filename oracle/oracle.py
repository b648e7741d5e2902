"""ctypes wrapper around oracle/_build/libcm3_oracle.so.  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and
--impl reference legs.  The product package (cm3_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libcm3_oracle.so")
MAX_AGENTS = 8


def build(force=False):
    if force or not os.path.isfile(LIB_PATH) or (
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "cm3_oracle.c"))):
        subprocess.check_call(["make", "-s", "-C", HERE] + (["-B"] if force else []))
    return LIB_PATH


class _CkCfg(C.Structure):
    _fields_ = [("n_rows", C.c_int), ("n_columns", C.c_int), ("n_obs", C.c_int),
                ("n_agents", C.c_int), ("max_steps", C.c_int),
                ("agents_r", C.c_int * MAX_AGENTS), ("agents_c", C.c_int * MAX_AGENTS)]


class _CkOut(C.Structure):
    _fields_ = [("grid", C.c_void_p), ("vec", C.c_void_p), ("obs_others", C.c_void_p),
                ("obs_self_t", C.c_void_p), ("obs_self_v", C.c_void_p),
                ("reward", C.c_void_p), ("local_rewards", C.c_void_p), ("done", C.c_void_p)]


class _PtCfg(C.Structure):
    _fields_ = [("n_agents", C.c_int), ("max_steps", C.c_int), ("dt", C.c_double),
                ("damping", C.c_double), ("contact_force", C.c_double),
                ("contact_margin", C.c_double), ("agent_size", C.c_double),
                ("mass", C.c_double), ("sensitivity", C.c_double),
                ("reach_thresh", C.c_double)]


class _PtOut(C.Structure):
    _fields_ = [("global_state", C.c_void_p), ("obs_others", C.c_void_p),
                ("obs_self", C.c_void_p), ("reward", C.c_void_p), ("reward_n", C.c_void_p),
                ("done", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.ock_create.restype = C.c_void_p
        L.ock_create.argtypes = [C.POINTER(_CkCfg), C.c_int]
        L.ock_destroy.argtypes = [C.c_void_p]
        L.ock_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_CkOut), C.c_int]
        L.ock_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_CkOut), C.c_int]
        L.ock_get_steps.argtypes = [C.c_void_p, C.c_void_p]
        L.opt_default_config.argtypes = [C.POINTER(_PtCfg), C.c_int, C.c_int]
        L.opt_create.restype = C.c_void_p
        L.opt_create.argtypes = [C.POINTER(_PtCfg), C.c_int]
        L.opt_destroy.argtypes = [C.c_void_p]
        L.opt_set_state.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.opt_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.opt_reset_to.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(_PtOut)]
        L.opt_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_PtOut), C.c_int]
        L.oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_philox_actions.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int64,
                                            C.c_int, C.c_int, C.c_void_p]
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads():
    return lib().oracle_max_threads()


class OracleCheckers(object):
    """B independent instances of the reference's Checkers env (float64 outputs)."""

    def __init__(self, num_envs, n_rows=3, n_columns=16, n_obs=2, agents_r=(0, 2),
                 agents_c=(16, 16), n_agents=1, max_steps=50, nthreads=1):
        cfg = _CkCfg(n_rows, n_columns, n_obs, n_agents, max_steps)
        for i in range(n_agents):
            cfg.agents_r[i] = int(agents_r[i])
            cfg.agents_c[i] = int(agents_c[i])
        self.B, self.N, self.R, self.Cc, self.O = num_envs, n_agents, n_rows, n_columns, n_obs
        self.nthreads = nthreads
        self._h = lib().ock_create(C.byref(cfg), num_envs)
        if not self._h:
            raise ValueError("bad Checkers config (n_rows odd, n_columns even required)")
        B, N, W = num_envs, n_agents, 2 * n_obs + 1
        L = 2 * max(N - 1, 1)
        self.out = dict(
            grid=np.zeros((B, n_rows, n_columns + 1, 2)), vec=np.zeros((B, N, 4)),
            obs_others=np.zeros((B, N, L)), obs_self_t=np.zeros((B, N, W, W, 3)),
            obs_self_v=np.zeros((B, N, 4)), reward=np.zeros(B),
            local_rewards=np.zeros((B, N)), done=np.zeros(B, dtype=np.uint8))
        self._o = _CkOut(*[_ptr(self.out[k]) for k in (
            "grid", "vec", "obs_others", "obs_self_t", "obs_self_v", "reward",
            "local_rewards", "done")])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ock_destroy(self._h)
            self._h = None

    def reset(self, goal_idx, mask=None):
        """goal_idx: [B, N] ints in {0, 1} (argmax of the reference's one-hot goals)."""
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(goal_idx, dtype=np.int32),
                                                 (self.B, self.N)))
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().ock_reset(self._h, _ptr(g), _ptr(m), C.byref(self._o), self.nthreads)
        if mask is None:
            self.out["reward"][:] = 0
            self.out["local_rewards"][:] = 0
        return self.out

    def step(self, actions):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.int32).reshape(self.B, self.N))
        lib().ock_step(self._h, _ptr(a), C.byref(self._o), self.nthreads)
        return self.out

    def steps(self):
        s = np.zeros(self.B, dtype=np.int32)
        lib().ock_get_steps(self._h, _ptr(s))
        return s


class OracleParticle(object):
    """B independent instances of MultiAgentEnv + multi-goal_spread (float64)."""

    def __init__(self, num_envs, n_agents, max_steps=50, nthreads=1, **overrides):
        cfg = _PtCfg()
        lib().opt_default_config(C.byref(cfg), n_agents, max_steps)
        for k, v in overrides.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self.B, self.N = num_envs, n_agents
        self.nthreads = nthreads
        self._h = lib().opt_create(C.byref(cfg), num_envs)
        if not self._h:
            raise ValueError("bad particle config")
        B, N = num_envs, n_agents
        L = 4 * max(N - 1, 1)
        self.out = dict(
            global_state=np.zeros((B, N, 4)), obs_others=np.zeros((B, N, L)),
            obs_self=np.zeros((B, N, 4)), reward=np.zeros(B), reward_n=np.zeros((B, N)),
            done=np.zeros(B, dtype=np.uint8))
        self._o = _PtOut(*[_ptr(self.out[k]) for k in (
            "global_state", "obs_others", "obs_self", "reward", "reward_n", "done")])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().opt_destroy(self._h)
            self._h = None

    def reset_to(self, pos, landmarks, mask=None):
        p = np.ascontiguousarray(pos, dtype=np.float64).reshape(self.B, self.N, 2)
        l = np.ascontiguousarray(landmarks, dtype=np.float64).reshape(self.B, self.N, 2)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().opt_reset_to(self._h, _ptr(p), _ptr(l), _ptr(m), C.byref(self._o))
        if mask is None:
            self.out["reward"][:] = 0
            self.out["reward_n"][:] = 0
        return self.out

    def set_state(self, pos=None, vel=None, landmarks=None, steps=None, collisions=None,
                  reached=None):
        def f(a, dt):
            return None if a is None else np.ascontiguousarray(a, dtype=dt)
        args = [f(pos, np.float64), f(vel, np.float64), f(landmarks, np.float64),
                f(steps, np.int32), f(collisions, np.int64), f(reached, np.uint8)]
        lib().opt_set_state(self._h, *[_ptr(a) for a in args])

    def get_state(self):
        B, N = self.B, self.N
        st = dict(pos=np.zeros((B, N, 2)), vel=np.zeros((B, N, 2)),
                  landmarks=np.zeros((B, N, 2)), steps=np.zeros(B, dtype=np.int32),
                  collisions=np.zeros(B, dtype=np.int64),
                  reached=np.zeros((B, N), dtype=np.uint8))
        lib().opt_get_state(self._h, *[_ptr(st[k]) for k in (
            "pos", "vel", "landmarks", "steps", "collisions", "reached")])
        return st

    def step(self, actions):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.int32).reshape(self.B, self.N))
        lib().opt_step(self._h, _ptr(a), C.byref(self._o), self.nthreads)
        return self.out


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def philox_actions(seed, env0, B, N, t0, T, n_actions=5):
    out = np.zeros((T, B, N), dtype=np.int8)
    lib().oracle_philox_actions(seed, env0, B, N, t0, T, n_actions, _ptr(out))
    return out
