"""Drop-in for the reference's env/checkers.py: same class, constructor, reset(goals) 5-tuple and
step(actions) 7-tuple (env/checkers.py:5-6, :291, :262), computed by the CUDA kernels of
libcm3env.so (B = 1, float64 outputs so values equal the reference's float64 exactly).
"""
import numpy as np


class Checkers(object):

    def __init__(self, n_rows=3, n_columns=16, n_obs=2, agents_r=[0, 2],
                 agents_c=[16, 16], n_agents=1, max_steps=50, device="cuda:0"):
        assert(n_rows % 2 == 1)      # checkers.py:16
        assert(n_columns % 2 == 0)   # checkers.py:17
        import torch
        from cm3_b200.vec_checkers import VecCheckers
        self.n_rows = n_rows
        self.n_columns = n_columns
        self.n_obs = n_obs
        self.total_rows = self.n_rows + 2 * self.n_obs          # :24
        self.total_columns = self.n_columns + 2 * self.n_obs + 1  # :25
        self.max_collectible = self.n_rows * self.n_columns      # :28
        self.n_agents = n_agents
        self.max_steps = max_steps
        self.agents_r = np.array(agents_r) + self.n_obs          # :34
        self.agents_c = np.array(agents_c) + self.n_obs          # :35
        self.goals = None
        self._vec = VecCheckers(1, n_rows, n_columns, n_obs, list(agents_r)[:n_agents],
                                list(agents_c)[:n_agents], n_agents, max_steps, device=device,
                                dtype=torch.float64)
        import ctypes
        # the reference's trainers never switch CUDA streams: the handle of the stream current at
        # construction is cached, one attribute look-up less on the per-step path
        self._stream = ctypes.c_void_p(torch.cuda.current_stream(self._vec.device).cuda_stream)

    # ------------------------------------------------------------------ tuple assembly
    _OBS = ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v")

    def _host(self, views, fields):
        # `views` are windows of the pinned host mirror (one device-to-host copy per call of the
        # facade); the reference hands out fresh arrays, so each field is copied out
        return {f: views[f][0].copy() for f in fields}

    def _obs_tuple(self, o):
        n = self.n_agents
        global_state = (o["grid"], [o["vec"][i] for i in range(n)])   # get_global_state, :79-94
        obs_others = [o["obs_others"][i] for i in range(n)]           # :151
        obs_self_t = [o["obs_self_t"][i] for i in range(n)]           # :141
        obs_self_v = [o["obs_self_v"][i] for i in range(n)]           # :142
        return global_state, obs_others, obs_self_t, obs_self_v

    def reset(self, goals):
        """Returns (global_state, obs_others, obs_self_t, obs_self_v, False), checkers.py:291."""
        self.goals = goals
        g = np.asarray(goals)
        if self.n_agents == 1:  # :271-276 - the start row follows the goal
            goal_idx = np.where(g[0] == 1)[0][0]
            self.agents_r = np.array([0 if goal_idx == 0 else 2]) + self.n_obs
        self._vec.reset(goals=g.reshape(self.n_agents, 2))
        o = self._host(self._vec.download(), self._OBS)
        return self._obs_tuple(o) + (False,)

    def step(self, actions):
        """Returns (global_state, obs_others, obs_self_t, obs_self_v, total_reward, local_rewards,
        done), checkers.py:262."""
        a = np.asarray(actions).reshape(1, self.n_agents)
        # one launch + one stream wait: the kernel reads the actions from and writes every field to
        # pinned host memory directly (VecCheckers.step_mapped)
        v = self._vec.step_mapped(a, stream=self._stream, copy=True)   # views of one fresh host copy
        o = {f: v[f][0] for f in self._OBS}
        local_rewards = [float(x) for x in v["local_rewards"][0]]
        return self._obs_tuple(o) + (np.float64(v["reward"][0]), local_rewards, bool(v["done"][0]))

    # ------------------------------------------------------------------ read-only views of state
    def _observe(self):
        self._vec.reset(mask=np.zeros(1, dtype=np.uint8))  # no env selected: observe only
        return self._host(self._vec.download(), self._OBS)

    def get_valid_grid(self):
        return self._observe()["grid"]

    def get_global_state(self):
        return self._obs_tuple(self._observe())[0]

    def get_local_observation(self):
        gs, oo, ot, ov = self._obs_tuple(self._observe())
        return oo, ot, ov

    def normalize(self, location):
        """checkers.py:112-125 (pure formatting helper)."""
        loc = np.array(location, dtype=float)
        if loc.ndim == 1:
            loc[0] = (loc[0] - self.total_rows / 2.0) / self.total_rows
            loc[1] = (loc[1] - self.total_columns / 2.0) / self.total_columns
        elif loc.ndim == 2:
            loc[:, 0] = (loc[:, 0] - self.total_rows / 2.0) / self.total_rows
            loc[:, 1] = (loc[:, 1] - self.total_columns / 2.0) / self.total_columns
        return loc

    @property
    def steps(self):
        return int(self._vec.unpack_state()["steps"][0])

    @property
    def agents_location(self):
        st = self._vec.unpack_state()
        return np.stack([st["r"][0], st["c"][0]], axis=1).astype(int)

    @property
    def agents_collected(self):
        st = self._vec.unpack_state()
        return np.stack([st["n_green"][0], st["n_orange"][0]], axis=1).astype(float)

    def _unsupported(self, *a, **k):
        raise NotImplementedError("the per-agent mutators of the reference (agent_act, get_reward, "
                                  "populate_world) are fused into the step kernel; use step()")
    agent_act = get_reward = populate_world = _unsupported
