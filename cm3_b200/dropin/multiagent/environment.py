"""Drop-in for multiagent/environment.py's MultiAgentEnv: same constructor, reset() 4-tuple
(environment.py:149) and step(action_n) 6-tuple (environment.py:123), for the multi-goal_spread
scenario.  One fused CUDA launch per step (B = 1, float64 state and outputs)."""
import numpy as np


class _Discrete(object):
    """Stand-in for gym.spaces.Discrete (only .n is ever read, environment.py:46,61-62)."""
    def __init__(self, n):
        self.n = n


class _Box(object):
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class MultiAgentEnv(object):
    metadata = {'render.modes': ['human', 'rgb_array']}

    def __init__(self, world, reset_callback=None, reward_callback=None,
                 observation_callback=None, info_callback=None,
                 done_callback=None, shared_viewer=True, max_steps=50, device="cuda:0"):
        import torch
        from cm3_b200.vec_particle import VecParticle
        self.world = world
        self.agents = self.world.policy_agents
        self.n = len(world.policy_agents)
        self.reset_callback = reset_callback
        self.reward_callback = reward_callback
        self.observation_callback = observation_callback
        self.info_callback = info_callback
        self.done_callback = done_callback
        self.discrete_action_space = True
        self.discrete_action_input = True
        self.force_discrete_action = world.discrete_action if hasattr(world, 'discrete_action') else False
        self.shared_reward = world.collaborative if hasattr(world, 'collaborative') else False
        self.time = 0
        self.max_steps = max_steps
        self.steps = 0

        scenario = getattr(world, "_cm3_scenario", None)
        callbacks_ok = scenario is not None and all(
            cb is None or getattr(cb, "__self__", None) is scenario
            for cb in (reset_callback, reward_callback, observation_callback, done_callback))
        if not callbacks_ok or reset_callback is None or reward_callback is None or \
                observation_callback is None or done_callback is None:
            raise NotImplementedError(
                "cm3_b200's MultiAgentEnv runs the multi-goal_spread callbacks as one fused CUDA "
                "kernel; pass scenario.reset_world / reward / observation / done of a Scenario "
                "loaded with multiagent.scenarios.load('multi-goal_spread.py')")
        if self.shared_reward:
            raise NotImplementedError("world.collaborative=True is not used by multi-goal_spread")
        self.scenario = scenario

        # spaces (environment.py:39-71); obs_dim probed through the callback like the reference
        self.action_space = []
        self.observation_space = []
        for agent in self.agents:
            self.action_space.append(_Discrete(world.dim_p * 2 + 1))
            obs_dim = len(observation_callback(agent, self.world))
            self.observation_space.append(_Box(-np.inf, +np.inf, (obs_dim,), np.float32))
            agent.action.c = np.zeros(self.world.dim_c)

        overrides = dict(dt=world.dt, damping=world.damping, contact_force=world.contact_force,
                         contact_margin=world.contact_margin, agent_size=world.agents[0].size,
                         mass=world.agents[0].mass)
        self._vec = VecParticle(1, self.n, scenario.config, prob_random=scenario.prob_random,
                                max_steps=max_steps, device=device, dtype=torch.float64, **overrides)
        self._results = None
        self._handed_out = None
        self._fresh = False
        # stock callbacks: the bound methods of the scenario itself, not overridden on the instance
        self._stock_callbacks = all(getattr(cb, "__func__", None) is getattr(type(scenario), name, None)
                                    for cb, name in ((reward_callback, "reward"), (observation_callback, "observation"),
                                                     (done_callback, "done")))
        import ctypes
        self._stream = ctypes.c_void_p(torch.cuda.current_stream(self._vec.device).cuda_stream)
        scenario._env = self
        self._upload_world()   # make_world() already drew an initial state (multi-goal_spread.py:62)

        self.shared_viewer = shared_viewer
        self.viewers = [None] if shared_viewer else [None] * self.n
        self._reset_render()

    # ------------------------------------------------------------------ host <-> device sync
    # The entity objects (world.agents[i].state.p_pos ...) are the reference's public state; the
    # device holds the authoritative copy.  After every launch the arrays handed to the entities are
    # READ-ONLY views of that launch's results, and the env remembers which array objects it handed
    # out: a caller that wants to move an entity assigns a new array (as reset_world does), which a
    # dozen identity checks detect - no value comparison, no device read, on the per-step path.
    def _host_arrays(self):
        pos = np.array([a.state.p_pos for a in self.world.agents], dtype=np.float64)
        vel = np.array([a.state.p_vel for a in self.world.agents], dtype=np.float64)
        lm = np.array([l.state.p_pos for l in self.world.landmarks], dtype=np.float64)
        return pos, vel, lm

    def _entity_arrays(self):
        w = self.world
        return [a.state.p_pos for a in w.agents] + [a.state.p_vel for a in w.agents] + [l.state.p_pos for l in w.landmarks]

    def _remember(self):
        self._handed_out = self._entity_arrays()
        self._handed_flags = ([bool(a.reached) for a in self.world.agents], self.scenario.collisions, self.steps)

    def _stale(self):
        h = self._handed_out
        if h is None:
            return True
        cur = self._entity_arrays()
        if len(cur) != len(h) or any(x is not y for x, y in zip(cur, h)):
            return True
        return self._handed_flags != ([bool(a.reached) for a in self.world.agents], self.scenario.collisions, self.steps)

    def _upload_world(self):
        """Host entity objects -> device state (after reset_world or a manual edit)."""
        pos, vel, lm = self._host_arrays()
        reached = np.array([[bool(a.reached) for a in self.world.agents]], dtype=np.uint8)
        self._vec.set_state(pos=pos[None], vel=vel[None], landmarks=lm[None],
                            steps=np.array([self.steps]), collisions=np.array([self.scenario.collisions]),
                            reached=reached)
        self._remember()
        self._results = None

    def _world_was_reset(self):
        self._handed_out = None
        self._results = None
        self._fresh = False

    def _ensure_synced(self):
        if self._stale():
            self._upload_world()

    def _adopt(self, views, with_reward, fresh=False):
        """Takes the results of the last launch into the entity objects and the scenario.  `views`:
        windows of the pinned host mirror (copied out here), or with fresh=True of a host copy that is
        already this call's own."""
        fields = ("global_state", "obs_others", "obs_self", "done") + (("reward", "reward_n", "collisions", "reached") if with_reward else ())
        res = {f: (views[f][0] if fresh else views[f][0].copy()) for f in fields}
        if not with_reward:
            res["reward"] = res["reward_n"] = None
        gs = res["global_state"]
        gs.flags.writeable = False
        for i, agent in enumerate(self.world.agents):
            st = agent.state
            st.p_vel = gs[i, 0:2]
            st.p_pos = gs[i, 2:4]
        if with_reward:   # a reset / observe-only launch leaves reached and the counter as they are
            reached = int(res["reached"])
            for i, agent in enumerate(self.world.agents):
                agent.reached = bool((reached >> i) & 1)
            self.scenario.collisions = int(res["collisions"])
        self._remember()
        self._results = res
        return res

    def _current_results(self):
        """Results for the scenario callbacks: the last launch's, or a fresh observe-only launch
        when the host state changed since."""
        if self._fresh and self._results is not None:
            return self._results
        if self._results is None or self._stale():
            keep = self._results
            self._upload_world()
            self._vec.reset(mask=np.zeros(1, dtype=np.uint8))  # observe only
            res = self._adopt(self._vec.download(), with_reward=False)
            if keep is not None and keep.get("reward_n") is not None:
                res["reward"], res["reward_n"] = keep["reward"], keep["reward_n"]
        return self._results

    # ------------------------------------------------------------------ reference API
    def step(self, action_n):
        """-> (global_state [N,4], obs_others_n, obs_n, reward, reward_n, done), environment.py:123"""
        self.agents = self.world.policy_agents
        self._ensure_synced()
        a = np.array([[int(x) for x in action_n[:self.n]]], dtype=np.int64)
        self.steps += 1
        # one launch + one stream wait: the kernel reads the actions from and writes every field (and
        # the reached / collision flags) to pinned host memory directly (VecParticle.step_mapped)
        res = self._adopt(self._vec.step_mapped(a, stream=self._stream, copy=True), with_reward=True, fresh=True)
        if self._stock_callbacks:
            # The callbacks are the scenario's own methods (checked in the constructor), whose results ARE
            # the kernel's outputs: hand them out directly, in the order environment.py:95-104 calls them
            obs_all, oth_all, rew_all = res["obs_self"], res["obs_others"], res["reward_n"]
            obs_n = [obs_all[i].copy() for i in range(self.n)]
            obs_others_n = [oth_all[i].copy() for i in range(self.n)]
            reward_n = [rew_all[i] for i in range(self.n)]
            done_n = [agent.reached for agent in self.agents]
        else:
            obs_n, obs_others_n, reward_n, done_n = [], [], [], []
            self._fresh = True          # nothing can have touched the entities since _adopt: the
            try:                        # callbacks below skip the staleness check
                for agent in self.agents:   # callback order of environment.py:95-104
                    obs_self, obs_others = self._get_obs(agent)
                    obs_n.append(obs_self)
                    obs_others_n.append(obs_others)
                    reward_n.append(self._get_reward(agent))
                    done_n.append(self._get_done(agent))
            finally:
                self._fresh = False
        reward = np.float64(res["reward"])   # np.sum(reward_n), evaluated on the device in index order
        global_state = res["global_state"].copy()
        done = bool(res["done"])
        assert done == (self.steps == self.max_steps or bool(np.all(done_n)))
        return global_state, obs_others_n, obs_n, reward, reward_n, done

    def reset(self):
        """-> (global_state, obs_others_n, obs_n, done), environment.py:149"""
        self.reset_callback(self.world)      # host RNG draws, like the reference
        self._reset_render()
        self.steps = 0
        pos, vel, lm = self._host_arrays()
        self._vec.reset(init_pos=pos[None], init_landmarks=lm[None])   # zeroes velocities, steps, counters
        res = self._adopt(self._vec.download(), with_reward=False)
        obs_n, obs_others_n, done_n = [], [], []
        self.agents = self.world.policy_agents
        for agent in self.agents:
            obs_self, obs_others = self._get_obs(agent)
            obs_n.append(obs_self)
            obs_others_n.append(obs_others)
            done_n.append(self._get_done(agent))
        return res["global_state"].copy(), obs_others_n, obs_n, np.any(done_n)

    def _get_info(self, agent):
        if self.info_callback is None:
            return {}
        return self.info_callback(agent, self.world)

    def _get_obs(self, agent):
        return self.observation_callback(agent, self.world)

    def _get_done(self, agent):
        return self.done_callback(agent, self.world)

    def _get_reward(self, agent):
        return self.reward_callback(agent, self.world)

    def _reset_render(self):
        self.render_geoms = None
        self.render_geoms_xform = None

    def render(self, mode='human'):
        raise NotImplementedError("rendering (pyglet) is out of scope; every training path of the "
                                  "reference runs with render=False (train_onpolicy.py:394)")
