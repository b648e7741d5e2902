"""Drop-in for multiagent/scenarios/multi-goal_spread.py: the same Scenario plugin surface
(make_world / reset_world / reward / observation / done / is_collision / benchmark_data).

reset_world draws the initial state on the HOST from the same global RNG streams, in the same
order, as the reference (multi-goal_spread.py:75-89: one random.random(), then np.random.uniform /
np.random.normal per entity), so a trainer that seeds `random` and `np.random` gets the same
episodes.  Everything that is computed per step - observation, reward, reached/done, collision
count - is produced by the fused CUDA kernel when MultiAgentEnv.step runs; the callbacks below
hand out those results per agent, which is how MultiAgentEnv consumes them (environment.py:95-104).
"""
import random

import numpy as np

from multiagent.core import World, Agent, Landmark
from multiagent.scenario import BaseScenario

# multi-goal_spread.py:7-16
colors = np.array([[221, 127, 106], [204, 169, 120], [191, 196, 139], [176, 209, 152],
                   [152, 209, 202], [152, 183, 209], [152, 152, 209], [185, 152, 209],
                   [209, 152, 203], [209, 152, 161]])


class Scenario(BaseScenario):
    def make_world(self, n_agents, config, prob_random):
        """multi-goal_spread.py:19-63"""
        world = World()
        world.dim_c = 0
        self.n_agents = n_agents
        self.agents_x = config['agents_x']
        self.agents_y = config['agents_y']
        self.landmarks_x = config['landmarks_x']
        self.landmarks_y = config['landmarks_y']
        self.initial_std = config['initial_std']
        self.prob_random = prob_random
        self.config = dict(config)
        world.collaborative = False
        world.agents = [Agent() for i in range(n_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.idx = i
            agent.collide = True
            agent.silent = True
            agent.size = 0.15
            agent.reached = False
        world.landmarks = [Landmark() for i in range(n_agents)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.idx = i
            landmark.collide = False
            landmark.movable = False
        self.colors = colors
        self.collisions = 0
        self._env = None          # set by MultiAgentEnv: the device-backed stepper
        world._cm3_scenario = self
        self.reset_world(world)
        return world

    def reset_world(self, world):
        """multi-goal_spread.py:65-93 - same draws, same order, same global RNGs."""
        for i, agent in enumerate(world.agents):
            agent.color = self.colors[i] / 256
        for i, landmark in enumerate(world.landmarks):
            landmark.color = self.colors[i] / 256
        rand_num = random.random()
        for i, agent in enumerate(world.agents):
            if rand_num < self.prob_random:
                agent.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            else:
                x = self.agents_x[i] + np.random.normal(0, self.initial_std)
                y = self.agents_y[i] + np.random.normal(0, self.initial_std)
                agent.state.p_pos = np.array([x, y])
            agent.state.p_vel = np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
            agent.reached = False
        for i, landmark in enumerate(world.landmarks):
            if rand_num < self.prob_random:
                landmark.state.p_pos = np.random.uniform(-1, +1, world.dim_p)
            else:
                landmark.state.p_pos = np.array([self.landmarks_x[i], self.landmarks_y[i]])
            landmark.state.p_vel = np.zeros(world.dim_p)
        self.collisions = 0
        if self._env is not None:
            self._env._world_was_reset()

    # ---- per-agent views of what the fused kernel computed -------------------------------
    def _results(self, world):
        if self._env is None:
            raise RuntimeError("scenario callbacks need a MultiAgentEnv bound to this world")
        return self._env._current_results()

    def observation(self, agent, world):
        """-> (obs_self [4], obs_others [4*max(N-1,1)]), multi-goal_spread.py:145-154"""
        if self._env is None:
            # MultiAgentEnv.__init__ probes the observation length before the env exists
            # (environment.py:69): shapes only
            n = max(self.n_agents - 1, 1)
            return (np.concatenate([agent.state.p_vel, agent.state.p_pos]), np.zeros(4 * n))
        res = self._results(world)
        return res["obs_self"][agent.idx].copy(), res["obs_others"][agent.idx].copy()

    def reward(self, agent, world):
        """Reward of `agent` for the last env.step (multi-goal_spread.py:121-138); the reached flag
        and the collision counter were updated by the same kernel launch."""
        res = self._results(world)
        if res.get("reward_n") is None:
            raise RuntimeError("reward() is defined after a step()")
        return res["reward_n"][agent.idx]

    def done(self, agent, world):
        """multi-goal_spread.py:140-143"""
        return True if agent.reached else False

    def is_collision(self, agent1, agent2):
        """multi-goal_spread.py:114-118 (host helper on the synced positions)"""
        delta_pos = agent1.state.p_pos - agent2.state.p_pos
        dist = np.sqrt(np.sum(np.square(delta_pos)))
        dist_min = agent1.size + agent2.size
        return True if dist < dist_min else False

    def benchmark_data(self, agent, world):
        """multi-goal_spread.py:95-111 - never wired by the trainers (info_callback=None)."""
        rew = 0
        collisions = 0
        occupied_landmarks = 0
        min_dists = 0
        for l in world.landmarks:
            dists = [np.sqrt(np.sum(np.square(a.state.p_pos - l.state.p_pos))) for a in world.agents]
            min_dists += min(dists)
            rew -= min(dists)
            if min(dists) < 0.1:
                occupied_landmarks += 1
        if agent.collide:
            for a in world.agents:
                if self.is_collision(a, agent):
                    rew -= 1
                    collisions += 1
        return (rew, collisions, min_dists, occupied_landmarks)
