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
    AGENT_SIZE = 0.15   # multi-goal_spread.py:47

    def make_world(self, n_agents, config, prob_random):
        """multi-goal_spread.py:19-63: N agents (collide, silent, size 0.15), N static landmarks
        that do not collide, preset positions from `config`, and one initial reset_world()."""
        self.n_agents, self.prob_random = n_agents, prob_random
        self.config = dict(config)
        for key in ("agents_x", "agents_y", "landmarks_x", "landmarks_y", "initial_std"):
            setattr(self, key, config[key])
        self.colors, self.collisions = colors, 0
        self._env = None                      # set by MultiAgentEnv: the device-backed stepper

        world = World()
        world.dim_c, world.collaborative = 0, False
        world._cm3_scenario = self
        for idx in range(n_agents):
            agent, landmark = Agent(), Landmark()
            agent.name, landmark.name = "agent %d" % idx, "landmark %d" % idx
            agent.idx = landmark.idx = idx
            agent.collide, agent.silent, agent.size, agent.reached = True, True, self.AGENT_SIZE, False
            landmark.collide, landmark.movable = False, False
            world.agents.append(agent)
            world.landmarks.append(landmark)
        self.reset_world(world)
        return world

    def _draw_initial_positions(self, dim_p):
        """The host RNG draws of multi-goal_spread.py:75-91, in the reference's order: one
        random.random(); then either uniform(-1, 1, 2) per agent followed by the same per landmark,
        or two scalar normal(0, initial_std) per agent (x, then y - drawn even when std == 0) with
        the landmarks on their presets.  Returns (agent_pos [N,2], landmark_pos [N,2])."""
        n = self.n_agents
        if random.random() < self.prob_random:
            agent_pos = np.stack([np.random.uniform(-1, +1, dim_p) for _ in range(n)])
            landmark_pos = np.stack([np.random.uniform(-1, +1, dim_p) for _ in range(n)])
        else:
            jitter = np.array([[np.random.normal(0, self.initial_std), np.random.normal(0, self.initial_std)]
                               for _ in range(n)])
            agent_pos = np.stack([self.agents_x[:n], self.agents_y[:n]], axis=1) + jitter
            landmark_pos = np.stack([self.landmarks_x[:n], self.landmarks_y[:n]], axis=1).astype(float)
        return agent_pos, landmark_pos

    def reset_world(self, world):
        """multi-goal_spread.py:65-93 - same draws, same order, same global RNG streams, so a
        trainer that seeds `random` and `np.random` sees the reference's episodes."""
        agent_pos, landmark_pos = self._draw_initial_positions(world.dim_p)
        for i, (agent, landmark) in enumerate(zip(world.agents, world.landmarks)):
            agent.color = landmark.color = self.colors[i] / 256
            agent.state.p_pos, landmark.state.p_pos = agent_pos[i].copy(), landmark_pos[i].copy()
            agent.state.p_vel, landmark.state.p_vel = np.zeros(world.dim_p), np.zeros(world.dim_p)
            agent.state.c = np.zeros(world.dim_c)
            agent.reached = False
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
        gap = np.sqrt(np.sum(np.square(agent1.state.p_pos - agent2.state.p_pos)))
        return bool(gap < agent1.size + agent2.size)

    def benchmark_data(self, agent, world):
        """multi-goal_spread.py:95-111 -> (rew, collisions, min_dists, occupied_landmarks).  Never
        wired by the trainers (info_callback=None, train_onpolicy.py:119); evaluated on the host
        from the synced positions."""
        pos = np.array([a.state.p_pos for a in world.agents])
        lms = np.array([l.state.p_pos for l in world.landmarks])
        nearest = np.sqrt(np.sum(np.square(pos[None, :, :] - lms[:, None, :]), axis=2)).min(axis=1)
        rew, min_dists = 0, 0
        for d in nearest:            # accumulate in landmark order like the reference's loop
            min_dists += d
            rew -= d
        occupied_landmarks = int((nearest < 0.1).sum())
        collisions = 0
        if agent.collide:
            collisions = sum(1 for a in world.agents if self.is_collision(a, agent))
            rew -= collisions
        return (rew, collisions, min_dists, occupied_landmarks)
