#!/usr/bin/env python
"""Golden-vector generator: runs the UNMODIFIED reference (/root/reference, through
oracle/ref_shims.py) on fixed op sequences and writes small fixtures to tests/golden/.

TEST INFRASTRUCTURE ONLY - runs in the build container (the reference tree does not
travel to the GPU box); the committed .npz files are what the tests read.

    python oracle/gen_golden.py            # regenerate everything
    python oracle/gen_golden.py --check    # regenerate in memory and diff against the files

A fixture is a batch of K single-env traces that share one aligned op sequence
(op_kind[s] = 0 reset / 1 step), so a batched implementation can replay all K traces as
K env instances.  Steps deliberately continue past `done` (the
reference does not auto-reset and `steps == max_steps` is an equality test).

Reference entry points exercised:
  env/checkers.py:265 (reset), :228 (step)
  multiagent/environment.py:125 (reset), :81 (step)
  multiagent/scenarios/multi-goal_spread.py:19,65,121,140,145
"""
import argparse
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 12341  # alg/config.json:6

RESET, STEP = 0, 1


def aligned_ops(n_episodes, steps_per_episode):
    ops = []
    for _ in range(n_episodes):
        ops.append(RESET)
        ops.extend([STEP] * steps_per_episode)
    return np.array(ops, dtype=np.int8)


# --------------------------------------------------------------------------- checkers
def checkers_trace(ck, ctor, goals_per_episode, ops, actions):
    """One env instance.  ctor = (n_rows, n_columns, n_obs, agents_r, agents_c, n_agents,
    max_steps); actions[s, n] is read at step ops."""
    n_rows, n_cols, n_obs, ar, ac, n, max_steps = ctor
    env = ck.Checkers(n_rows, n_cols, n_obs, list(ar), list(ac), n, max_steps)
    S = len(ops)
    W = 2 * n_obs + 1
    L = 2 * max(n - 1, 1)
    out = dict(
        grid=np.zeros((S, n_rows, n_cols + 1, 2)),
        vec=np.zeros((S, n, 4)),
        obs_others=np.zeros((S, n, L)),
        obs_self_t=np.zeros((S, n, W, W, 3)),
        obs_self_v=np.zeros((S, n, 4)),
        reward=np.zeros(S),
        local_rewards=np.zeros((S, n)),
        done=np.zeros(S, dtype=np.uint8),
        goals=np.zeros((S, n, 2)),
    )
    ep = -1
    for s, op in enumerate(ops):
        if op == RESET:
            ep += 1
            goals = np.array(goals_per_episode[ep], dtype=float)
            gs, oo, ot, ov, done = env.reset(goals)
            rew, lrew = 0.0, [0.0] * n
            out["goals"][s] = goals
        else:
            gs, oo, ot, ov, rew, lrew, done = env.step(actions[s])
            out["goals"][s] = out["goals"][s - 1]
        out["grid"][s] = np.array(gs[0])  # snapshot (H5: the reference returns a view)
        out["vec"][s] = np.array(gs[1])
        out["obs_others"][s] = np.array(oo)
        out["obs_self_t"][s] = np.array(ot)
        out["obs_self_v"][s] = np.array(ov)
        out["reward"][s] = rew
        out["local_rewards"][s] = np.array(lrew, dtype=float)
        out["done"][s] = bool(done)
    return out


def gen_checkers(ck, name, ctor, goal_choices, K, n_episodes, steps_per_episode, rng,
                 bad_action_rate=0.03):
    n = ctor[5]
    ops = aligned_ops(n_episodes, steps_per_episode)
    S = len(ops)
    traces = []
    actions = rng.integers(0, 5, size=(K, S, n)).astype(np.int32)
    bad = rng.random(size=actions.shape) < bad_action_rate
    actions[bad] = rng.choice([-1, 5, 7, -128, 127], size=int(bad.sum()))
    goals = []
    for k in range(K):
        gk = [goal_choices[rng.integers(len(goal_choices))] for _ in range(n_episodes)]
        goals.append(gk)
        traces.append(checkers_trace(ck, ctor, gk, ops, actions[k]))
    fix = {key: np.stack([t[key] for t in traces]) for key in traces[0]}
    fix["ops"] = ops
    fix["actions"] = actions
    fix["ctor_n_rows"] = np.int64(ctor[0])
    fix["ctor_n_columns"] = np.int64(ctor[1])
    fix["ctor_n_obs"] = np.int64(ctor[2])
    fix["ctor_agents_r"] = np.array(ctor[3], dtype=np.int64)
    fix["ctor_agents_c"] = np.array(ctor[4], dtype=np.int64)
    fix["ctor_n_agents"] = np.int64(n)
    fix["ctor_max_steps"] = np.int64(ctor[6])
    return name, fix


# --------------------------------------------------------------------------- particle
def seek_actions(world, rng, greedy_prob):
    """Goal-seeking scripted policy (drives agents through each other so the contact
    force, collision penalty and `reached` paths are exercised)."""
    acts = []
    for i, agent in enumerate(world.agents):
        if rng.random() >= greedy_prob:
            acts.append(int(rng.integers(0, 5)))
            continue
        d = world.landmarks[i].state.p_pos - agent.state.p_pos
        if abs(d[0]) >= abs(d[1]):
            acts.append(2 if d[0] > 0 else 1)   # environment.py:197-198: 1 -> -x, 2 -> +x
        else:
            acts.append(4 if d[1] > 0 else 3)   # environment.py:199-200: 3 -> -y, 4 -> +y
    return acts


def particle_trace(MAE, scenarios, n_agents, cfg, prob_random, max_steps, ops, actions,
                   seek_rng=None, greedy_prob=0.85):
    scenario = scenarios.load("multi-goal_spread.py").Scenario()
    world = scenario.make_world(n_agents, cfg, prob_random)
    env = MAE(world, scenario.reset_world, scenario.reward, scenario.observation, None,
              scenario.done, max_steps=max_steps)
    n = n_agents
    S = len(ops)
    L = 4 * max(n - 1, 1)
    out = dict(
        global_state=np.zeros((S, n, 4)),
        obs_others=np.zeros((S, n, L)),
        obs_self=np.zeros((S, n, 4)),
        reward=np.zeros(S),
        reward_n=np.zeros((S, n)),
        done=np.zeros(S, dtype=np.uint8),
        landmarks=np.zeros((S, n, 2)),
        collisions=np.zeros(S, dtype=np.int64),
        reached=np.zeros((S, n), dtype=np.uint8),
    )
    for s, op in enumerate(ops):
        if op == RESET:
            gs, oo, os_, done = env.reset()
            rew, rew_n = 0.0, [0.0] * n
        else:
            if seek_rng is not None:
                actions[s] = seek_actions(env.world, seek_rng, greedy_prob)
            gs, oo, os_, rew, rew_n, done = env.step(actions[s])
        out["global_state"][s] = gs
        out["obs_others"][s] = np.array(oo)
        out["obs_self"][s] = np.array(os_)
        out["reward"][s] = rew
        out["reward_n"][s] = np.array(rew_n, dtype=float)
        out["done"][s] = bool(done)
        out["landmarks"][s] = np.array([l.state.p_pos for l in env.world.landmarks])
        out["collisions"][s] = scenario.collisions
        out["reached"][s] = [bool(a.reached) for a in env.world.agents]
    return out


def gen_particle(MAE, scenarios, name, n_agents, cfg, prob_random, max_steps, K, n_episodes,
                 steps_per_episode, rng, seed, bad_action_rate=0.02, seek=False):
    ops = aligned_ops(n_episodes, steps_per_episode)
    S = len(ops)
    actions = rng.integers(0, 5, size=(K, S, n_agents)).astype(np.int32)
    bad = rng.random(size=actions.shape) < bad_action_rate
    actions[bad] = rng.choice([-1, 5, 7, -128, 127], size=int(bad.sum()))
    # the reference's reset_world draws from the global Python / NumPy RNGs
    # (multi-goal_spread.py:75-89); seed them so regeneration is reproducible.
    np.random.seed(seed)
    random.seed(seed)
    traces = [particle_trace(MAE, scenarios, n_agents, cfg, prob_random, max_steps, ops,
                             actions[k], seek_rng=rng if seek else None)
              for k in range(K)]
    fix = {key: np.stack([t[key] for t in traces]) for key in traces[0]}
    fix["ops"] = ops
    fix["actions"] = actions
    fix["n_agents"] = np.int64(n_agents)
    fix["max_steps"] = np.int64(max_steps)
    fix["prob_random"] = np.float64(prob_random)
    for key in ("agents_x", "agents_y", "landmarks_x", "landmarks_y"):
        fix["cfg_" + key] = np.array(cfg[key], dtype=np.float64)
    fix["cfg_initial_std"] = np.float64(cfg["initial_std"])
    fix["global_rng_seed"] = np.int64(seed)  # np.random.seed / random.seed before trace 0
    return name, fix


# --------------------------------------------------------------------------- driver
def generate_all():
    ck, MAE, scenarios = ref_shims.load_reference()
    rng = np.random.default_rng(SEED)
    out = []
    max_steps = ref_shims.reference_config("config.json")["max_steps"]  # 33

    c1 = ref_shims.reference_config("config_checkers_stage1.json")
    c2 = ref_shims.reference_config("config_checkers_stage2.json")

    def ctor_of(c, ms):
        i = c["init"]
        return (i["n_rows"], i["n_columns"], i["n_obs"], i["agents_r"], i["agents_c"],
                c["n_agents"], ms)

    out.append(gen_checkers(ck, "checkers_stage1", ctor_of(c1, max_steps),
                            [[[1, 0]], [[0, 1]]], K=16, n_episodes=2,
                            steps_per_episode=40, rng=rng))
    out.append(gen_checkers(ck, "checkers_stage2", ctor_of(c2, max_steps),
                            [np.eye(2).tolist(), [[0, 1], [1, 0]], [[1, 0], [1, 0]],
                             [[0, 1], [0, 1]]], K=16, n_episodes=2,
                            steps_per_episode=40, rng=rng))
    # tiny board: every cell gets collected well before max_steps -> "all collected" done
    out.append(gen_checkers(ck, "checkers_tiny_3x2", (3, 2, 2, [0, 2], [2, 2], 2, 50),
                            [np.eye(2).tolist(), [[0, 1], [1, 0]]], K=16, n_episodes=3,
                            steps_per_episode=60, rng=rng, bad_action_rate=0.0))
    out.append(gen_checkers(ck, "checkers_tiny_3x2_n1", (3, 2, 2, [0], [2], 1, 50),
                            [[[1, 0]], [[0, 1]]], K=16, n_episodes=3,
                            steps_per_episode=60, rng=rng, bad_action_rate=0.0))
    # the class defaults (checkers.py:5-6): 3x16 board, max_steps 50
    out.append(gen_checkers(ck, "checkers_default_3x16", (3, 16, 2, [0, 2], [16, 16], 2, 50),
                            [np.eye(2).tolist()], K=8, n_episodes=2,
                            steps_per_episode=55, rng=rng))

    p1 = ref_shims.reference_config("config_particle_stage1.json")
    pa = ref_shims.reference_config("config_particle_stage2_antipodal.json")
    pc = ref_shims.reference_config("config_particle_stage2_cross.json")
    pm = ref_shims.reference_config("config_particle_stage2_merge.json")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out.append(gen_particle(MAE, scenarios, "particle_stage1", 1, p1, 0.0, max_steps,
                                K=8, n_episodes=2, steps_per_episode=40, rng=rng, seed=SEED))
        out.append(gen_particle(MAE, scenarios, "particle_antipodal", 4, pa, 0.0, max_steps,
                                K=16, n_episodes=2, steps_per_episode=40, rng=rng,
                                seed=SEED + 1))
        out.append(gen_particle(MAE, scenarios, "particle_antipodal_n3", 3, pa, 0.0,
                                max_steps, K=8, n_episodes=2, steps_per_episode=40, rng=rng,
                                seed=SEED + 2))
        out.append(gen_particle(MAE, scenarios, "particle_cross", 4, pc, 0.0, max_steps,
                                K=16, n_episodes=2, steps_per_episode=40, rng=rng,
                                seed=SEED + 3))
        out.append(gen_particle(MAE, scenarios, "particle_merge", 2, pm, 0.0, max_steps,
                                K=16, n_episodes=2, steps_per_episode=40, rng=rng,
                                seed=SEED + 4))
        # README's stage-2 setting prob_random = 0.2; 1.0 forces the uniform-placement
        # branch (multi-goal_spread.py:77-78,88-89) in every episode
        out.append(gen_particle(MAE, scenarios, "particle_antipodal_random", 4, pa, 1.0,
                                max_steps, K=16, n_episodes=3, steps_per_episode=36, rng=rng,
                                seed=SEED + 5))
        # goal-seeking policies: agents cross at the centre -> many contacts, all-reached done
        out.append(gen_particle(MAE, scenarios, "particle_antipodal_seek", 4, pa, 0.0, 50,
                                K=16, n_episodes=2, steps_per_episode=50, rng=rng,
                                seed=SEED + 6, bad_action_rate=0.0, seek=True))
        out.append(gen_particle(MAE, scenarios, "particle_merge_seek", 2, pm, 0.0, 50,
                                K=16, n_episodes=2, steps_per_episode=50, rng=rng,
                                seed=SEED + 7, bad_action_rate=0.0, seek=True))
        out.append(gen_particle(MAE, scenarios, "particle_stage1_seek", 1, p1, 0.2, 50,
                                K=8, n_episodes=2, steps_per_episode=50, rng=rng,
                                seed=SEED + 8, bad_action_rate=0.0, seek=True))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    fixtures = generate_all()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    bad = 0
    for name, fix in fixtures:
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        if args.check:
            old = np.load(path)
            for key, val in fix.items():
                if not np.array_equal(old[key], val, equal_nan=True):
                    print("MISMATCH %s:%s" % (name, key))
                    bad += 1
            print("checked", name)
        else:
            np.savez_compressed(path, **fix)
            print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
