#!/usr/bin/env python
"""Golden vectors for the transition-batch layout: runs the UNMODIFIED
reference trainers' data path in the build container and commits what it produced.

TEST INFRASTRUCTURE ONLY.  Executes, from /root/reference (nothing is edited or copied):
  * the env (env/checkers.py, multiagent/*) stepped the way the trainers step it, with the
    transition tuples assembled field by field as alg/train_offpolicy.py:334-339 (Checkers, 16
    fields) and alg/train_onpolicy.py:329-338 (particle, 11 fields) assemble them;
  * alg/replay_buffer.py Replay_Buffer.add / sample_batch (ring of transitions);
  * alg/alg_credit_checkers.py Alg.process_batch / process_actions (:378-477) and
    alg/alg_credit.py Alg.process_batch / process_actions (:406-499), the same methods of
    alg_baseline*.py and alg_qmix*.py (which repeat different env-level fields per agent), and
    Alg.process_goals / process_global_state (:501-557) - pure NumPy methods, called
    unbound on a stand-in `self` that carries the dimension attributes they read.  The modules import
    TensorFlow 1.x at the top, which is absent here: an inert stub module named `tensorflow` satisfies
    the import (nothing of it is called by these two methods), the same way oracle/ref_shims.py
    stubs `gym`.
One more shim: the trainers build a transition with np.array([...]) over arrays of different shapes,
which NumPy < 1.24 silently turned into an object array and NumPy 2.x rejects; the generator passes
dtype=object explicitly, which is what the old behaviour was.

    python oracle/gen_golden_batch.py            # writes tests/golden/batch_{checkers,particle}.npz
    python oracle/gen_golden_batch.py --check    # regenerate in memory and diff against the files
"""
import argparse
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 12341  # alg/config.json:6


def load_alg_modules():
    if "tensorflow" not in sys.modules:
        # import-only stub, see the docstring: any attribute chain resolves to an inert placeholder
        # (alg/networks.py names tf.nn.relu in default arguments at import time)
        class _Inert(object):
            def __getattr__(self, name):
                return _Inert()

            def __call__(self, *a, **k):
                raise RuntimeError("TensorFlow is stubbed: only the NumPy methods of Alg may run")
        tf = types.ModuleType("tensorflow")
        tf.__getattr__ = lambda name: _Inert()
        sys.modules["tensorflow"] = tf
    alg_dir = os.path.join(ref_shims.REFERENCE_ROOT, "alg")
    if alg_dir not in sys.path:
        sys.path.insert(0, alg_dir)
    import alg_credit
    import alg_credit_checkers
    import replay_buffer
    import replay_buffer_dual
    return alg_credit, alg_credit_checkers, replay_buffer, replay_buffer_dual


def load_alg_variants():
    """The other two algorithm families' modules (same stub): {name: (particle module, checkers module)}."""
    load_alg_modules()
    import alg_baseline
    import alg_baseline_checkers
    import alg_qmix
    import alg_qmix_checkers
    return {"baseline": (alg_baseline, alg_baseline_checkers), "qmix": (alg_qmix, alg_qmix_checkers)}


def variant_outputs(fix, names, batch, alg_mods, attrs, state_key, l_state_one):
    """Adds to `fix` (i) what the baseline / qmix families' process_batch return for the same batch - only the
    arrays that differ from alg_credit*'s (asserted: nothing else does) - and (ii) the returns of
    process_goals / process_global_state (identical code in all six classes, called on alg_credit*'s outputs)."""
    credit = {k: fix["out_" + k] for k in names}
    for alg, mod in alg_mods.items():
        me = stand_in(mod.Alg, **attrs)
        out = mod.Alg.process_batch(me, batch.copy())
        assert len(out) == len(names)
        for k, v in zip(names, out):
            v = np.asarray(v)
            if v.shape != credit[k].shape or not np.array_equal(v, credit[k]):
                fix["out_%s_%s" % (alg, k)] = v
    n = attrs["n_agents"]
    me = types.SimpleNamespace(n_agents=n, l_goal=2, l_state_one_agent=l_state_one, l_state=n * l_state_one)
    any_alg = next(iter(alg_mods.values())).Alg
    n_steps = int(credit["n_steps"])
    gs, go = any_alg.process_goals(me, credit["goals"], n_steps)
    fix["out_goals_self"], fix["out_goals_others"] = np.asarray(gs), np.asarray(go)
    for suffix in ("", "_next"):
        one, others, state = any_alg.process_global_state(me, credit[state_key + suffix], n_steps)
        fix["out_v_global_one_agent" + suffix], fix["out_v_global_others" + suffix], fix["out_state" + suffix] = \
            np.asarray(one), np.asarray(others), np.asarray(state)


def transition(fields):
    """np.array([...]) of the trainers, with the object dtype spelled out (see the docstring)."""
    t = np.empty(len(fields), dtype=object)
    for i, f in enumerate(fields):
        t[i] = f
    return t


def stand_in(alg_cls, **attrs):
    """A `self` for the unbound NumPy methods of Alg: the attributes they read, and process_actions."""
    me = types.SimpleNamespace(**attrs)
    me.process_actions = types.MethodType(alg_cls.process_actions, me)
    return me


def checkers_batch(ck, alg_mod, rb_mod):
    cfg = ref_shims.reference_config("config_checkers_stage2.json")
    main = ref_shims.reference_config("config.json")["main"] if "main" in ref_shims.reference_config("config.json") else {}
    init, n = cfg["init"], cfg["n_agents"]
    max_steps = 33  # alg/config.json:61
    env = ck.Checkers(init["n_rows"], init["n_columns"], init["n_obs"], init["agents_r"], init["agents_c"], n, max_steps)
    np.random.seed(SEED)
    l_action = 5
    buf = rb_mod.Replay_Buffer(size=50)     # small ring: the 66 transitions below wrap it
    raw = []
    for ep in range(2):
        goals = np.eye(n)                                           # train_offpolicy.py:298
        global_state, local_others, local_self_t, local_self_v, done = env.reset(goals)
        actions_prev = np.zeros(n, dtype=int)                       # :300
        while not done:
            actions = np.random.randint(0, l_action, n)             # :315
            next_global_state, next_local_others, next_local_self_t, next_local_self_v, reward, local_rewards, done = env.step(actions)
            fields = [np.array(global_state[0]), np.array(global_state[1]), np.array(local_others), np.array(local_self_t),
                      np.array(local_self_v), actions_prev, actions, reward, local_rewards, np.array(next_global_state[0]),
                      np.array(next_global_state[1]), np.array(next_local_others), np.array(next_local_self_t),
                      np.array(next_local_self_v), done, goals]     # :337
            buf.add(transition(fields))
            raw.append([np.array(f, dtype=np.float64) if not isinstance(f, (bool, np.bool_)) else np.array(f) for f in fields])
            global_state, local_others, local_self_t, local_self_v = next_global_state, next_local_others, next_local_self_t, next_local_self_v
            actions_prev = actions
    # ring contents after 66 adds into 50 slots, in memory order (replay_buffer.py:11-16)
    memory = np.array(buf.memory)
    assert memory.shape == (50, 16)
    batch = buf.sample_batch(10 ** 6)      # len(memory) <= size: the whole memory, in order (:33-34)
    me = stand_in(alg_mod.Alg, experiment="checkers", n_agents=n, l_action=l_action, l_obs_others=2 * (n - 1),
                  rows_obs=5, columns_obs=5, channels_obs=3, l_obs_self=4)
    out = alg_mod.Alg.process_batch(me, batch)
    names = ["n_steps", "state_env", "state_agents", "obs_others", "obs_self_t", "obs_self_v", "actions_prev_1hot",
             "actions_1hot", "actions_others_1hot", "reward", "reward_local", "state_env_next", "state_agents_next",
             "obs_others_next", "obs_self_t_next", "obs_self_v_next", "done", "goals"]
    fix = {"out_" + k: np.asarray(v) for k, v in zip(names, out)}
    attrs = dict(experiment="checkers", n_agents=n, l_action=l_action, l_obs_others=2 * (n - 1),
                 rows_obs=5, columns_obs=5, channels_obs=3, l_obs_self=4)
    variant_outputs(fix, names, batch, {k: v[1] for k, v in load_alg_variants().items()}, attrs, "state_agents", 4)
    in_names = ["grid", "vec", "obs_others", "obs_self_t", "obs_self_v", "actions_prev", "actions", "reward", "local_rewards",
                "grid_next", "vec_next", "obs_others_next", "obs_self_t_next", "obs_self_v_next", "done", "goals"]
    for j, k in enumerate(in_names):      # the 66 transitions in the order they were added
        fix["in_" + k] = np.stack([r[j] for r in raw])
    fix["ring_capacity"] = np.array(50)
    # which added transition sits in each ring slot (derived from the reference's memory itself)
    def same(i, s):   # every one of the 16 fields
        return all(np.array_equal(np.asarray(raw[i][j], dtype=np.float64), np.asarray(memory[s][j], dtype=np.float64)) for j in range(16))
    fix["ring_slot_source"] = np.array([next(i for i in range(len(raw) - 1, -1, -1) if same(i, s)) for s in range(50)])
    return fix


def particle_batch(MultiAgentEnv, scenarios, alg_mod):
    cfg = ref_shims.reference_config("config_particle_stage2_antipodal.json")
    n = cfg["n_agents"]
    scenario = scenarios.load("multi-goal_spread.py").Scenario()
    random.seed(SEED)
    np.random.seed(SEED)
    world = scenario.make_world(n, cfg, 0.2)
    env = MultiAgentEnv(world, scenario.reset_world, scenario.reward, scenario.observation, None, scenario.done, max_steps=33)
    l_action, l_goal = 5, 2
    rows, raw, coll = [], [], []
    bench_rows, bench_pos, bench_lm = [], [], []
    for ep in range(2):
        global_state, local_others, local_self, done = env.reset()        # train_onpolicy.py:282
        goals = np.zeros([n, l_goal])
        for idx in range(n):
            goals[idx] = env.world.landmarks[idx].state.p_pos           # :283-285
        while not done:
            actions = np.random.randint(0, l_action, n)                   # :307
            next_global_state, next_local_others, next_local_self, reward, local_rewards, done = env.step(actions)
            fields = [global_state, np.array(local_others), np.array(local_self), actions, reward, local_rewards,
                      next_global_state, np.array(next_local_others), np.array(next_local_self), done, goals]   # :336
            rows.append(transition(fields))
            raw.append([np.array(f, dtype=np.float64) if not isinstance(f, (bool, np.bool_)) else np.array(f) for f in fields])
            # the info callback the trainers never wire (multi-goal_spread.py:95-111), on the same states
            bench_rows.append(np.array([scenario.benchmark_data(a, env.world) for a in env.world.agents], dtype=np.float64))
            bench_pos.append(np.array([a.state.p_pos for a in env.world.agents]))
            bench_lm.append(np.array([l.state.p_pos for l in env.world.landmarks]))
            global_state, local_others, local_self = next_global_state, next_local_others, next_local_self
        coll.append(scenario.collisions)                                  # :356 reads this per episode
    batch = np.array(rows)
    me = stand_in(alg_mod.Alg, experiment="particle", n_agents=n, l_action=l_action, l_obs_others=4 * (n - 1), l_obs=4)
    out = alg_mod.Alg.process_batch(me, batch)
    names = ["n_steps", "v_global", "obs_others", "v_local", "actions_1hot", "actions_others_1hot", "reward", "reward_local",
             "v_global_next", "obs_others_next", "v_local_next", "done", "goals"]
    fix = {"out_" + k: np.asarray(v) for k, v in zip(names, out)}
    attrs = dict(experiment="particle", n_agents=n, l_action=l_action, l_obs_others=4 * (n - 1), l_obs=4)
    variant_outputs(fix, names, batch, {k: v[0] for k, v in load_alg_variants().items()}, attrs, "v_global", 4)
    in_names = ["global_state", "obs_others", "obs_self", "actions", "reward", "reward_n", "global_state_next",
                "obs_others_next", "obs_self_next", "done", "goals"]
    for j, k in enumerate(in_names):
        fix["in_" + k] = np.stack([r[j] for r in raw])
    fix["episode_collisions"] = np.array(coll)
    fix["benchmark_data"] = np.stack(bench_rows)      # [T, N, 4] = (rew, collisions, min_dists, occupied_landmarks)
    fix["benchmark_pos"], fix["benchmark_landmarks"] = np.stack(bench_pos), np.stack(bench_lm)
    return fix


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    ck, MultiAgentEnv, scenarios = ref_shims.load_reference()
    alg_credit, alg_credit_checkers, replay_buffer, _ = load_alg_modules()
    fixtures = {"batch_checkers": checkers_batch(ck, alg_credit_checkers, replay_buffer),
                "batch_particle": particle_batch(MultiAgentEnv, scenarios, alg_credit)}
    bad = 0
    for name, fix in fixtures.items():
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        if args.check:
            old = np.load(path)
            same = set(old.files) == set(fix) and all(np.array_equal(old[k], fix[k]) for k in fix)
            print("%s: %s" % (name, "identical" if same else "DIFFERENT"))
            bad += 0 if same else 1
        else:
            np.savez_compressed(path, **fix)
            print("wrote %s (%d arrays, %d bytes)" % (path, len(fix), os.path.getsize(path)))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
