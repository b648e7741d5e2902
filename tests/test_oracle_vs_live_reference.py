"""CPU, build container only: the C oracle against FRESH traces of the live, unmodified reference
(/root/reference through oracle/ref_shims.py) - other seeds, board geometries and agent counts than
the committed fixtures under tests/golden/.  Skipped where the reference tree is absent (the GPU
box); the committed fixtures (test_oracle_golden.py) are what travels.

Importing the reference is confined to this file and oracle/gen_golden.py; nothing on the product
path does it."""
import os
import random
import sys
import warnings

import numpy as np
import pytest

import golden_util as gu
import oracle

sys.path.insert(0, os.path.dirname(os.path.abspath(oracle.__file__)))
import gen_golden  # noqa: E402
import ref_shims  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")


def _is_env_module(name):
    return name in ("multiagent", "env") or name.startswith("multiagent.") or name.startswith("env.")


@pytest.fixture(scope="module")
def ref():
    """The reference's modules, isolated from the drop-in packages of the same names
    (cm3_b200/dropin/{env,multiagent}) that other tests may have imported."""
    saved_modules = {k: sys.modules.pop(k) for k in list(sys.modules) if _is_env_module(k)}
    saved_path = list(sys.path)
    sys.path[:] = [p for p in sys.path if "dropin" not in p]
    try:
        yield ref_shims.load_reference()
    finally:
        for k in list(sys.modules):
            if _is_env_module(k):
                del sys.modules[k]
        sys.modules.update(saved_modules)
        sys.path[:] = saved_path


def replay_checkers(fix):
    ctor = gu.checkers_ctor(fix)
    env = oracle.OracleCheckers(fix["actions"].shape[0], **ctor)
    for s, op in enumerate(fix["ops"]):
        out = env.reset(gu.goal_idx_of(fix["goals"][:, s])) if op == gu.RESET else env.step(fix["actions"][:, s])
        for f in gu.CHECKERS_FIELDS:
            if op == gu.RESET and f in ("reward", "local_rewards"):
                continue
            np.testing.assert_array_equal(out[f], fix[f][:, s], err_msg="op %d %s" % (s, f))


# (n_rows, n_columns, n_obs, agents_r, agents_c, n_agents, max_steps): boards the kernels are
# compiled for, N = 1..4, and windows that reach past the border on every side
CHECKERS_CASES = [
    ((3, 8, 2, [0, 2], [8, 8], 2, 33), [np.eye(2).tolist(), [[0, 1], [1, 0]], [[1, 0], [1, 0]]]),
    ((3, 8, 2, [1], [8], 1, 33), [[[1, 0]], [[0, 1]]]),
    ((5, 8, 2, [0, 2, 4], [8, 8, 8], 3, 40), [[[1, 0], [0, 1], [1, 0]], [[0, 1], [0, 1], [1, 0]]]),
    ((3, 8, 1, [0, 1, 2, 1], [8, 8, 8, 7], 4, 25), [[[1, 0], [0, 1], [1, 0], [0, 1]]]),
    ((3, 8, 3, [0, 2], [8, 8], 2, 33), [np.eye(2).tolist()]),
    ((3, 4, 2, [0, 2], [4, 4], 2, 20), [np.eye(2).tolist(), [[0, 1], [1, 0]]]),
    ((3, 16, 2, [0, 2], [16, 16], 2, 50), [np.eye(2).tolist()]),
]


@pytest.mark.parametrize("case", range(len(CHECKERS_CASES)))
def test_checkers_oracle_matches_fresh_reference_traces(ref, case):
    ck = ref[0]
    ctor, goal_choices = CHECKERS_CASES[case]
    rng = np.random.default_rng(9000 + case)
    _, fix = gen_golden.gen_checkers(ck, "live", ctor, goal_choices, K=6, n_episodes=2,
                                     steps_per_episode=ctor[6] + 4, rng=rng, bad_action_rate=0.05)
    replay_checkers(fix)


PARTICLE_CASES = [  # (config file, n_agents, prob_random, max_steps, goal seeking policy)
    ("config_particle_stage2_antipodal.json", 4, 0.0, 33, False),
    ("config_particle_stage2_antipodal.json", 3, 0.2, 33, True),
    ("config_particle_stage2_cross.json", 4, 0.0, 50, True),
    ("config_particle_stage2_merge.json", 2, 0.0, 33, True),
    ("config_particle_stage2_merge.json", 2, 1.0, 33, False),
    ("config_particle_stage1.json", 1, 1.0, 33, True),
]


@pytest.mark.parametrize("case", range(len(PARTICLE_CASES)))
def test_particle_oracle_matches_fresh_reference_traces(ref, case):
    _, MAE, scenarios = ref
    cfg_name, n, prob_random, max_steps, seek = PARTICLE_CASES[case]
    cfg = ref_shims.reference_config(cfg_name)
    rng = np.random.default_rng(7000 + case)
    state = (np.random.get_state(), random.getstate())  # gen_particle seeds the global RNGs
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _, fix = gen_golden.gen_particle(MAE, scenarios, "live", n, cfg, prob_random, max_steps, K=6,
                                             n_episodes=2, steps_per_episode=max_steps + 3, rng=rng,
                                             seed=4242 + case, bad_action_rate=0.0 if seek else 0.03, seek=seek)
    finally:
        np.random.set_state(state[0])
        random.setstate(state[1])
    K = fix["actions"].shape[0]
    env = oracle.OracleParticle(K, n, max_steps=max_steps)
    steps = np.zeros(K, dtype=np.int32)
    ops = fix["ops"]
    for s in range(len(ops)):
        if ops[s] == gu.RESET:
            # free-running part: inject the reference's reset draw, then run on
            out = env.reset_to(fix["global_state"][:, s, :, 2:4], fix["landmarks"][:, s])
            steps[:] = 0
            continue
        out = env.step(fix["actions"][:, s])
        steps += 1
        for f in ("global_state", "obs_others", "obs_self", "reward", "reward_n"):
            np.testing.assert_allclose(out[f], fix[f][:, s], rtol=1e-9, atol=1e-11, err_msg="op %d %s" % (s, f))
        np.testing.assert_array_equal(out["done"], fix["done"][:, s], err_msg="op %d" % s)
        st = env.get_state()
        np.testing.assert_array_equal(st["collisions"], fix["collisions"][:, s])
        np.testing.assert_array_equal(st["reached"], fix["reached"][:, s])
