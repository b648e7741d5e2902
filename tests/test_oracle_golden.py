"""CPU: the C oracle (oracle/cm3_oracle.c) against every golden trace produced by the
live reference (oracle/gen_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import golden_util as gu
import oracle


@pytest.mark.parametrize("name", gu.fixtures("checkers"))
def test_checkers_oracle_bit_exact(name):
    fix = gu.load(name)
    ctor = gu.checkers_ctor(fix)
    K = fix["actions"].shape[0]
    env = oracle.OracleCheckers(K, **ctor)
    ops = fix["ops"]
    for s, op in enumerate(ops):
        if op == gu.RESET:
            out = env.reset(gu.goal_idx_of(fix["goals"][:, s]))
        else:
            out = env.step(fix["actions"][:, s])
        for f in gu.CHECKERS_FIELDS:
            if op == gu.RESET and f in ("reward", "local_rewards"):
                continue
            np.testing.assert_array_equal(out[f], fix[f][:, s], err_msg="%s op %d %s" % (name, s, f))


@pytest.mark.parametrize("name", gu.fixtures("particle"))
def test_particle_oracle_free_running(name):
    """Free-running float64 restatement; only libm exp/log1p/sqrt may differ from NumPy's,
    so the bound is a few ulp amplified by the stiff contact."""
    fix = gu.load(name)
    N = int(fix["n_agents"])
    K = fix["actions"].shape[0]
    env = oracle.OracleParticle(K, N, max_steps=int(fix["max_steps"]))
    for s, op in enumerate(fix["ops"]):
        if op == gu.RESET:
            out = env.reset_to(fix["global_state"][:, s, :, 2:4], fix["landmarks"][:, s])
        else:
            out = env.step(fix["actions"][:, s])
        for f in gu.PARTICLE_FIELDS:
            if op == gu.RESET and f in ("reward", "reward_n"):
                continue
            if f == "done":
                np.testing.assert_array_equal(out[f], fix[f][:, s], err_msg="%s op %d" % (name, s))
            else:
                np.testing.assert_allclose(out[f], fix[f][:, s], rtol=1e-9, atol=1e-11,
                                           err_msg="%s op %d %s" % (name, s, f))
        st = env.get_state()
        np.testing.assert_array_equal(st["collisions"], fix["collisions"][:, s])
        np.testing.assert_array_equal(st["reached"], fix["reached"][:, s])


@pytest.mark.parametrize("name", gu.fixtures("particle"))
def test_particle_oracle_teacher_forced(name):
    """One step from the reference's own state at every op: tight (no amplification)."""
    fix = gu.load(name)
    N = int(fix["n_agents"])
    K = fix["actions"].shape[0]
    env = oracle.OracleParticle(K, N, max_steps=int(fix["max_steps"]))
    ops = fix["ops"]
    steps = np.zeros(K, dtype=np.int32)
    for s in range(1, len(ops)):
        if ops[s - 1] == gu.RESET:
            steps[:] = 0
        if ops[s] != gu.STEP:
            continue
        gs = fix["global_state"][:, s - 1]
        env.set_state(pos=gs[:, :, 2:4], vel=gs[:, :, 0:2], landmarks=fix["landmarks"][:, s - 1],
                      steps=steps, collisions=fix["collisions"][:, s - 1])
        out = env.step(fix["actions"][:, s])
        steps += 1
        for f in ("global_state", "obs_others", "obs_self", "reward", "reward_n"):
            np.testing.assert_allclose(out[f], fix[f][:, s], rtol=1e-12, atol=1e-14,
                                       err_msg="%s op %d %s" % (name, s, f))
        np.testing.assert_array_equal(out["done"], fix["done"][:, s])


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    z = oracle.philox4x32_10([0, 0, 0, 0], [0, 0])
    assert [hex(int(x)) for x in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)
    assert [hex(int(x)) for x in f] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    p = oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                             [0xa4093822, 0x299f31d0])
    assert [hex(int(x)) for x in p] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_philox_actions_uniform():
    a = oracle.philox_actions(12341, 0, 4096, 2, 0, 8)
    assert a.min() == 0 and a.max() == 4
    counts = np.bincount(a.ravel(), minlength=5) / a.size
    assert np.all(np.abs(counts - 0.2) < 0.01)
    # keyed by global env id: a shard sees the same stream as the full batch
    b = oracle.philox_actions(12341, 1024, 512, 2, 3, 2)
    np.testing.assert_array_equal(b, a[3:5, 1024:1536])
