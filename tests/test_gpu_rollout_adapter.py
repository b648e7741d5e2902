"""GPU: the vectorised episode loop / transition batches (cm3_b200/rollout.py)
against a per-env replay of the reference's trainer loop (alg/train_offpolicy.py:299-368,
alg/train_onpolicy.py:282-350) driven on the oracle, and against a NumPy statement of
process_batch / process_actions (alg/alg_credit.py:406-499, alg/alg_credit_checkers.py:414-477)."""
import numpy as np
import pytest
import torch

import oracle
from cm3_b200 import VecCheckers, VecParticle, presets
from cm3_b200.rollout import TransitionCollector, evaluate_episodes, numpy_process_actions

pytestmark = pytest.mark.gpu


def trainer_loop_checkers(ctor, goal_idx, actions):
    """One env, the reference's loop: caller-side reset on done, actions_prev zeroed at reset."""
    T, n = actions.shape
    env = oracle.OracleCheckers(1, **ctor)
    obs = {k: v[0].copy() for k, v in env.reset(goal_idx[None]).items()}
    prev = np.zeros(n, dtype=np.int64)
    rows = []
    for t in range(T):
        out = {k: v[0].copy() for k, v in env.step(actions[t][None]).items()}
        rows.append(dict(cur=obs, act=actions[t].copy(), prev=prev.copy(), reward=out["reward"],
                         local=out["local_rewards"], done=bool(out["done"])))
        prev = actions[t].astype(np.int64)
        if out["done"]:
            out_next = {k: v[0].copy() for k, v in env.reset(goal_idx[None]).items()}
            prev = np.zeros(n, dtype=np.int64)
        else:
            out_next = out
        rows[-1]["next"] = out_next
        obs = out_next
    return rows


def test_checkers_transitions_match_the_trainer_loop():
    B, T = 48, 80
    ctor = dict(presets.CHECKERS["stage2"], max_steps=presets.MAX_STEPS)
    env = VecCheckers(B, **ctor)
    col = TransitionCollector(env, seed=7)
    col.reset(goals=np.eye(2))
    tr1 = {k: v.clone() for k, v in col.collect(T // 2).items()}           # fused launch, Philox actions
    tr2 = col.collect(T - T // 2, policy=lambda obs: (obs["vec"][:, :, 0].long() + 3) % 5)  # per-step policy
    tr = {k: torch.cat([tr1[k], tr2[k]]).cpu().numpy() for k in tr1}
    assert tr["actions"].shape == (T, B, 2) and tr["obs_self_t_next"].shape == (T, B, 2, 5, 5, 3)
    for b in (0, 1, 17, B - 1):
        rows = trainer_loop_checkers(ctor, np.array([0, 1]), tr["actions"][:, b])
        for t, r in enumerate(rows):
            assert np.array_equal(tr["actions_prev"][t, b], r["prev"]), (b, t)
            assert tr["reward"][t, b] == np.float32(r["reward"]) and bool(tr["done"][t, b]) == r["done"]
            assert np.array_equal(tr["local_rewards"][t, b], r["local"].astype(np.float32))
            for f in ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v"):
                assert np.array_equal(tr[f][t, b], r["cur"][f].astype(np.float32)), (f, b, t)
                assert np.array_equal(tr[f + "_next"][t, b], r["next"][f].astype(np.float32)), (f, b, t)
            assert np.array_equal(tr["goals"][t, b], np.eye(2))
    assert tr["done"].sum() >= 2 * B   # episodes of 33 steps: at least two boundaries per env


def test_process_batch_layout_checkers_and_particle():
    B, T, A = 6, 5, 5
    env = VecCheckers(B, **dict(presets.CHECKERS["stage2"], max_steps=presets.MAX_STEPS))
    col = TransitionCollector(env, l_action=A)
    col.reset(goals=np.eye(2))
    tr = col.collect(T)
    out = col.process_batch(tr)
    assert len(out) == 18 and out[0] == T * B                                   # alg_credit_checkers.py:477
    (n_steps, state_env, state_agents, obs_others, obs_self_t, obs_self_v, ap1, a1, ao1, reward, reward_local,
     state_env_n, state_agents_n, obs_others_n, obs_self_t_n, obs_self_v_n, done, goals) = out
    N = 2
    assert state_env.shape == (n_steps * N, 3, 9, 2) and state_agents.shape == (n_steps, N, 4)
    assert obs_self_t.shape == (n_steps * N, 5, 5, 3) and obs_others.shape == (n_steps * N, 2)
    assert reward.shape == (n_steps,) and reward_local.shape == (n_steps * N,) and done.shape == (n_steps * N,)
    assert goals.shape == (n_steps, N, 2)
    npy = {k: v.cpu().numpy() for k, v in tr.items()}
    # row (t, b, n) <-> index (t*B + b)*N + n; global quantities repeated per agent
    for (t, b, n) in [(0, 0, 0), (2, 3, 1), (T - 1, B - 1, 1)]:
        r = (t * B + b) * N + n
        assert np.array_equal(state_env[r].cpu().numpy(), npy["grid"][t, b])
        assert np.array_equal(state_env_n[r].cpu().numpy(), npy["grid_next"][t, b])
        assert np.array_equal(obs_self_t_n[r].cpu().numpy(), npy["obs_self_t_next"][t, b, n])
        assert np.array_equal(obs_self_v[r].cpu().numpy(), npy["obs_self_v"][t, b, n])
        assert reward_local[r].item() == npy["local_rewards"][t, b, n] and bool(done[r]) == bool(npy["done"][t, b])
    # one-hot formatting against the per-env NumPy statement, env by env
    a1 = a1.cpu().numpy().reshape(T, B, N, A)
    ao1 = ao1.cpu().numpy().reshape(T, B, N, N - 1, A)
    ap1 = ap1.cpu().numpy().reshape(T, B, N, A)
    for b in range(B):
        w1, wo = numpy_process_actions(npy["actions"][:, b], A)
        assert np.array_equal(a1[:, b].reshape(T * N, A), w1)
        assert np.array_equal(ao1[:, b].reshape(T * N, N - 1, A), wo)
        wp, _ = numpy_process_actions(npy["actions_prev"][:, b], A)
        assert np.array_equal(ap1[:, b].reshape(T * N, A), wp)

    penv = VecParticle(B, 4, presets.PARTICLE["antipodal"], max_steps=presets.MAX_STEPS)
    pc = TransitionCollector(penv)
    pc.reset(seed=3)
    ptr = pc.collect(T)
    pout = pc.process_batch(ptr)
    assert len(pout) == 13 and pout[0] == T * B                                 # alg_credit.py:499
    (n_steps, v_global, p_others, v_local, pa1, pao1, p_reward, p_reward_local, v_global_n, p_others_n, v_local_n,
     p_done, p_goals) = pout
    assert v_global.shape == (n_steps, 4, 4) and p_others.shape == (n_steps * 4, 12) and v_local.shape == (n_steps * 4, 4)
    assert p_reward.shape == (n_steps * 4,) and pao1.shape == (n_steps * 4, 3, 5) and p_goals.shape == (n_steps, 4, 2)
    lm = np.stack([presets.PARTICLE["antipodal"]["landmarks_x"], presets.PARTICLE["antipodal"]["landmarks_y"]], axis=1)
    assert np.allclose(p_goals[0].cpu().numpy(), lm)                            # train_onpolicy.py:283-285
    q = {k: v.cpu().numpy() for k, v in ptr.items()}
    r = (3 * B + 2) * 4 + 1
    assert np.array_equal(v_local_n[r].cpu().numpy(), q["obs_self_next"][3, 2, 1])
    assert p_reward[r].item() == q["reward"][3, 2] and p_reward_local[r].item() == q["reward_n"][3, 2, 1]
    # next of step t is current of step t+1 (same buffer, shifted view)
    assert np.array_equal(q["global_state_next"][:-1], q["global_state"][1:])


def test_particle_random_goals_are_tracked_per_step():
    """prob_random = 1: landmarks are redrawn at every in-kernel reset, so goals differ per step."""
    B, T = 32, 70
    env = VecParticle(B, 2, presets.PARTICLE["merge"], prob_random=1.0, max_steps=20)
    col = TransitionCollector(env)
    col.reset(seed=5)
    tr = col.collect(T)
    goals, done = tr["goals"].cpu().numpy(), tr["done"].cpu().numpy().astype(bool)
    assert goals.shape == (T, B, 2, 2)
    for b in range(4):
        for t in range(T - 1):
            changed = not np.array_equal(goals[t + 1, b], goals[t, b])
            assert changed == done[t, b], (b, t)


def test_evaluate_episodes_matches_the_reference_evaluation_loop():
    """evaluate.py:159-203 (test_checkers) and :87-123 (test_particle) replayed per env on the oracle
    with the same deterministic policy: per-agent and global reward sums up to each env's own first
    done, averaged over the envs; episode lengths; action distribution."""
    B = 96
    ctor = dict(presets.CHECKERS["stage2"], max_steps=presets.MAX_STEPS)
    env = VecCheckers(B, **ctor)
    step_no = {"t": 0}

    def policy(obs):  # depends on what the reference's actor sees: own vector, previous actions
        step_no["t"] += 1
        r = obs["vec"][:, :, 0].long() + obs["vec"][:, :, 1].long() + obs["actions_prev"].long() + step_no["t"]
        return r % 5

    rl, rg, info = evaluate_episodes(env, policy, reset_kwargs=dict(goals=np.eye(2)))
    orc = oracle.OracleCheckers(B, **ctor)
    out = orc.reset(np.array([[0, 1]]))
    prev = np.zeros((B, 2), dtype=np.int64)
    alive = np.ones(B, dtype=bool)
    loc, glob, length = np.zeros((B, 2)), np.zeros(B), np.zeros(B, dtype=np.int64)
    dist = np.zeros((2, 5))
    for t in range(1, presets.MAX_STEPS + 1):
        a = (out["vec"][:, :, 0].astype(np.int64) + out["vec"][:, :, 1].astype(np.int64) + prev + t) % 5
        out = orc.step(a.astype(np.int8))
        loc += out["local_rewards"] * alive[:, None]
        glob += out["reward"] * alive
        length += alive
        for n in range(2):
            dist[n] += np.bincount(a[alive, n], minlength=5)
        alive &= ~out["done"].astype(bool)
        prev = a
    assert not alive.any()  # every episode ends by max_steps (checkers.py:246)
    np.testing.assert_allclose(rl.cpu().numpy(), loc.mean(axis=0), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(float(rg), glob.mean(), rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(info["episode_len"].cpu().numpy(), length)
    np.testing.assert_allclose(info["action_distribution"].cpu().numpy(), dist / dist.sum(), rtol=1e-12)

    # particle: goals are the landmark positions (evaluate.py:96-98); policy walks towards them
    N, cfg = 4, presets.PARTICLE["cross"]   # episodes end between step 22 and max_steps, some contacts
    pe = VecParticle(B, N, cfg, max_steps=50, dtype=torch.float64)

    def seek(obs):
        d = obs["goals"] - obs["obs_self"][:, :, 2:4]
        horiz = d[:, :, 0].abs() >= d[:, :, 1].abs()
        return torch.where(horiz, torch.where(d[:, :, 0] > 0, 2, 1), torch.where(d[:, :, 1] > 0, 4, 3))

    pos = np.tile(np.stack([cfg["agents_x"], cfg["agents_y"]], axis=1), (B, 1, 1)).astype(np.float64)
    pos += np.random.default_rng(11).normal(0, 0.03, pos.shape)  # no exactly head-on (coincident) meetings
    lm = np.tile(np.stack([cfg["landmarks_x"], cfg["landmarks_y"]], axis=1), (B, 1, 1)).astype(np.float64)
    rl, rg, info = evaluate_episodes(pe, seek, reset_kwargs=dict(init_pos=pos, init_landmarks=lm))
    po = oracle.OracleParticle(B, N, max_steps=50)
    out = po.reset_to(pos, lm)
    alive = np.ones(B, dtype=bool)
    loc, glob, length = np.zeros((B, N)), np.zeros(B), np.zeros(B, dtype=np.int64)
    for t in range(50):
        d = lm - out["obs_self"][:, :, 2:4]
        horiz = np.abs(d[:, :, 0]) >= np.abs(d[:, :, 1])
        a = np.where(horiz, np.where(d[:, :, 0] > 0, 2, 1), np.where(d[:, :, 1] > 0, 4, 3)).astype(np.int8)
        out = po.step(a)
        loc += out["reward_n"] * alive[:, None]
        glob += out["reward"] * alive
        length += alive
        alive &= ~out["done"].astype(bool)
    assert np.isfinite(loc).all()
    np.testing.assert_array_equal(info["episode_len"].cpu().numpy(), length)
    np.testing.assert_allclose(rl.cpu().numpy(), loc.mean(axis=0), rtol=1e-7)
    np.testing.assert_allclose(float(rg), glob.mean(), rtol=1e-7)
    assert length.min() < 50 < length.max() + 1  # early all-reached endings and max_steps endings


def test_collected_particle_episodes_are_filed_good_or_bad_like_the_trainer_files_them():
    """train_onpolicy.py:329-356 on the device: the collector's transitions go through the episode
    router into the dual buffer; every transition must end up in memory_1 exactly when the episode it
    belongs to ended with scenario.collisions != 0 (the per-step `collisions` output, latched before
    the in-kernel reset).  Checked per env against a scan of done / collisions on the host."""
    from cm3_b200.replay import DeviceDualReplayBuffer, EpisodeRouter
    B, T, blocks = 256, 40, 3
    env = VecParticle(B, 2, presets.PARTICLE["merge"], max_steps=presets.MAX_STEPS)
    col = TransitionCollector(env, seed=5)
    col.reset()
    router, buf = EpisodeRouter(), DeviceDualReplayBuffer(size=10 ** 6)
    dones, colls, uids = [], [], []
    for k in range(blocks):
        tr = col.collect(T)
        uid = (torch.arange(T, device=env.device).unsqueeze(1) + k * T) * B + torch.arange(B, device=env.device)
        fields = {f: tr[f] for f in ("global_state", "obs_others", "obs_self", "actions", "reward", "reward_n",
                                     "global_state_next", "obs_others_next", "obs_self_next", "done", "goals")}
        fields["uid"] = uid
        ready, bad = router.push(fields, tr["done"], tr["collisions"])
        if ready:
            buf.add(ready, bad)
        dones.append(tr["done"].cpu().numpy()); colls.append(tr["collisions"].cpu().numpy()); uids.append(uid.cpu().numpy())
    done, coll, uid = np.concatenate(dones), np.concatenate(colls), np.concatenate(uids)
    want_bad, want_good = set(), set()
    for b in range(B):
        ep = []
        for t in range(T * blocks):
            ep.append(int(uid[t, b]))
            if done[t, b]:
                (want_bad if coll[t, b] != 0 else want_good).update(ep)
                ep = []
    got_bad = set(buf.memory_1.take()["uid"].tolist()) if len(buf.memory_1) else set()
    got_good = set(buf.memory_2.take()["uid"].tolist()) if len(buf.memory_2) else set()
    assert got_bad == want_bad and got_good == want_good
    assert want_bad and want_good                      # merge with random actions produces both kinds
    batch = buf.sample_batch(64)                       # half from each memory, memory_1 part first
    assert batch["uid"].shape[0] == 64 and set(batch["uid"][:32].tolist()) <= want_bad and set(batch["uid"][32:].tolist()) <= want_good
    assert batch["obs_others"].shape == (64, 2, 4)
