"""GPU parity tests for the Checkers kernels, through the C ABI (VecCheckers -> ctypes ->
libcm3env.so).  Bar: bit-exact against the reference goldens / the oracle (float32 outputs equal
float32(reference float64); float64 outputs equal the reference's float64)."""
import numpy as np
import pytest
import torch

import golden_util as gu
import oracle
from cm3_b200 import VecCheckers, presets

pytestmark = pytest.mark.gpu

CK2 = dict(presets.CHECKERS["stage2"], max_steps=presets.MAX_STEPS)
CK1 = dict(presets.CHECKERS["stage1"], max_steps=presets.MAX_STEPS)


def cmp_fields(out, ref, fields, cast, msg):
    for f in fields:
        got = out[f].cpu().numpy() if torch.is_tensor(out[f]) else out[f]
        want = ref[f]
        if f != "done":
            want = want.astype(cast)
        np.testing.assert_array_equal(got, want, err_msg="%s %s" % (msg, f))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", gu.fixtures("checkers"))
def test_golden_replay(name, dtype):
    fix = gu.load(name)
    ctor = gu.checkers_ctor(fix)
    K = fix["actions"].shape[0]
    env = VecCheckers(K, dtype=dtype, **ctor)
    cast = np.float32 if dtype == torch.float32 else np.float64
    for s, op in enumerate(fix["ops"]):
        ref = {f: fix[f][:, s] for f in gu.CHECKERS_FIELDS}
        if op == gu.RESET:
            out = env.reset(goals=fix["goals"][:, s])
            cmp_fields(out, ref, ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v", "done"),
                       cast, "%s op %d reset" % (name, s))
        else:
            out = env.step(fix["actions"][:, s])
            cmp_fields(out, ref, gu.CHECKERS_FIELDS, cast, "%s op %d step" % (name, s))


def run_oracle_vs_cuda(B, ctor, T, seed, goal_idx=None, dtype=torch.float32):
    rng = np.random.default_rng(seed)
    N = ctor["n_agents"]
    actions = rng.integers(0, 5, size=(T, B, N)).astype(np.int8)
    actions[rng.random(actions.shape) < 0.02] = 7
    if goal_idx is None:
        goal_idx = rng.integers(0, 2, size=(B, N)).astype(np.uint8)
    orc = oracle.OracleCheckers(B, nthreads=oracle.max_threads(), **ctor)
    env = VecCheckers(B, dtype=dtype, **ctor)
    cast = np.float32 if dtype == torch.float32 else np.float64
    ref = orc.reset(goal_idx)
    out = env.reset(goal_idx=goal_idx)
    cmp_fields(out, ref, ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v", "done"), cast, "reset")
    for t in range(T):
        ref = orc.step(actions[t])
        out = env.step(actions[t])
        cmp_fields(out, ref, gu.CHECKERS_FIELDS, cast, "t=%d" % t)
    return env, orc, actions, goal_idx


@pytest.mark.parametrize("B", [1, 15, 16, 17, 250, 4096])
def test_oracle_parity_ck2_ragged_batches(B):
    """Tile tails: B not a multiple of the 16 envs a warp owns takes the non-TMA flush path."""
    run_oracle_vs_cuda(B, CK2, 45, seed=B)


@pytest.mark.parametrize("B", [1, 31, 33, 1000])
def test_oracle_parity_ck1(B):
    run_oracle_vs_cuda(B, CK1, 45, seed=100 + B)


@pytest.mark.parametrize("N", [3, 4])
def test_oracle_parity_more_agents(N):
    ctor = dict(n_rows=3, n_columns=8, n_obs=2, agents_r=[0, 2, 1, 1][:N], agents_c=[8, 8, 8, 7][:N],
                n_agents=N, max_steps=40)
    run_oracle_vs_cuda(77, ctor, 60, seed=N)


@pytest.mark.parametrize("geom", [(3, 16, 2), (5, 8, 2), (3, 8, 1), (3, 8, 3), (3, 4, 2)])
def test_oracle_parity_other_geometries(geom):
    R, Cc, O = geom
    ctor = dict(n_rows=R, n_columns=Cc, n_obs=O, agents_r=[0, R - 1], agents_c=[Cc, Cc], n_agents=2,
                max_steps=50)
    run_oracle_vs_cuda(130, ctor, 70, seed=R * 100 + Cc)


def test_rollout_equals_stepping_and_is_bit_exact():
    """One fused T-step launch == T single-step launches == oracle (headline config, B=8192)."""
    B, T = 8192, presets.MAX_STEPS + 7
    rng = np.random.default_rng(7)
    actions = rng.integers(0, 5, size=(T, B, 2)).astype(np.int8)
    orc = oracle.OracleCheckers(B, nthreads=oracle.max_threads(), **CK2)
    env = VecCheckers(B, **CK2)
    orc.reset(np.array([[0, 1]]))
    env.reset(goals=np.eye(2))
    ro = env.rollout(T, actions=actions)
    for t in range(T):
        ref = orc.step(actions[t])
        cmp_fields({f: ro[f][t] for f in gu.CHECKERS_FIELDS}, ref, gu.CHECKERS_FIELDS, np.float32, "t=%d" % t)
    st = env.unpack_state()
    np.testing.assert_array_equal(st["steps"], orc.steps())


@pytest.mark.parametrize("N,B", [(1, 1024), (3, 264), (4, 512), (1, 1001), (3, 250)])
def test_rollout_equals_stepping_for_every_agent_count(N, B):
    """The multi-step launch (action rows streamed through shared memory) against single-step
    launches for the agent counts the CK2 test above does not cover.  B = 1001 / 250: ragged last
    tile, and for N = 3 rows that are not 4-byte aligned (direct action loads)."""
    ctor = dict(n_rows=3, n_columns=8, n_obs=2, agents_r=[0, 2, 1, 1][:N], agents_c=[8, 8, 8, 7][:N],
                n_agents=N, max_steps=40)
    T = 47  # runs past max_steps: without auto_reset both paths keep stepping the finished episode
    rng = np.random.default_rng(10 * N + B)
    actions = rng.integers(0, 5, size=(T, B, N)).astype(np.int8)
    goals = np.eye(2)[rng.integers(0, 2, size=N)]
    e1, e2 = VecCheckers(B, **ctor), VecCheckers(B, **ctor)
    e1.reset(goals=goals); e2.reset(goals=goals)
    ro = e1.rollout(T, actions=actions)
    for t in range(T):
        out = e2.step(actions[t])
        for f in gu.CHECKERS_FIELDS:
            assert torch.equal(out[f], ro[f][t]), (t, f)
    assert bool(ro["done"][39].all())
    for k in e1.state:
        assert torch.equal(e1.state[k], e2.state[k]), k


def test_philox_actions_match_cpu_twin_and_shard_invariance():
    B, T = 1024, 12
    env = VecCheckers(B, **CK2)
    env.reset(goals=np.eye(2))
    ro = env.rollout(T, actions=None, seed=presets.SEED, t0=5, record_actions=True)
    want = oracle.philox_actions(presets.SEED, 0, B, 2, 5, T)
    np.testing.assert_array_equal(ro["actions"].cpu().numpy(), want)
    # the same envs as a shard with env_id_offset see the same stream -> identical results
    half = VecCheckers(B // 2, env_id_offset=B // 2, **CK2)
    half.reset(goals=np.eye(2))
    rh = half.rollout(T, actions=None, seed=presets.SEED, t0=5)
    for f in gu.CHECKERS_FIELDS:
        assert torch.equal(rh[f], ro[f][:, B // 2:]), f


def test_auto_reset_properties():
    """auto_reset: after a done step the env restarts (observations of the fresh episode are
    returned); the episode that follows is identical to a manual reset + replay."""
    B, T = 512, 3 * presets.MAX_STEPS
    rng = np.random.default_rng(3)
    actions = rng.integers(0, 5, size=(T, B, 2)).astype(np.int8)
    env = VecCheckers(B, **CK2)
    first = {k: v.clone() for k, v in env.reset(goals=np.eye(2)).items()}
    ro = env.rollout(T, actions=actions, auto_reset=True)
    done = ro["done"].cpu().numpy()
    # random walks never clear the 3x8 board in 33 steps -> done exactly every max_steps
    want_done = np.zeros((T, B), dtype=np.uint8)
    want_done[presets.MAX_STEPS - 1::presets.MAX_STEPS] = 1
    np.testing.assert_array_equal(done, want_done)
    for t in np.nonzero(want_done[:, 0])[0]:
        for f in ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v"):
            assert torch.equal(ro[f][t], first[f]), (t, f)
    # second episode == fresh env fed the second episode's actions
    env2 = VecCheckers(B, **CK2)
    env2.reset(goals=np.eye(2))
    ro2 = env2.rollout(presets.MAX_STEPS - 1, actions=actions[presets.MAX_STEPS:2 * presets.MAX_STEPS - 1])
    for f in gu.CHECKERS_FIELDS:
        assert torch.equal(ro2[f], ro[f][presets.MAX_STEPS:2 * presets.MAX_STEPS - 1]), f


def test_masked_reset_and_state_roundtrip():
    B = 100
    rng = np.random.default_rng(11)
    env = VecCheckers(B, **CK2)
    orc = oracle.OracleCheckers(B, **CK2)
    g = np.tile(np.array([[0, 1]], dtype=np.uint8), (B, 1))
    env.reset(goal_idx=g)
    orc.reset(g)
    for t in range(10):
        a = rng.integers(0, 5, size=(B, 2))
        env.step(a)
        orc.step(a)
    mask = (rng.random(B) < 0.4).astype(np.uint8)
    g2 = rng.integers(0, 2, size=(B, 2)).astype(np.uint8)
    out = env.reset(goal_idx=g2, mask=mask)
    ref = orc.reset(g2, mask=mask)
    cmp_fields(out, ref, ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v"), np.float32, "masked reset")
    sd = env.state_dict()
    a = rng.integers(0, 5, size=(B, 2))
    o1 = {k: v.clone() for k, v in env.step(a).items()}
    cmp_fields(o1, orc.step(a), gu.CHECKERS_FIELDS, np.float32, "after masked reset")
    env.load_state_dict(sd)
    o2 = env.step(a)
    for f in gu.CHECKERS_FIELDS:
        assert torch.equal(o1[f], o2[f]), f


def test_host_state_roundtrip_through_the_abi():
    """cm3_checkers_get_state / set_state: a state saved to host arrays and put back reproduces
    the following steps bit for bit; a partial set leaves the other arrays alone."""
    B = 500
    rng = np.random.default_rng(8)
    env = VecCheckers(B, **CK2)
    env.reset(goals=np.eye(2))
    acts = rng.integers(0, 5, size=(6, B, 2)).astype(np.int8)
    for t in range(3):
        env.step(acts[t])
    saved = env.get_state_host()
    assert saved["remaining"].shape == (B,) and saved["agents"].shape == (B, 2)
    first = [{k: v.clone() for k, v in env.step(acts[t]).items()} for t in range(3, 6)]
    env.set_state_host(saved)
    for t in range(3, 6):
        out = env.step(acts[t])
        for f in gu.CHECKERS_FIELDS:
            assert torch.equal(out[f], first[t - 3][f]), (f, t)
    meta = env.state["meta"].clone()
    env.set_state_host({"remaining": saved["remaining"]})
    assert torch.equal(env.state["meta"], meta)
    assert np.array_equal(env.get_state_host()["remaining"], saved["remaining"])


def test_step_host_matches_device_step():
    B = 300
    rng = np.random.default_rng(5)
    env = VecCheckers(B, **CK2)
    orc = oracle.OracleCheckers(B, **CK2)
    env.reset(goals=np.eye(2))
    orc.reset(np.array([[0, 1]]))
    for t in range(5):
        a = rng.integers(0, 5, size=(B, 2)).astype(np.int8)
        got = env.step_host(a)
        cmp_fields(got, orc.step(a), gu.CHECKERS_FIELDS, np.float32, "host t=%d" % t)


@pytest.mark.parametrize("tile", [None, torch.int8])
def test_step_host_packed_is_one_copy_of_the_same_bytes(tile):
    """Batches that are a multiple of 32 envs keep the single-step outputs in one allocation and
    bring them back with cm3_checkers_step_host_packed; the result equals the per-field path and
    the oracle.  A device field outside the block is rejected."""
    import ctypes as C
    from cm3_b200 import _lib as L
    B = 320
    rng = np.random.default_rng(6)
    env = VecCheckers(B, tile_dtype=tile, **CK2)
    assert env.out.block is not None
    orc = oracle.OracleCheckers(B, **CK2)
    env.reset(goals=np.eye(2))
    orc.reset(np.array([[0, 1]]))
    for t in range(4):
        a = rng.integers(0, 5, size=(B, 2)).astype(np.int8)
        sd = env.state_dict()
        packed = {k: v.copy() for k, v in env.step_host(a).items()}
        assert env._host.block is not None
        env.load_state_dict(sd)
        some = env.step_host(a, fields=("grid", "reward", "done"))  # per-field copies
        ref = orc.step(a)
        for f in gu.CHECKERS_FIELDS:
            want = ref[f].astype(packed[f].dtype) if f != "done" else ref[f]
            np.testing.assert_array_equal(packed[f], want, err_msg="packed %s t=%d" % (f, t))
        for f in some:
            np.testing.assert_array_equal(some[f], packed[f], err_msg=f)
    other = torch.zeros(B, dtype=torch.float32, device=env.device)  # a reward buffer outside the block
    oc = L.CheckersOutputs(*[C.c_void_p((other if f == "reward" else env.out[f]).data_ptr()) for f in L.CheckersOutputs.FIELDS])
    rc = env.lib.cm3_checkers_step_host_packed(env._h, C.byref(env._st), C.c_void_p(env._host_actions.data_ptr()),
                                               C.c_void_p(env._actions_dev.data_ptr()), C.byref(oc),
                                               C.c_void_p(env.out.block.data_ptr()), C.c_void_p(env._host.block.data_ptr()),
                                               env.out.block.numel(), env._stream())
    assert rc == -1 and b"outside" in env.lib.cm3_last_error()


def test_headline_size_properties():
    """BASELINE config 2 at full size (65 536 envs x 2 agents, max_steps 33): size-independent
    invariants - cells only ever get collected, counts equal collected cells, rewards sum up,
    done exactly at max_steps or when the board is clear, and a sampled sub-batch is bit-exact
    against the oracle."""
    B, T = 65536, 40
    env = VecCheckers(B, **CK2)
    env.reset(goals=np.eye(2))
    ro = env.rollout(T, actions=None, seed=presets.SEED, record_actions=True)
    grid = ro["grid"]
    collected = (grid == 1).sum(dim=(2, 3, 4))        # [T,B]
    remaining = (grid == -1).sum(dim=(2, 3, 4))
    assert torch.all(collected + remaining == 24)
    assert torch.all(collected[1:] >= collected[:-1])
    counts = ro["vec"][:, :, :, 2:4].sum(dim=(2, 3))
    assert torch.equal(counts, collected.to(counts.dtype))
    assert torch.equal(ro["reward"], ro["local_rewards"].double().sum(dim=2).float())
    vals = torch.unique(ro["local_rewards"])
    assert set(vals.tolist()) <= {0.0, 1.0, -0.5, float(np.float32(-0.1))}
    done = ro["done"].bool()
    steps_done = torch.zeros(T, dtype=torch.bool, device=done.device)
    steps_done[presets.MAX_STEPS - 1] = True
    assert torch.equal(done, steps_done[:, None] | (remaining == 0))
    # sampled envs against the oracle with the recorded actions
    idx = np.random.default_rng(0).choice(B, 256, replace=False)
    acts = ro["actions"].cpu().numpy()[:, idx]
    orc = oracle.OracleCheckers(256, **CK2)
    orc.reset(np.array([[0, 1]]))
    for t in range(T):
        ref = orc.step(acts[t])
        cmp_fields({f: ro[f][t][idx] for f in gu.CHECKERS_FIELDS}, ref, gu.CHECKERS_FIELDS, np.float32, "t=%d" % t)


def test_errors():
    from cm3_b200 import Cm3Error
    with pytest.raises(AssertionError):
        VecCheckers(4, n_rows=4, n_columns=8)
    with pytest.raises(Cm3Error):
        VecCheckers(4, n_rows=7, n_columns=30, n_agents=2, agents_r=[0, 2], agents_c=[30, 30])
    with pytest.raises(Cm3Error):
        VecCheckers(4, 3, 8, 2, [0, 0], [8, 8], 2, 33)  # same start cell
    env = VecCheckers(4, **CK2)
    with pytest.raises(ValueError):
        env.reset(goals=np.zeros((2, 2)))


@pytest.mark.parametrize("B", [16, 1000, 4099])
def test_int8_tiles_are_the_same_numbers(B):
    """tile_dtype=int8: grid / obs_self_t as signed bytes - bit-identical values to the float
    tiles (and hence to the reference), through step, fused rollout and the host-buffer call."""
    rng = np.random.default_rng(B)
    T = 40
    actions = rng.integers(0, 5, size=(T, B, 2)).astype(np.int8)
    f32 = VecCheckers(B, **CK2)
    i8 = VecCheckers(B, tile_dtype=torch.int8, **CK2)
    assert i8.out["grid"].dtype == torch.int8 and i8.out["obs_self_t"].dtype == torch.int8
    assert i8.bytes_per_env_step() == 204 + 4 * 23 + 1 + 40 + 2 and f32.bytes_per_env_step() == 951
    a, b = f32.reset(goals=np.eye(2)), i8.reset(goals=np.eye(2))
    for t in range(T):
        for f in gu.CHECKERS_FIELDS:
            want = a[f].to(torch.int8) if f in ("grid", "obs_self_t") else a[f]
            assert torch.equal(b[f], want), (t, f)
        a, b = f32.step(actions[t]), i8.step(actions[t])
    f32.reset(goals=np.eye(2)); i8.reset(goals=np.eye(2))
    ra = f32.rollout(T, actions=actions, auto_reset=True)
    rb = i8.rollout(T, actions=actions, auto_reset=True)
    for f in gu.CHECKERS_FIELDS:
        want = ra[f].to(torch.int8) if f in ("grid", "obs_self_t") else ra[f]
        assert torch.equal(rb[f], want), f
    h = i8.step_host(actions[0])
    d = f32.step(actions[0])
    assert h["obs_self_t"].dtype == np.int8
    for f in gu.CHECKERS_FIELDS:
        assert np.array_equal(h[f], d[f].cpu().numpy().astype(h[f].dtype)), f
    with pytest.raises(Exception):
        VecCheckers(B, dtype=torch.float64, tile_dtype=torch.int8, **CK2)


def test_max_batch_sampled_parity():
    """The top of the batch sweep (1 048 576 envs, BASELINE configs[4]): a launch that large is
    still exact - 512 envs sampled across the whole id range replay bit for bit in the oracle -
    and the last (ragged) tile of a non-multiple batch is handled."""
    for B in (1 << 20, (1 << 20) - 5):
        T = 6
        env = VecCheckers(B, **CK2)
        env.reset(goals=np.eye(2))
        ro = env.rollout(T, actions=None, seed=3, record_actions=True, auto_reset=True)
        idx = np.unique(np.concatenate([np.random.default_rng(1).choice(B, 500, replace=False),
                                        [0, 1, 15, 16, B - 17, B - 16, B - 2, B - 1]]))
        acts = ro["actions"][:, idx].cpu().numpy()
        orc = oracle.OracleCheckers(len(idx), **CK2)
        orc.reset(np.array([[0, 1]]))
        for t in range(T):
            ref = orc.step(acts[t])
            cmp_fields({f: ro[f][t][idx] for f in gu.CHECKERS_FIELDS}, ref, gu.CHECKERS_FIELDS, np.float32,
                       "B=%d t=%d" % (B, t))
        del env, ro
        torch.cuda.empty_cache()
