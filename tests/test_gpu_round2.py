"""GPU parity tests for what round 2 added, all through the C ABI:

* chained single-step launches (cm3_*_step_chained: per-tile ticket words instead of a grid-wide
  dependency) == plain stream-ordered steps, directly and replayed from a CUDA graph;
* board geometry as data (any board the bitboards address) and 5..8 agents, Checkers and particle,
  bit-exact / 1e-5 against the oracle;
* stage-1 goal redraw on in-kernel reset (train_offpolicy.py:291-296) against the CPU Philox twin;
* masked resets keep the goals of the envs they touch;
* the `goal_idx`, `collisions` and `reached` output fields;
* the double-buffered host rollout (cm3_*_rollout_host) == a sequence of device steps;
* planned rollouts == rollout().
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import golden_util as gu
import oracle
from cm3_b200 import VecCheckers, VecParticle, presets
from cm3_b200.vec_checkers import REF_FIELDS as CK_REF
from cm3_b200.vec_particle import REF_FIELDS as PT_REF

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CK2 = dict(presets.CHECKERS["stage2"], max_steps=presets.MAX_STEPS)
CK1 = dict(presets.CHECKERS["stage1"], max_steps=presets.MAX_STEPS)
ANTI = presets.PARTICLE["antipodal"]
MERGE = presets.PARTICLE["merge"]


def _np(x):
    return x.cpu().numpy() if torch.is_tensor(x) else x


# ---------------------------------------------------------------------------------- chained steps
@pytest.mark.parametrize("B", [4096, 1000])
@pytest.mark.parametrize("game", ["ck2", "ck1", "pa4", "pm2"])
def test_chained_steps_equal_stream_ordered_steps(game, B):
    """T chained launches into a ring of output slots - issued directly and replayed from a CUDA
    graph (the ticket words make the chain replay-safe) - produce exactly what T plain launches
    produce, field by field, and leave the same state."""
    T, ring = 70, 35
    rng = np.random.default_rng(B)

    def make():
        if game.startswith("ck"):
            e = VecCheckers(B, **(CK2 if game == "ck2" else CK1))
            e.reset(goals=np.eye(2) if game == "ck2" else np.array([[0, 1]]))
        else:
            n, cfg = (4, ANTI) if game == "pa4" else (2, MERGE)
            e = VecParticle(B, n, cfg, max_steps=presets.MAX_STEPS)
            e.reset(seed=3)
        return e
    a, b, c = make(), make(), make()
    fields = CK_REF if game.startswith("ck") else PT_REF
    actions = torch.from_numpy(rng.integers(0, 5, size=(T, B, a.N)).astype(np.int8)).to(a.device)
    # reference: plain steps (T = 1 rollouts so that the in-kernel reset matches)
    want = a.alloc_outputs(T, fields=fields)
    for t in range(T):
        a.rollout(1, actions=actions[t:t + 1], auto_reset=True, t0=t, seed=5, out={k: v[t:t + 1] for k, v in want.items()})
    # chained, issued directly
    got = b.alloc_outputs(T, fields=fields)
    for t in range(T):
        b.step_chained(actions[t], {k: v[t] for k, v in got.items()}, seed=5, t0=t, auto_reset=True)
    torch.cuda.synchronize()
    for f in fields:
        assert torch.equal(got[f], want[f]), (game, f)
    for k in a.state:
        assert torch.equal(a.state[k], b.state[k]), k
    # chained, from a CUDA graph of `ring` launches replayed twice
    slots = c.alloc_outputs(ring, fields=fields)
    acts = [actions[t] for t in range(ring)]
    outs = [c._outputs_struct({k: v[t] for k, v in slots.items()}) for t in range(ring)]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        c.step_chained(acts[0], outs[0], seed=5, t0=0, auto_reset=True)   # warm-up outside capture
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    c.load_state_dict(make().state_dict())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for t in range(ring):
            c.step_chained(acts[t], outs[t], seed=5, t0=t, auto_reset=True)
    if game.startswith("ck"):
        # Checkers consumes no randomness on reset: the second replay continues the same trajectory
        # with the first `ring` actions again
        d = make()
        ref = d.alloc_outputs(ring, fields=fields)
        for rep in range(2):
            g.replay()
            for t in range(ring):
                d.rollout(1, actions=actions[t:t + 1], auto_reset=True, out={k: v[t:t + 1] for k, v in ref.items()})
            torch.cuda.synchronize()
            for f in fields:
                assert torch.equal(slots[f], ref[f]), (game, "replay %d" % rep, f)
    else:
        g.replay()
        torch.cuda.synchronize()
        for f in fields:
            assert torch.equal(slots[f], want[f][:ring]), (game, f)


@pytest.mark.parametrize("game,B", [("ck2", 40008), ("ck1", 60000)])
def test_chained_launches_with_two_tiles_per_block(game, B):
    """A chained Checkers launch with more tiles than resident slots gives every block two tiles (checkers.cu:
    launch_ck / step_tile; 40 008 envs x 2 agents = 2501 tiles, 60 000 envs x 1 agent = 3750 tiles, ragged last
    tile): a CUDA graph of such launches, replayed twice, equals plain stream-ordered steps in every field and
    leaves the same state."""
    from cm3_b200.graph import ChainedStepGraph
    ring = 12
    rng = np.random.default_rng(B)

    def make():
        e = VecCheckers(B, **(CK2 if game == "ck2" else CK1))
        e.reset(goals=np.eye(2) if game == "ck2" else np.array([[0, 1]]))
        return e
    c, d = make(), make()
    actions = torch.from_numpy(rng.integers(0, 5, size=(ring, B, c.N)).astype(np.int8)).to(c.device)
    slots = c.alloc_outputs(ring, fields=CK_REF)
    g = ChainedStepGraph(c, actions, slots, seed=5, t0=0, auto_reset=True)
    ref = d.alloc_outputs(ring, fields=CK_REF)
    for rep in range(2):
        g.replay()
        for t in range(ring):
            d.rollout(1, actions=actions[t:t + 1], auto_reset=True, out={k: v[t:t + 1] for k, v in ref.items()})
        torch.cuda.synchronize()
        for f in CK_REF:
            assert torch.equal(slots[f], ref[f]), (game, "replay %d" % rep, f)
    for k in c.state:
        assert torch.equal(c.state[k], d.state[k]), k


def test_chained_needs_sync_words():
    import ctypes as C
    from cm3_b200 import _lib as L
    env = VecCheckers(64, **CK2)
    env.reset(goals=np.eye(2))
    st = L.CheckersState(env._st.remaining, env._st.agents, env._st.meta, None)
    a = torch.zeros(64, 2, dtype=torch.int8, device=env.device)
    rc = env.lib.cm3_checkers_step_chained(env._h, C.byref(st), C.c_void_p(a.data_ptr()), 0, 0, 1, C.byref(env._out_c), env._stream())
    assert rc == -1 and b"sync" in env.lib.cm3_last_error()
    # without sync words the plain entry points still work (stream order alone)
    L.check(env.lib.cm3_checkers_step(env._h, C.byref(st), C.c_void_p(a.data_ptr()), C.byref(env._out_c), env._stream()))
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------------- geometry as data, many agents
def _run_ck(B, ctor, T, seed, dtype=torch.float32, **kw):
    rng = np.random.default_rng(seed)
    N = ctor["n_agents"]
    actions = rng.integers(0, 5, size=(T, B, N)).astype(np.int8)
    actions[rng.random(actions.shape) < 0.02] = -3
    goal_idx = rng.integers(0, 2, size=(B, N)).astype(np.uint8)
    orc = oracle.OracleCheckers(B, nthreads=oracle.max_threads(), **ctor)
    env = VecCheckers(B, dtype=dtype, **ctor, **kw)
    cast = np.float32 if dtype == torch.float32 else np.float64
    ref, out = orc.reset(goal_idx), env.reset(goal_idx=goal_idx)
    obs = ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v", "done")
    for f in obs:
        np.testing.assert_array_equal(_np(out[f]), ref[f] if f == "done" else ref[f].astype(cast), err_msg="reset " + f)
    for t in range(T):
        ref, out = orc.step(actions[t]), env.step(actions[t])
        for f in gu.CHECKERS_FIELDS:
            np.testing.assert_array_equal(_np(out[f]), ref[f] if f == "done" else ref[f].astype(cast), err_msg="t=%d %s" % (t, f))
        np.testing.assert_array_equal(_np(out["goal_idx"]), goal_idx)
    # the fused rollout takes the same kernels through their multi-step path
    env.reset(goal_idx=goal_idx)
    orc.reset(goal_idx)
    ro = env.rollout(T, actions=actions)
    for t in range(T):
        ref = orc.step(actions[t])
        for f in gu.CHECKERS_FIELDS:
            np.testing.assert_array_equal(_np(ro[f][t]), ref[f] if f == "done" else ref[f].astype(cast), err_msg="rollout t=%d %s" % (t, f))


@pytest.mark.parametrize("geom", [(7, 8, 2), (5, 12, 3), (9, 6, 1), (1, 2, 1), (3, 20, 2), (3, 2, 2)])
@pytest.mark.parametrize("N", [1, 2, 3])
def test_any_board_the_bitboards_address(geom, N):
    """Boards that never had a compiled kernel: the geometry travels as data (checkers.py:5-35)."""
    R, Cc, O = geom
    if N == 1 and R < 3:
        pytest.skip("stage-1 start row 2 needs three rows (checkers.py:276)")
    rows = [0, R - 1, R // 2][:N]
    cols = [Cc, Cc, Cc - 1][:N] if R == 1 else [Cc, Cc, Cc][:N]
    if R == 1:
        rows = [0, 0, 0][:N]
        cols = [Cc, Cc - 1, Cc - 2][:N]
        if Cc - (N - 1) < 0:
            pytest.skip("board too small for %d agents" % N)
    ctor = dict(n_rows=R, n_columns=Cc, n_obs=O, agents_r=rows, agents_c=cols, n_agents=N, max_steps=60)
    _run_ck(97, ctor, 64, seed=R * 1000 + Cc * 10 + N)


@pytest.mark.parametrize("N", [5, 6, 7, 8])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_checkers_up_to_eight_agents(N, dtype):
    ctor = dict(n_rows=5, n_columns=8, n_obs=2, agents_r=[0, 4, 1, 3, 2, 0, 4, 2][:N], agents_c=[8, 8, 8, 8, 8, 7, 7, 7][:N],
                n_agents=N, max_steps=50)
    _run_ck(70, ctor, 55, seed=N, dtype=dtype)


def test_checkers_int8_tiles_dynamic_geometry():
    ctor = dict(n_rows=5, n_columns=10, n_obs=3, agents_r=[0, 4], agents_c=[10, 10], n_agents=2, max_steps=50)
    rng = np.random.default_rng(2)
    B, T = 200, 40
    actions = rng.integers(0, 5, size=(T, B, 2)).astype(np.int8)
    a, b = VecCheckers(B, **ctor), VecCheckers(B, tile_dtype=torch.int8, **ctor)
    a.reset(goals=np.eye(2)); b.reset(goals=np.eye(2))
    ra, rb = a.rollout(T, actions=actions, auto_reset=True), b.rollout(T, actions=actions, auto_reset=True)
    for f in gu.CHECKERS_FIELDS:
        want = ra[f].to(torch.int8) if f in ("grid", "obs_self_t") else ra[f]
        assert torch.equal(rb[f], want), f


def test_static_and_dynamic_kernels_agree():
    """The stage-1 / stage-2 boards through the geometry-as-data kernels (CM3_CK_DYNAMIC=1 in a
    child process) against the oracle: same bar as the compiled-in geometry."""
    code = r"""
import sys
sys.path[:0] = [%r, %r, %r]
import numpy as np, torch, oracle, golden_util as gu
from cm3_b200 import VecCheckers, presets
for key, B in (("stage2", 530), ("stage1", 333)):
    ctor = dict(presets.CHECKERS[key], max_steps=33)
    N = ctor["n_agents"]
    rng = np.random.default_rng(B)
    acts = rng.integers(0, 5, size=(50, B, N)).astype(np.int8)
    gi = rng.integers(0, 2, size=(B, N)).astype(np.uint8)
    env, orc = VecCheckers(B, **ctor), oracle.OracleCheckers(B, **ctor)
    env.reset(goal_idx=gi); orc.reset(gi)
    ro = env.rollout(50, actions=acts)
    for t in range(50):
        ref = orc.step(acts[t])
        for f in gu.CHECKERS_FIELDS:
            want = ref[f] if f == "done" else ref[f].astype(np.float32)
            assert np.array_equal(ro[f][t].cpu().numpy(), want), (key, t, f)
print("dynamic ok")
""" % (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"))
    env = dict(os.environ, CM3_CK_DYNAMIC="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "dynamic ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("N", [5, 6, 7, 8])
def test_particle_up_to_eight_agents(N):
    """Teacher-forced float32 (1e-5, north_star's bar) and free-running float64 (1e-9) against the
    oracle for agent counts the reference's make_world accepts (multi-goal_spread.py:36-37)."""
    B, T = 96, 40
    rng = np.random.default_rng(N)
    ang = np.arange(N) * 2 * np.pi / N
    cfg = dict(n_agents=N, agents_x=list(0.8 * np.cos(ang)), agents_y=list(0.8 * np.sin(ang)),
               landmarks_x=list(-0.8 * np.cos(ang)), landmarks_y=list(-0.8 * np.sin(ang)), initial_std=0)
    pos = np.tile(np.stack([cfg["agents_x"], cfg["agents_y"]], axis=1), (B, 1, 1)) + rng.normal(0, 0.05, size=(B, N, 2))
    lm = np.tile(np.stack([cfg["landmarks_x"], cfg["landmarks_y"]], axis=1), (B, 1, 1))
    # goal-seeking actions: everyone crosses the centre, plenty of contacts
    def policy(st):
        d = lm - st["pos"]
        horiz = np.abs(d[..., 0]) > np.abs(d[..., 1])
        a = np.where(horiz, np.where(d[..., 0] > 0, 2, 1), np.where(d[..., 1] > 0, 4, 3))
        return np.where(rng.random(a.shape) < 0.2, rng.integers(0, 5, size=a.shape), a).astype(np.int8)
    po = oracle.OracleParticle(B, N, max_steps=T + 5)
    e64 = VecParticle(B, N, cfg, max_steps=T + 5, dtype=torch.float64)
    e32 = VecParticle(B, N, cfg, max_steps=T + 5)
    po.reset_to(pos, lm)
    e64.reset(init_pos=pos, init_landmarks=lm)
    contacts = 0
    for t in range(T):
        st = po.get_state()
        a = policy(st)
        e32.set_state(pos=st["pos"], vel=st["vel"], landmarks=st["landmarks"], steps=st["steps"],
                      collisions=st["collisions"], reached=st["reached"])
        ref = {k: v.copy() for k, v in po.step(a).items()}
        o32, o64 = e32.step(a), e64.step(a)
        # obs_others holds DIFFERENCES of agent states (multi-goal_spread.py:146-154): its float32 error
        # is 1e-5 of the operands, not of the (possibly cancelling) difference
        scale = float(np.abs(ref["global_state"]).max())
        for f in ("global_state", "obs_others", "obs_self", "reward", "reward_n"):
            atol = 1e-6 + (1e-5 * scale if f == "obs_others" else 0.0)
            np.testing.assert_allclose(_np(o32[f]), ref[f], rtol=1e-5, atol=atol, err_msg="f32 t=%d %s" % (t, f))
            np.testing.assert_allclose(_np(o64[f]), ref[f], rtol=1e-9, atol=1e-11, err_msg="f64 t=%d %s" % (t, f))
        np.testing.assert_array_equal(_np(o64["done"]), ref["done"])
        np.testing.assert_array_equal(_np(o64["collisions"]), po.get_state()["collisions"])
        contacts += int(po.get_state()["collisions"].sum())
    assert contacts > 0


# ---------------------------------------------------------------------------------- goals
def test_masked_reset_keeps_goals_and_first_reset_defaults():
    """a reset without goals must not overwrite the goals of the envs it touches."""
    B = 64
    env = VecCheckers(B, **CK2)
    out = env.reset()                                  # first reset: agent n -> goal n & 1
    np.testing.assert_array_equal(_np(out["goal_idx"]), np.tile([0, 1], (B, 1)))
    gi = np.random.default_rng(0).integers(0, 2, size=(B, 2)).astype(np.uint8)
    env.reset(goal_idx=gi)
    mask = np.zeros(B, dtype=np.uint8); mask[::3] = 1
    env.step(np.full((B, 2), 3, dtype=np.int8))
    out = env.reset(mask=mask)                         # masked, no goals: every env keeps its own
    np.testing.assert_array_equal(_np(out["goal_idx"]), gi)
    np.testing.assert_array_equal(env.unpack_state()["goal_bits"], gi[:, 0] | (gi[:, 1] << 1))
    assert (env.unpack_state()["steps"][mask == 1] == 0).all() and (env.unpack_state()["steps"][mask == 0] == 1).all()


def test_stage1_goal_redraw_on_in_kernel_reset():
    """cfg.random_goal: every in-kernel episode reset draws the new goal from Philox keyed by
    (seed; global env id, step) - the device form of train_offpolicy.py:291-296 - and the start
    row follows the goal (checkers.py:271-276).  Checked against the CPU Philox twin, step by
    step, and for balance."""
    B, T, seed, off = 2048, 200, 77, 1 << 20
    env = VecCheckers(B, random_goal=True, env_id_offset=off, **CK1)
    g0 = np.random.default_rng(1).integers(0, 2, size=(B, 1)).astype(np.uint8)
    env.reset(goal_idx=g0)
    ro = env.rollout(T, seed=seed, t0=10, auto_reset=True, record_actions=True)
    done, goal, vec = _np(ro["done"]), _np(ro["goal_idx"])[:, :, 0], _np(ro["vec"])[:, :, 0]
    cur = g0[:, 0].copy()
    n_new = 0
    for t in range(T):
        for b in np.nonzero(done[t])[0]:
            ctr = np.array([(off + b) & 0xFFFFFFFF, (off + b) >> 32, (10 + t + 1) & 0xFFFFFFFF, 0x60A10000], dtype=np.uint32)
            w = oracle.philox4x32_10(ctr, np.array([seed, 0], dtype=np.uint32))
            cur[b] = w[0] >> 31
            n_new += 1
            assert vec[t, b, 0] == (2 if cur[b] else 0) + 2, "start row follows the goal"
        np.testing.assert_array_equal(goal[t], cur, err_msg="t=%d" % t)
    assert n_new > 5 * B
    drawn = goal[1:][done[:-1] == 1]
    assert abs(drawn.mean() - 0.5) < 0.02
    # and the episodes are the reference's: replay every env on the oracle with the recorded actions
    orc = oracle.OracleCheckers(B, **CK1)
    orc.reset(g0)
    acts = _np(ro["actions"])
    cur = g0.copy()
    for t in range(60):
        ref = orc.step(acts[t])
        np.testing.assert_array_equal(_np(ro["reward"][t]), ref["reward"].astype(np.float32))
        np.testing.assert_array_equal(done[t], ref["done"])
        if done[t].any():
            cur[:, 0] = goal[t]
            orc.reset(cur, mask=done[t])
        np.testing.assert_array_equal(_np(ro["obs_self_t"][t]), orc.out["obs_self_t"].astype(np.float32))


def test_random_goal_is_stage1_only():
    from cm3_b200 import Cm3Error
    with pytest.raises(Cm3Error):
        VecCheckers(8, random_goal=True, **CK2)


# ---------------------------------------------------------------------------------- collisions / reached
def test_collision_count_and_reached_outputs_latch_the_finished_episode():
    """`collisions` [T,B] is scenario.collisions after each step, BEFORE the in-kernel reset clears
    it: at done steps it is the finished episode's total (train_onpolicy.py:356)."""
    B, T = 512, 100
    env = VecParticle(B, 2, MERGE, max_steps=presets.MAX_STEPS)
    env.reset(seed=9)
    st0 = env.get_state_host()
    ro = env.rollout(T, seed=4, auto_reset=True, record_actions=True)
    coll, reached, done, acts = (_np(ro[k]) for k in ("collisions", "reached", "done", "actions"))
    po = oracle.OracleParticle(B, 2, max_steps=presets.MAX_STEPS)
    po.set_state(pos=st0["sv"][:, :, 2:4].astype(np.float64), vel=st0["sv"][:, :, 0:2].astype(np.float64),
                 landmarks=st0["landmarks"].astype(np.float64), steps=st0["steps"], collisions=st0["collisions"].astype(np.int64),
                 reached=np.zeros((B, 2), dtype=np.uint8))
    # first episode of every env against the oracle (float32 vs float64 agree on collision counts
    # except at borderline distances: demand > 99.5 %)
    alive = np.ones(B, dtype=bool)
    agree = total = 0
    for t in range(presets.MAX_STEPS):
        po.step(acts[t])
        c = po.get_state()["collisions"]
        agree += int((coll[t][alive] == c[alive]).sum()); total += int(alive.sum())
        alive &= done[t] == 0
    assert agree / total > 0.995
    # latch: monotone inside an episode, restarts after done
    for t in range(1, T):
        cont = done[t - 1] == 0
        assert (coll[t][cont] >= coll[t - 1][cont]).all()
    fresh = done[:-1] == 1
    assert (coll[1:][fresh] <= 2).all()      # one step of a fresh 2-agent episode: 0 or 2
    assert coll[:-1][fresh].max() > 2        # merge: finished episodes carry their collisions
    # done <=> max_steps or everybody reached
    assert ((reached == 3) <= (done == 1)).all()


def test_explicit_and_in_kernel_reset_draws_are_different_streams():
    """explicit reset k and the in-kernel reset at step k - 1 used to share a Philox counter."""
    B = 256
    env = VecParticle(B, 2, MERGE, max_steps=1)   # every step ends the episode
    env.reset(seed=11, reset_counter=5)
    explicit = env.state["sv"][:, :, 2:4].clone()
    env.rollout(5, seed=11, t0=0, auto_reset=True)   # the reset after step 4 uses counter 5
    auto = env.state["sv"][:, :, 2:4]
    assert not torch.equal(explicit, auto)
    assert (explicit - auto).abs().max() > 1e-3


# ---------------------------------------------------------------------------------- host rollout, plans
@pytest.mark.parametrize("tile", [None, torch.int8])
def test_rollout_host_checkers_equals_device_steps(tile):
    B, T = 640, 12
    rng = np.random.default_rng(8)
    acts = rng.integers(0, 5, size=(T, B, 2)).astype(np.int8)
    a, b = VecCheckers(B, tile_dtype=tile, **CK2), VecCheckers(B, tile_dtype=tile, **CK2)
    a.reset(goals=np.eye(2)); b.reset(goals=np.eye(2))
    host = a.rollout_host(acts, auto_reset=True)
    dev = b.rollout(T, actions=acts, auto_reset=True)
    for f in gu.CHECKERS_FIELDS + ("goal_idx",):
        np.testing.assert_array_equal(host[f], _np(dev[f]), err_msg=f)
    if tile is torch.int8:
        assert host["grid"].dtype == np.int8
    # a second call continues the episodes
    host = {k: v.copy() for k, v in a.rollout_host(acts, auto_reset=True, t0=T).items()}
    dev = b.rollout(T, actions=acts, auto_reset=True, t0=T)
    for f in gu.CHECKERS_FIELDS:
        np.testing.assert_array_equal(host[f], _np(dev[f]), err_msg=f)


def test_rollout_host_particle_equals_device_steps():
    B, T = 1024, 10
    rng = np.random.default_rng(8)
    acts = rng.integers(0, 5, size=(T, B, 4)).astype(np.int8)
    a, b = VecParticle(B, 4, ANTI, max_steps=6), VecParticle(B, 4, ANTI, max_steps=6)
    a.reset(seed=1); b.reset(seed=1)
    host = a.rollout_host(acts, seed=2, auto_reset=True)
    dev = b.alloc_outputs(T)
    for t in range(T):
        b.rollout(1, actions=acts[t:t + 1], seed=2, t0=t, auto_reset=True, out={k: v[t:t + 1] for k, v in dev.items()})
    for f in gu.PARTICLE_FIELDS + ("collisions", "reached"):
        np.testing.assert_array_equal(host[f], _np(dev[f]), err_msg=f)


@pytest.mark.parametrize("game", ["ck2", "pa4"])
def test_planned_rollout_equals_rollout(game):
    B, T = 2048, presets.MAX_STEPS
    rng = np.random.default_rng(3)
    if game == "ck2":
        a, b = VecCheckers(B, **CK2), VecCheckers(B, **CK2)
        a.reset(goals=np.eye(2)); b.reset(goals=np.eye(2))
    else:
        a, b = VecParticle(B, 4, ANTI, max_steps=T), VecParticle(B, 4, ANTI, max_steps=T)
        a.reset(seed=1); b.reset(seed=1)
    acts = torch.from_numpy(rng.integers(0, 5, size=(T, B, a.N)).astype(np.int8)).to(a.device)
    plan = a.plan_rollout(T, actions=acts, auto_reset=True, seed=6)
    for rep in range(2):
        plan(rep * T)
        want = b.rollout(T, actions=acts, auto_reset=True, seed=6, t0=rep * T)
        torch.cuda.synchronize()
        for f in want:
            assert torch.equal(plan.out[f], want[f]), (rep, f)
    short = a.plan_rollout(7, actions=acts[:7].contiguous(), auto_reset=True, seed=6, out={f: v[:7] for f, v in plan.out.items()})
    short(2 * T)
    want = b.rollout(7, actions=acts[:7], auto_reset=True, seed=6, t0=2 * T)
    for f in want:
        assert torch.equal(plan.out[f][:7], want[f]), f
    with pytest.raises(ValueError):
        a.plan_rollout(T, actions=acts.to(torch.int32))


# ---------------------------------------------------------------------------------- N <= 2: kernel variants
_DUMP_CODE = r"""
import sys
sys.path[:0] = [%r]
import numpy as np, torch
from cm3_b200 import VecParticle, presets
dt = getattr(torch, %r)
N = %d
cfg = presets.PARTICLE["merge"] if N == 2 else presets.PARTICLE["stage1"]
out = {}
for B in %r:
    env = VecParticle(B, N, cfg, prob_random=0.3, max_steps=9, dtype=dt, env_id_offset=77)
    o = env.reset(seed=5)
    for k, v in o.items():
        if k not in ("reward", "reward_n", "collisions", "reached"):
            out["reset_%%d_%%s" %% (B, k)] = v.cpu().numpy().copy()
    rng = np.random.default_rng(B)
    for t in range(12):
        a = rng.integers(0, 5, size=(B, N)).astype(np.int8)
        if t %% 3 == 0:
            a[::7] = 9
        o = env.step(a)
        for k, v in o.items():
            out["step_%%d_%%d_%%s" %% (B, t, k)] = v.cpu().numpy().copy()
    ro = env.rollout(40, seed=3, t0=100, auto_reset=True, record_actions=True)
    for k, v in ro.items():
        out["roll_%%d_%%s" %% (B, k)] = v.cpu().numpy().copy()
    acts = rng.integers(0, 5, size=(20, B, N)).astype(np.int8)
    ro = env.rollout(20, actions=acts, seed=3, t0=140, auto_reset=True)
    for k, v in ro.items():
        out["roll2_%%d_%%s" %% (B, k)] = v.cpu().numpy().copy()
    # chained single-step launches into a ring, after plain launches on the same state
    ring = env.alloc_outputs(6)
    dacts = torch.from_numpy(acts[:6]).to(env.device)
    for t in range(6):
        env.step_chained(dacts[t], {k: v[t] for k, v in ring.items()}, seed=3, t0=200 + t, auto_reset=True)
    for k, v in ring.items():
        out["chain_%%d_%%s" %% (B, k)] = v.cpu().numpy().copy()
    for k, v in env.state.items():
        out["state_%%d_%%s" %% (B, k)] = v.cpu().numpy().copy()
np.savez(sys.argv[1], **out)
print("dumped", len(out))
"""


def _variants_agree(tmp_path, dtype, N, batches, env_a, env_b):
    res = {}
    for tag, env in (("a", env_a), ("b", env_b)):
        path = str(tmp_path / (tag + ".npz"))
        r = subprocess.run([sys.executable, "-c", _DUMP_CODE % (ROOT, dtype, N, batches), path], env=dict(os.environ, **env),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "dumped" in r.stdout, r.stdout + r.stderr
        res[tag] = np.load(path)
    assert set(res["a"].files) == set(res["b"].files) and len(res["a"].files) > 50
    for k in res["a"].files:
        a, b = res["a"][k], res["b"][k]
        assert a.dtype == b.dtype and a.shape == b.shape, k
        assert np.array_equal(a, b, equal_nan=True), (k, int((a != b).sum()))


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_pair_kernel_is_bit_identical_to_the_thread_per_env_kernel(dtype, tmp_path):
    """The one-lane-per-agent variant for two-agent envs (particle_pair.cu, CM3_PT_PAIR=1) performs the
    same operations in the same order as the one-thread-per-env kernel: every output of a reset,
    single steps (ragged batch), fused rollouts with in-kernel resets (Philox and given actions),
    chained steps and the final state must be IDENTICAL bit for bit."""
    _variants_agree(tmp_path, dtype, 2, (1000, 4096), dict(CM3_PT_PAIR="1", CM3_PT_DUO="0"), dict(CM3_PT_PAIR="0", CM3_PT_DUO="0"))


@pytest.mark.parametrize("N", [1, 2])
def test_duo_kernel_is_bit_identical_to_the_thread_per_env_kernel(N, tmp_path):
    """The two-envs-per-thread variant for the common launch of one- and two-agent envs
    (particle_duo.cu, CM3_PT_DUO=1; whole 64-env tiles, float, all outputs) against the default
    one-env-per-thread kernel: identical bits, including launches that alternate between the two kernels on one
    state (resets and the ragged batch always take particle.cu) and chained steps."""
    _variants_agree(tmp_path, "float32", N, (4096, 1000, 64), dict(CM3_PT_DUO="1"), dict(CM3_PT_DUO="0"))


@pytest.mark.parametrize("game", ["ck2", "pa4"])
def test_chained_step_graph_replays_continue_the_trajectory(game):
    """cm3_b200.graph.ChainedStepGraph: `ring` chained steps captured once; every replay continues the
    episodes exactly like the fused rollout over the same actions does (constructing the graph leaves
    no trace in the env state)."""
    from cm3_b200.graph import ChainedStepGraph
    B, ring = 2048, 16
    rng = np.random.default_rng(4)
    if game == "ck2":
        a, b = VecCheckers(B, **CK2), VecCheckers(B, **CK2)
        a.reset(goals=np.eye(2)); b.reset(goals=np.eye(2))
        fields = CK_REF
    else:
        a, b = VecParticle(B, 4, ANTI, max_steps=7), VecParticle(B, 4, ANTI, max_steps=7)
        a.reset(seed=2); b.reset(seed=2)
        fields = PT_REF
    acts = torch.from_numpy(rng.integers(0, 5, size=(ring, B, a.N)).astype(np.int8)).to(a.device)
    g = ChainedStepGraph(a, acts, a.alloc_outputs(ring, fields=fields), seed=9, t0=0, auto_reset=True)
    for k in a.state:
        assert torch.equal(a.state[k], b.state[k]), k
    for rep in range(3):
        got = g.replay()
        want = b.rollout(ring, actions=acts, seed=9, t0=0, auto_reset=True)
        torch.cuda.synchronize()
        for f in fields:
            assert torch.equal(got[f], want[f]), (game, rep, f)
    with pytest.raises(ValueError):
        ChainedStepGraph(a, acts[:1].contiguous(), a.alloc_outputs(1, fields=fields))


# ---------------------------------------------------------------------------------- 2-bit packed tiles
def _u2_equal_f32(u2env, out_u2, out_f32, msg):
    from cm3_b200.tiles import unpack_grid_u2, unpack_window_u2
    for f in gu.CHECKERS_FIELDS:
        got, want = out_u2[f], out_f32[f]
        if f == "grid":
            got = unpack_grid_u2(got, u2env.n_columns).to(torch.float32) if torch.is_tensor(got) else unpack_grid_u2(got, u2env.n_columns).astype(np.float32)
        elif f == "obs_self_t":
            got = unpack_window_u2(got, u2env.n_obs).to(torch.float32) if torch.is_tensor(got) else unpack_window_u2(got, u2env.n_obs).astype(np.float32)
        a = got.cpu().numpy() if torch.is_tensor(got) else got
        b = want.cpu().numpy() if torch.is_tensor(want) else want
        np.testing.assert_array_equal(a, b, err_msg="%s %s" % (msg, f))


@pytest.mark.parametrize("B", [16, 1000, 4099])
def test_u2_tiles_are_the_same_numbers(B):
    """tile_dtype="u2": grid / obs_self_t as 2 bits per cell (CM3_TILE_U2) decode to exactly the float
    tiles (and hence to the reference), through step, fused rollout and the host-buffer calls."""
    rng = np.random.default_rng(B)
    T = 40
    actions = rng.integers(0, 5, size=(T, B, 2)).astype(np.int8)
    f32, u2 = VecCheckers(B, **CK2), VecCheckers(B, tile_dtype="u2", **CK2)
    assert u2.out["grid"].shape == (B, 3, 2) and u2.out["obs_self_t"].shape == (B, 2, 5, 1)
    assert u2.bytes_per_env_step() == 24 + 40 + 4 * 23 + 1 + 40 + 2
    a, b = f32.reset(goals=np.eye(2)), u2.reset(goals=np.eye(2))
    for t in range(T):
        _u2_equal_f32(u2, b, a, "t=%d" % t)
        a, b = f32.step(actions[t]), u2.step(actions[t])
    f32.reset(goals=np.eye(2)); u2.reset(goals=np.eye(2))
    ra, rb = f32.rollout(T, actions=actions, auto_reset=True), u2.rollout(T, actions=actions, auto_reset=True)
    _u2_equal_f32(u2, rb, ra, "rollout")
    h, d = u2.step_host(actions[0]), f32.step(actions[0])
    _u2_equal_f32(u2, h, {k: v.cpu().numpy() for k, v in d.items()}, "step_host")
    if B % 32 == 0:
        f32.reset(goals=np.eye(2)); u2.reset(goals=np.eye(2))
        hh = u2.rollout_host(actions[:6], auto_reset=True)
        dd = f32.rollout(6, actions=actions[:6], auto_reset=True)
        _u2_equal_f32(u2, hh, {k: v.cpu().numpy() for k, v in dd.items()}, "rollout_host")


@pytest.mark.parametrize("geom,N", [((5, 12, 3), 2), ((3, 20, 2), 3), ((3, 8, 2), 1), ((7, 8, 1), 4), ((3, 16, 2), 2)])
def test_u2_tiles_on_other_boards(geom, N):
    """Two words per window row (n_obs = 3), three grid words per row (21 columns), one and several agents."""
    R, Cc, O = geom
    ctor = dict(n_rows=R, n_columns=Cc, n_obs=O, agents_r=[0, R - 1, R // 2, 1][:N], agents_c=[Cc, Cc, Cc, Cc - 1][:N],
                n_agents=N, max_steps=25)
    B, T = 200, 40
    rng = np.random.default_rng(R + Cc + N)
    actions = rng.integers(0, 5, size=(T, B, N)).astype(np.int8)
    gi = rng.integers(0, 2, size=(B, N)).astype(np.uint8)
    f32, u2 = VecCheckers(B, **ctor), VecCheckers(B, tile_dtype="u2", **ctor)
    f32.reset(goal_idx=gi); u2.reset(goal_idx=gi)
    ra, rb = f32.rollout(T, actions=actions, auto_reset=True), u2.rollout(T, actions=actions, auto_reset=True)
    _u2_equal_f32(u2, rb, ra, str(geom))
