"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (cm3_b200/sharding.py and
bench.py's multi-rank protocol).  No kernels run here; where env results are needed the oracle
stands in for the stepper, so what is tested is the sharding / gather plumbing: contiguous env-id
ranges, global-env-id keyed Philox streams (results independent of the number of ranks) and the
layout of the gathered rollout."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cm3_b200.sharding import EnvShard, all_gather_rollout, split_envs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_split_envs_properties():
    for total, world in [(8, 1), (10, 4), (65536, 8), (262144, 8), (7, 7), (1000003, 8)]:
        r = split_envs(total, world)
        assert len(r) == world and r[0][0] == 0
        assert sum(n for _, n in r) == total
        for (s0, n0), (s1, _) in zip(r, r[1:]):
            assert s0 + n0 == s1
        assert max(n for _, n in r) - min(n for _, n in r) <= 1
    with pytest.raises(ValueError):
        split_envs(3, 4)
    sh = EnvShard(262144, rank=3, world=8)
    assert (sh.start, sh.count, sh.uniform) == (98304, 32768, True)
    assert sh.owner_of(98304) == (3, 0) and sh.owner_of(262143) == (7, 32767)
    assert not EnvShard(10, rank=0, world=4).uniform


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle
        total, T, N = 64, 5, 2
        shard = EnvShard(total)
        assert (shard.rank, shard.world, shard.count) == (rank, world, total // world)

        # (1) layout of the gathered rollout: element value encodes (t, global env id, trailing index)
        def synth(t0, t1, e0, e1, trail):
            t = torch.arange(t0, t1).view(-1, 1, 1)
            e = torch.arange(e0, e1).view(1, -1, 1)
            k = torch.arange(trail).view(1, 1, -1)
            return (t * 1000000 + e * 100 + k).to(torch.float64)
        local = {"obs": synth(0, T, shard.start, shard.start + shard.count, 7),
                 "done": (synth(0, T, shard.start, shard.start + shard.count, 1)[..., 0] % 3).to(torch.uint8)}
        g = all_gather_rollout(local, shard)
        assert g["obs"].shape == (T, total, 7) and g["done"].shape == (T, total)
        assert torch.equal(g["obs"], synth(0, T, 0, total, 7))
        assert torch.equal(g["done"], (synth(0, T, 0, total, 1)[..., 0] % 3).to(torch.uint8))
        gr = all_gather_rollout(local, shard, time_major=False)
        assert gr["obs"].shape == (world, T, shard.count, 7)
        for r, (s, n) in enumerate(shard.ranges):
            assert torch.equal(gr["obs"][r], synth(0, T, s, s + n, 7))

        # (2) Philox action streams are keyed by the GLOBAL env id: the shards' streams concatenate
        # to the single-process stream
        seed = 12341
        mine = oracle.philox_actions(seed, shard.start, shard.count, N, 0, T)
        ga = all_gather_rollout({"a": torch.from_numpy(mine)}, shard)["a"].numpy()
        assert np.array_equal(ga, oracle.philox_actions(seed, 0, total, N, 0, T))

        # (3) sharded stepping == single-process stepping, env by env (the oracle stands in for the
        # stepper: env instances never interact, env/checkers.py keeps all state per object)
        ctor = dict(n_rows=3, n_columns=8, n_obs=2, agents_r=[0, 2], agents_c=[8, 8], n_agents=2, max_steps=33)
        env = oracle.OracleCheckers(shard.count, **ctor)
        env.reset(np.array([[0, 1]]))
        roll = {k: [] for k in ("obs_self_t", "reward", "done")}
        for t in range(T):
            out = env.step(mine[t])
            for k in roll:
                roll[k].append(torch.from_numpy(out[k].copy()))
        gathered = all_gather_rollout({k: torch.stack(v) for k, v in roll.items()}, shard)
        if rank == 0:
            full = oracle.OracleCheckers(total, **ctor)
            full.reset(np.array([[0, 1]]))
            for t in range(T):
                out = full.step(ga[t])
                for k in roll:
                    assert np.array_equal(gathered[k][t].numpy(), out[k]), (k, t)

        # (4) bench.py's max-over-ranks reduction
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
        with open(os.path.join(tmpdir, "ok_%d" % rank), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharding_and_gather(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok_0", "ok_1"]


def test_reference_arm_under_torchrun_prints_one_line():
    """bench.py --impl reference launched like the driver does for N > 1: rank 0 alone runs and
    prints the JSON line, the other rank exits 0 without work."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "5", "--warmup", "3",
           "--envs", "256"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
