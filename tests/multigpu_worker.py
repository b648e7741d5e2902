"""Worker for tests/test_gpu_multigpu.py, run as `python -m torch.distributed.run --nproc-per-node G`.
Every rank steps its env shard and gathers the rollout (NCCL all-gather and fused peer stores);
rank 0 additionally steps the WHOLE batch on its own GPU and checks that the gathered result is
identical, env by env, bit for bit: results do not depend on how the batch is sharded."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cm3_b200 import VecCheckers, VecParticle, presets  # noqa: E402
from cm3_b200.sharding import EnvShard, RolloutAllGather  # noqa: E402


def make(kind, B, device, offset):
    if kind == "ck2":
        env = VecCheckers(B, device=device, env_id_offset=offset, max_steps=33, **presets.CHECKERS["stage2"])
        env.reset(goals=np.eye(2))
    else:
        env = VecParticle(B, 2, presets.PARTICLE["merge"], prob_random=0.2, max_steps=33, device=device,
                          env_id_offset=offset)
        env.reset(seed=7, reset_counter=0)
    return env


def main():
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    total, T, seed = 4096, 40, 12341
    report = {}
    for kind in ("ck2", "pm2"):
        shard = EnvShard(total)
        refs = None
        if rank == 0:
            full = make(kind, total, shard.device, 0)
            refs = [{k: v.clone() for k, v in full.rollout(T, actions=None, seed=seed, t0=r * T, auto_reset=True).items()}
                    for r in range(2)]
            torch.cuda.synchronize()
        # peer_db: two symmetric buffers, rollout k + 1 enqueued while the stores of rollout k are published
        for mode in ("nccl", "peer", "peer_db"):
            env = make(kind, shard.count, shard.device, shard.start)
            try:
                g = RolloutAllGather(env, T, shard=shard, mode=mode.split("_")[0], double_buffer=mode.endswith("_db"))
            except Exception as e:  # noqa: BLE001
                report["%s_%s" % (kind, mode)] = "unavailable: %r" % (e,)
                if mode == "nccl":
                    raise
                continue
            outs = []
            for r in range(2):
                o = g.rollout(seed=seed, t0=r * T, auto_reset=True, wait=False)
                outs.append(o if mode.endswith("_db") else {k: v.clone() for k, v in o.items()})
            for o in outs:
                g.wait_ready(o) if mode.endswith("_db") else None
            torch.cuda.synchronize()
            dist.barrier()
            ok = True
            if rank == 0:
                for r in range(2):
                    for k, v in refs[r].items():
                        same = torch.equal(outs[r][k].reshape(v.shape), v)
                        ok = ok and same
                        if not same:
                            report["%s_%s_mismatch" % (kind, mode)] = "%s (rollout %d)" % (k, r)
            # every rank must hold the same gathered batch
            chk = torch.stack([o[k].double().sum() for o in outs for k in sorted(o)])
            lst = [torch.empty_like(chk) for _ in range(world)]
            dist.all_gather(lst, chk)
            same_everywhere = all(torch.equal(lst[0], x) for x in lst)
            report["%s_%s" % (kind, mode)] = "ok" if (ok and same_everywhere) else "MISMATCH"
            dist.barrier()
    if rank == 0:
        print("MULTIGPU_REPORT " + json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
