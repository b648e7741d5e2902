#!/usr/bin/env python
"""BASELINE.json configs[4]: batch sweep 256 ... 1 048 576 envs (Checkers stage 2 and particle
antipodal N=4) on one GPU - per-step launches (CUDA graph) and the fused 33-step rollout - as
agent-env-steps/s and fraction of the HBM roofline.  One JSON line per point.

    python tools/sweep.py [--out gpurun_out/sweep.jsonl] [--workloads ck2,pa4] [--max-envs 1048576]

Under torchrun the same sweep runs on every rank (weak scaling: the batch is per GPU) and rank 0
reports whole-job numbers from the max-over-ranks time.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def timed(fn, reps, world, device):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0.record()
    for r in range(reps):
        fn(r)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--workloads", default="ck2,pa4")
    ap.add_argument("--min-envs", type=int, default=256)
    ap.add_argument("--max-envs", type=int, default=1048576)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    peak, _ = bench.hbm_peak()
    T = bench.MAX_STEPS
    lines = []
    for wl in args.workloads.split(","):
        spec = bench.workload_spec(wl)
        B = args.min_envs
        while B <= args.max_envs:
            env = bench.make_env(spec, B, device, env_id_offset=rank * B)
            bpe = env.bytes_per_env_step()
            # ring sized to exceed L2 where the batch allows it (capped at 4 GB of outputs)
            ring = int(min(max(T, np.ceil(512e6 / (bpe * B))), max(T, 4e9 // (bpe * B)), 4096))
            runner = bench.StepRunner(env, spec, ring, bench.SEED + rank)
            runner.capture()
            k = max(ring, min(3300, ring * max(1, 3300 // ring)))
            runner.run(ring)
            ms = timed(lambda r: runner.graph.replay(), k // ring, world, device)
            steps = (k // ring) * ring
            v_step = world * B * spec["n"] * steps / (ms * 1e-3)
            gbs_step = bpe * B * steps / (ms * 1e-3) / 1e9
            del runner
            out = env.alloc_outputs(T)
            env.rollout(T, actions=None, seed=bench.SEED, auto_reset=True, out=out)
            reps = 10
            ms2 = timed(lambda r: env.rollout(T, actions=None, seed=bench.SEED, t0=r * T, auto_reset=True, out=out),
                        reps, world, device)
            out_b = env.out_bytes_per_env_step()
            fused_bpe = out_b + (bpe - out_b - spec["n"]) / T
            v_fused = world * B * spec["n"] * T * reps / (ms2 * 1e-3)
            gbs_fused = fused_bpe * B * T * reps / (ms2 * 1e-3) / 1e9
            line = {"workload": wl, "envs_per_gpu": B, "n_gpus": world, "n_agents": spec["n"],
                    "per_step": {"agent_env_steps_per_s": v_step, "us_per_step": ms * 1e3 / steps,
                                 "achieved_gbs": gbs_step, "frac": gbs_step / peak, "ring_slots": ring,
                                 "ring_mb": ring * bpe * B / 1e6},
                    "fused_T33": {"agent_env_steps_per_s": v_fused, "us_per_step": ms2 * 1e3 / (T * reps),
                                  "achieved_gbs": gbs_fused, "frac": gbs_fused / peak,
                                  "buffer_mb": out_b * B * T / 1e6},
                    "peak_gbs": peak}
            if rank == 0:
                print(json.dumps(line), flush=True)
                lines.append(line)
            del env, out
            torch.cuda.empty_cache()
            B *= 4
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
