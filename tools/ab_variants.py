#!/usr/bin/env python
"""A/B timing of the fused T-step rollout for one library build (CM3ENV_LIBRARY selects it):
per workload, action source (HBM stream / device Philox) and batch, us per step and the fraction of
the measured HBM peak.  Used to compare experimental builds (cm3_b200.build.build_library(defines=...))
on one GPU box; numbers are CUDA-event timings of 30 launches after 10 warm-up launches.

    CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_x.so python tools/ab_variants.py [workloads] [envs]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    wls = (sys.argv[1] if len(sys.argv) > 1 else "pa4,pa3,pm2,ck2,ck1").split(",")
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    T = bench.MAX_STEPS
    peak, _ = bench.hbm_peak()
    tag = os.path.basename(os.environ.get("CM3ENV_LIBRARY", "libcm3env.so"))
    for wl in wls:
        spec = bench.workload_spec(wl)
        env = bench.make_env(spec, B, "cuda:0")
        out = env.alloc_outputs(T)
        g = torch.Generator(device="cpu").manual_seed(1)
        acts = torch.randint(0, 5, (T, B, env.N), generator=g, dtype=torch.int8).cuda()
        out_b = env.out_bytes_per_env_step()
        bpe = out_b + spec["n"] + (env.bytes_per_env_step() - out_b - spec["n"]) / T
        res = {}
        for src in ("hbm", "philox"):
            def launch(i):
                if src == "hbm":
                    env.rollout(T, actions=acts, auto_reset=True, out=out)
                else:
                    env.rollout(T, seed=7, t0=i * T, auto_reset=True, out=out)
            for i in range(10):
                launch(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(30):
                launch(10 + i)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (30 * T)
            res[src] = (us, bpe * B / (us * 1e-6) / 1e9 / peak)
        print("%-22s %s B=%d  hbm-actions %.3f us/step frac %.3f | philox %.3f us/step frac %.3f" %
              (tag, wl, B, res["hbm"][0], res["hbm"][1], res["philox"][0], res["philox"][1]), flush=True)
        del env, out


if __name__ == "__main__":
    main()
