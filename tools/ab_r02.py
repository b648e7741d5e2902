#!/usr/bin/env python
"""A/B timing of one library build / switch setting (round 2): per workload the fused 33-step
launches (action stream from HBM) and the chained per-step launches, measured by bench.py's own
timed_steps().  CM3ENV_LIBRARY selects an experimental build, CM3_BALANCE / CM3_PDL ... the run-time
switches; --cutoff0 sets contact_cutoff = 0 (round 1's exact-zero criterion) for the particle envs.

    [CM3ENV_LIBRARY=...] python tools/ab_r02.py [--workloads pa4,pa3,pm2,ck2,ck1] [--envs 65536] [--cutoff0] [--tag x]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="pa4,pa3,pm2,ck2,ck1")
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--cutoff0", action="store_true")
    ap.add_argument("--tag", default="")
    ap.add_argument("--modes", default="fused,per_step_chained")
    args = ap.parse_args()
    peak, _ = bench.hbm_peak()
    tag = args.tag or os.path.basename(os.environ.get("CM3ENV_LIBRARY", "libcm3env.so"))
    if args.cutoff0:
        orig = bench.make_env

        def make_env(spec, B, device, env_id_offset=0):
            if spec["kind"] == "particle":
                from cm3_b200 import VecParticle
                env = VecParticle(B, spec["n"], spec["cfg"], prob_random=spec["prob_random"], max_steps=bench.MAX_STEPS,
                                  device=device, env_id_offset=env_id_offset, contact_cutoff=0.0)
                env.reset(seed=bench.SEED)
                return env
            return orig(spec, B, device, env_id_offset)
        bench.make_env = make_env
    for wl in args.workloads.split(","):
        r = bench.measure_workload(bench.workload_spec(wl), wl, args.envs, 660, 33, 0, 1, "cuda:0", 0, peak,
                                   modes=tuple(args.modes.split(",")), sample_clocks=False)
        parts = ["%s %.3f us/step frac %.3f" % (k, r[k]["us_per_step"], r[k]["frac"]) for k in ("fused", "per_step_chained", "per_step_stream_ordered") if k in r]
        print("%-28s %s B=%d  %s" % (tag, wl, args.envs, " | ".join(parts)), flush=True)


if __name__ == "__main__":
    main()
