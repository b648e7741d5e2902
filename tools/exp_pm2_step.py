#!/usr/bin/env python
"""Why does the chained per-step PM2 line read differently in bench.py's extras and in tools/ab_r02.py?
Varies K, the ring (steps per graph replay) and the clock sampler on one box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

peak, _ = bench.hbm_peak()
spec = bench.workload_spec("pm2")
orig_ring = bench.step_ring
for ring in (39, 66, 132, 264):
    bench.step_ring = lambda bpe, B, r=ring: r
    for K in (660, 3300):
        for sc in (False, True):
            if K < ring:
                continue
            r = bench.measure_workload(spec, "pm2", 65536, K, 33, 0, 1, "cuda:0", 0, peak, modes=("per_step_chained",), sample_clocks=sc)
            x = r["per_step_chained"]
            print("ring %3d K %4d (steps %4d) sampler %-5s  %.3f us/step frac %.3f" % (ring, K, x["steps"], sc, x["us_per_step"], x["frac"]), flush=True)
