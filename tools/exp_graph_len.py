#!/usr/bin/env python
"""Chained per-step launches: does the step time depend on the LENGTH of the replayed CUDA graph or on the size of
the output ring?  A ring of R slots is covered by R / G graphs of G chained launches each, replayed in turn.

    python tools/exp_graph_len.py [workloads] [envs]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cm3_b200.graph import ChainedStepGraph  # noqa: E402


class SplitRunner(object):
    def __init__(self, env, R, G, seed):
        self.env, self.R, self.G = env, R, G
        self.out = env.alloc_outputs(R, fields=bench.ref_fields(env))
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.actions = torch.randint(0, 5, (R, env.B, env.N), generator=g, dtype=torch.int8).to(env.device)
        self.graphs = []
        for s in range(0, R, G):
            self.graphs.append(ChainedStepGraph(env, self.actions[s:s + G].contiguous(), {k: v[s:s + G] for k, v in self.out.items()},
                                                seed=bench.SEED, t0=s, auto_reset=True).graph)
        self.launches = 0

    def run(self, k):
        for i in range(k // self.G):
            self.graphs[i % len(self.graphs)].replay()
        self.launches += k


def main():
    wls = (sys.argv[1] if len(sys.argv) > 1 else "pm2,pa4,ck2").split(",")
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    peak, _ = bench.hbm_peak()
    for wl in wls:
        spec = bench.workload_spec(wl)
        for R, Gs in ((40, (4, 8, 20, 40)), (120, (8, 24, 40, 120)), (360, (40, 360))):
            for G in Gs:
                env = bench.make_env(spec, B, "cuda:0")
                bpe = bench.bytes_per_env_step(env)
                if R * bpe * B > 30e9:
                    del env
                    continue
                r = SplitRunner(env, R, G, 7)
                K = max(R, 3300 // R * R)
                ms, _, _ = bench.timed_steps(r, K, R, 1, "cuda:0", 0, sample_clocks=False)
                us = ms * 1e3 / K
                print("%s B=%d ring %3d slots (%5.0f MB) graph %3d launches: %.3f us/step frac %.3f" % (
                    wl, B, R, R * env.out_bytes_per_env_step() * B / 1e6, G, us, bpe * B / (us * 1e-6) / 1e9 / peak), flush=True)
                del r, env
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
