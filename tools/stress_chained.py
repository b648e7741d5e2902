#!/usr/bin/env python
"""Stress check of the chained single-step launches under back-to-back issue (CUDA-graph replay): every output
field of every slot and the final state against plain stream-ordered steps, over games, batch sizes and
repetitions.  Prints one line per case; exit code 1 on any mismatch.

    python tools/stress_chained.py [reps]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cm3_b200 import VecCheckers, VecParticle, presets  # noqa: E402
from cm3_b200.graph import ChainedStepGraph  # noqa: E402

CK = {"ck2": dict(presets.CHECKERS["stage2"], max_steps=33), "ck1": dict(presets.CHECKERS["stage1"], max_steps=33)}
PT = {"pa4": (4, presets.PARTICLE["antipodal"]), "pm2": (2, presets.PARTICLE["merge"]), "pa3": (3, presets.PARTICLE["antipodal"])}


def make(game, B):
    if game in CK:
        e = VecCheckers(B, **CK[game])
        e.reset(goals=np.eye(2) if game == "ck2" else np.array([[0, 1]]))
    else:
        n, cfg = PT[game]
        e = VecParticle(B, n, cfg, max_steps=33)
        e.reset(seed=3)
    return e


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    ring, bad_total = 35, 0
    for game in ("ck2", "ck1", "pa4", "pa3", "pm2"):
        for B in (256, 4096, 32768, 65536):
            rng = np.random.default_rng(B)
            c = make(game, B)
            fields = tuple(k for k in c.field_shapes() if k not in ("goal_idx", "collisions", "reached"))
            actions = torch.from_numpy(rng.integers(0, 5, size=(ring, B, c.N)).astype(np.int8)).to(c.device)
            slots = c.alloc_outputs(ring, fields=fields)
            g = ChainedStepGraph(c, actions, slots, seed=5, t0=0, auto_reset=True)
            d = make(game, B)
            ref = d.alloc_outputs(ring, fields=fields)
            for t in range(ring):
                d.rollout(1, actions=actions[t:t + 1], auto_reset=True, t0=t, seed=5, out={k: v[t:t + 1] for k, v in ref.items()})
            torch.cuda.synchronize()
            fresh = make(game, B).state_dict()
            nbad = 0
            for rep in range(reps):
                c.load_state_dict(fresh)
                g.replay()
                torch.cuda.synchronize()
                for f in fields:
                    nbad += int((slots[f] != ref[f]).sum())
                sc, sd = c.state_dict(), d.state_dict()
                for k in sc:
                    if torch.is_tensor(sc[k]):
                        nbad += int((sc[k] != sd[k]).sum())
            print("%s B=%-6d %d graph replays of %d chained steps: %s" % (game, B, reps, ring, "ok" if nbad == 0 else "%d MISMATCHED elements" % nbad), flush=True)
            bad_total += nbad
            del c, d, g, slots, ref
            torch.cuda.empty_cache()
    sys.exit(1 if bad_total else 0)


if __name__ == "__main__":
    main()
