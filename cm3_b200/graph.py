"""Captured single-step launches: what a per-step inner loop pays per step is the kernel, not Python.

`ChainedStepGraph` records `ring` chained single-step launches (cm3_*_step_chained: tile i of step
k + 1 waits for tile i of step k only) - step t reading `actions[t]` and writing slot t
of a rollout ring - into one CUDA graph.  A replay re-issues the `ring` steps without any host work
per step; the ticket words that chain the launches live with the env state, so replays continue
seamlessly one after another.  Contract: `actions` (all `ring` slices) is complete before replay()
is called - a pre-generated stream, or the output of a policy that ran on the previous ring.

The reference has no counterpart (its loop is `env.step` once per Python iteration,
alg/train_onpolicy.py:302-350); this is the facade-level form of bench.py's per-step mode.
"""
import torch


class ChainedStepGraph(object):
    def __init__(self, env, actions, out_ring, seed=0, t0=0, auto_reset=True):
        ring = int(actions.shape[0])
        if actions.dtype != torch.int8 or tuple(actions.shape) != (ring, env.B, env.N) or not actions.is_contiguous() \
                or actions.device != env.device:
            raise ValueError("actions must be a contiguous int8 device tensor [ring, B, N]")
        if ring < 3:
            raise ValueError("a chained launch must not write the slots of the previous two launches: ring >= 3")
        for k, v in out_ring.items():
            if int(v.shape[0]) != ring or int(v.shape[1]) != env.B:
                raise ValueError("out_ring[%r] must be [ring, B, ...]" % k)
        self.env, self.ring, self.actions, self.out = env, ring, actions, out_ring
        # everything a launch needs is built here, once: per-slot output structs and action slices
        self._slots = [env._outputs_struct({k: v[t] for k, v in out_ring.items()}) for t in range(ring)]
        self._acts = [actions[t] for t in range(ring)]
        self._seed, self._t0, self._auto = seed, t0, auto_reset
        dev = env.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            state = env.state_dict()
            self._launch(0)                      # first launch outside capture (sets kernel attributes)
            env.load_state_dict(state)           # ... and leaves no trace in the env
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            for t in range(ring):
                self._launch(t)

    def _launch(self, t):
        self.env.step_chained(self._acts[t], self._slots[t], seed=self._seed, t0=self._t0 + t, auto_reset=self._auto)

    def replay(self):
        """`ring` env steps; returns the rollout ring (field -> [ring, B, ...], overwritten by the next replay)."""
        self.graph.replay()
        return self.out
