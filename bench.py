#!/usr/bin/env python
"""bench.py - agent-env-steps/s of the batched env stepper on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload ck2|pa4|pa3|pm2|ck1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over the whole env batch of a GPU: action apply -> move/collect
or dynamics+contact -> reward -> observation assembly for B envs, with in-kernel episode reset,
reading that step's int8 action slice from a pre-generated [33][B][N] stream resident in HBM and
writing EVERY output field of the step to HBM.

  --mode fused (default, the headline `value`): the 33 steps of an episode are one launch of the
      step kernel (state stays in registers between steps); outputs go to a rollout buffer
      [33][B][...] that is larger than L2 (CK2: 2.06 GB) and is overwritten by the next launch.
  --mode step (reported under extra.per_step_launches): one launch per step - what a policy in
      the loop needs - replayed from a CUDA graph, outputs into slot (t mod ring) of a ring > L2.

Timing: CUDA events on the launching stream around exactly K steps, barrier + synchronize on both
sides, max over ranks.

Prints ONE JSON line (rank 0).  Extra keys beyond the driver contract:
  roofline      dominant kernel vs the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the C oracle (port of the reference's step(), float64) on this box's host cores
  e2e           same metric through the host-buffer C-ABI call (H2D actions, D2H every output)
  extra         the fused T-step rollout kernel, the other env family, e2e variants
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "agent-env-steps/sec"
UNIT = "agent-env-steps/s"
B_PER_GPU = 65536
SEED = 12341        # alg/config.json:6
MAX_STEPS = 33      # alg/config.json:61
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
NVLINK_PEER_GBS = 770.0    # B200_PROFILING.md: measured peer copy, per direction per GPU (900 nominal)
FUSED_T = int(os.environ.get("CM3_BENCH_FUSED_T", "0"))  # experiment: steps per fused launch in measure_workload (0 = one episode)
SPIN_CYCLES = int(os.environ.get("CM3_BENCH_SPIN", "4000000"))  # ~2 ms of device spin queued ahead of a timed region (see timed_steps)
NUMA_CPUS = 0  # CPUs this process was bound to (cm3_b200.sharding.bind_host_to_gpu), 0 = not bound
REWARM = os.environ.get("CM3_BENCH_REWARM", "1") != "0"  # re-issue the warm-up steps behind the spin (timed_steps)


def workload_spec(name):
    from cm3_b200 import presets
    if name == "ck2":
        return dict(kind="checkers", ctor=dict(presets.CHECKERS["stage2"], max_steps=MAX_STEPS), n=2,
                    label="Checkers stage2 (config_checkers_stage2.json), %d envs x 2 agents per GPU, max_steps 33, goals eye(2)")
    if name == "ck1":
        return dict(kind="checkers", ctor=dict(presets.CHECKERS["stage1"], max_steps=MAX_STEPS), n=1,
                    label="Checkers stage1, %d envs x 1 agent per GPU, max_steps 33")
    if name in ("pa4", "pa3"):
        n = 4 if name == "pa4" else 3
        return dict(kind="particle", cfg=presets.PARTICLE["antipodal"], n=n, prob_random=0.0,
                    label="Particle antipodal N=%d (config_particle_stage2_antipodal.json), %%d envs per GPU, max_steps 33, prob_random 0" % n)
    if name == "pm2":
        return dict(kind="particle", cfg=presets.PARTICLE["merge"], n=2, prob_random=0.0,
                    label="Particle merge N=2 (config_particle_stage2_merge.json, initial_std 0.05), %d envs per GPU, max_steps 33")
    raise SystemExit("unknown workload %r" % name)


def workload_config(spec, B, l2="n/a (host memory)"):
    """The `config` of the JSON line: the workload, under the same keys for the GPU arm and the reference arm."""
    return {"workload": spec["label"] % B, "envs_per_gpu": B, "n_agents": spec["n"],
            "actions": "uniform integers in 0..4, a [33][B][N] stream generated before the timed region and read by every step",
            "episodes": "max_steps 33; an env whose step returns done starts a fresh episode",
            "l2": l2}


def make_env(spec, B, device, env_id_offset=0):
    from cm3_b200 import VecCheckers, VecParticle
    if spec["kind"] == "checkers":
        env = VecCheckers(B, device=device, env_id_offset=env_id_offset, tile_dtype=spec.get("tile_dtype"), **spec["ctor"])
        env.reset(goals=np.eye(2) if spec["n"] == 2 else np.array([[1, 0]]))
    else:
        env = VecParticle(B, spec["n"], spec["cfg"], prob_random=spec["prob_random"], max_steps=MAX_STEPS,
                          device=device, env_id_offset=env_id_offset)
        env.reset(seed=SEED)
    return env


def make_oracle(spec, B, nthreads):
    import oracle
    if spec["kind"] == "checkers":
        env = oracle.OracleCheckers(B, nthreads=nthreads, **spec["ctor"])
        goal = np.array([[0, 1]]) if spec["n"] == 2 else np.array([[0]])

        def reset():
            env.reset(goal)
    else:
        env = oracle.OracleParticle(B, spec["n"], max_steps=MAX_STEPS, nthreads=nthreads)
        cfg, n = spec["cfg"], spec["n"]
        pos = np.tile(np.stack([cfg["agents_x"][:n], cfg["agents_y"][:n]], axis=1), (B, 1, 1)).astype(np.float64)
        lm = np.tile(np.stack([cfg["landmarks_x"][:n], cfg["landmarks_y"][:n]], axis=1), (B, 1, 1)).astype(np.float64)

        def reset():
            env.reset_to(pos, lm)
    reset()
    return env, reset


def time_oracle(spec, B, nthreads, min_seconds, max_steps_total=None):
    """Runs the CPU port the way the trainers drive the reference (train_onpolicy.py:282-323):
    random actions, caller-side reset every max_steps.  Returns (agent-env-steps/s, steps run)."""
    env, reset = make_oracle(spec, B, nthreads)
    rng = np.random.default_rng(SEED)
    actions = rng.integers(0, 5, size=(MAX_STEPS, B, spec["n"])).astype(np.int32)
    env.step(actions[0])  # warm caches / thread pool
    reset()
    n_steps, t0 = 0, time.perf_counter()
    while True:
        for t in range(MAX_STEPS):
            env.step(actions[t])
        reset()
        n_steps += MAX_STEPS
        el = time.perf_counter() - t0
        if el >= min_seconds or (max_steps_total and n_steps >= max_steps_total):
            break
    return B * spec["n"] * n_steps / el, n_steps, el


class ClockSampler(object):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.ok, self.stop_flag, self.max_mhz = False, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _loop(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.002)

    def sample(self):
        if not self.ok:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:  # noqa: BLE001
            pass

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        self.thread.join()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable: %s" % getattr(self, "err", "no samples")}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key):
    """Per-launch dram bytes of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f).get(kernel_key)
    return None


REF_FIELDS = None  # env facades' REF_FIELDS: the reference's return tuple, nothing else, is written and counted


def ref_fields(env):
    from cm3_b200 import vec_checkers, vec_particle
    return vec_checkers.REF_FIELDS if hasattr(env, "n_rows") else vec_particle.REF_FIELDS


class StepRunner(object):
    """K single-step launches into a rollout ring > L2, replayed from a CUDA graph of `ring` steps.
    chained=True: cm3_*_step_chained (tile i of step k+1 waits for tile i of step k only);
    chained=False: plain stream order + programmatic dependent launch (round 1's per-step mode)."""

    def __init__(self, env, spec, ring, seed, chained=True):
        import torch
        self.torch, self.env, self.ring, self.chained = torch, env, ring, chained
        self.ring_out = env.alloc_outputs(ring, fields=ref_fields(env))
        g = torch.Generator(device="cpu").manual_seed(seed)
        acts = torch.randint(0, 5, (ring, env.B, env.N), generator=g, dtype=torch.int8)
        self.actions = acts.to(env.device)
        self.graph = None
        self.launches = 0

    def launch(self, t):
        # one kernel launch: one step with in-kernel episode reset, outputs -> ring slot
        i = t % self.ring
        if self.chained:
            self.env.step_chained(self.actions[i], {k: v[i] for k, v in self.ring_out.items()}, seed=SEED, t0=t, auto_reset=True)
        else:
            self.env.rollout(1, actions=self.actions[i:i + 1], auto_reset=True, t0=t,
                             out={k: v[i:i + 1] for k, v in self.ring_out.items()})

    def capture(self):
        torch = self.torch
        if self.chained:  # the facade's own captured form (cm3_b200/graph.py)
            from cm3_b200.graph import ChainedStepGraph
            self.graph = ChainedStepGraph(self.env, self.actions, self.ring_out, seed=SEED, auto_reset=True).graph
            return
        self.graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(device=self.env.device)
        s.wait_stream(torch.cuda.current_stream(self.env.device))
        with torch.cuda.stream(s):
            self.launch(0)  # first launch outside capture (sets kernel attributes)
        torch.cuda.current_stream(self.env.device).wait_stream(s)
        torch.cuda.synchronize(self.env.device)
        with torch.cuda.graph(self.graph):
            for t in range(self.ring):
                self.launch(t)

    def run(self, k):
        n_full, rem = divmod(k, self.ring)
        for _ in range(n_full):
            self.graph.replay()
        for t in range(rem):
            self.launch(t)
        self.launches += k


class FusedRunner(object):
    """K steps as K/T launches of the T-step kernel reading the pre-generated action stream.  The
    launches are planned (ctypes arguments built once, VecCheckers.plan_rollout), so the host path of
    the timed region is the foreign call alone."""

    def __init__(self, env, T, seed):
        import torch
        self.env, self.T = env, T
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.actions = torch.randint(0, 5, (T, env.B, env.N), generator=g, dtype=torch.int8).to(env.device)
        self.out = env.alloc_outputs(T, fields=ref_fields(env))
        self.plans = {T: env.plan_rollout(T, actions=self.actions, out=self.out, auto_reset=True, seed=SEED)}
        self.launches = 0
        self.t = 0

    def plan_for(self, n):
        if n not in self.plans:
            self.plans[n] = self.env.plan_rollout(n, actions=self.actions[:n].contiguous(), auto_reset=True, seed=SEED,
                                                  out={f: v[:n] for f, v in self.out.items()})
        return self.plans[n]

    def capture(self):
        pass

    def prepare(self, k):
        if k % self.T:
            self.plan_for(k % self.T)

    def run(self, k):
        n_full, rem = divmod(k, self.T)
        full = self.plans[self.T]
        for _ in range(n_full):
            full(self.t)
            self.t += self.T
        if rem:
            self.plan_for(rem)(self.t)
            self.t += rem
        self.launches += n_full + (1 if rem else 0)


def step_ring(bpe, B):
    """Slots of the rollout ring of the per-step modes (= steps per CUDA-graph replay): larger than L2 - at least 33
    slots and 512 MB of outputs.  (Longer rings were tried so that one replay is ~1 ms of device work even for the
    4 us steps of the small particle envs: the PM2 line got SLOWER with the ring, 3.95 -> 4.4 -> 4.9 us per step at
    39 / 66 / 132+ slots, tools/exp_pm2_step.py; the host is not what limits it.)"""
    return int(min(4096, max(MAX_STEPS, np.ceil(512e6 / (bpe * B)))))


def bytes_per_env_step(env):
    return env.bytes_per_env_step()


def timed_steps(runner, K, W, world, device, local_rank, sample_clocks=True):
    """W warm-up steps, then exactly K timed steps.  Returns (ms max over ranks, clocks, per-rank ms).

    The event pair brackets device time only: a spin kernel (torch.cuda._sleep) is queued ahead
    of the first event, so that every launch of the K steps is already enqueued when the GPU
    reaches it - host enqueue latency and Python overhead are outside the pair.  (Round 1 recorded
    the first event on an idle stream: with the driver's K = 20 the single 190 us launch was timed
    together with the host's call path and read 0.59 of the roofline where ncu showed 0.98.)"""
    import torch
    import torch.distributed as dist

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if hasattr(runner, "prepare"):
        runner.prepare(W)
        runner.prepare(K)
    runner.run(W)
    sampler = ClockSampler(local_rank) if sample_clocks else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.start()
    torch.cuda._sleep(SPIN_CYCLES)
    if REWARM:
        # a few more untimed steps behind the spin: the timed launches then follow a launch of the SAME kernel
        # on the device (same shared-memory carve-out, warm instruction / constant / descriptor caches), as
        # every launch of a long run does, instead of following a one-thread spin kernel
        runner.run(W)
    e0.record()
    runner.run(K)
    e1.record()
    if sampler:
        sampler.sample()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    per_rank = [ms]
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        allt = torch.zeros(world, device=device, dtype=torch.float64)
        dist.all_gather_into_tensor(allt, t)
        per_rank = [float(x) for x in allt.tolist()]
        ms = max(per_rank)
    return ms, clocks, per_rank


class GatherRunner(object):
    """BASELINE.json configs[3]: every launch is one fused T-step rollout of this rank's shard
    (device Philox actions, in-kernel episode reset) whose outputs are delivered to EVERY GPU as
    [T, total_envs, ...] - by ncclAllGather ("nccl") or by the kernel's own peer stores over
    NVLink ("peer", cm3_*_rollout_gather).  One "step" is still one env step of the whole batch."""

    def __init__(self, env, shard, T, mode, overlap=True):
        from cm3_b200.sharding import RolloutAllGather
        self.env, self.T = env, T
        self.g = RolloutAllGather(env, T, shard=shard, mode=mode, fields=ref_fields(env), double_buffer=overlap)
        self.mode = self.g.mode
        self.launches = 0

    def capture(self):
        pass

    def run(self, k):
        for i in range(k // self.T):
            self.g.rollout(seed=SEED, t0=self.launches * self.T, auto_reset=True, time_major=False)
            self.launches += 1


KERNEL_NAMES = {"ck2": "checkers_kernel<StaticGeo<3,8,2>,2,float,float>", "ck1": "checkers_kernel<StaticGeo<3,8,2>,1,float,float>",
                "pa4": "particle_kernel<4,float,0,1>", "pa3": "particle_kernel<3,float,0,1>", "pm2": "particle_kernel<2,float,0,1>"}


def setup_dist():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from cm3_b200.sharding import bind_host_to_gpu
    global NUMA_CPUS
    NUMA_CPUS = bind_host_to_gpu(local_rank)   # pinned host buffers of the e2e path on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: whatever NCCL prints while the communicator comes up
        # (the "NCCL version ..." line under NCCL_DEBUG=VERSION) is sent to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            torch.cuda.set_device(local_rank)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.cuda.set_device(local_rank)
    return rank, world, local_rank, "cuda:%d" % local_rank


def fused_bytes(env, spec, T, with_actions=True):
    """Algorithmic bytes per env-step of a T-step launch: outputs (+ the action stream) + state / T."""
    out_b = env.out_bytes_per_env_step()
    state_b = env.bytes_per_env_step() - out_b - spec["n"]
    return out_b + (spec["n"] if with_actions else 0) + state_b / T


def measure_workload(spec, wl, B, K, W, rank, world, device, local_rank, peak, modes=("fused", "step"), sample_clocks=True):
    """One workload on this rank's GPU (all ranks call it together): fused 33-step launches with the
    action stream from HBM, then one chained launch per step from a CUDA graph.  Returns a dict of
    whole-job numbers (max over ranks)."""
    import torch
    T = FUSED_T or MAX_STEPS
    env = make_env(spec, B, device, env_id_offset=rank * B)
    bpe = bytes_per_env_step(env)
    res = {"workload": spec["label"] % B, "envs_per_gpu": B, "n_agents": spec["n"], "n_gpus": world,
           "kernel": KERNEL_NAMES.get(wl)}
    if "fused" in modes:
        runner = FusedRunner(env, T, SEED + rank)
        Kf = max(T, K // T * T)
        ms, clocks, per_rank = timed_steps(runner, Kf, max(W, T), world, device, local_rank, sample_clocks)
        fb = fused_bytes(env, spec, T)
        ach = fb * B * Kf / (ms * 1e-3) / 1e9
        res["fused"] = {"value": world * B * spec["n"] * Kf / (ms * 1e-3), "unit": UNIT, "steps": Kf, "launches": Kf // T,
                        "us_per_step": ms * 1e3 / Kf, "achieved_gbs": ach, "frac": ach / peak,
                        "algorithmic_bytes_per_env_step": fb, "per_rank_ms": per_rank, "clocks": clocks}
        del runner
    ring = step_ring(bpe, B)
    for name, chained in (("per_step_chained", True), ("per_step_stream_ordered", False)):
        if name in modes or "step" in modes:
            runner = StepRunner(env, spec, ring, SEED + rank, chained=chained)
            runner.capture()
            Ks = max(ring, K // ring * ring)
            ms, clocks, per_rank = timed_steps(runner, Ks, ring, world, device, local_rank, sample_clocks)
            ach = bpe * B * Ks / (ms * 1e-3) / 1e9
            res[name] = {"value": world * B * spec["n"] * Ks / (ms * 1e-3), "unit": UNIT, "steps": Ks, "launches": Ks,
                         "us_per_step": ms * 1e3 / Ks, "achieved_gbs": ach, "frac": ach / peak,
                         "algorithmic_bytes_per_env_step": bpe, "rollout_ring_slots": ring, "per_rank_ms": per_rank,
                         "clocks": clocks}
            del runner
    del env
    torch.cuda.empty_cache()
    return res


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank, world, local_rank, device = setup_dist()
    spec = workload_spec(args.workload)
    B, K, W = args.envs, args.steps, args.warmup
    env = make_env(spec, B, device, env_id_offset=rank * B)
    bpe = bytes_per_env_step(env)
    ring = step_ring(bpe, B)
    gather_mode = None
    T = MAX_STEPS
    W = max(W, 3)
    out_b = env.out_bytes_per_env_step()
    state_b = bpe - out_b - spec["n"]
    fused_bpe = fused_bytes(env, spec, T)
    launch_cfg = None
    if args.gather != "none":
        from cm3_b200.sharding import EnvShard
        K = max(T, K // T * T)
        W = (W + T - 1) // T * T
        runner = GatherRunner(env, EnvShard(world * B, rank=rank, world=world, local_rank=local_rank), T, args.gather,
                              overlap=not args.no_overlap)
        gather_mode = runner.mode
        mode = "gather"
    elif args.mode == "fused":
        runner = FusedRunner(env, T, SEED + rank)
        mode = "fused"
    else:
        runner = StepRunner(env, spec, ring, SEED + rank, chained=args.mode == "step")
        mode = args.mode
    runner.capture()
    l0 = runner.launches
    ms, clocks, per_rank = timed_steps(runner, K, W, world, device, local_rank)
    # launches inside the timed region: everything the runner issued minus the warm-up's share
    timed_launches = {"fused": (K + T - 1) // T, "gather": K // T}.get(mode, K)
    value = world * B * spec["n"] * K / (ms * 1e-3)
    peak, peak_src = hbm_peak()
    step_us = ms * 1e3 / K
    if mode in ("step", "step_ordered"):
        eff_bpe, launches, launch_us = bpe, K, step_us
        launch_cfg = {"launch": "1 kernel launch per step (%s), replayed from a CUDA graph of %d steps" % (
                          "cm3_*_step_chained: per-tile ticket chaining" if mode == "step" else "stream order + programmatic dependent launch", ring),
                      "rollout_ring_slots": ring,
                      "l2": "outputs go to a %d-slot rollout ring of %.0f MB (> 126 MB L2); no explicit flush" % (ring, ring * bpe * B / 1e6)}
    else:
        # the bytes of the launches actually issued: whole 33-step launches + one shorter remainder
        n_full, rem = divmod(K, T)
        launches = timed_launches
        eff_bpe = (out_b + spec["n"]) + state_b * launches / K
        launch_us = ms * 1e3 / launches
        launch_cfg = {"launch": "1 kernel launch per %d steps (one episode): state stays in registers between the steps of a launch; K = %d steps = %d launch(es) of %d%s" % (
                          T, K, n_full, T, (" + 1 of %d" % rem) if rem else ""),
                      "l2": "every step writes all its outputs to a [%d][B] rollout buffer of %.0f MB (> 126 MB L2), overwritten by the next launch; no explicit flush" % (T, out_b * B * T / 1e6)}
    achieved = eff_bpe * B / (step_us * 1e-6) / 1e9
    kernel_key = "%s_%s" % (args.workload, "step" if mode.startswith("step") else "rollout")
    kname = KERNEL_NAMES[args.workload]

    config = workload_config(spec, B, launch_cfg.pop("l2"))
    launch_config = dict({"mode": mode, "actions": "int8 in HBM", "auto_reset": "in-kernel",
                          "parallelism": "env batch sharded over %d GPU(s), no per-step collective" % world,
                          "timing": "CUDA events on the launching stream behind a %d-cycle device spin (all launches enqueued before the first event fires); barrier + synchronize both sides; max over ranks" % SPIN_CYCLES,
                          "host_affinity": ("each rank bound to the %d CPUs next to its GPU (NVML affinity mask)" % NUMA_CPUS) if NUMA_CPUS else "not bound"},
                         **launch_cfg)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 bitboard state, f32 outputs" if spec["kind"] == "checkers" else "f32",
        "data": "synthetic",
        "config": config,
        "launch_config": launch_config,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(kernel_key),
                     "peak_source": peak_src, "algorithmic_bytes_per_env_step": eff_bpe,
                     "bytes_per_launch": eff_bpe * B * (K / launches), "launch_us": launch_us, "kernel": kname},
        "gpu_launches": launches,
        "per_rank_ms": per_rank,
        "clocks": clocks,
    }
    if gather_mode is not None:
        sent = out_b * B * (world - 1)  # bytes this GPU delivers to its peers per env step
        out["gpu_launches"] = K // T
        out["config"]["actions"] = "Philox4x32-10 on the device, keyed by (seed, global env id, step)"
        out["launch_config"].update({"launch": "1 launch per %d fused env steps, device Philox actions" % T,
                              "actions": "Philox4x32-10 on the device, keyed by (seed, global env id, step)",
                              "rollout_all_gather": gather_mode, "overlap": "two symmetric buffers: rollout k+1 runs while the stores of rollout k drain" if not args.no_overlap else "none",
                              })
        out["config"]["l2"] = "gathered rollout buffer [%d][%d envs] = %.0f MB per GPU (> 126 MB L2 when >= 2 GPUs at the default size)" % (T, world * B, out_b * world * B * T / 1e6)
        out["roofline"] = gather_roofline(out_b, state_b, T, B, world, step_us, peak, peak_src, kname, gather_mode)
    else:
        out["e2e"] = measure_e2e(spec, args, world, rank, device)   # all ranks take part
        if not args.no_extras:
            extra = measure_extras(env, spec, args, peak, mode, world, rank, device, local_rank, ring)
            if rank == 0:
                out["extra"] = extra
                if world == 1:
                    out["cpu_baseline"] = cpu_baseline(spec, args)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def gather_roofline(out_b, state_b, T, B, world, step_us, peak, peak_src, kname, gather_mode):
    sent = out_b * B * (world - 1)
    hbm_written = (out_b + state_b / T) * world * B / (step_us * 1e-6) / 1e9  # the whole gathered batch lands in this GPU's HBM
    nv = sent / (step_us * 1e-6) / 1e9 if world > 1 else 0.0
    return {"bound": "nvlink" if world > 1 else "hbm", "achieved": nv if world > 1 else hbm_written,
            "peak": NVLINK_PEER_GBS if world > 1 else peak, "unit": "GB/s",
            "frac": (nv / NVLINK_PEER_GBS) if world > 1 else hbm_written / peak, "traffic": None,
            "peak_source": "measured peer-copy reference, B200_PROFILING.md (770 GB/s per direction per GPU; 900 nominal)" if world > 1 else peak_src,
            "kernel": "%s rollout_gather (%s)" % (kname, gather_mode),
            "nvlink_bytes_sent_per_gpu_per_step": sent, "nvlink_gbs_per_gpu": nv,
            "hbm_gbs_written_per_gpu": hbm_written, "hbm_frac": hbm_written / peak,
            "algorithmic_bytes_per_env_step": out_b + state_b / T,
            "bytes_per_launch": (out_b + state_b / T) * world * B * T, "launch_us": step_us * T,
            "note": "every output byte of a shard is delivered to each of the other %d GPU(s): the exchange, not the stepper, bounds this configuration" % (world - 1)}


def _max_over_ranks(x, world, device):
    import torch
    import torch.distributed as dist
    if world > 1:
        t = torch.tensor([x], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x


def measure_e2e(spec, args, world=1, rank=0, device="cuda:0"):
    """The same metric end to end through the host-buffer API: every step copies that step's
    actions from pinned host memory to the device, runs the step kernel and copies EVERY output
    field of the step back into pinned host memory.

    Headline `value`: VecCheckers/VecParticle.rollout_host -> cm3_*_rollout_host, which double
    buffers the device side (the D2H copy of step t overlaps the kernel of step t + 1), with the
    most compact lossless encoding for Checkers (grid / obs_self_t, whose values are always in
    {-1, 0, +1}, packed 2 bits per cell); `int8_tiles` is the same call with one byte per cell in the
    reference's array shapes; `unpipelined_fp32` is round 1's call, step_host with fp32 tiles: copy,
    kernel, copy in series.  With world > 1 every rank drives its own GPU (own PCIe link) at the same
    time; values are the whole job's, from the slowest rank."""
    import torch
    import torch.distributed as dist
    from cm3_b200 import VecCheckers
    B, N = args.envs, spec["n"]
    T = MAX_STEPS
    rng = np.random.default_rng(SEED)
    acts = rng.integers(0, 5, size=(T, B, N)).astype(np.int8)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    res = {}
    reps = max(1, min(3, args.e2e_steps // T)) if args.e2e_steps >= T else 1
    Th = T if args.e2e_steps >= T else max(3, args.e2e_steps)
    a = acts[:Th]
    steps = reps * Th
    API = ("Vec*.rollout_host -> cm3_*_rollout_host: per step H2D of the step's actions (pinned), kernel, D2H of every output "
           "field (pinned); the D2H of step t overlaps the kernel of step t+1")

    def pipelined(tile_dtype, enc):
        if spec["kind"] == "checkers":
            env = VecCheckers(B, device=device, tile_dtype=tile_dtype, env_id_offset=rank * B, **spec["ctor"])
            env.reset(goals=np.eye(2) if spec["n"] == 2 else np.array([[1, 0]]))
        else:
            env = make_env(spec, B, device, env_id_offset=rank * B)
        env.rollout_host(acts)  # warm-up (allocates the pinned areas)
        env.rollout_host(a)
        sync_all()
        t0 = time.perf_counter()
        for r in range(reps):
            out = env.rollout_host(a, t0=r * Th)
        float(out["reward"][-1, 0])
        el = _max_over_ranks(time.perf_counter() - t0, world, device)
        bo = sum(v[0].nbytes for v in out.values())
        del env
        return {"value": world * B * N * steps / el, "unit": UNIT, "h2d_bytes_per_step": world * B * N,
                "d2h_bytes_per_step": world * int(bo), "n_gpus": world, "steps": steps, "ms_per_step": el * 1e3 / steps,
                "api": API, "encoding": enc, "d2h_gbs_per_gpu": bo * steps / el / 1e9,
                "note": "in-kernel episode reset on; PCIe-bound: %.1f MB D2H per step per GPU" % (bo / 1e6)}

    # (a) pipelined, with the most compact lossless encoding the library offers: Checkers packs the two tile
    # outputs 2 bits per cell (CM3_TILE_U2, include/cm3env.h; cm3_b200/tiles.py decodes); (a') the same with
    # int8 tiles (arrays of the reference's shapes, one byte per cell).  Particle outputs are fp32 as they are.
    if spec["kind"] == "checkers":
        res.update(pipelined("u2", "grid / obs_self_t packed 2 bits per cell in 32-bit words (CM3_TILE_U2: lossless, their values are "
                                   "always in {-1, 0, +1}; cm3_b200/tiles.py decodes on the consumer's side), vectors and rewards fp32"))
        res["int8_tiles"] = pipelined(torch.int8, "grid / obs_self_t int8 in the reference's array shapes (lossless), vectors and rewards fp32")
    else:
        res.update(pipelined(None, "fp32"))
    # (b) round 1's path: step_host, fp32 tiles, copy / kernel / copy in series
    env = make_env(spec, B, device, env_id_offset=rank * B)
    steps = max(3, min(args.e2e_steps, 30))
    for t in range(3):
        env.step_host(acts[t % T])
    sync_all()
    t0 = time.perf_counter()
    for t in range(steps):
        out = env.step_host(acts[t % T])
    float(out["reward"][0])
    el = _max_over_ranks(time.perf_counter() - t0, world, device)
    bo = sum(v.nbytes for v in out.values())
    res["unpipelined_fp32"] = {"value": world * B * N * steps / el, "unit": UNIT, "d2h_bytes_per_step": world * int(bo),
                               "steps": steps, "ms_per_step": el * 1e3 / steps,
                               "api": "Vec*.step_host -> cm3_*_step_host_packed (fp32 tiles; no auto-reset: reference semantics)"}
    del env
    torch.cuda.empty_cache()
    return res


def measure_dropin_latency(device):
    """B = 1 drop-ins (the only form alg/train_onpolicy.py:302-350 can use unchanged): microseconds per
    env.step through the reference's own API, beside the reference's Python step."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "cm3_b200", "dropin"))
    res = {}
    try:
        from cm3_b200 import presets
        from env.checkers import Checkers   # the way a trainer imports it (train_offpolicy.py:24), dropin/ on sys.path
        ck = presets.CHECKERS["stage2"]
        env = Checkers(ck["n_rows"], ck["n_columns"], ck["n_obs"], ck["agents_r"], ck["agents_c"], ck["n_agents"], MAX_STEPS)
        rng = np.random.default_rng(0)
        env.reset(np.eye(2))
        n = 600
        acts = rng.integers(0, 5, size=(n, 2))
        for t in range(50):
            env.step(acts[t])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(n):
            if t % MAX_STEPS == 0:
                env.reset(np.eye(2))
            env.step(acts[t])
        res["checkers_stage2_us_per_step"] = (time.perf_counter() - t0) * 1e6 / n
        res["checkers_reference_python_us_per_step"] = 94.0
    except Exception as e:  # noqa: BLE001
        res["checkers_error"] = repr(e)
    try:
        from multiagent.environment import MultiAgentEnv
        import multiagent.scenarios as scenarios
        from cm3_b200 import presets
        scenario = scenarios.load("multi-goal_spread.py").Scenario()
        world = scenario.make_world(4, presets.PARTICLE["antipodal"], 0.0)
        env = MultiAgentEnv(world, scenario.reset_world, scenario.reward, scenario.observation, None, scenario.done, max_steps=MAX_STEPS)
        env.reset()
        rng = np.random.default_rng(0)
        n = 600
        acts = rng.integers(0, 5, size=(n, 4))
        for t in range(50):
            env.step(acts[t])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(n):
            if t % MAX_STEPS == 0:
                env.reset()
            env.step(acts[t])
        res["particle_antipodal_us_per_step"] = (time.perf_counter() - t0) * 1e6 / n
        res["particle_reference_python_us_per_step"] = 380.0
    except Exception as e:  # noqa: BLE001
        res["particle_error"] = repr(e)
    res["note"] = "B = 1 drop-in facades stepping through the reference's API (resets included); reference figures: BASELINE.md section 2, probed in the build container"
    return res


def measure_gather(wl, B, mode, rank, world, device, local_rank, peak, peak_src, reps=6, overlap=True):
    """BASELINE configs[3] as an `extra` line at world > 1 (all ranks call it)."""
    from cm3_b200.sharding import EnvShard
    spec = workload_spec(wl)
    T = MAX_STEPS
    env = make_env(spec, B, device, env_id_offset=rank * B)
    try:
        runner = GatherRunner(env, EnvShard(world * B, rank=rank, world=world, local_rank=local_rank), T, mode, overlap=overlap)
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}
    ms, clocks, per_rank = timed_steps(runner, reps * T, 2 * T, world, device, local_rank, sample_clocks=False)
    step_us = ms * 1e3 / (reps * T)
    out_b = env.out_bytes_per_env_step()
    state_b = env.bytes_per_env_step() - out_b - spec["n"]
    r = gather_roofline(out_b, state_b, T, B, world, step_us, peak, peak_src, KERNEL_NAMES[wl], runner.mode)
    return {"value": world * B * spec["n"] * reps * T / (ms * 1e-3), "unit": UNIT, "envs_per_gpu": B, "total_envs": world * B,
            "mode": runner.mode, "overlap": overlap, "us_per_step": step_us, "launches": reps, "per_rank_ms": per_rank,
            "nvlink_gbs_per_gpu": r["nvlink_gbs_per_gpu"], "frac_of_peer_copy_770": r["frac"],
            "hbm_gbs_written_per_gpu": r["hbm_gbs_written_per_gpu"]}


def measure_sweep(rank, world, device, local_rank, peak, workloads=("ck2", "pa4"), max_envs=1048576):
    """BASELINE configs[4]: batch sweep 256 ... 1 048 576 envs per GPU, fused 33-step launches and
    chained per-step launches, whole-job numbers at this world size."""
    lines = []
    for wl in workloads:
        spec = workload_spec(wl)
        B = 256
        while B <= max_envs:
            K = 330 if B <= 262144 else 99
            r = measure_workload(spec, wl, B, K, 33, rank, world, device, local_rank, peak, modes=("fused", "per_step_chained"),
                                 sample_clocks=False)
            line = {"workload": wl, "envs_per_gpu": B, "n_gpus": world}
            for k in ("fused", "per_step_chained"):
                if k in r:
                    line[k] = {x: r[k][x] for x in ("value", "us_per_step", "achieved_gbs", "frac")}
            lines.append(line)
            B *= 4
    return lines


def measure_extras(env, spec, args, peak, mode="fused", world=1, rank=0, device="cuda:0", local_rank=0, ring=MAX_STEPS):
    """Everything beside the headline, measured by the same timed_steps(): all ranks call this."""
    import torch
    extra = {}
    B, N = env.B, env.N
    K = max(args.steps, 330)
    peak_src = hbm_peak()[1]
    # (0) this workload, one launch PER STEP (what a policy in the loop needs)
    r = measure_workload(spec, args.workload, B, K, 33, rank, world, device, local_rank, peak, modes=("step",))
    extra["per_step_launches"] = dict(r["per_step_chained"], note="one cm3_*_step_chained launch per step from a CUDA graph: tile i of step k+1 waits only for tile i of step k (ticket words), so the store phase of a step overlaps the compute phase of the next; outputs to a rollout ring > L2")
    extra["per_step_stream_ordered"] = dict(r["per_step_stream_ordered"], note="round 1's per-step mode: plain stream order + programmatic dependent launch (griddepcontrol.wait on the whole previous grid)")
    # (1) fused T-step rollout kernel with device Philox actions
    T = MAX_STEPS
    out = env.alloc_outputs(T, fields=ref_fields(env))
    plan = env.plan_rollout(T, actions=None, out=out, auto_reset=True, seed=SEED)

    class _P(object):
        launches = 0

        def run(self, k):
            for i in range(k // T):
                plan(self.launches * T)
                self.launches += 1
    reps = max(3, min(30, K // T))
    ms, _, _ = timed_steps(_P(), reps * T, 3 * T, world, device, local_rank, sample_clocks=False)
    fb = fused_bytes(env, spec, T, with_actions=False)
    ach = fb * B * T * reps / (ms * 1e-3) / 1e9
    extra["fused_rollout_T33_philox"] = {"value": world * B * N * T * reps / (ms * 1e-3), "unit": UNIT, "launches": reps,
                                         "ms_per_launch": ms / reps, "achieved_gbs": ach, "frac": ach / peak,
                                         "algorithmic_bytes_per_env_step": fb,
                                         "note": "one launch = 33 env steps, device Philox actions, in-kernel reset"}
    del out, plan
    torch.cuda.empty_cache()
    # (1b) Checkers with the lossless compact tile encodings: the same step, 1/4 and 1/16 of the tile bytes
    if spec["kind"] == "checkers":
        for tag, td in (("int8", torch.int8), ("u2", "u2")):
            r = measure_workload(dict(spec, tile_dtype=td), args.workload, B, K, 33, rank, world, device, local_rank, peak,
                                 modes=("fused",), sample_clocks=False)
            extra["fused_%s_tiles" % tag] = dict(r["fused"], note="grid / obs_self_t as %s: the roofline fraction is against THIS layout's algorithmic bytes" % (
                "int8" if tag == "int8" else "2 bits per cell (CM3_TILE_U2)"))
    # (2) the other env family and the other configs named by BASELINE.json, same batch per GPU
    others = [w for w in ("ck2", "pa4", "pa3", "pm2", "ck1") if w != args.workload]
    extra["workloads"] = {}
    for wl in others:
        extra["workloads"][wl] = measure_workload(workload_spec(wl), wl, B, K, 33, rank, world, device, local_rank, peak,
                                                  modes=("fused", "per_step_chained"))
    # (3) configs[3]: PM2 rollout all-gather (32 768 envs per GPU), by NCCL and by the kernel's own peer stores
    if world > 1:
        for m in ("nccl", "peer"):
            extra["gather_pm2_%s" % m] = measure_gather("pm2", 32768, m, rank, world, device, local_rank, peak, peak_src)
        extra["gather_pm2_peer_no_overlap"] = measure_gather("pm2", 32768, "peer", rank, world, device, local_rank, peak, peak_src, overlap=False)
    # (4) configs[4]: batch sweep at this world size
    if not args.no_sweep:
        extra["sweep"] = measure_sweep(rank, world, device, local_rank, peak)
    # (5) B = 1 drop-in latency
    if rank == 0:
        extra["dropin_b1"] = measure_dropin_latency(device)
    return extra


def cpu_baseline(spec, args):
    import oracle
    Bs = 4096
    v, n_steps, el = time_oracle(spec, Bs, 1, args.cpu_seconds)
    return {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d envs x %d steps of the same workload (random actions, caller-side reset every 33 steps), %.1f s on 1 thread of %d host cores; C float64 port of the reference's step() (oracle/cm3_oracle.c)" % (Bs, n_steps, el, os.cpu_count()),
            "python_reference_note": "the reference's own pure-Python step() ran at ~1.75e4 (Checkers stage2) / ~1.05e4 (particle N=4) agent-env-steps/s on one core of the build container (BASELINE.md section 2); it cannot run on this box"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path on all host threads (the oracle port;
    the Python reference itself does not travel to this box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    spec = workload_spec(args.workload)
    # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1: not a limit we accept
    # for the baseline)
    nthreads = max(oracle.max_threads(), len(os.sched_getaffinity(0)))
    Bs = args.envs if args.ref_envs <= 0 else min(args.envs, args.ref_envs)   # default: the FULL batch per step
    K, W = args.steps, max(args.warmup, 3)
    env, reset = make_oracle(spec, Bs, nthreads)
    rng = np.random.default_rng(SEED)
    actions = rng.integers(0, 5, size=(MAX_STEPS, Bs, spec["n"])).astype(np.int32)
    budget_s = 120.0
    for t in range(W):
        env.step(actions[t % MAX_STEPS])
    reset()
    t0 = time.perf_counter()
    done_steps = 0
    for t in range(K):
        env.step(actions[t % MAX_STEPS])
        if (t + 1) % MAX_STEPS == 0:
            reset()
        done_steps += 1
        if time.perf_counter() - t0 > budget_s:
            break
    el = time.perf_counter() - t0
    value = Bs * spec["n"] * done_steps / el
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample = "%d envs per step (of the %d-env workload) x %d steps, %d threads (OpenMP over envs), caller-side reset every 33 steps" % (Bs, args.envs, done_steps, nthreads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": world, "steps": done_steps, "warmup": W,
        "ms_per_step": el * 1e3 / done_steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the same workload description as the GPU arm prints (the driver compares the two)
        "config": workload_config(spec, args.envs),
        "launch_config": {"mode": "cpu", "parallelism": "OpenMP over envs, %d host threads, one process" % nthreads,
                          "timing": "time.perf_counter around the timed steps", "launch": "one C call per step; caller-side reset every 33 steps",
                          "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3300)
    ap.add_argument("--warmup", type=int, default=99)
    ap.add_argument("--impl", default="cm3_b200", choices=["cm3_b200", "reference"])
    ap.add_argument("--workload", default="ck2")
    ap.add_argument("--envs", type=int, default=B_PER_GPU, help="env instances per GPU")
    ap.add_argument("--e2e-steps", type=int, default=99)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="--gather: single symmetric buffer, no overlap of rollout k+1 with the drain of k")
    ap.add_argument("--ref-envs", type=int, default=0,
                    help="--impl reference: envs per step (0 = the full --envs batch; a smaller sample stays L3-resident)")
    ap.add_argument("--mode", default="fused", choices=["fused", "step", "step_ordered"],
                    help="fused: one launch per 33-step episode (headline); step: one chained launch per step; step_ordered: round 1's stream-ordered per-step launches")
    ap.add_argument("--gather", default="none", choices=["none", "nccl", "peer", "auto"],
                    help="all-gather the rollout buffers to every GPU (BASELINE.json configs[3])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
