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


def make_env(spec, B, device, env_id_offset=0):
    from cm3_b200 import VecCheckers, VecParticle
    if spec["kind"] == "checkers":
        env = VecCheckers(B, device=device, env_id_offset=env_id_offset, **spec["ctor"])
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


class StepRunner(object):
    """K single-step launches into a rollout ring, replayed from a CUDA graph."""

    def __init__(self, env, spec, ring, seed):
        import torch
        self.torch, self.env, self.ring = torch, env, ring
        self.ring_out = env.alloc_outputs(ring)
        g = torch.Generator(device="cpu").manual_seed(seed)
        acts = torch.randint(0, 5, (ring, env.B, env.N), generator=g, dtype=torch.int8)
        self.actions = acts.to(env.device)
        self.slots = [{k: v[t] for k, v in self.ring_out.items()} for t in range(ring)]
        self.graph = None

    def launch(self, t):
        # one kernel launch: T=1 rollout with in-kernel episode reset, outputs -> ring slot
        self.env.rollout(1, actions=self.actions[t % self.ring:t % self.ring + 1], auto_reset=True,
                         out={k: v.unsqueeze(0) for k, v in self.slots[t % self.ring].items()})

    def capture(self):
        torch = self.torch
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


class FusedRunner(object):
    """K steps as K/T launches of the T-step kernel reading the pre-generated action stream."""

    def __init__(self, env, T, seed):
        import torch
        self.env, self.T = env, T
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.actions = torch.randint(0, 5, (T, env.B, env.N), generator=g, dtype=torch.int8).to(env.device)
        self.out = env.alloc_outputs(T)
        self.launches = 0

    def capture(self):
        pass

    def run(self, k):
        n_full, rem = divmod(k, self.T)
        for _ in range(n_full):
            self.env.rollout(self.T, actions=self.actions, auto_reset=True, out=self.out)
        if rem:
            self.env.rollout(rem, actions=self.actions[:rem], auto_reset=True,
                             out={f: v[:rem] for f, v in self.out.items()})
        self.launches += n_full + (1 if rem else 0)


def bytes_per_env_step(env):
    return env.bytes_per_env_step()


def timed_steps(runner, K, W, world, device, local_rank, sample_clocks=True):
    """W warm-up steps, then exactly K timed steps; returns (ms max over ranks, clocks)."""
    import torch
    import torch.distributed as dist

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    runner.run(W)
    sampler = ClockSampler(local_rank) if sample_clocks else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.start()
    e0.record()
    runner.run(K)
    e1.record()
    if sampler:
        sampler.sample()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, clocks


class GatherRunner(object):
    """BASELINE.json configs[3]: every launch is one fused T-step rollout of this rank's shard
    (device Philox actions, in-kernel episode reset) whose outputs are delivered to EVERY GPU as
    [T, total_envs, ...] - by ncclAllGather ("nccl") or by the kernel's own peer stores over
    NVLink ("peer", cm3_*_rollout_gather).  One "step" is still one env step of the whole batch."""

    def __init__(self, env, shard, T, mode):
        from cm3_b200.sharding import RolloutAllGather
        self.env, self.T = env, T
        self.g = RolloutAllGather(env, T, shard=shard, mode=mode)
        self.mode = self.g.mode
        self.launches = 0

    def capture(self):
        pass

    def run(self, k):
        for i in range(k // self.T):
            self.g.rollout(seed=SEED, t0=self.launches * self.T, auto_reset=True, time_major=False)
            self.launches += 1


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
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
    device = "cuda:%d" % local_rank
    spec = workload_spec(args.workload)
    B, K, W = args.envs, args.steps, args.warmup
    env = make_env(spec, B, device, env_id_offset=rank * B)
    bpe = bytes_per_env_step(env)
    # ring larger than L2 (126 MB): at least 33 slots and >= 512 MB of outputs
    ring = max(MAX_STEPS, int(np.ceil(512e6 / (bpe * B))))
    gather_mode = None
    T = MAX_STEPS
    W = max(W, 3)
    out_b = env.out_bytes_per_env_step()
    state_b = bpe - out_b - spec["n"]
    fused_bpe = out_b + spec["n"] + state_b / T
    launch_cfg = None
    if args.gather != "none":
        from cm3_b200.sharding import EnvShard
        K = max(T, K // T * T)
        W = (W + T - 1) // T * T
        runner = GatherRunner(env, EnvShard(world * B, rank=rank, world=world, local_rank=local_rank), T, args.gather)
        gather_mode = runner.mode
        mode = "gather"
    elif args.mode == "fused":
        runner = FusedRunner(env, T, SEED + rank)
        mode = "fused"
    else:
        runner = StepRunner(env, spec, ring, SEED + rank)
        mode = "step"
    runner.capture()
    ms, clocks = timed_steps(runner, K, W, world, device, local_rank)
    value = world * B * spec["n"] * K / (ms * 1e-3)
    peak, peak_src = hbm_peak()
    step_us = ms * 1e3 / K
    if mode == "step":
        eff_bpe, launches, launch_us = bpe, K, step_us
        launch_cfg = {"launch": "1 kernel launch per step, replayed from a CUDA graph of %d steps" % ring,
                      "rollout_ring_slots": ring,
                      "l2": "outputs go to a %d-slot rollout ring of %.0f MB (> 126 MB L2); no explicit flush" % (ring, ring * bpe * B / 1e6)}
    else:
        eff_bpe, launches = fused_bpe, (K + T - 1) // T
        launch_us = ms * 1e3 / launches
        launch_cfg = {"launch": "1 kernel launch per %d steps (one episode): state stays in registers between the steps of a launch" % T,
                      "l2": "every step writes all its outputs to a [%d][B] rollout buffer of %.0f MB (> 126 MB L2), overwritten by the next launch; no explicit flush" % (T, out_b * B * T / 1e6)}
    achieved = eff_bpe * B / (step_us * 1e-6) / 1e9
    kernel_key = "%s_%s" % (args.workload, "step" if mode == "step" else "rollout")
    kname = {"ck2": "checkers_kernel<3,8,2,2,float,float>", "ck1": "checkers_kernel<3,8,2,1,float,float>",
             "pa4": "particle_kernel<4,float>", "pa3": "particle_kernel<3,float>", "pm2": "particle_kernel<2,float>"}[args.workload]

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 bitboard state, f32 outputs" if spec["kind"] == "checkers" else "f32",
        "data": "synthetic",
        "config": dict({"workload": spec["label"] % B, "envs_per_gpu": B, "n_agents": spec["n"], "mode": mode,
                        "actions": "uniform int8 in 0..4, a [33][B][N] stream pre-generated in HBM and read by every step",
                        "auto_reset": True,
                        "parallelism": "env batch sharded over %d GPU(s), no per-step collective" % world}, **launch_cfg),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(kernel_key),
                     "peak_source": peak_src, "algorithmic_bytes_per_env_step": eff_bpe,
                     "bytes_per_launch": eff_bpe * B * (K / launches), "launch_us": launch_us, "kernel": kname},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if gather_mode is not None:
        sent = out_b * B * (world - 1)  # bytes this GPU delivers to its peers per env step
        launch_us = step_us
        out["gpu_launches"] = K // T
        out["config"].update({"launch": "1 launch per %d fused env steps, device Philox actions" % T,
                              "actions": "Philox4x32-10 on the device, keyed by (seed, global env id, step)",
                              "rollout_all_gather": gather_mode,
                              "l2": "gathered rollout buffer [%d][%d envs] = %.0f MB per GPU (> 126 MB L2 when >= 2 GPUs at the default size)" % (T, world * B, out_b * world * B * T / 1e6)})
        hbm_written = (out_b + state_b / T) * world * B / (step_us * 1e-6) / 1e9  # the whole gathered batch lands in this GPU's HBM
        nv = sent / (step_us * 1e-6) / 1e9 if world > 1 else 0.0
        out["roofline"] = {"bound": "nvlink" if world > 1 else "hbm", "achieved": nv if world > 1 else hbm_written,
                           "peak": NVLINK_PEER_GBS if world > 1 else peak, "unit": "GB/s",
                           "frac": (nv / NVLINK_PEER_GBS) if world > 1 else hbm_written / peak, "traffic": None,
                           "peak_source": "measured peer-copy reference, B200_PROFILING.md (770 GB/s per direction per GPU; 900 nominal)" if world > 1 else peak_src,
                           "kernel": "%s rollout_gather (%s)" % (kname, gather_mode),
                           "nvlink_bytes_sent_per_gpu_per_step": sent, "nvlink_gbs_per_gpu": nv,
                           "hbm_gbs_written_per_gpu": hbm_written, "hbm_frac": hbm_written / peak,
                           "algorithmic_bytes_per_env_step": out_b + state_b / T,
                           "bytes_per_launch": (out_b + state_b / T) * world * B * T, "launch_us": step_us * T,
                           "note": "every output byte of a shard is delivered to each of the other %d GPU(s): the exchange, not the stepper, bounds this configuration" % (world - 1)}
    if args.gather != "none":
        pass  # the collective run reports the kernel-side number only
    else:
        out["e2e"] = measure_e2e(env, spec, args, world)   # all ranks take part
        if rank == 0 and not args.no_extras:
            out["extra"] = measure_extras(env, spec, args, peak, mode, world, device, local_rank, ring)
            if spec["kind"] == "checkers":
                out["extra"]["e2e_int8_tiles"] = measure_e2e_int8(spec, args, device)
            if world == 1:
                out["cpu_baseline"] = cpu_baseline(spec, args)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def measure_e2e(env, spec, args, world=1):
    """The same metric through the host-buffer entry point (cm3_*_step_host): each step copies the
    step's actions from pinned host memory to the device, launches the kernel and copies EVERY
    output field back into pinned host memory, then waits.  With world > 1 every rank drives its
    own GPU (own PCIe link) at the same time; the value is the whole job's, from the slowest rank."""
    import torch
    import torch.distributed as dist
    B, N = env.B, env.N
    rng = np.random.default_rng(SEED)
    acts = rng.integers(0, 5, size=(8, B, N)).astype(np.int8)
    steps = max(3, min(args.e2e_steps, args.steps))
    for t in range(3):
        env.step_host(acts[t % 8])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for t in range(steps):
        out = env.step_host(acts[t % 8])
    float(out["reward"][0])
    el = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([el], device=env.device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        el = float(tt.item())
    bo = sum(v.nbytes for v in out.values())
    return {"value": world * B * N * steps / el, "unit": UNIT, "h2d_bytes_per_step": world * B * N,
            "d2h_bytes_per_step": world * int(bo), "n_gpus": world,
            "steps": steps, "ms_per_step": el * 1e3 / steps,
            "api": "VecCheckers/VecParticle.step_host -> cm3_*_step_host (host actions in, all output fields out, pinned)",
            "note": "no auto-reset on this path (reference semantics); PCIe-bound: %.1f MB D2H per step" % (bo / 1e6)}


def measure_e2e_int8(spec, args, device):
    """Checkers only: the same host-buffer call with grid / obs_self_t delivered as int8 (lossless:
    their values are always in {-1, 0, +1}) - a quarter of the tile bytes over PCIe."""
    import torch
    from cm3_b200 import VecCheckers
    env = VecCheckers(args.envs, device=device, tile_dtype=torch.int8, **spec["ctor"])
    env.reset(goals=np.eye(2) if spec["n"] == 2 else np.array([[1, 0]]))
    res = measure_e2e(env, spec, args)
    res["api"] = "VecCheckers(tile_dtype=int8).step_host -> cm3_checkers_step_host, cfg.tile = CM3_TILE_I8"
    return res


def measure_extras(env, spec, args, peak, mode="fused", world=1, device="cuda:0", local_rank=0, ring=MAX_STEPS):
    import torch
    extra = {}
    B, N = env.B, env.N
    bpe = bytes_per_env_step(env)
    if mode == "fused":
        # (0) the same workload as one launch PER STEP (what a policy in the loop needs)
        runner = StepRunner(env, spec, ring, SEED)
        runner.capture()
        K = max(ring, args.steps // ring * ring)
        ms, _ = timed_steps(runner, K, ring, 1, device, local_rank, sample_clocks=False)
        ach = bpe * B * K / (ms * 1e-3) / 1e9
        extra["per_step_launches"] = {"value": B * N * K / (ms * 1e-3), "unit": UNIT, "launches": K,
                                      "us_per_step": ms * 1e3 / K, "achieved_gbs": ach, "frac": ach / peak,
                                      "algorithmic_bytes_per_env_step": bpe, "rollout_ring_slots": ring,
                                      "note": "one launch per step from a CUDA graph with programmatic dependent launch; outputs to a %d-slot ring of %.0f MB (> L2); this GPU only" % (ring, ring * bpe * B / 1e6)}
        del runner
    # (1) fused T-step rollout kernel: state in registers, Philox actions, one launch per 33 steps
    T = MAX_STEPS
    out = env.alloc_outputs(T)
    for _ in range(3):
        env.rollout(T, actions=None, seed=SEED, auto_reset=True, out=out)
    reps = max(3, min(30, args.steps // T))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for r in range(reps):
        env.rollout(T, actions=None, seed=SEED, t0=r * T, auto_reset=True, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out_bytes = env.out_bytes_per_env_step()
    state_bytes = bpe - out_bytes - N
    fused_bpe = out_bytes + state_bytes / T  # actions come from Philox: no action bytes
    ach = fused_bpe * B * T * reps / (ms * 1e-3) / 1e9
    extra["fused_rollout_T33_philox"] = {"value": B * N * T * reps / (ms * 1e-3), "unit": UNIT, "launches": reps,
                                  "ms_per_launch": ms / reps, "achieved_gbs": ach, "frac": ach / peak,
                                  "algorithmic_bytes_per_env_step": fused_bpe,
                                  "note": "one launch = 33 env steps, device Philox actions, in-kernel reset, outputs to a [33][B] rollout buffer (%.0f MB > L2)" % (out_bytes * B * T / 1e6)}
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
    t_probe = time.perf_counter()
    for t in range(W):
        env.step(actions[t % MAX_STEPS])
    per_step = (time.perf_counter() - t_probe) / W
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
    sample = "%d envs per step (of the %d-env workload) x %d steps, %d threads (OpenMP over envs), caller-side reset every 33 steps" % (Bs, args.envs, done_steps, nthreads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": done_steps, "warmup": W,
        "ms_per_step": el * 1e3 / done_steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["label"] % args.envs, "sample": sample},
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
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--ref-envs", type=int, default=0,
                    help="--impl reference: envs per step (0 = the full --envs batch; a smaller sample stays L3-resident)")
    ap.add_argument("--mode", default="fused", choices=["fused", "step"],
                    help="fused: one launch per 33-step episode (headline); step: one launch per step")
    ap.add_argument("--gather", default="none", choices=["none", "nccl", "peer", "auto"],
                    help="all-gather the rollout buffers to every GPU (BASELINE.json configs[3])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
