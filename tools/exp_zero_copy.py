#!/usr/bin/env python
"""Experiment: host-buffer stepping (e2e path) variants on one GPU.
 A  step_host: H2D actions, kernel into device outputs, D2H every field, sync (cm3_*_step_host)
 B  zero-copy outputs: kernel stores straight into pinned (device-mapped) host buffers over PCIe
 C  B + zero-copy actions: the kernel also reads the actions from pinned host memory
 D  A split into k chunks on k streams (copy/compute overlap)
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cm3_b200 import VecCheckers, VecParticle, presets, _lib as L  # noqa: E402


def make(kind, B, tile=None, off=0):
    if kind == "ck2":
        env = VecCheckers(B, tile_dtype=tile, env_id_offset=off, max_steps=33, **presets.CHECKERS["stage2"])
        env.reset(goals=np.eye(2))
    else:
        env = VecParticle(B, 4, presets.PARTICLE["antipodal"], max_steps=33, env_id_offset=off)
        env.reset(seed=1)
    return env


def timeit(fn, steps=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


def main():
    B = 65536
    for kind, tile in (("ck2", None), ("ck2", torch.int8), ("pa4", None)):
        env = make(kind, B, tile)
        N = env.N
        acts = np.random.default_rng(0).integers(0, 5, size=(B, N)).astype(np.int8)
        nbytes = env.out_bytes_per_env_step() * B
        tag = "%s%s" % (kind, "_i8" if tile is not None else "")
        tA = timeit(lambda: env.step_host(acts))
        print("%s A step_host            %.3f ms  %.3g agent-steps/s  D2H %.1f GB/s" % (tag, tA * 1e3, B * N / tA, nbytes / tA / 1e9), flush=True)
        # B: zero-copy outputs
        host = env.alloc_outputs(T=1, pinned_host=True)
        ptrs = [{f: host[f].data_ptr() for f in host}]
        hact = torch.zeros(1, B, N, dtype=torch.int8).pin_memory()
        dact = torch.zeros(1, B, N, dtype=torch.int8, device=env.device)
        stream = torch.cuda.current_stream()

        def fB():
            hact.numpy()[0] = acts
            dact.copy_(hact, non_blocking=True)
            env.rollout_gather(1, ptrs, B, 0, actions=dact)
            stream.synchronize()
        try:
            tB = timeit(fB)
            print("%s B zero-copy outputs    %.3f ms  %.3g agent-steps/s  PCIe %.1f GB/s" % (tag, tB * 1e3, B * N / tB, nbytes / tB / 1e9), flush=True)
            # verify against device step
            dev = env.step(acts)
            fB()
        except Exception as e:  # noqa: BLE001
            print(tag, "B failed:", repr(e))
        # C: zero-copy actions as well (raw pointer through the C ABI)
        lib = env.lib
        is_ck = kind == "ck2"
        Out = L.CheckersOutputs if is_ck else L.ParticleOutputs
        fn = lib.cm3_checkers_rollout_gather if is_ck else lib.cm3_particle_rollout_gather
        FIELDS = Out.FIELDS
        arr = (Out * 1)()
        arr[0] = Out(*[C.c_void_p(host[f].data_ptr()) for f in FIELDS])

        def fC():
            hact.numpy()[0] = acts
            L.check(fn(env._h, C.byref(env._st), C.c_void_p(hact.data_ptr()), 0, 0, 1, 0, None, 1, arr, B, 0,
                       C.c_void_p(stream.cuda_stream)))
            stream.synchronize()
        try:
            tC = timeit(fC)
            print("%s C zero-copy in+out     %.3f ms  %.3g agent-steps/s  PCIe %.1f GB/s" % (tag, tC * 1e3, B * N / tC, nbytes / tC / 1e9), flush=True)
        except Exception as e:  # noqa: BLE001
            print(tag, "C failed:", repr(e))
        del env
        # D: chunked step_host on k streams
        for k in (2, 4):
            Bc = B // k
            envs = [make(kind, Bc, tile, off=i * Bc) for i in range(k)]
            streams = [torch.cuda.Stream() for _ in range(k)]
            for e in envs:
                e.step_host(acts[:Bc])   # allocate pinned buffers
            hosts = [e._host for e in envs]

            def fD():
                for i, (e, s) in enumerate(zip(envs, streams)):
                    with torch.cuda.stream(s):
                        e._host_actions.numpy()[...] = acts[i * Bc:(i + 1) * Bc]
                        e._actions_dev.copy_(e._host_actions, non_blocking=True)
                        e.step(e._actions_dev)
                        for f, v in e.out.items():
                            hosts[i][f].copy_(v, non_blocking=True)
                for s in streams:
                    s.synchronize()
            tD = timeit(fD)
            print("%s D %d chunks/streams      %.3f ms  %.3g agent-steps/s  D2H %.1f GB/s" % (tag, k, tD * 1e3, B * N / tD, nbytes / tD / 1e9), flush=True)
            del envs


if __name__ == "__main__":
    main()
