#!/bin/bash
# round 2, first GPU call: parity suite on the new kernels, the default bench line and the
# driver-style short bench (--steps 20 --warmup 5), the reference arm.
set -u
mkdir -p gpurun_out
TAG=${1:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ck2_driverlike_$TAG.json 2> gpurun_out/bench_ck2_driverlike_$TAG.err ) 2>&1 | tail -4; echo "bench driverlike rc=$?"; tail -5 gpurun_out/bench_ck2_driverlike_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_driverlike_$TAG.json
( time timeout 900 python bench.py --no-sweep > gpurun_out/bench_ck2_$TAG.json 2> gpurun_out/bench_ck2_$TAG.err ) 2>&1 | tail -4; echo "bench rc=$?"; tail -5 gpurun_out/bench_ck2_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_$TAG.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_ck2_$TAG.json 2>&1; tail -c 600 gpurun_out/bench_ref_ck2_$TAG.json
