#!/bin/bash
# Runs on the GPU box (through gpurun): parity tests, bench lines, ncu launch list and one full
# ncu capture of each dominant kernel.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_ck2_$TAG.json 2> gpurun_out/bench_ck2_$TAG.err; echo "bench ck2 rc=$?"; cat gpurun_out/bench_ck2_$TAG.json; tail -3 gpurun_out/bench_ck2_$TAG.err
python bench.py --workload pa4 > gpurun_out/bench_pa4_$TAG.json 2> gpurun_out/bench_pa4_$TAG.err; echo "bench pa4 rc=$?"; cat gpurun_out/bench_pa4_$TAG.json; tail -3 gpurun_out/bench_pa4_$TAG.err
python bench.py --impl reference --steps 200 > gpurun_out/bench_ref_ck2_$TAG.json 2>&1; cat gpurun_out/bench_ref_ck2_$TAG.json
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
# launch list (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ck2_$TAG.csv \
    python bench.py --steps 99 --warmup 3 --no-extras > gpurun_out/ncu_launch_ck2_$TAG.log 2>&1; echo "ncu launches ck2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pa4_$TAG.csv \
    python bench.py --workload pa4 --steps 99 --warmup 3 --no-extras > gpurun_out/ncu_launch_pa4_$TAG.log 2>&1; echo "ncu launches pa4 rc=$?"
# full capture of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:checkers_kernel -s 20 -c 3 -f -o gpurun_out/prof_ck2_$TAG \
    python bench.py --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_ck2_$TAG.log 2>&1; echo "ncu full ck2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:particle_kernel -s 20 -c 3 -f -o gpurun_out/prof_pa4_$TAG \
    python bench.py --workload pa4 --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_pa4_$TAG.log 2>&1; echo "ncu full pa4 rc=$?"
ls -la gpurun_out | head -40
