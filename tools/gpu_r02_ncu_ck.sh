#!/bin/bash
# ncu launch lists + full captures of the Checkers kernels (fused CK2 / CK1, chained per-step CK2 / CK1) and the sanitizer runs on the chained tests
set -u
mkdir -p gpurun_out
TAG=${1:-r02x2}
for wl in ck2 ck1; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${wl}_fused_$TAG.csv \
      python bench.py --workload $wl --steps 330 --warmup 33 --no-extras > gpurun_out/ncu_launch_${wl}_fused_$TAG.log 2>&1; echo "ncu launches fused $wl rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:checkers_kernel -s 4 -c 1 -f -o gpurun_out/prof_${wl}_fused_$TAG \
      python bench.py --workload $wl --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_${wl}_fused_$TAG.log 2>&1; echo "ncu full fused $wl rc=$?"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${wl}_step_$TAG.csv \
      python bench.py --workload $wl --mode step --steps 99 --warmup 3 --no-extras > gpurun_out/ncu_launch_${wl}_step_$TAG.log 2>&1; echo "ncu launches step $wl rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:checkers_kernel -s 40 -c 1 -f -o gpurun_out/prof_${wl}_step_$TAG \
      python bench.py --workload $wl --mode step --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_${wl}_step_$TAG.log 2>&1; echo "ncu full step $wl rc=$?"
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_checkers.py -x -q -k "chained or graph or ragged or u2" > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_$TAG.log
CM3_CHAIN_TPB=2 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -x -q -k "chained" > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_$TAG.log
