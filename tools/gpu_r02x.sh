#!/bin/bash
# r02x: the bench lines of the final kernels again (after the per-step ring went back to its original size):
# default, driver-style, per workload, reference arms.  ncu captures, sanitizer logs and the parity suite: r02w.
set -u
mkdir -p gpurun_out
TAG=${1:-r02x}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv; nproc
( time timeout 900 python bench.py > gpurun_out/bench_ck2_$TAG.json 2> gpurun_out/bench_ck2_$TAG.err ) 2>&1 | grep real; echo "bench ck2 rc=$?"; tail -3 gpurun_out/bench_ck2_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_$TAG.json 2>/dev/null | head -16
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ck2_driverlike_$TAG.json 2> gpurun_out/bench_ck2_driverlike_$TAG.err ) 2>&1 | grep real; echo "bench driverlike rc=$?"
python tools/bench_summary.py gpurun_out/bench_ck2_driverlike_$TAG.json 2>/dev/null | head -16
for wl in pa4 pa3 pm2 ck1; do
  timeout 600 python bench.py --workload $wl --no-extras > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_${wl}_$TAG.err
  python tools/bench_summary.py gpurun_out/bench_${wl}_$TAG.json 2>/dev/null | head -4
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_ck2_$TAG.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/bench_ref_ck2_$TAG.json')); print('ref ck2', d['value'], d['cpu_baseline']['cores'])"
timeout 300 python bench.py --impl reference --workload pa4 --steps 200 > gpurun_out/bench_ref_pa4_$TAG.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/bench_ref_pa4_$TAG.json')); print('ref pa4', d['value'], d['cpu_baseline']['cores'])"
