#!/bin/bash
# bench lines of the round-2 reference run (same commands as tools/gpu_r02_final.sh, without the ncu / sanitizer part)
set -u
mkdir -p gpurun_out
TAG=${1:-r02z}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory --format=csv; nproc
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "graph or chained" > gpurun_out/pytest_gpu_graph_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_graph_$TAG.log
( time timeout 900 python bench.py > gpurun_out/bench_ck2_$TAG.json 2> gpurun_out/bench_ck2_$TAG.err ) 2>&1 | grep real; echo "bench ck2 rc=$?"; tail -3 gpurun_out/bench_ck2_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_$TAG.json 2>/dev/null | head -14
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ck2_driverlike_$TAG.json 2> gpurun_out/bench_ck2_driverlike_$TAG.err ) 2>&1 | grep real; echo "bench driverlike rc=$?"
python tools/bench_summary.py gpurun_out/bench_ck2_driverlike_$TAG.json 2>/dev/null | head -8
for wl in pa4 pa3 pm2 ck1; do
  timeout 600 python bench.py --workload $wl --no-extras > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_${wl}_$TAG.err
  python tools/bench_summary.py gpurun_out/bench_${wl}_$TAG.json 2>/dev/null | head -3
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_ck2_$TAG.json 2>&1; python -c "import json; d=json.load(open('gpurun_out/bench_ref_ck2_$TAG.json')); print('ref ck2', d['value'], d['cpu_baseline']['cores'], sorted(d['config']))"
timeout 300 python bench.py --impl reference --workload pa4 --steps 200 > gpurun_out/bench_ref_pa4_$TAG.json 2>&1; python -c "import json; d=json.load(open('gpurun_out/bench_ref_pa4_$TAG.json')); print('ref pa4', d['value'], d['cpu_baseline']['cores'])"
python tools/bench_summary.py gpurun_out/bench_ck2_$TAG.json 2>/dev/null | sed -n 3,12p
