#!/bin/bash
# r02r: full parity suite on the current defaults, double-buffered staging for N = 3 (CM3_PT_STAGE_MAXN=3 build),
# the complete default bench line (wall time of the whole run, e2e with the packed tiles as the headline).
set -u
mkdir -p gpurun_out
TAG=${1:-r02r}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
AB=gpurun_out/ab_$TAG.txt
: > $AB
for rep in 1 2; do
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_stage3.so python tools/ab_r02.py --tag "two staging sets N<=3" --workloads pa3 >> $AB 2>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag "two staging sets N<=2 (default)" --workloads pa3 >> $AB 2>>gpurun_out/ab_$TAG.err
done
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_stage3.so python tools/ab_r02.py --tag "two staging sets N<=3 262k" --workloads pa3 --envs 262144 --modes fused >> $AB 2>>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag "default 262k" --workloads pa3 --envs 262144 --modes fused >> $AB 2>>gpurun_out/ab_$TAG.err
cat $AB; tail -3 gpurun_out/ab_$TAG.err
( time timeout 900 python bench.py > gpurun_out/bench_ck2_$TAG.json 2> gpurun_out/bench_ck2_$TAG.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench_ck2_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_$TAG.json 2>/dev/null | head -20
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ck2_driverlike_$TAG.json 2> gpurun_out/bench_ck2_driverlike_$TAG.err ) 2>&1 | grep real
python tools/bench_summary.py gpurun_out/bench_ck2_driverlike_$TAG.json 2>/dev/null | head -5
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_ck2_$TAG.json 2>/dev/null ) 2>&1 | grep real
python -c "import json; d=json.load(open('gpurun_out/bench_ref_ck2_$TAG.json')); print('ref ck2', d['value'], d['cpu_baseline']['cores'])"
