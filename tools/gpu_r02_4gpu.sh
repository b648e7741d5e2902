#!/bin/bash
# 4-GPU call: the driver-style bench at N = 4 (extras included) and the reference arm under torchrun.
set -u
mkdir -p gpurun_out
TAG=${1:-r02v}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( time timeout 1200 $TR --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_ck2_n4_$TAG.json 2> gpurun_out/bench_ck2_n4_$TAG.err ) 2>&1 | grep real; echo "bench n4 rc=$?"; tail -3 gpurun_out/bench_ck2_n4_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_n4_$TAG.json 2>/dev/null | head -18
( time timeout 600 $TR --master-port 29532 bench.py --impl reference --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_ref_ck2_n4_$TAG.json 2>/dev/null ) 2>&1 | grep real
python -c "import json; d=json.load(open('gpurun_out/bench_ref_ck2_n4_$TAG.json')); print('ref n4', d['value'], d['n_gpus'], d['cpu_baseline']['cores'])"
