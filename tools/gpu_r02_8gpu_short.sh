#!/bin/bash
# 8-GPU call on the final kernels: multi-GPU parity, the driver-style bench at N = 8 (extras: workloads, configs[3]
# gather, configs[4] sweep) and the long weak-scaling line.
set -u
mkdir -p gpurun_out
TAG=${1:-r02v}
timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q -s > gpurun_out/pytest_multigpu_8_$TAG.log 2>&1; echo "pytest multigpu rc=$?"; tail -4 gpurun_out/pytest_multigpu_8_$TAG.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 1500 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_ck2_n8_$TAG.json 2> gpurun_out/bench_ck2_n8_$TAG.err ) 2>&1 | grep real; echo "bench n8 rc=$?"; tail -3 gpurun_out/bench_ck2_n8_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_n8_$TAG.json | head -20
( time timeout 900 $TR --master-port 29524 bench.py --gpus 8 --steps 3300 --warmup 99 --no-extras > gpurun_out/scale_ck2_n8_long_$TAG.json 2> gpurun_out/scale_ck2_n8_long_$TAG.err ) 2>&1 | grep real
python tools/bench_summary.py gpurun_out/scale_ck2_n8_long_$TAG.json | head -3
( time timeout 600 $TR --master-port 29525 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_ref_ck2_n8_$TAG.json 2>/dev/null ) 2>&1 | grep real
python -c "import json; d=json.load(open('gpurun_out/bench_ref_ck2_n8_$TAG.json')); print('ref n8', d['value'], d['n_gpus'], d['cpu_baseline']['cores'])"
