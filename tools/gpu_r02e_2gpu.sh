#!/bin/bash
# 2-GPU call: multi-GPU parity (NCCL / peer / double-buffered peer gather) and the driver-style bench at N = 2
set -u
mkdir -p gpurun_out
TAG=${1:-r02e}
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q -s > gpurun_out/pytest_multigpu_$TAG.log 2>&1; echo "pytest multigpu rc=$?"; tail -8 gpurun_out/pytest_multigpu_$TAG.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_ck2_n2_$TAG.json 2> gpurun_out/bench_ck2_n2_$TAG.err ) 2>&1 | tail -4; echo "bench n2 rc=$?"; tail -5 gpurun_out/bench_ck2_n2_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_n2_$TAG.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload pm2 --envs 32768 --gather peer --steps 330 --warmup 33 > gpurun_out/gather_pm2_peer_n2_$TAG.json 2> gpurun_out/gather_pm2_peer_n2_$TAG.err; echo "gather peer rc=$?"; tail -3 gpurun_out/gather_pm2_peer_n2_$TAG.err
python tools/bench_summary.py gpurun_out/gather_pm2_peer_n2_$TAG.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload pm2 --envs 32768 --gather peer --no-overlap --steps 330 --warmup 33 > gpurun_out/gather_pm2_peer_noov_n2_$TAG.json 2> gpurun_out/gather_pm2_peer_noov_n2_$TAG.err; echo "gather peer no-overlap rc=$?"
python tools/bench_summary.py gpurun_out/gather_pm2_peer_noov_n2_$TAG.json
