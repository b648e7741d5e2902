#!/bin/bash
# 8-GPU call: multi-GPU parity, the driver-style bench at N = 8 (weak scaling, extras: workloads, configs[3]
# gather by NCCL / peer stores, configs[4] sweep), and configs[3] as headline lines.
set -u
mkdir -p gpurun_out
TAG=${1:-r02j}
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q -s > gpurun_out/pytest_multigpu_8_$TAG.log 2>&1; echo "pytest multigpu rc=$?"; tail -4 gpurun_out/pytest_multigpu_8_$TAG.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 1500 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_ck2_n8_$TAG.json 2> gpurun_out/bench_ck2_n8_$TAG.err ) 2>&1 | grep real; echo "bench n8 rc=$?"; tail -3 gpurun_out/bench_ck2_n8_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_n8_$TAG.json
for m in peer nccl; do
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --workload pm2 --envs 32768 --gather $m --steps 660 --warmup 66 > gpurun_out/gather_pm2_${m}_n8_$TAG.json 2> gpurun_out/gather_pm2_${m}_n8_$TAG.err; echo "gather $m rc=$?"; tail -2 gpurun_out/gather_pm2_${m}_n8_$TAG.err
python tools/bench_summary.py gpurun_out/gather_pm2_${m}_n8_$TAG.json
done
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --workload pm2 --envs 32768 --gather peer --no-overlap --steps 660 --warmup 66 > gpurun_out/gather_pm2_peer_noov_n8_$TAG.json 2> gpurun_out/gather_pm2_peer_noov_n8_$TAG.err; echo "gather peer no-overlap rc=$?"
python tools/bench_summary.py gpurun_out/gather_pm2_peer_noov_n8_$TAG.json
( time timeout 900 $TR --master-port 29524 bench.py --gpus 8 --steps 3300 --warmup 99 --no-extras > gpurun_out/scale_ck2_n8_long_$TAG.json 2> gpurun_out/scale_ck2_n8_long_$TAG.err ) 2>&1 | grep real
python tools/bench_summary.py gpurun_out/scale_ck2_n8_long_$TAG.json | head -3
