#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02l}
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "duo or pair or chained" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
AB=gpurun_out/ab_$TAG.txt
: > $AB
for rep in 1 2; do
python tools/ab_r02.py --tag "duo phase-wise" --workloads pm2 >> $AB 2>gpurun_out/ab_$TAG.err
CM3_PT_DUO=0 python tools/ab_r02.py --tag "one env/thread" --workloads pm2 >> $AB 2>>gpurun_out/ab_$TAG.err
done
python tools/ab_r02.py --tag "duo phase-wise 262k" --workloads pm2 --envs 262144 >> $AB 2>>gpurun_out/ab_$TAG.err
cat $AB; tail -3 gpurun_out/ab_$TAG.err
