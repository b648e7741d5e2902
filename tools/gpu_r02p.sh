#!/bin/bash
# r02p: branch-free logaddexp0 (A/B against the literal three-way branch), early state store / publish of
# chained launches (CM3_CHAIN_EARLY=0|1|2), warm-up re-issued behind the spin in bench.py's timed region.
set -u
mkdir -p gpurun_out
TAG=${1:-r02p}
timeout 1200 python -m pytest tests/test_gpu_particle.py tests/test_gpu_round2.py tests/test_gpu_rollout_adapter.py -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
for m in 1 2; do
CM3_CHAIN_EARLY=$m timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_rollout_adapter.py tests/test_gpu_checkers.py -m gpu -x -q -k "chained or graph or collector or evaluate" > gpurun_out/pytest_gpu_early${m}_$TAG.log 2>&1; echo "pytest early=$m rc=$?"; tail -2 gpurun_out/pytest_gpu_early${m}_$TAG.log
done
AB=gpurun_out/ab_$TAG.txt
: > $AB
for rep in 1 2; do
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_laebranchy.so python tools/ab_r02.py --tag "logaddexp branchy" --workloads pm2,pa3,pa4 --modes fused >> $AB 2>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag "logaddexp select" --workloads pm2,pa3,pa4 --modes fused >> $AB 2>>gpurun_out/ab_$TAG.err
done
for rep in 1 2; do
for m in 0 1 2; do
CM3_CHAIN_EARLY=$m python tools/ab_r02.py --tag "chain_early=$m" --modes per_step_chained >> $AB 2>>gpurun_out/ab_$TAG.err
done
done
cat $AB; tail -3 gpurun_out/ab_$TAG.err
for rw in 0 1 0 1; do
CM3_BENCH_REWARM=$rw python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('rewarm=$rw K=20 frac %.3f us/step %.3f' % (d['roofline']['frac'], d['ms_per_step']*1e3))" | tee -a $AB
done
