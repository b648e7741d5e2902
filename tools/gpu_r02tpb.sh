#!/bin/bash
# Chained Checkers launches with two tiles per block (one resident wave): parity + stress, then A/B against one tile per block.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_rollout_adapter.py tests/test_gpu_checkers.py -m gpu -x -q > gpurun_out/pytest_gpu_tpb.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_tpb.log
CM3_CHAIN_TPB=2 timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "chained or graph" 2>&1 | tail -1
python tools/stress_chained.py 5 2>&1 | tail -21
CM3_CHAIN_TPB=2 python tools/stress_chained.py 3 2>&1 | grep -v " ok" | tail -5; echo "forced tpb=2 stress done"
AB=gpurun_out/ab_tpb.txt
: > $AB
for rep in 1 2; do
for B in 65536 131072; do
CM3_CHAIN_TPB=1 python tools/ab_r02.py --tag "one tile per block" --workloads ck2,ck1 --envs $B --modes per_step_chained >> $AB 2>/dev/null
python tools/ab_r02.py --tag "two tiles per block when one wave" --workloads ck2,ck1 --envs $B --modes per_step_chained >> $AB 2>/dev/null
done
done
python tools/ab_r02.py --tag "fused (same kernel)" --workloads ck2,ck1 --modes fused >> $AB 2>/dev/null
cat $AB
