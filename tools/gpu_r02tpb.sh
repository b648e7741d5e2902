#!/bin/bash
# Chained Checkers launches with several tiles per block (loop around the tile body, barrier before the early state store):
# whole parity suite, graph-replay stress (rule / forced 2 / forced 4 tiles), then A/B against one tile per block.
set -u
mkdir -p gpurun_out
TAG=${1:-r02x4}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
python tools/stress_chained.py 5 > gpurun_out/stress_$TAG.log 2>&1; echo "stress rc=$? ok lines: $(grep -c ' ok' gpurun_out/stress_$TAG.log)"
for t in 2 4; do CM3_CHAIN_TPB=$t python tools/stress_chained.py 3 2>&1 | grep "ck" | grep -c " ok"; done
for e in 1 2; do CM3_CHAIN_EARLY=$e python tools/stress_chained.py 2 2>&1 | grep -c " ok"; done
AB=gpurun_out/ab_$TAG.txt
: > $AB
for rep in 1 2; do
for B in 65536 131072 262144; do
CM3_CHAIN_TPB=1 python tools/ab_r02.py --tag "one tile per block" --workloads ck2,ck1 --envs $B --modes per_step_chained >> $AB 2>/dev/null
python tools/ab_r02.py --tag "tiles per block: one wave" --workloads ck2,ck1 --envs $B --modes per_step_chained >> $AB 2>/dev/null
done
done
python tools/ab_r02.py --tag "fused (same kernel)" --workloads ck2,ck1 --modes fused >> $AB 2>/dev/null
python tools/ab_r02.py --tag "32768 envs" --workloads ck2,ck1 --envs 32768 --modes per_step_chained >> $AB 2>/dev/null
cat $AB
