#!/bin/bash
# round 2, third GPU call: parity incl. tightened particle tests, A/B of chained partial grids and the contact cutoff
set -u
mkdir -p gpurun_out
TAG=${1:-r02c}
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; grep -a "drift after" gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
AB=gpurun_out/ab_$TAG.txt
python tools/ab_r02.py --tag "default(parts=2)" > $AB 2>gpurun_out/ab_$TAG.err
CM3_CHAIN_PARTS=1 python tools/ab_r02.py --tag parts1 --modes per_step_chained >> $AB 2>>gpurun_out/ab_$TAG.err
CM3_CHAIN_PARTS=4 python tools/ab_r02.py --tag parts4 --modes per_step_chained >> $AB 2>>gpurun_out/ab_$TAG.err
CM3_CHAIN_PARTS=3 python tools/ab_r02.py --tag parts3 --modes per_step_chained --workloads pa4,pa3,pm2 >> $AB 2>>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag cutoff0 --cutoff0 --workloads pa4,pa3,pm2 --modes fused >> $AB 2>>gpurun_out/ab_$TAG.err
cat $AB; tail -3 gpurun_out/ab_$TAG.err
python - <<'PY' > gpurun_out/dropin_$TAG.txt 2>&1
import bench, json
print(json.dumps(bench.measure_dropin_latency("cuda:0"), indent=1))
PY
cat gpurun_out/dropin_$TAG.txt
