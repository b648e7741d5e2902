#!/bin/bash
# GPU experiment N (1 GPU): FULL-specialised particle kernel - parity with both instantiations, A/B timing.
set -u
mkdir -p gpurun_out
TAG=${1:-r01n}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
CM3_PT_FULL=0 python -m pytest tests/test_gpu_particle.py -m gpu -x -q > gpurun_out/pytest_gpu_nofull_$TAG.log 2>&1; echo "pytest CM3_PT_FULL=0 rc=$?"; tail -3 gpurun_out/pytest_gpu_nofull_$TAG.log
python tools/ab_variants.py pa4,pa3,pm2 2>&1 | sed 's/^/full   /' | tee -a gpurun_out/ab_$TAG.txt
CM3_PT_FULL=0 python tools/ab_variants.py pa4,pa3,pm2 2>&1 | sed 's/^/nofull /' | tee -a gpurun_out/ab_$TAG.txt
CM3_TMA=0 python tools/ab_variants.py pa4,pa3,pm2 2>&1 | sed 's/^/notma  /' | tee -a gpurun_out/ab_$TAG.txt
python tools/ab_variants.py ck2,ck1 2>&1 | tee -a gpurun_out/ab_$TAG.txt
for wl in pa4 pa3 ck2; do
  python bench.py --workload $wl --mode step --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step $wl', d['value'], d['ms_per_step']*1e3, d['roofline']['frac'])"
done
for wl in pa4 pa3; do
ncu --set full --clock-control none --import-source on -k regex:particle_kernel -s 4 -c 1 -f -o gpurun_out/prof_${wl}_fused_$TAG \
    python bench.py --workload $wl --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_${wl}_fused_$TAG.log 2>&1; echo "ncu full fused $wl rc=$?"
done
