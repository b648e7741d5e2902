#!/bin/bash
# GPU experiment O2 (1 GPU): evict-last L2 prefetch of the action rows.
set -u
mkdir -p gpurun_out
TAG=${1:-r01o2}
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_pf.so timeout 120 python tools/ab_variants.py ck2,pa4,pa3 2>&1 | tail -4 | tee -a gpurun_out/ab_$TAG.txt
python tools/ab_variants.py ck2,pa4,pa3 2>&1 | tail -3 | tee -a gpurun_out/ab_$TAG.txt
python -m pytest tests/test_gpu_particle.py -m gpu -x -q 2>&1 | tail -2
