#!/bin/bash
# GPU experiment R2 (1 GPU): confirm pa3 with single-buffered staging + LDG/STS action stream.
set -u
mkdir -p gpurun_out
TAG=${1:-r01r2}
python tools/ab_variants.py pa4,pa3,pm2 2>&1 | tee -a gpurun_out/ab_$TAG.txt
python -m pytest tests/test_gpu_particle.py -m gpu -x -q 2>&1 | tail -2
