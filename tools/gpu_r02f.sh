#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02f}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
AB=gpurun_out/ab_$TAG.txt
python tools/ab_r02.py --tag "elect-leader" > $AB 2>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag "elect-leader" >> $AB 2>>gpurun_out/ab_$TAG.err
cat $AB; tail -3 gpurun_out/ab_$TAG.err
