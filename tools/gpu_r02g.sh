#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02g}
AB=gpurun_out/ab_$TAG.txt
: > $AB
for rep in 1 2; do
python tools/ab_r02.py --tag "elect-leader" --modes fused >> $AB 2>gpurun_out/ab_$TAG.err
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_lane0.so python tools/ab_r02.py --modes fused >> $AB 2>>gpurun_out/ab_$TAG.err
done
cat $AB; tail -3 gpurun_out/ab_$TAG.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
