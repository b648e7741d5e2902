#!/bin/bash
# r02t: register bounds (20 / 24 resident blocks per SM) for the particle kernels again, now that chained launches
# publish their tile before the output phase - do two co-resident grids pipeline?
set -u
mkdir -p gpurun_out
TAG=${1:-r02t}
AB=gpurun_out/ab_$TAG.txt
: > $AB
for rep in 1 2; do
for m in 20 24; do
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_minb$m.so python tools/ab_r02.py --tag "min blocks/SM $m" --workloads pa4,pa3,pm2 >> $AB 2>gpurun_out/ab_$TAG.err
done
python tools/ab_r02.py --tag "default" --workloads pa4,pa3,pm2 >> $AB 2>>gpurun_out/ab_$TAG.err
done
cat $AB; tail -3 gpurun_out/ab_$TAG.err
