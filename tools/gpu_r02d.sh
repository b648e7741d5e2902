#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02d}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
AB=gpurun_out/ab_$TAG.txt
python tools/ab_r02.py --tag "direct-stores(N<=2)" --workloads pm2,pa4,ck2 > $AB 2>gpurun_out/ab_$TAG.err
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_staged.so python tools/ab_r02.py --workloads pm2 >> $AB 2>>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag "direct-stores(N<=2) 262k" --workloads pm2 --envs 262144 >> $AB 2>>gpurun_out/ab_$TAG.err
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_staged.so python tools/ab_r02.py --workloads pm2 --envs 262144 >> $AB 2>>gpurun_out/ab_$TAG.err
cat $AB; tail -3 gpurun_out/ab_$TAG.err
ncu --set full --clock-control none --import-source on -k regex:particle_kernel -s 4 -c 1 -f -o gpurun_out/prof_pm2_fused_$TAG \
      python bench.py --workload pm2 --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_pm2_fused_$TAG.log 2>&1; echo "ncu full fused pm2 rc=$?"
