#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02k}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
AB=gpurun_out/ab_$TAG.txt
: > $AB
for rep in 1 2; do
python tools/ab_r02.py --tag "duo(2 envs/thread)" --workloads pm2 >> $AB 2>gpurun_out/ab_$TAG.err
CM3_PT_DUO=0 python tools/ab_r02.py --tag "one env/thread" --workloads pm2 >> $AB 2>>gpurun_out/ab_$TAG.err
done
python tools/ab_r02.py --tag "duo 262k" --workloads pm2 --envs 262144 >> $AB 2>>gpurun_out/ab_$TAG.err
CM3_PT_DUO=0 python tools/ab_r02.py --tag "one env/thread 262k" --workloads pm2 --envs 262144 >> $AB 2>>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag "duo 16k" --workloads pm2 --envs 16384 >> $AB 2>>gpurun_out/ab_$TAG.err
CM3_PT_DUO=0 python tools/ab_r02.py --tag "one env/thread 16k" --workloads pm2 --envs 16384 >> $AB 2>>gpurun_out/ab_$TAG.err
CM3_BENCH_FUSED_T=99 python tools/ab_r02.py --tag "fused T=99" --workloads pa4,pa3,pm2,ck2,ck1 --modes fused >> $AB 2>>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag "fused T=33" --workloads pa4,pa3,ck2,ck1 --modes fused >> $AB 2>>gpurun_out/ab_$TAG.err
cat $AB; tail -3 gpurun_out/ab_$TAG.err
ncu --set full --clock-control none --import-source on -k regex:particle_duo_kernel -s 4 -c 1 -f -o gpurun_out/prof_pm2_fused_$TAG \
      python bench.py --workload pm2 --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_pm2_fused_$TAG.log 2>&1; echo "ncu full fused pm2 rc=$?"
