#!/bin/bash
# round 2, second GPU call: parity of the fixed / new paths, A/B of the round-2 switches, ncu of the laggards
set -u
mkdir -p gpurun_out
TAG=${1:-r02b}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log
AB=gpurun_out/ab_$TAG.txt
python tools/ab_r02.py --tag default > $AB 2>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag cutoff0 --cutoff0 --workloads pa4,pa3,pm2 >> $AB 2>>gpurun_out/ab_$TAG.err
CM3_BALANCE=0 python tools/ab_r02.py --tag balance0 --workloads ck1,ck2,pa4 >> $AB 2>>gpurun_out/ab_$TAG.err
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_minboff.so python tools/ab_r02.py --workloads pa4,pa3,pm2 >> $AB 2>>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag default262k --envs 262144 --workloads pa4,pm2,ck1 >> $AB 2>>gpurun_out/ab_$TAG.err
cat $AB; tail -3 gpurun_out/ab_$TAG.err
python - <<'PY' > gpurun_out/dropin_$TAG.txt 2>&1
import bench, json
print(json.dumps(bench.measure_dropin_latency("cuda:0"), indent=1))
PY
cat gpurun_out/dropin_$TAG.txt
for wl in pm2 ck1 pa3; do
  K=particle_kernel; [ $wl = ck1 ] && K=checkers_kernel
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/prof_${wl}_fused_$TAG \
      python bench.py --workload $wl --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_${wl}_fused_$TAG.log 2>&1; echo "ncu full fused $wl rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:particle_kernel -s 40 -c 1 -f -o gpurun_out/prof_pa4_step_$TAG \
      python bench.py --workload pa4 --mode step --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_pa4_step_$TAG.log 2>&1; echo "ncu full step pa4 rc=$?"
ls -la gpurun_out/*.ncu-rep
