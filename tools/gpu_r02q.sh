#!/bin/bash
# r02q: where the chained per-step mode stands after early publish - the step time against the batch size
# (chain latency vs slot limit), ticket atomic consumed late (CM3_TICKET_LATE build), ncu of one chained launch.
set -u
mkdir -p gpurun_out
TAG=${1:-r02q}
AB=gpurun_out/ab_$TAG.txt
: > $AB
for B in 8192 16384 32768 65536 131072; do
python tools/ab_r02.py --tag "default envs=$B" --workloads pa4,pm2,ck1,ck2 --envs $B --modes per_step_chained >> $AB 2>gpurun_out/ab_$TAG.err
done
for B in 16384 32768; do
CM3_CHAIN_EARLY=0 python tools/ab_r02.py --tag "early=0 envs=$B" --workloads pa4,pm2 --envs $B --modes per_step_chained >> $AB 2>>gpurun_out/ab_$TAG.err
done
for rep in 1 2; do
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_ticketlate.so python tools/ab_r02.py --tag "ticket late" >> $AB 2>>gpurun_out/ab_$TAG.err
python tools/ab_r02.py --tag "ticket early (default)" >> $AB 2>>gpurun_out/ab_$TAG.err
done
cat $AB; tail -3 gpurun_out/ab_$TAG.err
CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/libcm3env_ticketlate.so timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_checkers.py -m gpu -x -q -k "chained or graph or rollout_equals" > gpurun_out/pytest_gpu_ticketlate_$TAG.log 2>&1; echo "pytest ticketlate rc=$?"; tail -2 gpurun_out/pytest_gpu_ticketlate_$TAG.log
for wl in pa4 pm2; do
ncu --set full --clock-control none --import-source on -k regex:particle_kernel -s 40 -c 1 -f -o gpurun_out/prof_${wl}_step_$TAG \
    python bench.py --workload $wl --mode step --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_${wl}_step_$TAG.log 2>&1; echo "ncu full step $wl rc=$?"
done
