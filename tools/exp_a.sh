#!/bin/bash
# GPU experiment A (round 1, session 2): parity after the particle-kernel restructuring, then A/B of
# programmatic dependent launch and warps-per-block on the per-step launch path.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_expA.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_expA.log
run() { # name, env..., -- args
  local name=$1; shift
  env "$@" > /dev/null 2>&1 || true
}
ab() { # tag lib pdl workload
  local tag=$1 lib=$2 pdl=$3 wl=$4
  CM3ENV_LIBRARY=$lib CM3_PDL=$pdl python bench.py --workload $wl --no-extras --steps 3300 --warmup 99 \
      > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  echo "$tag rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/ab_$tag.json')); print('value=%.4g ms/step=%.5f frac=%.3f sm=%s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))" 2>&1 | tail -1)"
}
D=$PWD/cm3_b200/csrc
for rep in 1 2; do
ab ck2_wpb2_pdl1_$rep $D/libcm3env.so 1 ck2
ab ck2_wpb2_pdl0_$rep $D/libcm3env.so 0 ck2
ab ck2_wpb1_pdl1_$rep $D/libcm3env_wpb1.so 1 ck2
ab ck2_wpb1_pdl0_$rep $D/libcm3env_wpb1.so 0 ck2
ab ck2_wpb4_pdl1_$rep $D/libcm3env_wpb4.so 1 ck2
ab pa4_pdl1_$rep $D/libcm3env.so 1 pa4
ab pa4_pdl0_$rep $D/libcm3env.so 0 pa4
done
ab pa3_pdl1 $D/libcm3env.so 1 pa3
ab pm2_pdl1 $D/libcm3env.so 1 pm2
ab ck1_pdl1 $D/libcm3env.so 1 ck1
python bench.py --workload pa4 > gpurun_out/bench_pa4_expA.json 2> gpurun_out/bench_pa4_expA.err; echo "bench pa4 rc=$?"; cat gpurun_out/bench_pa4_expA.json
ncu --set full --clock-control none --import-source on -k regex:particle_kernel -s 20 -c 2 -f -o gpurun_out/prof_pa4_expA \
    python bench.py --workload pa4 --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_pa4_expA.log 2>&1; echo "ncu full pa4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:checkers_kernel -s 20 -c 2 -f -o gpurun_out/prof_ck2_expA \
    python bench.py --workload ck2 --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_ck2_expA.log 2>&1; echo "ncu full ck2 rc=$?"
ls -la gpurun_out | head -60
