#!/bin/bash
# Round-2 reference run on one B200: parity suite + smoke, every bench line (default, driver-style,
# per workload), the reference arm, ncu launch lists + full captures of the dominant kernels (fused and
# chained per-step), compute-sanitizer memcheck / racecheck.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02z}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv; nproc
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; grep -a "drift after" gpurun_out/pytest_gpu_$TAG.log | head -1; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
( time timeout 900 python bench.py > gpurun_out/bench_ck2_$TAG.json 2> gpurun_out/bench_ck2_$TAG.err ) 2>&1 | grep real; echo "bench ck2 rc=$?"; tail -3 gpurun_out/bench_ck2_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_$TAG.json
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ck2_driverlike_$TAG.json 2> gpurun_out/bench_ck2_driverlike_$TAG.err ) 2>&1 | grep real; echo "bench driverlike rc=$?"
python tools/bench_summary.py gpurun_out/bench_ck2_driverlike_$TAG.json | head -8
for wl in pa4 pa3 pm2 ck1; do
  timeout 600 python bench.py --workload $wl --no-extras > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_${wl}_$TAG.err
  python tools/bench_summary.py gpurun_out/bench_${wl}_$TAG.json | head -4
done
timeout 300 python bench.py --impl reference --steps 200 > gpurun_out/bench_ref_ck2_$TAG.json 2>&1; python -c "import json; d=json.load(open('gpurun_out/bench_ref_ck2_$TAG.json')); print('ref ck2', d['value'], d['cpu_baseline']['cores'])"
timeout 300 python bench.py --impl reference --workload pa4 --steps 200 > gpurun_out/bench_ref_pa4_$TAG.json 2>&1; python -c "import json; d=json.load(open('gpurun_out/bench_ref_pa4_$TAG.json')); print('ref pa4', d['value'], d['cpu_baseline']['cores'])"
for wl in ck2 pa4 pm2 ck1; do
  K=checkers_kernel; [ $wl = pa4 ] && K=particle_kernel; [ $wl = pm2 ] && K=particle_kernel
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${wl}_fused_$TAG.csv \
      python bench.py --workload $wl --steps 330 --warmup 33 --no-extras > gpurun_out/ncu_launch_${wl}_fused_$TAG.log 2>&1; echo "ncu launches fused $wl rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/prof_${wl}_fused_$TAG \
      python bench.py --workload $wl --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_${wl}_fused_$TAG.log 2>&1; echo "ncu full fused $wl rc=$?"
done
for wl in ck2 pa4; do
  K=checkers_kernel; [ $wl = pa4 ] && K=particle_kernel
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${wl}_step_$TAG.csv \
      python bench.py --workload $wl --mode step --steps 99 --warmup 3 --no-extras > gpurun_out/ncu_launch_${wl}_step_$TAG.log 2>&1; echo "ncu launches step $wl rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 1 -f -o gpurun_out/prof_${wl}_step_$TAG \
      python bench.py --workload $wl --mode step --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_${wl}_step_$TAG.log 2>&1; echo "ncu full step $wl rc=$?"
done
SEL="ragged or rollout_equals or masked or auto_reset or int8 or teacher_forced_f32_large or packed or chained or any_board or eight_agents or rollout_host or goal_redraw or pair_kernel"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_checkers.py tests/test_gpu_particle.py tests/test_gpu_round2.py -x -q \
    -k "$SEL" > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_$TAG.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_checkers.py tests/test_gpu_particle.py tests/test_gpu_round2.py -x -q \
    -k "rollout_equals or ragged_batches or chained or eight_agents or rollout_host" > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_$TAG.log
