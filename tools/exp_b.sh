#!/bin/bash
# GPU experiment B: full parity suite (incl. drop-in facades), bench lines with the restructured
# kernels, batch sweep, ncu launch lists + full captures, compute-sanitizer on a small case.
set -u
mkdir -p gpurun_out
TAG=${1:-r01b}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
for wl in ck2 pa4; do
  python bench.py --workload $wl > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_$TAG.json"))
print("$wl value=%.4g us/step=%.3f frac=%.3f e2e=%.4g fused=%.4g (%.3f) cpu=%.4g" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"], d["e2e"]["value"], d["extra"]["fused_rollout_T33"]["value"], d["extra"]["fused_rollout_T33"]["frac"], d["cpu_baseline"]["value"]))
PY
done
for wl in pa3 pm2 ck1; do
  python bench.py --workload $wl --no-extras > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
  python -c "import json; d=json.load(open('gpurun_out/bench_${wl}_$TAG.json')); print('$wl value=%.4g us/step=%.3f frac=%.3f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac']))"
done
python bench.py --impl reference --steps 200 > gpurun_out/bench_ref_ck2_$TAG.json 2>&1; tail -c 600 gpurun_out/bench_ref_ck2_$TAG.json; echo
python tools/sweep.py --out gpurun_out/sweep_$TAG.jsonl > gpurun_out/sweep_$TAG.log 2>&1; echo "sweep rc=$?"
python - <<PY
import json
for l in open("gpurun_out/sweep_$TAG.jsonl"):
    d=json.loads(l); print("%s B=%8d step %.3g (%.3f) %.2fus | fused %.3g (%.3f)" % (d["workload"], d["envs_per_gpu"], d["per_step"]["agent_env_steps_per_s"], d["per_step"]["frac"], d["per_step"]["us_per_step"], d["fused_T33"]["agent_env_steps_per_s"], d["fused_T33"]["frac"]))
PY
for wl in ck2 pa4; do
  K=checkers_kernel; [ $wl = pa4 ] && K=particle_kernel
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${wl}_$TAG.csv \
      python bench.py --workload $wl --steps 99 --warmup 3 --no-extras > gpurun_out/ncu_launch_${wl}_$TAG.log 2>&1; echo "ncu launches $wl rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 2 -f -o gpurun_out/prof_${wl}_$TAG \
      python bench.py --workload $wl --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_${wl}_$TAG.log 2>&1; echo "ncu full $wl rc=$?"
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_checkers.py tests/test_gpu_particle.py -x -q \
    -k "ragged or rollout_equals or masked or auto_reset" > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_$TAG.log
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
