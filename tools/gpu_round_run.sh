#!/bin/bash
# 1-GPU reference run of a round - parity suite, every bench line, reference
# arm, batch sweep, ncu launch lists + full captures (fused and per-step), compute-sanitizer.
set -u
mkdir -p gpurun_out
TAG=${1:-r01s}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
except Exception as e:
    print(f, "NO JSON", e); sys.exit(0)
r=d["roofline"]; x=d.get("extra",{})
s="%s n=%d %s value=%.4g us/step=%.3f frac=%.3f" % (f.split("/")[-1], d["n_gpus"], d["config"].get("mode"), d["value"], d["ms_per_step"]*1e3, r["frac"])
if "e2e" in d: s+=" e2e=%.4g" % d["e2e"]["value"]
if "per_step_launches" in x: s+=" step=%.4g(%.3f)" % (x["per_step_launches"]["value"], x["per_step_launches"]["frac"])
if "fused_rollout_T33_philox" in x: s+=" philox=%.4g(%.3f)" % (x["fused_rollout_T33_philox"]["value"], x["fused_rollout_T33_philox"]["frac"])
if "e2e_int8_tiles" in x: s+=" e2e_i8=%.4g" % x["e2e_int8_tiles"]["value"]
if "cpu_baseline" in d: s+=" cpu=%.4g" % d["cpu_baseline"]["value"]
print(s)
PY
}
for wl in ck2 pa4 pa3 pm2 ck1; do
  python bench.py --workload $wl > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; tail -3 gpurun_out/bench_${wl}_$TAG.err
  summ gpurun_out/bench_${wl}_$TAG.json
done
python bench.py --impl reference --steps 200 > gpurun_out/bench_ref_ck2_$TAG.json 2>&1; python -c "import json; d=json.load(open('gpurun_out/bench_ref_ck2_$TAG.json')); print('ref ck2', d['value'], d['cpu_baseline']['cores'])"
python bench.py --impl reference --workload pa4 --steps 200 > gpurun_out/bench_ref_pa4_$TAG.json 2>&1; python -c "import json; d=json.load(open('gpurun_out/bench_ref_pa4_$TAG.json')); print('ref pa4', d['value'], d['cpu_baseline']['cores'])"
python tools/sweep.py --out gpurun_out/sweep_$TAG.jsonl > gpurun_out/sweep_$TAG.log 2>&1; echo "sweep rc=$?"
for wl in ck2 pa4 pa3; do
  K=checkers_kernel; [ $wl != ck2 ] && K=particle_kernel
  # the default (fused) command: launch list + one full capture of the dominant kernel
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${wl}_fused_$TAG.csv \
      python bench.py --workload $wl --steps 330 --warmup 33 --no-extras > gpurun_out/ncu_launch_${wl}_fused_$TAG.log 2>&1; echo "ncu launches fused $wl rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/prof_${wl}_fused_$TAG \
      python bench.py --workload $wl --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_${wl}_fused_$TAG.log 2>&1; echo "ncu full fused $wl rc=$?"
  # per-step launches
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${wl}_step_$TAG.csv \
      python bench.py --workload $wl --mode step --steps 99 --warmup 3 --no-extras > gpurun_out/ncu_launch_${wl}_step_$TAG.log 2>&1; echo "ncu launches step $wl rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 2 -f -o gpurun_out/prof_${wl}_step_$TAG \
      python bench.py --workload $wl --mode step --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_${wl}_step_$TAG.log 2>&1; echo "ncu full step $wl rc=$?"
done
SEL="ragged or rollout_equals or masked or auto_reset or int8 or teacher_forced_f32_large or packed"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_checkers.py tests/test_gpu_particle.py -x -q \
    -k "$SEL" > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_$TAG.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_checkers.py tests/test_gpu_particle.py -x -q \
    -k "rollout_equals or ragged_batches" > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_$TAG.log
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
