#!/bin/bash
# GPU experiment K (1 GPU): swizzled-TMA particle kernel - parity (both store paths), bench lines,
# source-level ncu capture.
set -u
mkdir -p gpurun_out
TAG=${1:-r01k}
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
print(s)
PY
}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
CM3_TMA=0 python -m pytest tests/test_gpu_particle.py -m gpu -x -q > gpurun_out/pytest_gpu_notma_$TAG.log 2>&1; echo "pytest CM3_TMA=0 rc=$?"; tail -3 gpurun_out/pytest_gpu_notma_$TAG.log
for wl in pa4 pa3 pm2 ck2 ck1; do
  python bench.py --workload $wl --cpu-seconds 2 > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; tail -3 gpurun_out/bench_${wl}_$TAG.err
  summ gpurun_out/bench_${wl}_$TAG.json
done
for wl in pa4 pa3 pm2; do
CM3_TMA=0 python bench.py --workload $wl --no-extras > gpurun_out/bench_${wl}_notma_$TAG.json 2>/dev/null; summ gpurun_out/bench_${wl}_notma_$TAG.json
done
for wl in pa4 pa3; do
ncu --set full --clock-control none --import-source on -k regex:particle_kernel -s 4 -c 1 -f -o gpurun_out/prof_${wl}_fused_$TAG \
    python bench.py --workload $wl --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_${wl}_fused_$TAG.log 2>&1; echo "ncu full fused $wl rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:checkers_kernel -s 4 -c 1 -f -o gpurun_out/prof_ck2_fused_$TAG \
    python bench.py --workload ck2 --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_ck2_fused_$TAG.log 2>&1; echo "ncu full fused ck2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:checkers_kernel -s 40 -c 1 -f -o gpurun_out/prof_ck2_step_$TAG \
    python bench.py --workload ck2 --mode step --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_ck2_step_$TAG.log 2>&1; echo "ncu full step ck2 rc=$?"
