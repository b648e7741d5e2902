#!/bin/bash
# GPU experiment G (1 GPU): source-level ncu captures of the particle kernel (fused, PA4 and PA3)
# and of the per-step kernels, plus the per-step timing bound with the inter-launch dependency
# removed (libcm3env_nowait.so: griddepcontrol.wait compiled out - racy results, timing only).
set -u
mkdir -p gpurun_out
TAG=${1:-r01g}
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
except Exception as e:
    print(f, "NO JSON", e); sys.exit(0)
r=d["roofline"]
print("%s %s value=%.4g us/step=%.3f frac=%.3f" % (f.split("/")[-1], d["config"].get("mode"), d["value"], d["ms_per_step"]*1e3, r["frac"]))
PY
}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
for wl in ck2 pa4; do
  for lib in libcm3env.so libcm3env_nowait.so; do
    CM3ENV_LIBRARY=$PWD/cm3_b200/csrc/$lib python bench.py --workload $wl --mode step --no-extras > gpurun_out/step_${wl}_${lib%.so}_$TAG.json 2> gpurun_out/step_${wl}_${lib%.so}_$TAG.err
    echo $lib; summ gpurun_out/step_${wl}_${lib%.so}_$TAG.json
  done
done
for wl in pa4 pa3; do
  ncu --set full --clock-control none --import-source on -k regex:particle_kernel -s 4 -c 1 -f -o gpurun_out/prof_${wl}_fused_$TAG \
      python bench.py --workload $wl --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_${wl}_fused_$TAG.log 2>&1; echo "ncu full fused $wl rc=$?"
done
for wl in ck2 pa4; do
  K=checkers_kernel; [ $wl = pa4 ] && K=particle_kernel
  ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 2 -f -o gpurun_out/prof_${wl}_step_$TAG \
      python bench.py --workload $wl --mode step --steps 99 --warmup 33 --no-extras > gpurun_out/ncu_full_${wl}_step_$TAG.log 2>&1; echo "ncu full step $wl rc=$?"
done
ls -la gpurun_out | head -40
