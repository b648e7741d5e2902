#!/bin/bash
# GPU experiment C: thread-per-env particle kernel (parity + throughput), int8 tiles.
set -u
mkdir -p gpurun_out
TAG=${1:-r01c}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
for wl in pa4 pa3 pm2; do
  python bench.py --workload $wl > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; tail -3 gpurun_out/bench_${wl}_$TAG.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_$TAG.json"))
print("$wl value=%.4g us/step=%.3f frac=%.3f e2e=%.4g fused=%.4g (%.3f) cpu=%.4g" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"], d["e2e"]["value"], d["extra"]["fused_rollout_T33"]["value"], d["extra"]["fused_rollout_T33"]["frac"], d["cpu_baseline"]["value"]))
PY
done
python bench.py --workload ck2 > gpurun_out/bench_ck2_$TAG.json 2> gpurun_out/bench_ck2_$TAG.err; echo "bench ck2 rc=$?"; tail -3 gpurun_out/bench_ck2_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_ck2_$TAG.json"))
print("ck2 value=%.4g us/step=%.3f frac=%.3f e2e=%.4g e2e_i8=%.4g fused=%.4g (%.3f)" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"], d["e2e"]["value"], d["extra"]["e2e_int8_tiles"]["value"], d["extra"]["fused_rollout_T33"]["value"], d["extra"]["fused_rollout_T33"]["frac"]))
PY
python tools/sweep.py --workloads pa4 --out gpurun_out/sweep_pa4_$TAG.jsonl > gpurun_out/sweep_pa4_$TAG.log 2>&1; echo "sweep rc=$?"
python - <<PY
import json
for l in open("gpurun_out/sweep_pa4_$TAG.jsonl"):
    d=json.loads(l); print("%s B=%8d step %.3g (%.3f) %.2fus | fused %.3g (%.3f)" % (d["workload"], d["envs_per_gpu"], d["per_step"]["agent_env_steps_per_s"], d["per_step"]["frac"], d["per_step"]["us_per_step"], d["fused_T33"]["agent_env_steps_per_s"], d["fused_T33"]["frac"]))
PY
ncu --set full --clock-control none --import-source on -k regex:particle_kernel -s 20 -c 2 -f -o gpurun_out/prof_pa4_$TAG \
    python bench.py --workload pa4 --steps 66 --warmup 3 --no-extras > gpurun_out/ncu_full_pa4_$TAG.log 2>&1; echo "ncu full pa4 rc=$?"
