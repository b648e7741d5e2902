#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02n}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_$TAG.log
python - <<'PY'
import bench, json
print(json.dumps({k: v for k, v in bench.measure_dropin_latency("cuda:0").items() if k != "note"}))
PY
( time timeout 900 python bench.py --no-sweep > gpurun_out/bench_ck2_$TAG.json 2> gpurun_out/bench_ck2_$TAG.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench_ck2_$TAG.err
python tools/bench_summary.py gpurun_out/bench_ck2_$TAG.json 2>/dev/null | head -16
for f in 0.5 0.9; do
echo "CM3_BALANCE_FRAC=$f"
for rep in 1 2; do
CM3_BALANCE_FRAC=$f python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('  K=20   frac %.3f us/step %.3f' % (d['roofline']['frac'], d['ms_per_step']*1e3))"
done
CM3_BALANCE_FRAC=$f python bench.py --steps 3300 --warmup 99 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('  K=3300 frac %.3f us/step %.3f' % (d['roofline']['frac'], d['ms_per_step']*1e3))"
done
