#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02m}
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/pytest_dropin_$TAG.log 2>&1; echo "pytest dropin rc=$?"; tail -3 gpurun_out/pytest_dropin_$TAG.log
python - <<'PY'
import bench, json
print(json.dumps({k: v for k, v in bench.measure_dropin_latency("cuda:0").items() if k != "note"}))
PY
for f in 0.5 0.9; do
echo "CM3_BALANCE_FRAC=$f"
for rep in 1 2; do
CM3_BALANCE_FRAC=$f python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('  K=20   frac %.3f us/step %.3f' % (d['roofline']['frac'], d['ms_per_step']*1e3))"
done
CM3_BALANCE_FRAC=$f python bench.py --steps 3300 --warmup 99 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('  K=3300 frac %.3f us/step %.3f' % (d['roofline']['frac'], d['ms_per_step']*1e3))"
done
