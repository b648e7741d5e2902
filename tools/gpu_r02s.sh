#!/bin/bash
# r02s: is the per-step line of short steps (PM2, ~4 us per step) host-limited inside bench.py?  Spin length and clock sampler A/B.
set -u
mkdir -p gpurun_out
TAG=${1:-r02s}
for spin in 4000000 16000000 4000000 16000000; do
CM3_BENCH_SPIN=$spin python bench.py --workload pm2 --mode step --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('spin=$spin pm2 step K=%d us/step %.3f frac %.3f' % (d['steps'], d['ms_per_step']*1e3, d['roofline']['frac']))"
done
for K in 660 3300; do
python bench.py --workload pm2 --mode step --steps $K --warmup 33 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('K=$K pm2 step us/step %.3f frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))"
done
python tools/ab_r02.py --tag "ab K=660" --workloads pm2 --modes per_step_chained
