#!/bin/bash
# r02u: host thread / pinned buffers bound to the GPU's NUMA node (cm3_b200.sharding.bind_host_to_gpu): e2e A/B at 1 GPU
set -u
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -i "numa\|socket\|model name" | head -8
for b in 0 1 0 1; do
CM3_BIND_NUMA=$b python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('bind=$b e2e u2 %.4g (%.1f GB/s) int8 %.4g fp32 %.4g  affinity: %s' % (e['value'], e['d2h_gbs_per_gpu'], e['int8_tiles']['value'], e['unpipelined_fp32']['value'], d['launch_config']['host_affinity']))"
done
