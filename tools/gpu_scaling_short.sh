#!/bin/bash
# Weak-scaling lines of the headline (CK2) and of PA3 at N = 2 .. the GPUs of this box (fused mode,
# no extras); N = 1 and N = 8 come from tools/gpu_round_run.sh / tools/gpu_8gpu_short.sh.
set -u
mkdir -p gpurun_out
TAG=${1:-r01t}
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
n=2
while [ $n -le $NG ]; do
  for wl in ck2 pa3; do
    timeout 200 $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --workload $wl --no-extras > gpurun_out/scale_${wl}_n${n}_$TAG.json 2> gpurun_out/scale_${wl}_n${n}_$TAG.err
    python - gpurun_out/scale_${wl}_n${n}_$TAG.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("%s n=%d value=%.4g us/step=%.3f frac=%.3f e2e=%.4g" % (sys.argv[1].split("/")[-1], d["n_gpus"], d["value"], d["ms_per_step"]*1e3, d["roofline"]["frac"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "NO JSON", e)
PY
  done
  n=$((n*2))
done
