#!/bin/bash
# GPU experiment J (8 GPUs of one box, kept short: box time is charged per GPU): weak scaling of
# the headline at N = 8, BASELINE configs[3] (PM2, 262 144 envs over 8 GPUs, rollout all-gather by
# NCCL and by fused peer stores), the multi-GPU parity test.
set -u
mkdir -p gpurun_out
TAG=${1:-r01j}
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
except Exception as e:
    print(f, "NO JSON", e); sys.exit(0)
r=d["roofline"]
s="%s n=%d %s value=%.4g us/step=%.3f frac=%.3f (%s)" % (f.split("/")[-1], d["n_gpus"], d["config"].get("mode"), d["value"], d["ms_per_step"]*1e3, r["frac"], r["bound"])
if "e2e" in d: s+=" e2e=%.4g" % d["e2e"]["value"]
if "nvlink_gbs_per_gpu" in r: s+=" gather=%s nvlink=%.1f GB/s/GPU" % (d["config"].get("rollout_all_gather"), r["nvlink_gbs_per_gpu"])
print(s)
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multigpu.py -x -q -s > gpurun_out/pytest_multigpu_$TAG.log 2>&1; echo "pytest multigpu rc=$?"; tail -4 gpurun_out/pytest_multigpu_$TAG.log
for g in nccl peer; do
  timeout 300 $TR --nproc-per-node $NG --master-port 29513 bench.py --gpus $NG --workload pm2 --envs 32768 --gather $g --steps 3300 --warmup 99 \
      > gpurun_out/gather_pm2_${g}_n${NG}_$TAG.json 2> gpurun_out/gather_pm2_${g}_n${NG}_$TAG.err; echo "gather pm2 $g n=$NG rc=$?"; tail -2 gpurun_out/gather_pm2_${g}_n${NG}_$TAG.err | cut -c1-300
  summ gpurun_out/gather_pm2_${g}_n${NG}_$TAG.json
done
# one weak-scaling line of the headline (the driver measures 1/2/4/8 itself at round end)
timeout 300 $TR --nproc-per-node $NG --master-port 29511 bench.py --gpus $NG --no-extras > gpurun_out/scale_ck2_n${NG}_$TAG.json 2> gpurun_out/scale_ck2_n${NG}_$TAG.err; echo "scale ck2 n=$NG rc=$?"; tail -2 gpurun_out/scale_ck2_n${NG}_$TAG.err | cut -c1-300
summ gpurun_out/scale_ck2_n${NG}_$TAG.json
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1; head -12 gpurun_out/topo_$TAG.txt; nproc
