#!/bin/bash
# GPU experiment F (8 GPUs of one box): weak scaling 1..8 in the headline mode, rollout all-gather
# (BASELINE configs[3]: PM2, 262 144 envs over 8 GPUs) by NCCL and by fused peer stores, the
# multi-GPU parity test, the batch sweep at 8 GPUs, the reference arm under torchrun.
set -u
mkdir -p gpurun_out
TAG=${1:-r01f}
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
python -m pytest tests/test_gpu_multigpu.py -x -q -s > gpurun_out/pytest_multigpu_$TAG.log 2>&1; echo "pytest multigpu rc=$?"; tail -4 gpurun_out/pytest_multigpu_$TAG.log
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
python bench.py --no-extras > gpurun_out/scale_ck2_n1_$TAG.json 2> gpurun_out/scale_ck2_n1_$TAG.err; summ gpurun_out/scale_ck2_n1_$TAG.json
python bench.py --workload pa4 --no-extras > gpurun_out/scale_pa4_n1_$TAG.json 2> gpurun_out/scale_pa4_n1_$TAG.err; summ gpurun_out/scale_pa4_n1_$TAG.json
n=2
while [ $n -le $NG ]; do
  $TR --nproc-per-node $n --master-port 29511 bench.py --gpus $n --no-extras > gpurun_out/scale_ck2_n${n}_$TAG.json 2> gpurun_out/scale_ck2_n${n}_$TAG.err; echo "scale ck2 n=$n rc=$?"; tail -2 gpurun_out/scale_ck2_n${n}_$TAG.err | cut -c1-300
  summ gpurun_out/scale_ck2_n${n}_$TAG.json
  $TR --nproc-per-node $n --master-port 29512 bench.py --gpus $n --workload pa4 --no-extras > gpurun_out/scale_pa4_n${n}_$TAG.json 2> gpurun_out/scale_pa4_n${n}_$TAG.err; echo "scale pa4 n=$n rc=$?"
  summ gpurun_out/scale_pa4_n${n}_$TAG.json
  for g in nccl peer; do
    $TR --nproc-per-node $n --master-port 29513 bench.py --gpus $n --workload pm2 --envs 32768 --gather $g --steps 3300 --warmup 99 \
        > gpurun_out/gather_pm2_${g}_n${n}_$TAG.json 2> gpurun_out/gather_pm2_${g}_n${n}_$TAG.err; echo "gather pm2 $g n=$n rc=$?"; tail -2 gpurun_out/gather_pm2_${g}_n${n}_$TAG.err | cut -c1-300
    summ gpurun_out/gather_pm2_${g}_n${n}_$TAG.json
  done
  n=$((n*2))
done
for g in nccl peer; do
  $TR --nproc-per-node $NG --master-port 29514 bench.py --gpus $NG --workload ck2 --envs 16384 --gather $g --steps 660 --warmup 33 \
      > gpurun_out/gather_ck2_${g}_n${NG}_$TAG.json 2> gpurun_out/gather_ck2_${g}_n${NG}_$TAG.err; echo "gather ck2 $g n=$NG rc=$?"; tail -2 gpurun_out/gather_ck2_${g}_n${NG}_$TAG.err | cut -c1-300
  summ gpurun_out/gather_ck2_${g}_n${NG}_$TAG.json
done
$TR --nproc-per-node $NG --master-port 29516 tools/sweep.py --out gpurun_out/sweep_n${NG}_$TAG.jsonl > gpurun_out/sweep_n${NG}_$TAG.log 2>&1; echo "sweep n=$NG rc=$?"; tail -3 gpurun_out/sweep_n${NG}_$TAG.log | cut -c1-250
$TR --nproc-per-node $NG --master-port 29515 bench.py --impl reference --gpus $NG --steps 100 > gpurun_out/ref_n${NG}_$TAG.json 2> gpurun_out/ref_n${NG}_$TAG.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/ref_n${NG}_$TAG.json') if l.startswith('{')][-1]); print('ref', d['value'], d['cpu_baseline']['cores'], d['config']['sample'])"
python bench.py --impl reference --workload pa4 --steps 100 > gpurun_out/bench_ref_pa4_$TAG.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_ref_pa4_$TAG.json')); print('ref pa4', d['value'], d['cpu_baseline']['cores'])"
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1; head -12 gpurun_out/topo_$TAG.txt; nproc
