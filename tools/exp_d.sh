#!/bin/bash
# GPU experiment D (N GPUs of one box): full parity suite incl. the multi-GPU test, bench lines in
# the fused headline mode, weak scaling 1..N, rollout all-gather by NCCL and by fused peer stores.
set -u
mkdir -p gpurun_out
TAG=${1:-r01d}
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
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
if "e2e_int8_tiles" in x: s+=" e2e_i8=%.4g" % x["e2e_int8_tiles"]["value"]
if "nvlink_gbs_per_gpu" in r: s+=" gather=%s nvlink=%.1f GB/s/GPU" % (d["config"].get("rollout_all_gather"), r["nvlink_gbs_per_gpu"])
print(s)
PY
}
for wl in ck2 pa4; do
  python bench.py --workload $wl > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; tail -3 gpurun_out/bench_${wl}_$TAG.err
  summ gpurun_out/bench_${wl}_$TAG.json
done
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
n=2
while [ $n -le $NG ]; do
  $TR --nproc-per-node $n --master-port 29511 bench.py --gpus $n --no-extras > gpurun_out/scale_ck2_n${n}_$TAG.json 2> gpurun_out/scale_ck2_n${n}_$TAG.err; echo "scale ck2 n=$n rc=$?"; tail -2 gpurun_out/scale_ck2_n${n}_$TAG.err
  summ gpurun_out/scale_ck2_n${n}_$TAG.json
  $TR --nproc-per-node $n --master-port 29512 bench.py --gpus $n --workload pa4 --no-extras > gpurun_out/scale_pa4_n${n}_$TAG.json 2> gpurun_out/scale_pa4_n${n}_$TAG.err; echo "scale pa4 n=$n rc=$?"
  summ gpurun_out/scale_pa4_n${n}_$TAG.json
  for g in nccl peer; do
    $TR --nproc-per-node $n --master-port 29513 bench.py --gpus $n --workload pm2 --envs 32768 --gather $g --steps 3300 --warmup 99 \
        > gpurun_out/gather_pm2_${g}_n${n}_$TAG.json 2> gpurun_out/gather_pm2_${g}_n${n}_$TAG.err; echo "gather pm2 $g n=$n rc=$?"; tail -2 gpurun_out/gather_pm2_${g}_n${n}_$TAG.err
    summ gpurun_out/gather_pm2_${g}_n${n}_$TAG.json
    $TR --nproc-per-node $n --master-port 29514 bench.py --gpus $n --workload ck2 --envs 16384 --gather $g --steps 660 --warmup 33 \
        > gpurun_out/gather_ck2_${g}_n${n}_$TAG.json 2> gpurun_out/gather_ck2_${g}_n${n}_$TAG.err; echo "gather ck2 $g n=$n rc=$?"; tail -2 gpurun_out/gather_ck2_${g}_n${n}_$TAG.err
    summ gpurun_out/gather_ck2_${g}_n${n}_$TAG.json
  done
  n=$((n*2))
done
$TR --nproc-per-node $NG --master-port 29515 bench.py --impl reference --gpus $NG --steps 100 > gpurun_out/ref_n${NG}_$TAG.json 2>&1; tail -c 300 gpurun_out/ref_n${NG}_$TAG.json; echo
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1; head -12 gpurun_out/topo_$TAG.txt
