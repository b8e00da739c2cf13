#!/bin/bash
# One multi-GPU gpurun call: host-ceiling probe at 1/2/4/../N ranks + bench at the listed rank counts.
# usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> "<bench rank counts>" [bench args]
TAG=$1; N=$2; BN=$3; shift; shift; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
nproc > $OUT/${TAG}_host.txt; free -g | head -2 >> $OUT/${TAG}_host.txt; lscpu | grep -i "model name\|numa\|socket" >> $OUT/${TAG}_host.txt
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/h2d_probe.py 2>/dev/null | tail -1 | tee -a $OUT/${TAG}_h2d.jsonl
  fi
done
for n in $BN; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 "$@" 2> $OUT/${TAG}_bench_${n}gpu.err | tail -1 > $OUT/${TAG}_bench_${n}gpu.json
  echo "bench N=$n rc $?"; tail -2 $OUT/${TAG}_bench_${n}gpu.err
  python - <<PY
import json
r = json.loads(open("$OUT/${TAG}_bench_${n}gpu.json").read().strip().splitlines()[-1])
print("N=$n value", round(r["value"]), "e2e", round(r["e2e"]["value"]), "pcl16", round(r.get("e2e_pcl16", {}).get("value", 0)))
for k, v in (r.get("workloads") or {}).items():
    print("   ", k, round(v["value"]), "e2e", round(v.get("e2e", 0)))
PY
done
