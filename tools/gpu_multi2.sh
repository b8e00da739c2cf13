#!/bin/bash
# bench at several rank counts on ONE multi-GPU box (no host probes); usage (under gpurun --gpus N): bash tools/gpu_multi2.sh <tag> "<rank counts>"
TAG=$1; BN=$2
OUT=gpurun_out; mkdir -p $OUT
for n in $BN; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 2> $OUT/${TAG}_bench_${n}gpu.err | tail -1 > $OUT/${TAG}_bench_${n}gpu.json
  echo "bench N=$n rc $?"; tail -2 $OUT/${TAG}_bench_${n}gpu.err
  python - <<PY
import json
try:
    r = json.loads(open("$OUT/${TAG}_bench_${n}gpu.json").read().strip().splitlines()[-1])
    print("N=$n value", round(r["value"]), "e2e", round(r["e2e"]["value"]), "pcl16", round(r.get("e2e_pcl16", {}).get("value", 0)))
    for k, v in (r.get("workloads") or {}).items():
        print("   ", k, round(v["value"]), "e2e", round(v.get("e2e", 0)))
except Exception as e:
    print("no line:", e)
PY
done
