#!/bin/bash
# One multi-GPU gpurun call: host-ceiling probe + bench at N ranks. usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> [bench args]
TAG=$1; N=$2; shift; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
for n in $(seq 1 $N); do
  if [ $n = 1 ] || [ $n = 2 ] || [ $n = 4 ] || [ $n = 8 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/h2d_probe.py 2>/dev/null | tail -1 | tee -a $OUT/${TAG}_h2d.jsonl
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 "$@" > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
echo "bench rc $?"; tail -3 $OUT/${TAG}_bench_${N}gpu.err
python - <<PY
import json
r = json.loads(open("$OUT/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("value", round(r["value"]), "e2e", round(r["e2e"]["value"]), "pcl16", round(r.get("e2e_pcl16", {}).get("value", 0)))
for k, v in (r.get("workloads") or {}).items():
    print(" ", k, round(v["value"]), "e2e", round(v.get("e2e", 0)))
PY
