#!/bin/bash
# One short gpurun call: GPU parity suite + one bench line (no ncu).
# usage (under gpurun): bash tools/gpu_quick.sh <tag> [pytest -k expression]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > $OUT/${TAG}_pytest.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
fi
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
LPL_BENCH_KERNELS=$OUT/${TAG}_kernels.json timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    r = json.load(open("$OUT/${TAG}_bench.json"))
    print("value", round(r["value"]), "e2e", round(r["e2e"]["value"]), "ms/step", round(r["ms_per_step"], 3))
    ks = json.load(open("$OUT/${TAG}_kernels.json"))
    for k in ks:
        print(f'  {k["kernel"]:22s} {k["ms_per_launch"]:.3f} ms x{k["launches"]}  share {k["share"]:.3f}')
except Exception as e:
    print("no bench json:", e)
PY
