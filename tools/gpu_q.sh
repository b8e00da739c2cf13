#!/bin/bash
# usage under gpurun: bash tools/gpu_q.sh <tag> [pytest -k expr]   -> GPU tests, quick device-resident rate of kitti154
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > $OUT/${TAG}_pytest.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
fi
tail -4 $OUT/${TAG}_pytest.log
for w in kitti154 cloud2m; do
LPL_WORKLOAD=$w timeout 300 python tools/kernel_times.py > $OUT/${TAG}_kt_$w.txt 2>&1
head -${3:-14} $OUT/${TAG}_kt_$w.txt
done
timeout 300 python tools/step_time.py 2>&1 | tee $OUT/${TAG}_step.txt
