#!/bin/bash
# One gpurun call for the round's evidence: GPU parity suite, bench line, ncu launch list, ncu --set full of the top
# kernels (with source, report kept) and of EVERY kernel of one step (summary + DRAM traffic only).
# usage (under gpurun): bash tools/gpu_round2.sh <tag> [skip-tests]
# The ncu passes run the batch as ONE stream (LPL_SPLIT=1): one launch per kernel and batch, as in the bench's own
# per-kernel table. Numbers printed under ncu are never bench values.
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -3 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; cut -c1-400 $OUT/${TAG}_bench.json
export LPL_SPLIT=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/profile_step.py --frames 154 --steps 2 > $OUT/${TAG}_ncu1.log 2>&1
python tools/ncu_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt
head -12 $OUT/${TAG}_launches_summary.txt
REGEX=$(python tools/ncu_launches.py $OUT/${TAG}_launches.csv --top-regex 6)
echo "full capture with source of: $REGEX"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$REGEX" -c 6 \
    -o $OUT/${TAG}_full -f python tools/profile_step.py --frames 154 --steps 1 > $OUT/${TAG}_ncu2.log 2>&1
# every kernel of one step (the first step of a context also clears the hash planes: same kernels)
timeout 1500 ncu --set full --clock-control none -o /tmp/${TAG}_all -f python tools/profile_step.py --frames 154 --steps 1 > $OUT/${TAG}_ncu3.log 2>&1
cp profiles/traffic.json $OUT/${TAG}_traffic.json 2>/dev/null
python tools/ncu_summary.py /tmp/${TAG}_all.ncu-rep --traffic $OUT/${TAG}_traffic.json > $OUT/${TAG}_ncu_all_summary.txt 2>&1
ls -la $OUT | tail -12
