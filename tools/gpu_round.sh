#!/bin/bash
# One gpurun call: GPU parity suite, bench line, ncu launch list, ncu --set full of the top kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
# Everything lands in gpurun_out/<tag>_*; numbers printed under ncu are never bench values.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -3 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; cut -c1-600 $OUT/${TAG}_bench.json
# launch list: one warm step skipped by -s is not possible without the count, so profile both steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/profile_step.py --frames 154 --steps 2 > $OUT/${TAG}_ncu1.log 2>&1
python tools/ncu_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt
head -40 $OUT/${TAG}_launches_summary.txt
REGEX=$(python tools/ncu_launches.py $OUT/${TAG}_launches.csv --top-regex 6)
echo "full capture of: $REGEX"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$REGEX" \
    -o $OUT/${TAG}_full -f python tools/profile_step.py --frames 154 --steps 1 > $OUT/${TAG}_ncu2.log 2>&1
ls -la $OUT
