#!/bin/bash
# usage under gpurun: bash tools/gpu_sanitize.sh <tag>
TAG=${1:-san}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck synccheck initcheck racecheck; do
  echo "== $tool" >> $OUT/${TAG}_sanitizer.txt
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_run.py > $OUT/${TAG}_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|sanitize_run ok|smoke ok|Traceback|AssertionError" $OUT/${TAG}_$tool.log | cut -c1-220 | sort | uniq -c | head -20 >> $OUT/${TAG}_sanitizer.txt
done
cat $OUT/${TAG}_sanitizer.txt
