#!/bin/bash
# per-kernel times of every workload shape (and of a single KITTI frame); usage under gpurun: bash tools/gpu_kt.sh <tag>
TAG=${1:-kt}
OUT=gpurun_out
mkdir -p $OUT
for w in kitti154 synth128 cloud2m; do
  LPL_WORKLOAD=$w timeout 300 python tools/kernel_times.py > $OUT/${TAG}_kt_$w.txt 2>&1
done
LPL_WORKLOAD=kitti154 LPL_FRAMES=1 timeout 300 python tools/kernel_times.py > $OUT/${TAG}_kt_kitti1.txt 2>&1
head -60 $OUT/${TAG}_kt_cloud2m.txt
