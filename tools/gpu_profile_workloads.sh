#!/bin/bash
# ncu launch lists of the other BASELINE shapes and a --set full capture of the dominant kernel of the 2M-point cloud
# usage under gpurun: bash tools/gpu_profile_workloads.sh <tag>
TAG=${1:-wl}
OUT=gpurun_out; mkdir -p $OUT
export LPL_SPLIT=1
for w in cloud2m synth128; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_$w.csv \
      python tools/profile_step.py --workload $w --steps 2 > $OUT/${TAG}_ncu_$w.log 2>&1
  python tools/ncu_launches.py $OUT/${TAG}_launches_$w.csv > $OUT/${TAG}_launches_summary_$w.txt
  head -8 $OUT/${TAG}_launches_summary_$w.txt
done
timeout 900 ncu --set full --clock-control none -k "regex:(k_dror_query|k_dror_grid_scatter|k_dror_near|k_clu_union_sm)" -c 8 -o /tmp/${TAG}_cloud2m -f \
    python tools/profile_step.py --workload cloud2m --steps 2 > $OUT/${TAG}_ncu_full_cloud2m.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_cloud2m.ncu-rep > $OUT/${TAG}_ncu_full_summary_cloud2m.txt 2>&1
grep -c "^==" $OUT/${TAG}_ncu_full_summary_cloud2m.txt
