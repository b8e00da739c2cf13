#!/bin/bash
# Build library variants with different -D tuning knobs (the LPL_* macros with #ifndef defaults in csrc/)
# into lidar_processing_v2_b200/variants/lib_<name>.so; time them on a GPU box with
#   LPL_B200_LIBRARY=.../lib_<name>.so python tools/kernel_times.py <kernel substrings>
# usage: tools/variants.sh "name:-DFLAG=1 -DOTHER=2" "name2:..."
set -e
cd "$(dirname "$0")/../lidar_processing_v2_b200/csrc"
mkdir -p ../variants
rm -f ../variants/*.so
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off \
       -shared -cudart static --threads 4 $flags -o ../variants/lib_$name.so \
       capi.cu ring_dror.cu segment.cu cluster.cu hull.cu obb.cu ingest.cu split.cu knn.cu 2>&1 | grep -i "error" || true
done
ls ../variants
