#!/bin/bash
# A/B of the two BVH render paths (per-warp-queue kernel vs global-queue wavefront pipeline) on the large workloads.
# usage (under gpurun): bash tools/ab_paths.sh <tag> ["<workload> <w> <h> <spp>" ...]
TAG=${1:-ab}; shift
mkdir -p gpurun_out
OUT=gpurun_out/ab_paths_$TAG.log; : > $OUT
if [ $# -eq 0 ]; then set -- "c4 480 270 16" "c5 480 270 8" "c2_view 1200 900 100"; fi
for cfg in "$@"; do
  for path in queue stream; do
    echo "=== GORT_PATH=$path $cfg" >> $OUT
    GORT_PATH=$path timeout 600 python tools/prof_scene.py $cfg 3 >> $OUT 2>&1
  done
done
cat $OUT
