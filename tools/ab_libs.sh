#!/bin/bash
# A/B of library builds (same ABI) on the wavefront pipeline: bash tools/ab_libs.sh <tag> "<lib list>" ["<workload> <w> <h> <spp>" ...]
TAG=${1:-ab}; LIBS=${2:-"libgort.so"}; shift; shift
mkdir -p gpurun_out
OUT=gpurun_out/ab_libs_$TAG.log; : > $OUT
if [ $# -eq 0 ]; then set -- "c4 1920 1080 16" "c5 1920 1080 8"; fi
for rep in 1 2; do
for lib in $LIBS; do
  for cfg in "$@"; do
    echo -n "$lib $cfg: " >> $OUT
    GORT_LIB=$PWD/concurrent-raytracer-go_b200/lib/$lib timeout 600 python tools/prof_scene.py $cfg 3 2>&1 | grep trace_ms | tail -2 | awk '{printf "%s ", $6}' >> $OUT
    echo >> $OUT
  done
done
done
cat $OUT
