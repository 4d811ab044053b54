#!/bin/bash
# A/B of walk variants on the BVH workloads (under gpurun): bash tools/ab_walk.sh <tag> "<lib list>"
TAG=${1:-ab}; LIBS=${2:-"libgort.so libgort_lane.so"}
mkdir -p gpurun_out
OUT=gpurun_out/ab_walk_$TAG.log; : > $OUT
for lib in $LIBS; do
  export GORT_LIB=$PWD/concurrent-raytracer-go_b200/lib/$lib
  echo "=== $lib" >> $OUT
  timeout 300 python tools/prof_scene.py c4 480 270 16 3 >> $OUT 2>&1
  timeout 300 python tools/prof_scene.py c5 480 270 8 3 >> $OUT 2>&1
  timeout 300 python tools/prof_scene.py c2_view 1200 900 100 3 >> $OUT 2>&1
done
unset GORT_LIB
cat $OUT
