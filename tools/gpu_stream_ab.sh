#!/bin/bash
# one GPU call: A/B of the wavefront pipeline's switches (tools/ab_stream.py) on the full-size C4 and the 4K C5 frame at 32 spp
# usage (under gpurun): bash tools/gpu_sort_ab.sh [settings]    default: GORT_SORT=0,GORT_SORT=1,GORT_SORT=2,GORT_SORT=3
SET=${1:-GORT_SORT=0,GORT_SORT=1,GORT_SORT=2,GORT_SORT=3}
mkdir -p gpurun_out
OUT=gpurun_out/ab_stream_$(echo $SET | tr ',=' '__' | cut -c1-40).log; : > $OUT
timeout 60 python tools/ab_stream.py c4 1920 1080 64 2 $SET >> $OUT 2>&1; echo "c4 rc=$?" >> $OUT
timeout 90 python tools/ab_stream.py c5 3840 2160 32 2 $SET >> $OUT 2>&1; echo "c5 rc=$?" >> $OUT
cat $OUT
