#!/bin/bash
# one short GPU call: shared-memory staging of the top of the wide tree (GORT_TOP: 1 bulk copy, 2 plain loads) against the default
mkdir -p gpurun_out
L=gpurun_out/ab_stream_top.log
GORT_STREAM_DEBUG=1 AB_NO_STATS=1 timeout 40 python tools/ab_stream.py c4 1920 1080 64 2 GORT_TOP=0,GORT_TOP=1,GORT_TOP=2 > $L 2>&1; echo "c4 rc=$?" >> $L
AB_NO_STATS=1 timeout 40 python tools/ab_stream.py c5 3840 2160 32 1 GORT_TOP=0,GORT_TOP=1,GORT_TOP=2 >> $L 2>&1; echo "c5 rc=$?" >> $L
cat $L
