#!/bin/bash
# One GPU call of the round: parity tests, bench line, ncu launch list + full capture of trace_kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [workload]
TAG=${1:-r1}; WL=${2:-c1_view}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --workload $WL > gpurun_out/bench_${TAG}_$WL.json 2> gpurun_out/bench_${TAG}_$WL.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}_$WL.json
python bench.py --impl reference --steps 3 --warmup 1 --workload $WL > gpurun_out/bench_${TAG}_${WL}_reference.json 2>/dev/null; cat gpurun_out/bench_${TAG}_${WL}_reference.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $WL"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_$WL.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_trace_${TAG}_$WL $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
echo done
