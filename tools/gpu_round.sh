#!/bin/bash
# One GPU call of the round: parity tests, bench line, ncu launch list, measured-FLOP counters and a full capture of the top kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [workload] [top-kernel regex]
TAG=${1:-r2}; WL=${2:-c1_view}; TOP=${3:-trace_kernel}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log; fi
python bench.py --workload $WL ${STEPS:+--steps $STEPS} > gpurun_out/bench_${TAG}_$WL.json 2> gpurun_out/bench_${TAG}_$WL.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}_$WL.json
python bench.py --impl reference --steps 3 --warmup 1 --workload $WL > gpurun_out/bench_${TAG}_${WL}_reference.json 2>/dev/null; cat gpurun_out/bench_${TAG}_${WL}_reference.json
CMD="python bench.py --steps ${NCU_STEPS:-2} --warmup 3 --no-cpu-baseline --no-scale-c5 --workload $WL"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}_$WL.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
# SURVEY 8d source (1): measured FP32 instruction counts of the trace kernels (ncu --set full does not collect them)
ncu --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum \
    --clock-control none -c 4000 --csv --log-file gpurun_out/flops_${TAG}_$WL.csv $CMD > gpurun_out/ncu_m_$TAG.log 2>&1
python tools/ncu_flops.py gpurun_out/flops_${TAG}_$WL.csv $WL > gpurun_out/flops_${TAG}_$WL.txt 2>&1; tail -2 gpurun_out/flops_${TAG}_$WL.txt
cp profiles/flops.json gpurun_out/flops.json
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$TOP -s ${SKIP:-3} -c ${COUNT:-1} -f -o gpurun_out/prof_${TAG}_$WL $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
echo done
