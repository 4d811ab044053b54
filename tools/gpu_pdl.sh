#!/bin/bash
# A/B of the chained frame (programmatic dependent launches, no timing events, no memset node) against the previous library,
# then the C2 bench lines and the GPU tests.  bash tools/gpu_pdl.sh
mkdir -p gpurun_out
L=$PWD/concurrent-raytracer-go_b200/lib
for rep in 1 2; do
  for lib in libgort_prev.so libgort.so; do
    [ -f $L/$lib ] || continue
    GORT_LIB=$L/$lib timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-scale-c5 2> gpurun_out/pdl_${lib}_$rep.err | grep '^{' > gpurun_out/pdl_${lib}_$rep.json
  done
done
GORT_NO_PDL=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-scale-c5 2>/dev/null | grep '^{' > gpurun_out/pdl_nopdl.json
for w in c2_view c2_faithful; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 2> gpurun_out/bench_$w.err | grep '^{' > gpurun_out/bench_$w.json
  GORT_LIB=$L/libgort_prev.so timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_${w}_prev.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/pdl_*.json') + glob.glob('gpurun_out/bench_c2*.json')):
    try:
        d = json.loads(open(f).read())
        print(f, 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'launches', d['gpu_launches'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log; fi
