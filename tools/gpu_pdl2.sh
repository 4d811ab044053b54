#!/bin/bash
# the chained frame with stamp-based kernel times: bench twice per library, then the GPU tests
mkdir -p gpurun_out
L=$PWD/concurrent-raytracer-go_b200/lib
for rep in 1 2; do
  for lib in libgort_prev.so libgort.so; do
    [ -f $L/$lib ] || continue
    GORT_LIB=$L/$lib timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-scale-c5 2> gpurun_out/pdl2_${lib}_$rep.err | grep '^{' > gpurun_out/pdl2_${lib}_$rep.json
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/pdl2_*.json')):
    try:
        d = json.loads(open(f).read())
        print(f, 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'kernel_ms', round(d['roofline']['kernel_ms'], 5), 'cull_ms', round(d['roofline']['cull_ms'], 5))
    except Exception as e:
        print(f, 'unreadable', e)
PY
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log; fi
