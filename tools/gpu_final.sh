#!/bin/bash
# last GPU call of the round at N=1: bench lines, launch list + one full capture of trace_kernel (C1-view), smoke, GPU tests
mkdir -p gpurun_out
timeout 400 python bench.py --steps 40 --warmup 5 > gpurun_out/final_c1view.json 2> gpurun_out/final_c1view.err; echo "bench rc=$?"
for w in c2_view c2_faithful; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 2> gpurun_out/final_$w.err | grep '^{' > gpurun_out/final_$w.json
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scale-c5"
$CMD > gpurun_out/plain_final.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final_c1_view.csv $CMD > gpurun_out/ncu_l_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_final_c1_view $CMD > gpurun_out/ncu_f_final.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_final.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4), d.get('scale_c5', {}).get('ms_per_frame'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
