#!/bin/bash
# two GPUs: frame link with the handshake folded into the cull / resolve kernels against the previous library and the one-thread kernels
mkdir -p gpurun_out
L=$PWD/concurrent-raytracer-go_b200/lib
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 --no-scale-c5 2> gpurun_out/n2_$tag.err | grep '^{' > gpurun_out/n2_$tag.json
}
for rep in 1 2; do
  run new_$rep GORT_LIB=$L/libgort.so
  [ -f $L/libgort_prev.so ] && run prev_$rep GORT_LIB=$L/libgort_prev.so
done
run unfused GORT_LINK_UNFUSED=1
run nochain GORT_LINK_NO_CHAIN=1
run nopdl GORT_NO_PDL=1
for e in "" GORT_EARLY_DEVICE=1; do
  env $e timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-scale-c5 2>/dev/null | grep '^{' > gpurun_out/n1_early_${e:-off}.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/n2_*.json') + glob.glob('gpurun_out/n1_early_*.json')):
    try:
        d = json.loads(open(f).read())
        print(f, 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'launches', d['gpu_launches'], d['run'].get('frame_link_vs_nccl_gather'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
timeout 300 python -m pytest tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -3
