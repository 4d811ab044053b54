# A/B runs: tools/ab.sh "<lib list>" "<workloads>" [steps]
LIBS=${1:-"libgort.so"}; WLS=${2:-"c1_view c2_view"}; STEPS=${3:-20}
for lib in $LIBS; do for w in $WLS; do GORT_LIB=$PWD/concurrent-raytracer-go_b200/lib/$lib timeout 600 python bench.py --steps $STEPS --warmup 3 --workload $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$w', 'ms', round(d['ms_per_step'],4), 'trace_ms', round(d['roofline']['kernel_ms'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'gflop', round(d['roofline']['algorithmic_flops_per_launch']/1e9,3))"; done; done
