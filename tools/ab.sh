# A/B runs: tools/ab.sh "<lib list>" "<workloads>" "<urgent depths>"
LIBS=${1:-"libgort.so libgort_mc3.so"}; WLS=${2:-"c1_view c2_view"}; UDS=${3:-"2"}
for lib in $LIBS; do for w in $WLS; do for ud in $UDS; do GORT_URGENT_DEPTH=$ud GORT_LIB=$PWD/concurrent-raytracer-go_b200/lib/$lib timeout 300 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$w', 'urgent', '$ud', 'ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))"; done; done; done
