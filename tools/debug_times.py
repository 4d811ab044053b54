"""Per-warp timeline of the trace kernel (needs `make -C concurrent-raytracer-go_b200 debug`):
GORT_DEBUG_TIMES=1 GORT_LIB=$PWD/concurrent-raytracer-go_b200/lib/libgort_dbg.so python tools/debug_times.py [workloads]"""
import importlib
import os
import sys

os.environ.setdefault("GORT_DEBUG_TIMES", "1")
os.environ.setdefault("GORT_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "concurrent-raytracer-go_b200", "lib", "libgort_dbg.so"))
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common as Cm
G = importlib.import_module("concurrent-raytracer-go_b200")
r = G.NewParallelRenderer(1)
r.SetSeed(20240601)
want = sys.argv[1:] or ["c1_view", "c2_view"]
for name, d, W, H in (("c1_view", Cm.c1_view(), 800, 600), ("c2_view", Cm.c2_view(), 1200, 900)):
    if name not in want:
        continue
    sc = G.SceneFromDict(d, 1)
    for i in range(3):
        r.Render(sc, W, H)
    print(name, r.lastStats.trace_ms, flush=True)
    r.SetCollectStats(True); r.Render(sc, W, H); r.SetCollectStats(False)
    s = r.lastStats.as_dict()
    print({k: s[k] for k in ("closest_queries", "shadow_queries", "shaded_hits", "soft_shadow_rays", "light_evals", "paths_depth_ge5", "paths_depth_ge20", "paths_depth_max", "nodes_visited", "sphere_tests", "tri_tests", "rng_blocks", "cone_tests")}, flush=True)
