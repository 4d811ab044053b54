import sys, importlib, os, json
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import common as Cm
G = importlib.import_module("concurrent-raytracer-go_b200")
r = G.NewParallelRenderer(1)
r.SetSeed(20240601)
for name, d, W, H in (("c1_view", Cm.c1_view(), 800, 600), ("c2_view", Cm.c2_view(), 1200, 900)):
    sc = G.SceneFromDict(d, 1)
    for i in range(3):
        r.Render(sc, W, H)
    print(name, r.lastStats.trace_ms, flush=True)
    r.SetCollectStats(True); r.Render(sc, W, H); r.SetCollectStats(False)
    s = r.lastStats.as_dict()
    print({k: s[k] for k in ("closest_queries","shadow_queries","shaded_hits","soft_shadow_rays","light_evals","paths_depth_ge5","paths_depth_ge20","paths_depth_max","nodes_visited","sphere_tests","tri_tests","rng_blocks")}, flush=True)
