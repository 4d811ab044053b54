"""Render one bench workload at a chosen size a few times (for ncu / timeline runs on big scenes).
usage: python tools/prof_scene.py <workload> <width> <height> <spp> [frames]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B
G = importlib.import_module("concurrent-raytracer-go_b200")
name, W, H, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
frames = int(sys.argv[5]) if len(sys.argv) > 5 else 3
kind, _, _, _, depth, options, desc = B.WORKLOADS[name]
flat = B.Workload(kind, options).flat(G)
r = G.NewParallelRenderer(1)
r.SetSamples(spp); r.SetMaxDepth(depth); r.SetSeed(20240601)
r.UploadScene(flat)
for i in range(frames):
    r.Render(flat, W, H)
    print(name, W, H, spp, "trace_ms %.3f" % r.lastStats.trace_ms, "bvh_ms %.1f" % r.lastStats.bvh_build_ms, flush=True)
r.SetCollectStats(True); r.Render(flat, W, H)
s = r.lastStats.as_dict()
print({k: s[k] for k in ("closest_queries", "shadow_queries", "shaded_hits", "soft_shadow_rays", "light_evals", "nodes_visited", "sphere_tests", "tri_tests", "cone_tests", "soft_pairs_skipped", "pairs_backfacing", "diffuse_evals", "bvh_nodes", "bvh_bytes")})
for k, site in enumerate(("FILL", "EXTEND", "SHADE hard", "SHADE soft (no candidates)")):
    if s["walk_warp_visits"][k]:
        print("walk %-28s lane visits %12d  lane utilisation %.3f" % (site, s["walk_lane_visits"][k], s["walk_lane_visits"][k] / s["walk_warp_visits"][k]))
