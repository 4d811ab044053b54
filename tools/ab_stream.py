"""A/B of the wavefront pipeline's switches (GORT_SORT: queue read in Morton order of the hit points; GORT_TOP: top of the wide
tree staged in shared memory) in one process: same scene, same seed, frames rendered alternately under each setting; the exact
accumulators must be equal.
usage: python tools/ab_stream.py <workload> <width> <height> <spp> [reps] [settings, e.g. GORT_SORT=0,GORT_SORT=1,GORT_TOP=1]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B
G = importlib.import_module("concurrent-raytracer-go_b200")
name, W, H, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
modes = sys.argv[6].split(",") if len(sys.argv) > 6 else ["GORT_SORT=0", "GORT_SORT=1"]


def apply(mode):
    os.environ["GORT_SORT"] = "0"; os.environ["GORT_TOP"] = "0"
    k, v = mode.split("=")
    os.environ[k] = v


kind, _, _, _, depth, options, desc = B.WORKLOADS[name]
t0 = time.time()
flat = B.Workload(kind, options).flat(G)
r = G.NewParallelRenderer(1)
r.SetSamples(spp); r.SetMaxDepth(depth); r.SetSeed(20240601)
os.environ["GORT_PATH"] = "stream"
r.UploadScene(flat)
print("%s: scene + upload %.1f s" % (name, time.time() - t0), flush=True)
rad = {}
for rep in range(reps):
    for mode in modes:
        apply(mode)
        r.Render(flat, W, H)
        assert r.lastStats.render_path == 2
        print(name, W, H, spp, mode, "trace_ms %.3f" % r.lastStats.trace_ms, "launches", r.lastStats.kernel_launches, flush=True)
        if rep == 0:
            rad[mode] = r.ReadRadiance(W, H)
base = rad[modes[0]]
for mode in modes[1:]:
    d = np.abs(rad[mode] - base)
    print(name, mode, "vs", modes[0], ": max |diff| %.3g, pixels that differ %d of %d, frame max %.3g" % (d.max(), int((d.max(axis=-1) > 0).sum()), W * H, base.max()), flush=True)

r.SetCollectStats(True)
for mode in ([] if os.environ.get("AB_NO_STATS") else modes[:2]):
    apply(mode)
    r.Render(flat, W, H)
    s = r.lastStats.as_dict()
    print(name, mode, "stats frame trace_ms %.3f" % r.lastStats.trace_ms, {k: s[k] for k in ("closest_queries", "shadow_queries", "shaded_hits", "nodes_visited")}, flush=True)
    for k, site in enumerate(("primary", "extension", "hard shadow", "soft shadow")):
        if s["walk_warp_visits"][k]:
            print("   walk %-12s lane visits %12d  lane utilisation %.3f" % (site, s["walk_lane_visits"][k], s["walk_lane_visits"][k] / s["walk_warp_visits"][k]), flush=True)
