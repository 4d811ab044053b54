"""Edge cases of the staged top-of-tree block (GORT_TOP) on tiny and small trees: every scene rendered through the wavefront
pipeline with the block off and on must give the same exact accumulators.  No torch, no oracle: runs in a few seconds."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as Cm
G = importlib.import_module("concurrent-raytracer-go_b200")


def few(n):
    d = Cm.c1_view()
    d["objects"] = d["objects"][:n]
    return d


scenes = [("1 sphere", few(1), 0), ("2 spheres", few(2), 0), ("3 spheres", few(3), 0), ("c1_view", Cm.c1_view(), 0), ("c3 cubes", Cm.c3(), 0),
          ("c2_view prisms", Cm.c2_view(), 1), ("40 spheres", Cm.random_sphere_scene(40, 3, cam_z=13.0), 0),
          ("300 spheres", Cm.random_sphere_scene(300, 5, cam_z=13.0), 0), ("1500 spheres", Cm.random_sphere_scene(1500, 77, cam_z=13.0), 0),
          ("20000 spheres", Cm.random_sphere_scene(20000, 9, cam_z=13.0), 0)]
os.environ["GORT_PATH"] = "stream"
r = G.NewParallelRenderer(1)
r.SetSamples(2); r.SetMaxDepth(6); r.SetSeed(11)
bad = 0
for name, d, opt in scenes:
    for bvh in ("host", "device") if len(d["objects"]) >= 40 else ("host",):
        os.environ["GORT_BVH"] = bvh
        sc = G.SceneFromDict(d, opt) if opt else G.SceneFromDict(d)
        rad = {}
        try:
            for top in ("0", "1", "2"):
                os.environ["GORT_TOP"] = top
                r.Render(sc, 256, 192)
                assert r.lastStats.render_path == 2
                rad[top] = r.ReadRadiance(256, 192)
        except Exception as e:  # keep going: the other scenes still say something
            print("%-16s bvh %-6s FAILED: %r" % (name, bvh, e), flush=True)
            bad += 1
            continue
        same = bool((rad["0"] == rad["1"]).all() and (rad["0"] == rad["2"]).all())
        bad += not same
        print("%-16s bvh %-6s lit %.3f  top block on == off: %s" % (name, bvh, float((rad["0"].sum(-1) > 0).mean()), same), flush=True)
print("edge check:", "OK" if bad == 0 else "%d MISMATCHES" % bad)
