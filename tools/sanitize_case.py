"""Smallest end-to-end case for compute-sanitizer (one tool per gpurun call): tiny-scene and BVH kernels,
soft shadows, glass, prisms, fog, sharded slab + unswizzle."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as Cm
G = importlib.import_module("concurrent-raytracer-go_b200")
r = G.NewParallelRenderer(1)
r.SetSamples(4); r.SetMaxDepth(12); r.SetSeed(3)
a = r.Render(G.SceneFromDict(Cm.c1_view()), 160, 120)
b = r.Render(G.SceneFromDict(Cm.c2_view(), 3), 160, 120)
c = r.Render(G.SceneFromDict(Cm.random_sphere_scene(400, 5)), 96, 64)
r.SetShard(1, 3)
d = r.Render(G.SceneFromDict(Cm.c2_view(), 1), 100, 70)
print("ok", int(a.sum()), int(b.sum()), int(c.sum()), int(d.sum()))
