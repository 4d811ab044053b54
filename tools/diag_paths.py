"""Diagnostic: where do the two BVH render paths (GORT_PATH=queue / stream) differ on a sphere cloud?
usage: python tools/diag_paths.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as Cm
G = importlib.import_module("concurrent-raytracer-go_b200")
r = G.NewParallelRenderer(1)


def both(sc, W, H):
    out = []
    for path in ("queue", "stream"):
        os.environ["GORT_PATH"] = path
        r.Render(sc, W, H)
        out.append(r.ReadRadiance(W, H))
    return out


for only in (None, "metal", "glass", "dielectric"):
    for soft in (False,):
        for depth in (2, 3, 4, 16):
            d = Cm.random_sphere_scene(4000, 5, cam_z=26.0)
            if only:
                keep = [o for o in d["objects"] if o["material"]["type"] == only]
                for o in d["objects"]:
                    if o["material"]["type"] != only:
                        o["material"] = dict(keep[0]["material"])
            sc = G.SceneFromDict(d)
            r.SetSamples(5); r.SetMaxDepth(depth); r.SetSeed(3); r.SetSoftShadows(soft)
            a, b = both(sc, 384, 216)
            rel = np.abs(a - b).max(-1) / (1.0 + np.abs(a).max(-1))
            print("only %-10s soft %d depth %2d: differing pixels >1e-6: %.5f  >1e-4: %.5f  >1e-2: %.5f  max %.3g  mean a %.4f mean b %.4f" % (
                only, soft, depth, (rel > 1e-6).mean(), (rel > 1e-4).mean(), (rel > 1e-2).mean(), rel.max(), a.mean(), b.mean()), flush=True)
