import sys, os, time, json, importlib
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np
import oracle as O
import common as Cm
G = importlib.import_module("concurrent-raytracer-go_b200")
r = G.NewParallelRenderer(1)
print("regs ok; fp32 peak", r.MeasureFp32Peak())
# C3 deterministic
d = Cm.c3()
r.SetSamples(1); r.SetMaxDepth(8); r.SetAntiAliasing(False); r.SetSoftShadows(False)
img = r.Render(G.SceneFromDict(d), 800, 600)
ref, _, _ = O.Scene(d).render(800, 600, samples=1, max_depth=8, jitter=False, soft_shadows=False)
print("C3 within1", Cm.within_one(img, ref), "mae", Cm.mae(img, ref), "nonblack", (img[...,:3].sum(-1)>0).mean(), (ref[...,:3].sum(-1)>0).mean(), "kernel_ms", r.lastStats.kernel_ms)
# C1-view same-stream
d = Cm.c1_view()
r.SetSamples(8); r.SetMaxDepth(50); r.SetAntiAliasing(True); r.SetSoftShadows(True); r.SetSeed(7)
img = r.Render(G.SceneFromDict(d), 800, 600)
ref, _, _ = O.Scene(d).render(800, 600, samples=8, max_depth=50, rng_mode=O.RNG_PHILOX, seed=7)
print("C1 8spp within1", Cm.within_one(img, ref), "mae", Cm.mae(img, ref), "psnr", Cm.psnr(img, ref), "kernel_ms", r.lastStats.kernel_ms, r.lastStats.total_ms)
r.SetSamples(100)
for i in range(3):
    img = r.Render(G.SceneFromDict(d), 800, 600)
    print("C1 100spp kernel_ms", r.lastStats.kernel_ms, "trace", r.lastStats.trace_ms, "total", r.lastStats.total_ms)
r.SetCollectStats(True)
img = r.Render(G.SceneFromDict(d), 800, 600)
print(json.dumps(r.lastStats.as_dict()))
