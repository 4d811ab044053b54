import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import common as Cm, synth, oracle as O
G = importlib.import_module("concurrent-raytracer-go_b200")
a = synth.c4_arrays()
flat = synth.to_gort(a); osc = synth.to_oracle(a)
W, H = 480, 270
crop = (W // 2 - 24, H // 2 - 16, W // 2 + 24, H // 2 + 16)
x0, y0, x1, y1 = crop
r = G.NewParallelRenderer(1); r.UploadScene(flat)
for spp, depth, soft, jit, rec in ((1, 1, False, False, True), (1, 1, True, False, True), (1, 16, False, False, True), (4, 16, True, True, True), (1, 2, False, False, True), (1, 3, False, False, True)):
    for nocull in (0, 1):
        if nocull: os.environ["GORT_NO_CONE_CULL"] = "1"
        else: os.environ.pop("GORT_NO_CONE_CULL", None)
        r.SetSamples(spp); r.SetMaxDepth(depth); r.SetSoftShadows(soft); r.SetAntiAliasing(jit); r.SetSeed(9)
        img = r.Render(flat, W, H)
        rad = r.ReadRadiance(W, H)
        ref, rrad, _ = osc.render(W, H, samples=spp, max_depth=depth, jitter=jit, soft_shadows=soft, rng_mode=O.RNG_PHILOX, seed=9, crop=crop, use_accel=True, threads=8, want_radiance=True)
        p, q = img[y0:y1, x0:x1], ref[y0:y1, x0:x1]
        dr = np.abs(rad[y0:y1, x0:x1] - rrad[y0:y1, x0:x1]).max(-1)
        print("spp %d depth %d soft %d jitter %d nocull %d: within1 %.4f psnr %.1f  radiance |diff| median %.2e p90 %.2e max %.2e frac>1e-3 %.3f" % (
            spp, depth, soft, jit, nocull, Cm.within_one(p, q), Cm.psnr(p, q), np.median(dr), np.quantile(dr, 0.9), dr.max(), (dr > 1e-3).mean()), flush=True)
