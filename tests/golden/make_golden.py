"""Generates the committed golden fixtures from the float64 oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference itself cannot be run here (Go, no toolchain;
unseeded RNG), so the fixtures are oracle outputs: they pin the CUDA path AND guard the oracle
against accidental edits."""
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
import common as Cm  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    img, _, _ = O.Scene(Cm.c3()).render(800, 600, samples=1, max_depth=8, jitter=False, soft_shadows=False)
    Image.fromarray(img, "RGBA").save(os.path.join(OUT, "c3_two_red_cubes_800x600_1spp_d8.png"), optimize=True)
    # C1-view at 200x150, 4 spp, Philox seed 1: small same-stream fixture + radiance samples
    img, rad, cnt = O.Scene(Cm.c1_view()).render(200, 150, samples=4, max_depth=50, rng_mode=O.RNG_PHILOX, seed=1, want_radiance=True)
    Image.fromarray(img, "RGBA").save(os.path.join(OUT, "c1_view_200x150_4spp_seed1.png"), optimize=True)
    ys, xs = np.nonzero(rad.sum(-1) > 0)
    pick = list(range(0, len(ys), max(1, len(ys) // 40)))[:40]
    json.dump({"config": "c1_view 200x150 4spp depth50 philox seed1", "counters": cnt,
               "radiance": [{"x": int(xs[i]), "y": int(ys[i]), "rgb": rad[ys[i], xs[i]].tolist()} for i in pick]},
              open(os.path.join(OUT, "c1_view_200x150_4spp_seed1.json"), "w"), indent=1)
    # a BVH scene for both BVH render paths: the 1 500-sphere cloud (metal / glass / dielectric, 3 lights, soft shadows) at
    # 320x180, 2 spp, depth 4, Philox seed 9
    d = Cm.random_sphere_scene(1500, 77, cam_z=13.0)
    img, _, _ = O.Scene(d).render(320, 180, samples=2, max_depth=4, rng_mode=O.RNG_PHILOX, seed=9, use_accel=True)
    Image.fromarray(img, "RGBA").save(os.path.join(OUT, "sphere_cloud_1500_320x180_2spp_d4_seed9.png"), optimize=True)


if __name__ == "__main__":
    main()
