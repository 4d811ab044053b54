"""The oracle must keep reproducing the committed golden fixtures (tests/golden/make_golden.py) — guards the
checker itself; the GPU tests compare the CUDA path with the same files."""
import json
import os

import numpy as np
from PIL import Image

import common as Cm

GOLD = os.path.join(Cm.ROOT, "tests", "golden")


def test_oracle_reproduces_c3_fixture(oracle):
    ref = np.array(Image.open(os.path.join(GOLD, "c3_two_red_cubes_800x600_1spp_d8.png")).convert("RGBA"))
    img, _, _ = oracle.Scene(Cm.c3()).render(800, 600, samples=1, max_depth=8, jitter=False, soft_shadows=False)
    assert (img == ref).all()


def test_oracle_reproduces_c1_fixture(oracle):
    ref = np.array(Image.open(os.path.join(GOLD, "c1_view_200x150_4spp_seed1.png")).convert("RGBA"))
    meta = json.load(open(os.path.join(GOLD, "c1_view_200x150_4spp_seed1.json")))
    img, rad, cnt = oracle.Scene(Cm.c1_view()).render(200, 150, samples=4, max_depth=50, rng_mode=oracle.RNG_PHILOX, seed=1,
                                                      want_radiance=True, threads=3)
    assert (img == ref).all()
    assert cnt == meta["counters"]  # Philox mode is independent of the thread count / tile schedule
    for s in meta["radiance"]:
        assert np.allclose(rad[s["y"], s["x"]], s["rgb"], rtol=1e-12, atol=0)


def test_oracle_reproduces_sphere_cloud_fixture(oracle):
    ref = np.array(Image.open(os.path.join(GOLD, "sphere_cloud_1500_320x180_2spp_d4_seed9.png")).convert("RGBA"))
    d = Cm.random_sphere_scene(1500, 77, cam_z=13.0)
    img, _, _ = oracle.Scene(d).render(320, 180, samples=2, max_depth=4, rng_mode=oracle.RNG_PHILOX, seed=9, use_accel=True)
    assert (img == ref).all()
    # the oracle's own BVH answers like its linear scan (renderer.go:333-346): same image on a crop
    lin, _, _ = oracle.Scene(d).render(320, 180, samples=2, max_depth=4, rng_mode=oracle.RNG_PHILOX, seed=9, crop=(120, 60, 200, 110))
    assert (lin[60:110, 120:200] == ref[60:110, 120:200]).all()
