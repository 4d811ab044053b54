"""The oracle's optional BVH (test speed-up for large scenes) must return exactly what the reference's
linear scan returns (hitWorld renderer.go:333-346), including the last-wins tie rule."""
import numpy as np

import common as Cm


def test_accel_equals_linear_scan(oracle):
    d = Cm.random_sphere_scene(400, 11)
    rng = np.random.default_rng(0)
    for k in range(40):
        d["objects"].append({"type": "cube", "position": rng.uniform(-8, 8, 3).tolist(), "size": rng.uniform(0.3, 2, 3).tolist(),
                             "material": {"type": "metal", "color": [0.5, 0.5, 0.5]}})
    s = oracle.Scene(d)
    n_hit = 0
    for i in range(1500):
        o = rng.uniform(-15, 15, 3)
        dr = rng.normal(size=3) * rng.uniform(0.2, 3)
        a = s.hit_world(o, dr)
        b = s.hit_world(o, dr, use_accel=True)
        assert (a is None) == (b is None)
        if a is not None:
            n_hit += 1
            assert a["t"] == b["t"] and a["prim"] == b["prim"] and a["material"] == b["material"]
            assert a["normal"].tolist() == b["normal"].tolist()
    assert n_hit > 100


def test_accel_render_identical(oracle):
    d = Cm.random_sphere_scene(300, 12)
    s = oracle.Scene(d)
    kw = dict(samples=2, max_depth=6, rng_mode=oracle.RNG_PHILOX, seed=3, want_radiance=True)
    a, ra, _ = s.render(64, 36, **kw)
    b, rb, _ = s.render(64, 36, use_accel=True, **kw)
    assert (a == b).all() and (ra == rb).all()
