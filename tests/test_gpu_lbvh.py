"""The device-side BVH builder (csrc/lbvh.cu: Morton order + Karras' parallel radix tree) against the host builder (binned SAH)
and the reference's linear scan.  The reference has no acceleration structure (hitWorld scans every hittable,
/root/reference internal/renderer/renderer.go:333-346): whatever tree is used, the closest hit must be the scan's, so the
two builders must give the same frame."""
import os

import numpy as np
import pytest

import common as Cm
from test_gpu_parity import configure

pytestmark = pytest.mark.gpu


class bvh_builder:
    def __init__(self, which):
        self.which = which

    def __enter__(self):
        self.old = os.environ.get("GORT_BVH")
        os.environ["GORT_BVH"] = self.which

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("GORT_BVH", None)
        else:
            os.environ["GORT_BVH"] = self.old


def mixed_scene(n, seed):
    d = Cm.random_sphere_scene(n, seed, cam_z=26.0)
    rng = np.random.default_rng(seed + 1)
    for i in range(0, len(d["objects"]), 7):  # every seventh object becomes a cube (12 triangles)
        o = d["objects"][i]
        d["objects"][i] = {"type": "cube", "position": o["position"], "size": (rng.uniform(0.3, 1.0, 3)).tolist(), "material": o["material"]}
    return d


@pytest.mark.parametrize("n,seed", [(3000, 1), (20000, 2)])
def test_device_tree_closest_hit_equals_linear_scan(gort, oracle, n, seed):
    d = mixed_scene(n, seed)
    with bvh_builder("device"):
        r = gort.NewParallelRenderer(1)
        r.UploadScene(gort.SceneFromDict(d))
    osc = oracle.Scene(d)
    rng = np.random.default_rng(seed)
    m = 400
    o = rng.uniform(-12, 12, (m, 3))
    dirs = rng.normal(size=(m, 3)) * rng.uniform(0.2, 3.0, (m, 1))
    o32, d32 = o.astype(np.float32).astype(np.float64), dirs.astype(np.float32).astype(np.float64)
    t, order = r.TraceRays(o32, d32)
    hits = bad = 0
    for i in range(m):
        h = osc.hit_world(o32[i], d32[i])
        if h is None:
            bad += t[i] >= 0
            continue
        hits += 1
        bad += t[i] < 0 or abs(t[i] - h["t"]) > 2e-4 * max(1.0, h["t"])
    r.close()
    assert hits > m // 3 and bad <= 2, (hits, bad)


@pytest.mark.parametrize("path", ["queue", "stream"])
def test_device_tree_renders_the_host_trees_frame(gort, path):
    d = mixed_scene(6000, 4)
    os.environ["GORT_PATH"] = path
    try:
        out = []
        for which in ("host", "device"):
            with bvh_builder(which):
                r = gort.NewParallelRenderer(1)
                configure(r, 4, 10, seed=12)
                img = r.Render(gort.SceneFromDict(d), 384, 216).copy()
                out.append((img, r.ReadRadiance(384, 216), r.lastStats.bvh_nodes, r.lastStats.bvh_build_ms))
                r.close()
    finally:
        del os.environ["GORT_PATH"]
    (a_img, a, na, _), (b_img, b, nb, ms) = out
    assert a.max() > 0
    assert nb == 6000 + 11 * ((6000 + 6) // 7) - 1  # one primitive per leaf: n - 1 inner nodes
    # same hits -> same radiance (a tie between two primitives at equal t is settled by the scan order in both trees)
    same = float((np.abs(a - b).max(axis=-1) <= 1e-6 * (1.0 + np.abs(a).max(axis=-1))).mean())
    assert same >= 0.9999, same
    assert float((a_img == b_img).all(axis=-1).mean()) >= 0.9999


def test_device_tree_with_coincident_primitives(gort, oracle):
    """equal Morton codes (concentric and duplicated spheres) are split by their position in the sorted order"""
    objs = []
    for k in range(40):
        objs.append({"type": "sphere", "position": [0, 0, 0], "radius": 0.5 + 0.01 * k, "material": {"type": "glass", "color": [0.9, 0.9, 1.0], "refractionIndex": 1.3}})
    for k in range(40):
        objs.append({"type": "sphere", "position": [1.5, 0.2, 0], "radius": 0.4, "material": {"type": "metal", "color": [0.8, 0.6, 0.2], "roughness": 0.1}})
    objs.append({"type": "cube", "position": [0, -1.5, 0], "size": [8, 0.4, 8], "material": {"type": "lambertian", "color": [0.6, 0.6, 0.6]}})
    d = {"camera": {"position": [0.5, 0.3, 5], "aspectRatio": 1.5}, "objects": objs,
         "lights": [{"type": "point", "position": [4, 6, 5], "color": [1, 1, 1], "intensity": 40}]}
    with bvh_builder("device"):
        r = gort.NewParallelRenderer(1)
        configure(r, 2, 6, seed=3)
        img = r.Render(gort.SceneFromDict(d), 300, 200)
        r.close()
    ref, _, _ = oracle.Scene(d).render(300, 200, samples=2, max_depth=6, rng_mode=oracle.RNG_PHILOX, seed=3)
    from test_gpu_parity import check
    check(img, ref)
