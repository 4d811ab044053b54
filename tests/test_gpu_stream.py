"""The two BVH render paths of libgort give the same frame.

Large scenes run the global-queue wavefront pipeline (csrc/stream.cu: one bounce at a time over queues in HBM, a
traversal kernel that refills finished lanes); small ones the per-warp-queue kernel (csrc/kernels.cu).  Both follow
traceRay / calculateDirectLighting (/root/reference internal/renderer/renderer.go:165-331) with the same fp32
arithmetic, the same Philox counters and commutative fixed-point accumulators, so GORT_PATH=queue and
GORT_PATH=stream must produce the same radiance for the same (scene, params, seed) — and each must meet the oracle
bar on its own."""
import os

import numpy as np
import pytest

import common as Cm
from test_gpu_parity import _contact_scene, _random_scene, all_materials_scene, check, configure

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer(gort):
    r = gort.NewParallelRenderer(1)
    yield r
    r.close()


class forced_path:
    def __init__(self, path, **env):
        self.env = dict(env, GORT_PATH=path)

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.env}
        os.environ.update({k: str(v) for k, v in self.env.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def render_both(renderer, scene, W, H, **env):
    with forced_path("queue"):
        a_img = renderer.Render(scene, W, H).copy()
        assert renderer.lastStats.render_path == 1
        a = renderer.ReadRadiance(W, H)
    with forced_path("stream", **env):
        b_img = renderer.Render(scene, W, H).copy()
        assert renderer.lastStats.render_path == 2
        b = renderer.ReadRadiance(W, H)
    return a_img, a, b_img, b


def assert_same(a_img, a, b_img, b):
    assert a.max() > 0
    # same arithmetic per ray; the compiler may contract a multiply-add differently in the two kernels
    same = float((np.abs(a - b).max(axis=-1) <= 1e-6 * (1.0 + np.abs(a).max(axis=-1))).mean())
    assert same >= 0.9999, same
    assert float((a_img == b_img).all(axis=-1).mean()) >= 0.9999


@pytest.mark.parametrize("soft,recursive,depth,jitter", [(True, True, 12, True), (False, True, 50, True), (True, False, 50, True),
                                                         (True, True, 1, False), (True, True, 2, True)])
def test_stream_equals_queue_all_materials(gort, renderer, soft, recursive, depth, jitter):
    sc = gort.SceneFromDict(all_materials_scene())
    configure(renderer, 3, depth, soft=soft, recursive=recursive, jitter=jitter, seed=21)
    assert_same(*render_both(renderer, sc, 640, 400))


@pytest.mark.parametrize("seed", [1, 2, 4, 5, 7, 8, 10, 11])
def test_stream_equals_queue_random_scenes(gort, renderer, seed):
    """spheres / cubes / prisms, every material, 0..11 lights (up to three light chunks in the wavefront pipeline)"""
    d = _random_scene(seed)
    if len(d["objects"]) <= 12 and all(o["type"] == "sphere" for o in d["objects"]) and len(d["lights"]) <= 4:
        d["objects"].append({"type": "cube", "position": [0, -4, 0], "size": [1, 1, 1], "material": {"type": "lambertian", "color": [0.5, 0.5, 0.5]}})
    rng = np.random.default_rng(1000 + seed)
    depth, soft, spp = int(rng.choice([1, 2, 3, 6, 12])), bool(rng.integers(0, 2)), int(rng.choice([1, 2, 4]))
    configure(renderer, spp, depth, soft=soft, seed=seed)
    a_img, a, b_img, b = render_both(renderer, gort.SceneFromDict(d, 1), 300, 200)
    if a.max() == 0:
        pytest.skip("nothing in frame")
    assert_same(a_img, a, b_img, b)


def test_stream_equals_queue_sphere_cloud_with_fog(gort, renderer):
    """C4/C5-style cloud (metal / glass / dielectric), fog on the primary hit, 3 lights, soft shadows with overflowing
    candidate lists (the pool_trace<SOFT> source), several batches and a partial last batch."""
    d = Cm.random_sphere_scene(4000, 5, cam_z=26.0)
    d["fog"] = {"enabled": True, "density": 0.01, "color": [0.25, 0.25, 0.25], "type": "exponential"}
    sc = gort.SceneFromDict(d, gort.LOAD_FOG)
    configure(renderer, 5, 16, seed=3)
    one = render_both(renderer, sc, 384, 216)
    assert_same(*one)
    many = render_both(renderer, sc, 384, 216, GORT_STREAM_BATCH=200000)
    assert np.array_equal(one[3], many[3])  # batching never changes the integer accumulators


def test_stream_counters_equal_queue_counters(gort, renderer):
    """Same rays, same tests: the device counters of the two paths agree (walk order inside a leaf aside)."""
    d = Cm.random_sphere_scene(3000, 9, cam_z=26.0)
    sc = gort.SceneFromDict(d)
    configure(renderer, 4, 8, seed=5)
    renderer.SetCollectStats(True)
    try:
        with forced_path("queue"):
            renderer.Render(sc, 320, 180)
            q = renderer.lastStats.as_dict()
        with forced_path("stream"):
            renderer.Render(sc, 320, 180)
            s = renderer.lastStats.as_dict()
    finally:
        renderer.SetCollectStats(False)
    # (how a lit pair's 16 soft rays are answered — empty cone, candidate list or BVH walks — is each path's own policy: the
    # pipeline gives a cone walk 16 node visits; the answers are the same, the counters of that stage are not)
    for k in ("primary_generated", "closest_queries", "shaded_hits", "light_evals", "pairs_backfacing", "diffuse_evals", "specular_evals"):
        assert abs(q[k] - s[k]) <= 2e-4 * max(1, q[k]), (k, q[k], s[k])
    # (node visits differ by design: the pipeline walks a 4-wide collapse of the tree — four slab tests per visit, counted as
    # two binary visits — defers its leaf tests, and sends more pairs' soft rays through the BVH)
    assert 0.5 * q["nodes_visited"] < s["nodes_visited"] < 4.0 * q["nodes_visited"]
    assert s["primary_generated"] > 0 and s["primary_generated"] <= s["primary_rays"]
    # lane refill: the walk's SIMT use per call site (primary, extension, hard, soft) is well above the static batches'
    for site in range(1, 4):  # (primary rays are coherent either way)
        if s["walk_warp_visits"][site] > 100000:
            util_s = s["walk_lane_visits"][site] / s["walk_warp_visits"][site]
            util_q = q["walk_lane_visits"][site] / max(1, q["walk_warp_visits"][site])
            print("site %d lane utilisation queue %.3f stream %.3f" % (site, util_q, util_s))
            assert util_s > 1.2 * util_q  # (a frame this small cannot fill the refill pools: 0.75-0.85 at full size)


def test_stream_same_stream_as_oracle(gort, oracle, renderer):
    """The wavefront pipeline against the float64 oracle on its own (same Philox draws): the stochastic-config bar."""
    d = Cm.random_sphere_scene(1500, 77, cam_z=13.0)
    configure(renderer, 2, 16, seed=9)
    with forced_path("stream"):
        img = renderer.Render(gort.SceneFromDict(d), 320, 180)
    ref, _, _ = oracle.Scene(d).render(320, 180, samples=2, max_depth=16, rng_mode=oracle.RNG_PHILOX, seed=9, use_accel=True)
    check(img, ref, within=0.998, max_bad_lit=90)  # (the dense cloud: see test_random_spheres_bvh_scene_same_stream)
    d = _contact_scene(False)
    configure(renderer, 4, 12, seed=17)
    with forced_path("stream"):
        img = renderer.Render(gort.SceneFromDict(d), 480, 320)
    ref, _, _ = oracle.Scene(d).render(480, 320, samples=4, max_depth=12, rng_mode=oracle.RNG_PHILOX, seed=17)
    check(img, ref)


def test_stream_deterministic_c3(gort, oracle, renderer):
    d = Cm.c3()
    configure(renderer, 1, 8, jitter=False, soft=False)
    with forced_path("stream"):
        img = renderer.Render(gort.SceneFromDict(d), 800, 600)
    ref, _, _ = oracle.Scene(d).render(800, 600, samples=1, max_depth=8, jitter=False, soft_shadows=False)
    check(img, ref)


def test_stream_shards_compose(gort, renderer):
    """tile shards rendered by the wavefront pipeline compose to the unsharded frame, bit for bit"""
    d = Cm.random_sphere_scene(2500, 3, cam_z=26.0)
    sc = gort.SceneFromDict(d)
    configure(renderer, 3, 6, seed=8)
    with forced_path("stream"):
        full = renderer.Render(sc, 200, 136).copy()
        out = np.zeros_like(full)
        for rank in range(3):
            renderer.SetShard(rank, 3)
            renderer.Render(sc, 200, 136, out=out)
        renderer.SetShard(0, 1)
    assert np.array_equal(full, out)


@pytest.mark.parametrize("objects,bvh", [(1, "host"), (3, "host"), (300, "host"), (300, "device"), (6000, "device")])
def test_stream_switches_keep_the_frame(gort, renderer, objects, bvh):
    """The staged top-of-tree block (GORT_TOP: the first four levels of the 4-wide tree in shared memory, one bulk copy per
    CTA; 2 = plain loads) and the sorted mode (GORT_SORT: scatter reads the queue in Morton order of the hit points) change
    where a node is read from and the order paths are processed in — never a path's arithmetic: the exact accumulators
    are equal, from a one-sphere tree (top block = the root alone) to one deeper than the block."""
    if objects <= 3:
        d = Cm.c1_view()
        d["objects"] = d["objects"][:objects]
    else:
        d = Cm.random_sphere_scene(objects, 40 + objects, cam_z=13.0 if objects < 1000 else 26.0)
    configure(renderer, 2, 6, seed=11)
    with forced_path("stream", GORT_BVH=bvh, GORT_TOP=0, GORT_SORT=0):
        sc = gort.SceneFromDict(d)
        renderer.Render(sc, 320, 240)
        assert renderer.lastStats.render_path == 2
        base = renderer.ReadRadiance(320, 240)
    assert base.max() > 0
    for env in ({"GORT_TOP": 1}, {"GORT_TOP": 2}, {"GORT_TOP": 0, "GORT_SORT": 1}, {"GORT_TOP": 1, "GORT_SORT": 1}, {"GORT_SORT": 3}):
        with forced_path("stream", GORT_BVH=bvh, **env):
            renderer.Render(sc, 320, 240)
            assert np.array_equal(renderer.ReadRadiance(320, 240), base), env


@pytest.mark.parametrize("path", ["queue", "stream"])
def test_sphere_cloud_golden_fixture(gort, renderer, path):
    """both BVH render paths against the committed oracle image of the 1 500-sphere cloud (tests/golden/make_golden.py)"""
    from PIL import Image
    ref = np.array(Image.open(os.path.join(Cm.ROOT, "tests", "golden", "sphere_cloud_1500_320x180_2spp_d4_seed9.png")).convert("RGBA"))
    configure(renderer, 2, 4, seed=9)
    with forced_path(path):
        img = renderer.Render(gort.SceneFromDict(Cm.random_sphere_scene(1500, 77, cam_z=13.0)), 320, 180)
    assert (ref[..., :3].sum(-1) > 0).mean() > 0.2
    check(img, ref, within=0.998, max_bad_lit=90)  # (the dense cloud's bar: see test_random_spheres_bvh_scene_same_stream)
