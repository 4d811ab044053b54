"""Host logic of the product (no GPU): the C++ scene loader (mirror of scene.LoadFromFile/GetHittables,
/root/reference internal/scene/scene.go:45-190) against the oracle's independent expansion, the BVH
builder's invariants, and the error behaviour where the reference would panic."""
import json

import numpy as np
import pytest

import common as Cm


@pytest.mark.parametrize("name,prisms", [("sphere_reflections_light.json", False), ("final_silver_prism_purple_cube_.json", False),
                                         ("final_silver_prism_purple_cube_.json", True), ("two_red_cubes_scene.json", False)])
def test_loader_matches_oracle_expansion(gort, oracle, name, prisms):
    d = Cm.load_scene_dict(name)
    hs = gort.HostScene(json.dumps(d), gort.LOAD_PRISMS if prisms else 0)
    os_ = oracle.Scene(d, prisms=prisms)
    hc, oc = hs.counts(), os_.counts()
    assert (hc["spheres"], hc["triangles"], hc["hittables"], hc["lights"]) == (oc["spheres"], oc["triangles"], oc["hittables"], oc["lights"])
    for i in range(hc["triangles"]):
        v9, mat, order = hs.triangle(i)
        o12, omat = os_.triangle(i)
        assert v9.tolist() == o12[:9].tolist() and mat == omat
    for i in range(hc["materials"]):
        t, v = hs.material(i)
        ot, ov = os_.material(i)
        assert t == ot
        assert v[:3].tolist() == ov[:3].tolist()
        if t in (1, 2):  # metal / shiny carry all four scalars
            assert v[3:].tolist() == ov[3:].tolist()
        if t in (4, 5):
            assert v[6] == ov[6]
    # scan order is a permutation: spheres and triangles interleaved in object order
    orders = [hs.sphere(i)[3] for i in range(hc["spheres"])] + [hs.triangle(i)[2] for i in range(hc["triangles"])]
    assert sorted(orders) == list(range(len(orders)))


def test_missing_color_default_and_camera(gort):
    hs = gort.HostScene(open(Cm.SCENES + "/sphere_reflections_light.json").read())
    t, v = hs.material(1)
    assert t == 1 and v[:3].tolist() == [1, 1, 1] and v[4] == 1.0  # F5 default albedo, metallic default 1
    cam = hs.camera()
    assert cam[:3].tolist() == [0, 0, -8] and cam[9] == 60 and cam[10] == 1.33


def test_vec3_object_form(gort):  # Vec3.UnmarshalJSON accepts {"X":..,"Y":..,"Z":..} (vector.go:185-192)
    d = {"camera": {"position": {"X": 1, "Y": 2, "Z": 3}}, "objects": [], "lights": [{"position": [1, 2, 3], "color": {"X": 0.5, "Y": 0.25, "Z": 1}, "intensity": 2}]}
    hs = gort.HostScene(json.dumps(d))
    assert hs.camera()[:3].tolist() == [1, 2, 3]
    assert hs.light(0).tolist() == [1, 2, 3, 0.5, 0.25, 1, 2]


@pytest.mark.parametrize("bad", [
    '{"objects": [{"type": "sphere", "radius": 1, "material": {"color": [1,1,1]}}]}',            # no material.type -> panic scene.go:105
    '{"objects": [{"type": "sphere", "radius": 1}]}',                                               # nil material map
    '{"objects": [{"type": "cube", "material": {"type": "metal", "color": "red"}}]}',              # colour not an array
    '{"objects": [{"type": "sphere", "material": {"type": "metal", "color": [1,1,1], "roughness": "x"}}]}',
    '{"objects": [{"type": "sphere", "position": [1,2], "material": {"type": "metal", "color": [1,1,1]}}]}',  # Vec3 needs 3
    '{"objects": [',
    '[]',
])
def test_loader_errors_where_reference_fails(gort, bad):
    with pytest.raises(gort.GortError) as e:
        gort.HostScene(bad)
    assert e.value.code == -6


def test_unknown_object_with_bad_material_is_skipped(gort):
    # GetHittables only calls createMaterial for sphere/cube (scene.go:69-83)
    hs = gort.HostScene('{"objects": [{"type": "torus"}, {"type": "triangularPrism", "material": 5}]}')
    assert hs.counts()["hittables"] == 0


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (5, 2), (64, 3), (1000, 4), (20000, 5)])
def test_bvh_invariants_random_spheres(gort, n, seed):
    d = Cm.random_sphere_scene(n, seed)
    info = gort.HostScene(json.dumps(d)).bvh_validate()
    assert info["max_depth"] <= 56
    assert info["leaves"] >= (n + 3) // 4


def test_bvh_invariants_mixed_and_degenerate(gort):
    rng = np.random.default_rng(9)
    objs = []
    for i in range(300):
        m = {"type": "lambertian", "color": [1, 1, 1]}
        if i % 2:
            objs.append({"type": "cube", "position": rng.uniform(-5, 5, 3).tolist(), "size": rng.uniform(0.1, 1, 3).tolist(), "material": m})
        else:
            objs.append({"type": "sphere", "position": rng.uniform(-5, 5, 3).tolist(), "radius": 0.3, "material": m})
    # degenerate: many identical centroids, zero-size cube, zero-radius sphere
    for _ in range(40):
        objs.append({"type": "sphere", "position": [1, 1, 1], "radius": 0.5, "material": {"type": "glass", "color": [1, 1, 1]}})
    objs.append({"type": "cube", "position": [0, 0, 0], "size": [0, 0, 0], "material": {"type": "metal", "color": [1, 1, 1]}})
    objs.append({"type": "sphere", "position": [0, 0, 0], "radius": 0.0, "material": {"type": "metal", "color": [1, 1, 1]}})
    hs = gort.HostScene(json.dumps({"objects": objs}))
    info = hs.bvh_validate()
    assert hs.counts()["triangles"] == 151 * 12 and info["max_depth"] <= 56


def test_flat_scene_round_trip_types(gort):
    hs = gort.HostScene(open(Cm.SCENES + "/final_silver_prism_purple_cube_.json").read(), gort.LOAD_PRISMS)
    flat = hs.to_flat()
    assert flat.desc.n_triangles == 40 and flat.desc.n_materials == 4 and flat.desc.n_lights == 3
    assert flat.desc.abi_version == gort.ABI_VERSION


def test_renderer_block_is_parsed_but_never_applied(gort):
    """README.md:285-291 advertises a per-scene "renderer" block that the reference loader drops (scene.go:12-16).
    The loader keeps it as hints for hosts that opt in (raytracer -scene-settings); nothing else changes."""
    import json
    d = Cm.load_scene_dict("final_silver_prism_purple_cube_.json")
    assert d["renderer"]["samples"] == 200 and d["renderer"]["maxDepth"] == 20
    a = gort.HostScene(json.dumps(d))
    del d["renderer"]
    b = gort.HostScene(json.dumps(d))
    assert a.counts() == b.counts()


def test_parallel_bvh_build_equals_sequential(gort):
    """Scenes of >= 20 000 primitives build the lower subtrees on worker threads; every split depends only on its own
    primitive range, so the tree (inner nodes, depth, leaves, bytes) is the sequential one and passes the invariant check."""
    import json
    import os
    # 30 000: worker threads build the subtrees; 70 000: the top of the tree also scans its ranges (>= 65 536 primitives) in
    # parallel chunks and the flatten pass fills the arrays in parallel
    for n, seed in ((30000, 11), (70000, 12)):
        d = Cm.random_sphere_scene(n, seed, extent=40.0)
        hs = gort.HostScene(json.dumps(d))
        infos = []
        for th in ("1", "7"):
            os.environ["GORT_BVH_THREADS"] = th
            try:
                infos.append(hs.bvh_validate())
            finally:
                del os.environ["GORT_BVH_THREADS"]
        assert infos[0] == infos[1] and infos[0]["leaves"] == infos[0]["nodes"] + 1


def test_scene_from_desc_host_side(gort):
    """gort_scene_upload's host half (gort_host_scene_from_desc: the same validation and copy, no device): a description of
    > 100 000 primitives is copied by all host threads into arrays that resize() leaves untouched — every element must come
    out as it went in —, a rebuilt scene reuses its arrays whatever the sizes, bad descriptions are refused."""
    import synth
    a = synth.scene_arrays(110_000, 1_500, 50.0, 5, 120.0, fog={"density": 0.01, "color": (0.2, 0.3, 0.4)})
    flat = synth.to_gort(a)
    hs = gort.HostScene.from_desc(flat)
    c = hs.counts()
    assert c == {"spheres": 110_000, "triangles": 18_000, "materials": 111_500, "lights": 3, "hittables": 110_000 + 1_500}
    rng = np.random.default_rng(3)
    for i in rng.integers(0, 110_000, 400).tolist() + [0, 109_999]:
        ctr, r, m, o = hs.sphere(i)
        assert np.array_equal(ctr, a["spheres"][i][0]) and (r, m, o) == a["spheres"][i][1:]
    for i in rng.integers(0, 18_000, 400).tolist() + [0, 17_999]:
        v, m, o = hs.triangle(i)
        assert np.array_equal(v, np.asarray(a["triangles"][i][0]).reshape(-1)) and (m, o) == a["triangles"][i][1:]
    for i in rng.integers(0, 111_500, 400).tolist() + [0, 111_499]:
        t, v = hs.material(i)
        want = a["materials"][i]
        assert t == want["type"] and v[6] == want.get("ior", 1.5)
        if t == gort.MAT_TYPES["perfectmirror"]:
            assert v[4] == 1.0  # GetMetallic (advanced_materials.go:165)
        if t != gort.MAT_TYPES["dielectric"]:
            assert np.array_equal(v[:3], want.get("color", (1, 1, 1)))
    assert np.array_equal(hs.light(2), np.array([0.0, 80.0, 0.0, 1, 1, 1, 3000.0]))
    assert hs.bvh_validate()["nodes"] > 0
    # rebuilt in place: smaller, then larger than ever (arrays shrink and grow), and the same answers as a fresh scene
    small = synth.to_gort(synth.scene_arrays(50, 3, 10.0, 6, 30.0))
    gort.HostScene.from_desc(small, into=hs)
    assert hs.counts()["spheres"] == 50 and hs.counts()["triangles"] == 36 and hs.counts()["lights"] == 3
    fresh = gort.HostScene.from_desc(small)
    assert all(np.array_equal(hs.sphere(i)[0], fresh.sphere(i)[0]) for i in range(50))
    b = synth.scene_arrays(130_000, 0, 50.0, 7, 120.0)
    gort.HostScene.from_desc(synth.to_gort(b), into=hs)
    assert hs.counts()["spheres"] == 130_000 and hs.counts()["triangles"] == 0
    assert np.array_equal(hs.sphere(129_999)[0], b["spheres"][129_999][0])
    # refused: an order that is not a permutation, a material index out of range, a missing array, another ABI
    for field, value, msg in (("sphere_order", 1, "permutation"), ("sphere_material", 10**6, "material index")):
        bad = synth.to_gort(synth.scene_arrays(200, 0, 10.0, 8, 30.0))
        getattr(bad.desc, field)[7] = value
        with pytest.raises(gort.GortError, match=msg):
            gort.HostScene.from_desc(bad)
    bad = synth.to_gort(synth.scene_arrays(200, 0, 10.0, 8, 30.0))
    bad.desc.sphere_radius = None
    with pytest.raises(gort.GortError, match="sphere arrays missing"):
        gort.HostScene.from_desc(bad)
    bad = synth.to_gort(synth.scene_arrays(200, 0, 10.0, 8, 30.0))
    bad.desc.abi_version = gort.ABI_VERSION + 1
    with pytest.raises(gort.GortError, match="abi_version"):
        gort.HostScene.from_desc(bad)


def test_loader_survives_hostile_json(gort):
    """Scene text can come from the network (chunk farm): whatever it is, the loader answers with a scene or an error.
    Nesting is bounded (the parser recurses per level; Go's decoder has a bound too), truncated and binary input is refused."""
    for n, ok in ((100, True), (999, True), (1000, False), (300_000, False)):
        text = '{"objects":' + "[" * n + "]" * n + "}"
        if ok:
            assert gort.HostScene(text).counts()["hittables"] == 0
        else:
            with pytest.raises(gort.GortError, match="exceeded max depth"):
                gort.HostScene(text)
    with pytest.raises(gort.GortError, match="exceeded max depth"):
        gort.HostScene('{"a":' * 5000)
    for text in ('{"objects":[', "[" * 50 + "1" + "]" * 49, "\x00\xff", "", '{"objects":[{"type":"sphere"', '{"lights":[{"position":"x"}]}',
                 '{"objects":[{"type":"cube","position":[0,0,0],"size":[1,1],"material":{"type":"metal","color":[1,1,1]}}]}'):
        with pytest.raises(gort.GortError):
            gort.HostScene(text)

    from hypothesis import given, settings, strategies as st

    leaves = st.one_of(st.none(), st.booleans(), st.floats(allow_nan=False, allow_infinity=False), st.integers(-10**6, 10**6), st.text(max_size=8))
    values = st.recursive(leaves, lambda c: st.one_of(st.lists(c, max_size=4), st.dictionaries(
        st.sampled_from(["type", "position", "radius", "size", "material", "color", "vertices", "intensity", "x"]), c, max_size=5)), max_leaves=25)
    scenes = st.fixed_dictionaries({}, optional={"camera": values, "objects": st.lists(values, max_size=4), "lights": st.lists(values, max_size=3),
                                                 "fog": values, "sky": values, "renderer": values})

    @settings(max_examples=300, deadline=None)
    @given(scenes, st.integers(0, 7))
    def structured(doc, options):
        try:
            gort.HostScene(json.dumps(doc), options).counts()
        except gort.GortError:
            pass

    @settings(max_examples=200, deadline=None)
    @given(st.text(alphabet='{}[]":,0123456789.eE-+tfn aobjectslighpr\\u', max_size=60))
    def soup(text):
        try:
            gort.HostScene(text)
        except gort.GortError:
            pass

    structured()
    soup()
