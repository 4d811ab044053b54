"""Edge cases of the render path through the C ABI: empty and ragged inputs, extreme parameters,
sharding properties at full size, determinism, and error codes."""
import numpy as np
import pytest
import torch

import common as Cm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer(gort):
    r = gort.NewParallelRenderer(1)
    yield r
    r.close()


def setup(r, spp=2, depth=10, seed=1):
    r.SetSamples(spp); r.SetMaxDepth(depth); r.SetAntiAliasing(True); r.SetSoftShadows(True); r.SetRecursiveReflections(True)
    r.SetSeed(seed); r.SetCameraMode(0); r.SetShard(0, 1); r.SetCollectStats(False)


def test_empty_scene_is_black(gort, renderer):
    setup(renderer)
    img = renderer.Render(gort.SceneFromDict({"camera": {"position": [0, 0, 5], "aspectRatio": 1.0}, "objects": [], "lights": []}), 100, 60)
    assert (img[..., :3] == 0).all() and (img[..., 3] == 255).all()


def test_no_lights_gives_ambient_only(gort, oracle, renderer):
    d = {"camera": {"position": [0, 0, 5], "aspectRatio": 1.0}, "lights": [],
         "objects": [{"type": "sphere", "position": [0, 0, 0], "radius": 1.5, "material": {"type": "lambertian", "color": [0.5, 0.5, 0.5]}}]}
    setup(renderer, 1, 1)
    renderer.SetAntiAliasing(False)
    img = renderer.Render(gort.SceneFromDict(d), 64, 64)
    assert tuple(img[32, 32]) == (87, 87, 87, 255)  # ambient 0.1 -> 87 (tone-map table)
    assert tuple(img[0, 0]) == (0, 0, 0, 255)


@pytest.mark.parametrize("w,h", [(1, 1), (31, 33), (70, 45), (257, 3), (32, 32)])
def test_ragged_sizes_match_oracle(gort, oracle, renderer, w, h):
    d = Cm.c1_view()
    d["camera"]["position"] = [0, 0, 3.5]
    setup(renderer, 3, 6, seed=8)
    img = renderer.Render(gort.SceneFromDict(d), w, h)
    ref, _, _ = oracle.Scene(d).render(w, h, samples=3, max_depth=6, rng_mode=oracle.RNG_PHILOX, seed=8)
    assert img.shape == (h, w, 4)
    assert Cm.within_one(img, ref) >= 0.99 and (img[..., 3] == 255).all()


def test_max_depth_zero_is_black(gort, renderer):
    setup(renderer, 2, 0)
    img = renderer.Render(gort.SceneFromDict(Cm.c1_view()), 200, 150)
    assert (img[..., :3] == 0).all()


def test_determinism_and_seed(gort, renderer):
    sc = gort.SceneFromDict(Cm.c1_view())
    setup(renderer, 16, 50, seed=5)
    a = renderer.Render(sc, 800, 600).copy()
    b = renderer.Render(sc, 800, 600).copy()
    assert (a == b).all()  # fixed-point accumulation: bit-reproducible for any schedule
    renderer.SetSeed(6)
    c = renderer.Render(sc, 800, 600)
    assert (a != c).any() and Cm.psnr(a, c) > 30


def test_shards_compose_to_the_full_frame(gort, renderer):
    """Property at the full C1 size: the union of the statically interleaved shards (slab -> unswizzle)
    is bit-identical to the single-shard frame for 2, 3 and 8 shards; host-destination shards too."""
    W, H = 800, 600
    sc = gort.SceneFromDict(Cm.c1_view())
    setup(renderer, 8, 50, seed=3)
    full = renderer.Render(sc, W, H).copy()
    for n in (2, 3, 8):
        per = gort.shard_slab_bytes(W, H, n)
        slabs = torch.zeros(n * per, dtype=torch.uint8, device="cuda")
        for rank in range(n):
            renderer.SetShard(rank, n)
            renderer.RenderShardDevice(W, H, slabs.data_ptr() + rank * per)
        out = torch.zeros(H * W * 4, dtype=torch.uint8, device="cuda")
        renderer.UnswizzleDevice(slabs.data_ptr(), n, W, H, out.data_ptr())
        torch.cuda.synchronize()
        got = out.cpu().numpy().reshape(H, W, 4)
        assert (got == full).all()
        # numpy restatement of the un-swizzle used by the gloo tests agrees with the kernel
        assert (gort.unswizzle_host(slabs.cpu().numpy(), n, W, H) == full).all()
    # host destination: each shard writes only its own tiles
    acc = np.zeros((H, W, 4), dtype=np.uint8)
    for rank in range(4):
        renderer.SetShard(rank, 4)
        renderer.Render(sc, W, H, out=acc)
    assert (acc == full).all()
    renderer.SetShard(0, 1)


def test_render_device_matches_host_path(gort, renderer):
    W, H = 640, 360
    sc = gort.SceneFromDict(Cm.c2_view())
    setup(renderer, 2, 8, seed=2)
    host = renderer.Render(sc, W, H).copy()
    out = torch.zeros(H * W * 4, dtype=torch.uint8, device="cuda")
    st = renderer.RenderDevice(W, H, out.data_ptr(), want_stats=True)
    assert st.kernel_ms > 0 and st.primary_rays == W * H * 2
    # the kernel split comes from stamps the kernels write (no event between them): consistent with the frame's event time
    assert st.trace_ms > 0 and st.cull_ms > 0 and st.resolve_ms >= 0
    assert abs(st.cull_ms + st.trace_ms + st.resolve_ms - st.kernel_ms) < 1e-3 and st.trace_ms < st.kernel_ms
    assert (out.cpu().numpy().reshape(H, W, 4) == host).all()


def test_stats_counters(gort, renderer):
    sc = gort.SceneFromDict(Cm.c1_view())
    setup(renderer, 4, 50, seed=1)
    renderer.SetCollectStats(True)
    a = renderer.Render(sc, 400, 300).copy()
    st = renderer.lastStats
    assert st.primary_rays == 400 * 300 * 4
    # the cull pass removes pixel blocks that cannot see geometry, so fewer closest-hit queries than samples
    assert 0 < st.closest_queries < st.primary_rays and st.shadow_queries == st.light_evals + st.soft_shadow_rays
    assert st.soft_shadow_rays % 16 == 0 and st.shaded_hits > 0 and st.algorithmic_flops > 0
    assert st.sphere_tests > 0 and st.tri_tests == 0
    # the exact culls of the shade stage: pairs facing away from their light, lit pairs with an empty shadow cone
    assert st.pairs_backfacing > 0 and st.soft_pairs_skipped > 0
    assert st.pairs_backfacing + st.light_evals + st.soft_pairs_skipped <= st.shaded_hits * 2  # two lights
    renderer.SetCollectStats(False)
    b = renderer.Render(sc, 400, 300)
    assert (a == b).all()  # the counting variant renders the same image


def test_benchmark_data_schema(gort, tmp_path):
    """BenchmarkData (renderer.go:31-42) is recorded by every Render and written by SaveBenchmarkData (renderer.go:119-126)."""
    import json
    r = gort.NewParallelRenderer(1)
    try:
        assert r.benchmarkData == {}
        setup(r, 2, 5, seed=3)
        r.Render(gort.SceneFromDict(Cm.c1_view()), 64, 48)
        b = r.benchmarkData
        assert sorted(b) == sorted(["scene_name", "resolution", "render_time_seconds", "samples", "max_depth", "num_workers", "objects",
                                    "lights", "timestamp", "features"])
        assert b["resolution"] == "64x48" and b["samples"] == 2 and b["max_depth"] == 5 and b["objects"] == 5 and b["lights"] == 2
        assert b["render_time_seconds"] > 0 and len(b["features"]) == 4
        path = str(tmp_path / "benchmark_data.json")
        r.SaveBenchmarkData(path)
        assert json.load(open(path))["resolution"] == "64x48"
    finally:
        r.close()


def test_error_codes(gort, renderer):
    sc = gort.SceneFromDict(Cm.c3())
    setup(renderer)
    for bad in [(0, 10), (10, -1)]:
        with pytest.raises(gort.GortError) as e:
            renderer.Render(sc, *bad, out=np.zeros((1, 1, 4), dtype=np.uint8)) if False else renderer.RenderDevice(bad[0], bad[1], 0)
        assert e.value.code == -1
    renderer.SetSamples(0)
    with pytest.raises(gort.GortError):
        renderer.Render(sc, 16, 16)
    renderer.SetSamples(1)
    r2 = gort.NewParallelRenderer(1)
    with pytest.raises(gort.GortError) as e:
        r2.RenderDevice(16, 16, 0)
    assert e.value.code == -4  # GORT_ERR_NO_SCENE
    with pytest.raises(gort.GortError) as e:
        r2.UploadScene(gort.Scene('{"objects": [{"type": "sphere", "material": {}}]}'))
    assert e.value.code == -6  # GORT_ERR_PARSE (the reference panics here)
    # float64 coordinates that leave the fp32 range of the device path are refused at upload, not mis-rendered; the context
    # has no scene afterwards
    far = Cm.c1_view()
    far["objects"][0]["position"] = [1e35, 0, 0]
    with pytest.raises(gort.GortError, match="fp32") as e:
        r2.UploadScene(gort.SceneFromDict(far))
    assert e.value.code == -1
    with pytest.raises(gort.GortError) as e:
        r2.RenderDevice(16, 16, 0)
    assert e.value.code == -4
    r2.close()
    with pytest.raises(gort.GortError):
        gort.LoadFromFile("/nonexistent/scene.json")


def test_frame_link_assembles_the_frame_in_the_owners_memory(gort, renderer):
    """Frame link (include/gort.h): every rank's resolve kernel stores its tiles straight into the owner's row-major
    frame.  On this single-GPU box the three ranks are three contexts of one process (gort_link_open_local stands in
    for the CUDA-IPC mapping) run one after the other, peers first, so no kernel ever waits for another."""
    W, H, n = 200, 150, 3
    sc = gort.SceneFromDict(Cm.c2_view(), 1)
    renderer.SetSamples(4); renderer.SetMaxDepth(10); renderer.SetSeed(17); renderer.SetShard(0, 1)
    full = renderer.Render(sc, W, H).copy()
    ranks = [gort.NewParallelRenderer(1) for _ in range(n)]
    for r in ranks:
        r.SetSamples(4); r.SetMaxDepth(10); r.SetSeed(17)
        r.UploadScene(sc)
    link0, handle = ranks[0].LinkCreate(W, H, n)
    assert len(handle) == 64
    links = [link0] + [ranks[k].LinkOpenLocal(link0, k) for k in range(1, n)]
    for k in (2, 1, 0):  # peers first: their tiles and arrival counts are in place when the owner waits for them
        st = ranks[k].RenderLinked(W, H, links[k], want_stats=True)
        assert st.n_tiles == len(gort.tiles_of_shard(W, H, k, n))
    img = ranks[0].LinkRead(link0, W, H)
    assert np.array_equal(img, full)
    # frame 2 through the same links; then the wrong order (owner first) is refused on the host — nothing is enqueued that
    # would spin on the GPU waiting for ranks that cannot run beside it — and the right order still works afterwards
    for k in (1, 2, 0):
        ranks[k].RenderLinked(W, H, links[k])
    assert np.array_equal(ranks[0].LinkRead(link0, W, H), full)
    with pytest.raises(gort.GortError) as e:
        ranks[0].RenderLinked(W, H, links[0])
    assert "before the owner" in str(e.value)
    for k in (2, 1, 0):
        ranks[k].RenderLinked(W, H, links[k])
    assert np.array_equal(ranks[0].LinkRead(link0, W, H), full)
    for k in (2, 1, 0):
        ranks[k].LinkClose(links[k])
        ranks[k].close()


def test_multi_device_context_matches_single_device(gort, renderer):
    """gort_create with several devices (single process): tiles interleaved over the GPUs, every device resolving
    straight into the lead device's frame over NVLink peer access (or slabs + peer copies + un-swizzle when
    GORT_NO_PEER_DIRECT is set) — bit-identical to one GPU.  Needs >= 2 GPUs."""
    import os
    n = gort.load_library().gort_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    W, H = 330, 250
    sc = gort.SceneFromDict(Cm.c1_view())
    renderer.SetSamples(8); renderer.SetMaxDepth(50); renderer.SetSeed(23); renderer.SetShard(0, 1)
    full = renderer.Render(sc, W, H).copy()
    for nd in sorted({2, min(n, 8)}):
        for no_direct in (False, True):
            if no_direct:
                os.environ["GORT_NO_PEER_DIRECT"] = "1"
            try:
                r = gort.NewParallelRenderer(nd, devices=list(range(nd)))
            finally:
                os.environ.pop("GORT_NO_PEER_DIRECT", None)
            r.SetSamples(8); r.SetMaxDepth(50); r.SetSeed(23)
            img = r.Render(sc, W, H)
            assert r.lastStats.n_devices == nd
            assert np.array_equal(img, full), (nd, no_direct)
            r.close()


def test_host_frame_kinds_agree(gort, renderer):
    """gort_render into (a) ordinary pageable memory (staged through the ctx's page-locked frame), (b) a page-locked
    HostFrame (resolve stores straight into it, culled blocks early on a second stream), (c) the zero-copy path
    switched off (device frame + D2H copy): the same bytes."""
    import os
    sc = gort.SceneFromDict(Cm.c1_view())
    renderer.SetSamples(6); renderer.SetMaxDepth(50); renderer.SetSeed(31); renderer.SetShard(0, 1)
    W, H = 333, 257
    a = renderer.Render(sc, W, H).copy()
    hf = gort.HostFrame(W, H)
    hf.array[:] = 7
    b = renderer.Render(sc, W, H, out=hf.array).copy()
    os.environ["GORT_NO_ZERO_COPY"] = "1"
    try:
        c = renderer.Render(sc, W, H).copy()
    finally:
        del os.environ["GORT_NO_ZERO_COPY"]
    hf.close()
    assert (a[..., 3] == 255).all() and a[..., :3].max() > 0
    assert np.array_equal(a, b) and np.array_equal(a, c)
