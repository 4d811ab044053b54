"""The reference's distributed-render protocol (/root/reference internal/distributed/distributed_renderer.go): RenderChunk in,
RemoteResult out over POST /render; NodeInfo over GET /status; a dispatcher that farms the chunks of a frame out to nodes.
CPU part: the protocol itself against nodes with a stand-in renderer.  GPU part: a node backed by libgort returns the
very pixels of the full frame, whatever the chunking."""
import importlib
import json
import urllib.error
import urllib.request

import numpy as np
import pytest

import common as Cm

CF = importlib.import_module("concurrent-raytracer-go_b200.chunkfarm")


def fake_render(chunk):
    """deterministic stand-in: r = x, g = y, b = chunk id, a = 255"""
    w, h = chunk["end_x"] - chunk["start_x"], chunk["end_y"] - chunk["start_y"]
    ys, xs = np.mgrid[chunk["start_y"]:chunk["end_y"], chunk["start_x"]:chunk["end_x"]]
    out = np.zeros((h, w, 4), dtype=np.uint8)
    out[..., 0], out[..., 1], out[..., 2], out[..., 3] = xs & 255, ys & 255, chunk["id"] & 255, 255
    return out


@pytest.fixture()
def nodes():
    ns = [CF.ChunkNode(fake_render).start() for _ in range(2)]
    yield ns
    for n in ns:
        n.stop()


def test_render_and_status_wire_format(nodes):
    n = nodes[0]
    addr = "127.0.0.1:%d" % n.port
    chunk = {"id": 7, "start_x": 4, "end_x": 9, "start_y": 2, "end_y": 5, "width": 16, "height": 8, "scene": "s.json", "priority": 1}
    dr = CF.DistributedRenderer([addr])
    res = dr.RenderChunkRemotely(chunk, addr)
    assert set(res) == {"chunk_id", "pixels", "duration", "node_id"}  # RemoteResult (:41-47); "error" is omitempty
    assert res["chunk_id"] == 7 and res["node_id"] == "node-%d" % n.port and res["duration"] >= 0
    assert len(res["pixels"]) == 5 * 3
    assert res["pixels"][0] == {"x": 4, "y": 2, "r": 4, "g": 2, "b": 7, "a": 255}
    assert res["pixels"][-1] == {"x": 8, "y": 4, "r": 8, "g": 4, "b": 7, "a": 255}
    info = dr.GetNodeInfo(addr)
    assert set(info) == {"id", "cpu_usage", "memory_usage", "active_jobs", "max_jobs", "load_average"}  # NodeInfo (:54-61)
    assert info["id"] == "node-%d" % n.port and info["active_jobs"] == 0 and info["max_jobs"] == 8 and info["memory_usage"] > 0
    # the compact extension encoding carries the same rectangle
    b64 = dr.RenderChunkRemotely(dict(chunk, encoding="rgba_b64"), addr)
    assert np.array_equal(CF.decode_pixels(b64, chunk), CF.decode_pixels(res, chunk))


def test_http_errors_follow_the_reference_server(nodes):
    base = "http://127.0.0.1:%d" % nodes[0].port
    with pytest.raises(urllib.error.HTTPError) as e:  # GET /render: "Method not allowed" (:259-262)
        urllib.request.urlopen(base + "/render")
    assert e.value.code == 405
    req = urllib.request.Request(base + "/render", data=b"{not json", method="POST")
    with pytest.raises(urllib.error.HTTPError) as e:  # "Invalid request body" (:264-267)
        urllib.request.urlopen(req)
    assert e.value.code == 400
    # a well-formed body with a bad rectangle is answered with RemoteResult.Error, not an HTTP error
    bad = {"id": 1, "start_x": 5, "end_x": 3, "start_y": 0, "end_y": 1, "width": 8, "height": 8, "scene": "", "priority": 0}
    req = urllib.request.Request(base + "/render", data=json.dumps(bad).encode(), method="POST")
    res = json.loads(urllib.request.urlopen(req).read())
    assert "outside the frame" in res["error"] and res["pixels"] == []
    # what a request may ask of a node is bounded: frame size, body size
    huge = dict(bad, start_x=0, end_x=1, width=65535, height=65535)
    req = urllib.request.Request(base + "/render", data=json.dumps(huge).encode(), method="POST")
    assert "frame too large" in json.loads(urllib.request.urlopen(req).read())["error"]
    req = urllib.request.Request(base + "/render", data=b"{}", method="POST", headers={"Content-Length": str(CF.MAX_BODY_BYTES + 1)})
    with pytest.raises(urllib.error.HTTPError) as e:
        urllib.request.urlopen(req)
    assert e.value.code == 413


def test_scene_paths_stay_inside_the_nodes_scene_directory(tmp_path):
    """a request names its scene by path: the node opens it under its scene root only"""
    root = tmp_path / "scenes"
    (root / "sub").mkdir(parents=True)
    (root / "sub" / "a.json").write_text("{}")
    (tmp_path / "secret.json").write_text("{}")
    assert CF.resolve_scene_path(str(root), "sub/a.json") == str((root / "sub" / "a.json").resolve())
    for name in ("../secret.json", "sub/../../secret.json", str(tmp_path / "secret.json"), "/etc/passwd"):
        with pytest.raises(ValueError):
            CF.resolve_scene_path(str(root), name)
    (root / "link.json").symlink_to(tmp_path / "secret.json")
    with pytest.raises(ValueError):
        CF.resolve_scene_path(str(root), "link.json")


def test_dispatcher_assembles_the_frame_from_two_nodes(nodes):
    W, H = 70, 45  # ragged against the 32x16 chunks
    chunks = CF.make_chunks(W, H, 32, 16, "scene.json")
    assert len(chunks) == 3 * 3 and chunks[-1]["end_x"] == W and chunks[-1]["end_y"] == H
    dr = CF.DistributedRenderer(["127.0.0.1:%d" % n.port for n in nodes])
    results = dr.DistributeWork(chunks)
    img = CF.assemble(chunks, results, W, H)
    want = np.zeros((H, W, 4), dtype=np.uint8)
    for c in chunks:
        want[c["start_y"]:c["end_y"], c["start_x"]:c["end_x"]] = fake_render(c)
    assert np.array_equal(img, want)
    st = dr.GetStats()
    assert st["remote_jobs"] == 9 and st["failed_jobs"] == 0 and st["success_rate"] == 100.0 and st["total_nodes"] == 2
    assert {r["node_id"] for r in results} == {"node-%d" % n.port for n in nodes}  # both nodes took work


def test_failed_node_is_counted(nodes):
    dr = CF.DistributedRenderer(["127.0.0.1:9"], timeout=0.5)  # nothing listens there
    with pytest.raises(Exception):
        dr.RenderChunkRemotely(CF.make_chunks(8, 8, 8, 8, "s")[0], "127.0.0.1:9")
    assert dr.GetStats()["failed_jobs"] == 1 and dr.GetStats()["success_rate"] == 0.0


@pytest.mark.gpu
def test_gpu_node_returns_the_full_frames_pixels(gort):
    """a libgort-backed node: the chunks of a frame, rendered one region at a time, are the full frame bit for bit"""
    d = Cm.c2_view()
    W, H = 240, 180
    r = gort.NewParallelRenderer(1)
    r.SetSamples(4); r.SetMaxDepth(10); r.SetSeed(5)
    full = r.Render(gort.SceneFromDict(d), W, H).copy()
    # region render through the C ABI: only the blocks that touch the rectangle are traced, the rest is black
    r.SetCrop(100, 70, 150, 110)
    part = r.Render(gort.SceneFromDict(d), W, H).copy()
    r.SetCrop()
    assert np.array_equal(part[70:110, 100:150], full[70:110, 100:150])
    assert (part[:64, :96, :3] == 0).all() and (full[..., :3].sum(-1) > 0).any()
    r.close()
    node = CF.ChunkNode(CF.GpuChunkRenderer(samples=4, max_depth=10, seed=5)).start()
    try:
        chunks = CF.make_chunks(W, H, 64, 48, json.dumps(d))  # the scene travels inline
        dr = CF.DistributedRenderer(["127.0.0.1:%d" % node.port])
        img = CF.assemble(chunks, dr.DistributeWork(chunks), W, H)
    finally:
        node.stop()
        node.render.close()
    assert np.array_equal(img, full)
