"""HTTP chunk-farm compatibility: a B200 node that speaks the reference's distributed-render protocol.

The reference ships a dispatcher and a stub worker (/root/reference internal/distributed/distributed_renderer.go):
`DistributedRenderer.RenderChunkRemotely` POSTs a `RenderChunk` as JSON to `http://<node>/render` (:76-106) and expects a
`RemoteResult` back (:41-47); `GetNodeInfo` GETs `/status` and expects a `NodeInfo` (:54-61, :132-150).  The reference's own
`RemoteRenderServer` sleeps 100 ms and returns no pixels (:258-283).  `ChunkNode` below is that server backed by libgort: the
chunk's pixel rectangle is rendered on the GPU (region render, `gort_render_params.crop_*`; pixels are bit-identical to the
same pixels of a full frame, whatever the chunking) and returned in the `pixels` array.

Wire format notes
  * request  : `{"id", "start_x", "end_x", "start_y", "end_y", "width", "height", "scene", "priority"}` (:29-39).  `scene` is a
    string the reference never interprets; here it is a scene-JSON path on the node, or the scene JSON text itself.
    Optional extension keys (absent in the reference): `samples`, `max_depth`, `seed`, `options` (loader option bits),
    `encoding: "rgba_b64"` (the rectangle as one base64 string of row-major RGBA8 instead of one JSON object per pixel).
  * response : `{"chunk_id", "pixels": [{"x","y","r","g","b","a"}, ...], "duration", "node_id"}` (+ `"error"` on failure).  The
    reference declares `Pixel` as `X, Y int `json:"x, y"`` / `R, G, B, A uint8 `json:"r, g, b, a"`` (:49-52): one tag shared by
    several fields, which encoding/json resolves by dropping them all — its pixels would marshal as `{}`.  The keys used here are
    the ones those tags evidently meant.
  * status   : `{"id", "cpu_usage", "memory_usage", "active_jobs", "max_jobs", "load_average"}` (:54-61), with real values.

`DistributedRenderer` mirrors the dispatcher side (same method names) so that the protocol can be exercised end to end
without Go; `assemble` turns the results into an image.RGBA-style [H, W, 4] array.
"""
from __future__ import annotations

import base64
import json
import os
import threading
import time
import urllib.request
from concurrent.futures import ThreadPoolExecutor
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Callable, Optional, Sequence

import numpy as np

CHUNK_KEYS = ("id", "start_x", "end_x", "start_y", "end_y", "width", "height", "scene", "priority")
# A node is reachable from the network: what a request may ask of it is bounded (the reference's stub renders nothing, so it
# has no such limits to mirror).
MAX_FRAME_PIXELS = 1 << 26   # width x height of the frame a chunk belongs to (the node allocates the frame per request)
MAX_BODY_BYTES = 16 << 20    # POST /render body (a scene JSON sent inline is a few MB at most)


def make_chunks(width: int, height: int, chunk_w: int, chunk_h: int, scene: str) -> list:
    """row-major rectangles covering the frame, as RenderChunk dicts"""
    out, k = [], 0
    for y in range(0, height, chunk_h):
        for x in range(0, width, chunk_w):
            out.append({"id": k, "start_x": x, "end_x": min(x + chunk_w, width), "start_y": y, "end_y": min(y + chunk_h, height),
                        "width": width, "height": height, "scene": scene, "priority": 0})
            k += 1
    return out


def validate_chunk(c: dict) -> Optional[str]:
    for k in CHUNK_KEYS[:7]:
        if not isinstance(c.get(k), int) or isinstance(c.get(k), bool):
            return "chunk field %s must be an integer" % k
    if c["width"] <= 0 or c["height"] <= 0:
        return "width/height must be positive"
    if c["width"] > 65535 or c["height"] > 65535 or c["width"] * c["height"] > MAX_FRAME_PIXELS:
        return "frame too large for this node (at most %d pixels)" % MAX_FRAME_PIXELS
    if not (0 <= c["start_x"] < c["end_x"] <= c["width"] and 0 <= c["start_y"] < c["end_y"] <= c["height"]):
        return "chunk rectangle outside the frame"
    if not isinstance(c.get("scene", ""), str):
        return "scene must be a string"
    return None


def encode_pixels(rect: np.ndarray, x0: int, y0: int, encoding: str = "json"):
    """rect: [h, w, 4] uint8 of the chunk -> the `pixels` value of a RemoteResult"""
    h, w = rect.shape[:2]
    if encoding == "rgba_b64":
        return base64.b64encode(np.ascontiguousarray(rect).tobytes()).decode()
    flat = rect.reshape(-1, 4).tolist()
    return [{"x": x0 + (i % w), "y": y0 + (i // w), "r": p[0], "g": p[1], "b": p[2], "a": p[3]} for i, p in enumerate(flat)]


def decode_pixels(result: dict, chunk: dict) -> np.ndarray:
    w, h = chunk["end_x"] - chunk["start_x"], chunk["end_y"] - chunk["start_y"]
    px = result["pixels"]
    if isinstance(px, str):
        return np.frombuffer(base64.b64decode(px), dtype=np.uint8).reshape(h, w, 4)
    rect = np.zeros((h, w, 4), dtype=np.uint8)
    for p in px:
        rect[p["y"] - chunk["start_y"], p["x"] - chunk["start_x"]] = (p["r"], p["g"], p["b"], p["a"])
    return rect


def resolve_scene_path(root: str, name: str) -> str:
    """A request names its scene file; the node opens it only inside its own scene directory (no absolute paths, no `..` or
    symlinks that lead out of it)."""
    root = os.path.realpath(root)
    path = os.path.realpath(os.path.join(root, name))
    if os.path.isabs(name) or os.path.commonpath([root, path]) != root:
        raise ValueError("scene path outside the node's scene directory")
    return path


class GpuChunkRenderer:
    """RenderChunk -> [h, w, 4] uint8 through libgort (one renderer per node; calls are serialised by ChunkNode)."""

    def __init__(self, samples: int = 100, max_depth: int = 50, seed: int = 0, options: int = 0, device: int = 0, scene_root: str = "."):
        from . import NewParallelRenderer, SceneFromDict  # noqa: F401  (raises GortError without a CUDA device: no CPU fallback)
        self._mod = __import__(__package__, fromlist=["x"])
        self.r = NewParallelRenderer(1, devices=[device])
        self.scene_root = os.path.realpath(scene_root)  # `scene` paths of a request are resolved inside this directory only
        self.defaults = {"samples": samples, "max_depth": max_depth, "seed": seed, "options": options}
        self._scene_key = None
        self._scene = None

    def _scene_for(self, text: str, options: int):
        key = (text, options)
        if key != self._scene_key:
            if text.lstrip().startswith("{"):
                self._scene = self._mod.Scene(text, options)
            else:
                self._scene = self._mod.LoadFromFile(resolve_scene_path(self.scene_root, text), options)
            self._scene_key = key
        return self._scene

    def __call__(self, chunk: dict) -> np.ndarray:
        cfg = {k: int(chunk.get(k, v)) for k, v in self.defaults.items()}
        sc = self._scene_for(chunk["scene"], cfg["options"])
        r = self.r
        r.SetSamples(cfg["samples"]); r.SetMaxDepth(cfg["max_depth"]); r.SetSeed(cfg["seed"])
        r.SetCrop(chunk["start_x"], chunk["start_y"], chunk["end_x"], chunk["end_y"])
        try:
            img = r.Render(sc, chunk["width"], chunk["height"])
        finally:
            r.SetCrop()
        return img[chunk["start_y"]:chunk["end_y"], chunk["start_x"]:chunk["end_x"]].copy()

    def close(self):
        self.r.close()


class ChunkNode:
    """RemoteRenderServer (distributed_renderer.go:238-302) with a real renderer behind /render."""

    def __init__(self, render: Callable[[dict], np.ndarray], port: int = 0, host: str = "127.0.0.1", max_jobs: int = 8, node_id: Optional[str] = None):
        self.render = render
        self.max_jobs = max_jobs
        self._lock = threading.Lock()  # a gort_ctx is externally synchronised: one Render at a time
        self._active = 0
        self._count_lock = threading.Lock()
        node = self

        class Handler(BaseHTTPRequestHandler):
            def log_message(self, *a):  # quiet
                pass

            def _json(self, code: int, obj):
                body = json.dumps(obj).encode()
                self.send_response(code)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def _plain(self, code: int, text: str):  # http.Error
                body = (text + "\n").encode()
                self.send_response(code)
                self.send_header("Content-Type", "text/plain; charset=utf-8")
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def do_POST(self):
                if self.path.split("?")[0] != "/render":
                    return self._plain(404, "404 page not found")
                try:
                    n = int(self.headers.get("Content-Length", "0"))
                    if n < 0 or n > MAX_BODY_BYTES:
                        return self._plain(413, "Request body too large")
                    chunk = json.loads(self.rfile.read(n))
                    if not isinstance(chunk, dict):
                        raise ValueError
                except (ValueError, json.JSONDecodeError):
                    return self._plain(400, "Invalid request body")  # :264-267
                self._json(200, node.handle_render(chunk))

            def do_GET(self):
                p = self.path.split("?")[0]
                if p == "/status":
                    return self._json(200, node.handle_status())
                if p == "/render":
                    return self._plain(405, "Method not allowed")  # :259-262
                self._plain(404, "404 page not found")

            def do_PUT(self):
                self._plain(405, "Method not allowed")

            do_DELETE = do_PUT

        self.server = ThreadingHTTPServer((host, port), Handler)
        self.port = self.server.server_address[1]
        self.node_id = node_id or "node-%d" % self.port  # "node-" + port, :276
        self._thread = None

    def handle_render(self, chunk: dict) -> dict:
        start = time.time()
        out = {"chunk_id": chunk.get("id", 0) if isinstance(chunk.get("id", 0), int) else 0, "pixels": [], "duration": 0.0, "node_id": self.node_id}
        err = validate_chunk(chunk)
        if err is None:
            with self._count_lock:
                self._active += 1
            try:
                with self._lock:
                    rect = self.render(chunk)
                out["pixels"] = encode_pixels(rect, chunk["start_x"], chunk["start_y"], str(chunk.get("encoding", "json")))
            except Exception as e:  # RemoteResult.Error (:46)
                err = "%s: %s" % (type(e).__name__, e)
            finally:
                with self._count_lock:
                    self._active -= 1
        if err is not None:
            out["error"] = err
        out["duration"] = time.time() - start
        return out

    def handle_status(self) -> dict:
        try:
            load = os.getloadavg()[0]
        except OSError:
            load = 0.0
        rss = 0
        try:
            with open("/proc/self/statm") as f:
                rss = int(f.read().split()[1]) * os.sysconf("SC_PAGE_SIZE")
        except (OSError, ValueError, IndexError):
            pass
        ncpu = os.cpu_count() or 1
        return {"id": self.node_id, "cpu_usage": min(100.0, 100.0 * load / ncpu), "memory_usage": rss, "active_jobs": self._active,
                "max_jobs": self.max_jobs, "load_average": load}

    def start(self):
        self._thread = threading.Thread(target=self.server.serve_forever, daemon=True)
        self._thread.start()
        return self

    def stop(self):
        self.server.shutdown()
        self.server.server_close()
        if self._thread:
            self._thread.join(timeout=5)


class DistributedRenderer:
    """Dispatcher side, mirroring distributed_renderer.go:14-236 (RenderChunkRemotely, GetNodeInfo, UpdateNodeLoad,
    GetOptimalNode, DistributeWork, GetStats).  DistributeWork assigns chunks to the least-loaded node."""

    def __init__(self, nodes: Sequence[str], timeout: float = 30.0):
        self.nodes = list(nodes)
        self.timeout = timeout  # http.Client{Timeout: 30 s}, :68
        self.nodeLoads = {n: 0 for n in self.nodes}
        self.remoteJobs = self.failedJobs = self.localJobs = 0
        self.startTime = time.time()
        self._lock = threading.Lock()

    def RenderChunkRemotely(self, chunk: dict, nodeAddr: str) -> dict:
        req = urllib.request.Request("http://%s/render" % nodeAddr, data=json.dumps(chunk).encode(), headers={"Content-Type": "application/json"}, method="POST")
        try:
            with urllib.request.urlopen(req, timeout=self.timeout) as resp:
                result = json.loads(resp.read())
        except Exception:
            with self._lock:
                self.failedJobs += 1
            raise
        with self._lock:
            self.remoteJobs += 1
        return result

    def GetNodeInfo(self, nodeAddr: str) -> dict:
        with urllib.request.urlopen("http://%s/status" % nodeAddr, timeout=self.timeout) as resp:
            return json.loads(resp.read())

    def UpdateNodeLoad(self, nodeID: str, load: int):
        with self._lock:
            self.nodeLoads[nodeID] = load

    def GetOptimalNode(self) -> str:
        with self._lock:
            return min(self.nodeLoads, key=self.nodeLoads.get) if self.nodeLoads else ""

    def DistributeWork(self, chunks: Sequence[dict]) -> list:
        def one(chunk):
            node = self.GetOptimalNode()
            if not node:
                raise RuntimeError("no available nodes")
            with self._lock:
                self.nodeLoads[node] += 1
            try:
                return self.RenderChunkRemotely(chunk, node)
            finally:
                with self._lock:
                    self.nodeLoads[node] -= 1

        with ThreadPoolExecutor(max_workers=max(1, 2 * len(self.nodes))) as ex:
            return list(ex.map(one, chunks))

    def GetStats(self) -> dict:
        total = self.remoteJobs + self.failedJobs
        return {"remote_jobs": self.remoteJobs, "local_jobs": self.localJobs, "failed_jobs": self.failedJobs, "total_nodes": len(self.nodes),
                "elapsed_time": time.time() - self.startTime, "success_rate": 100.0 if total == 0 else 100.0 * self.remoteJobs / total}


def assemble(chunks: Sequence[dict], results: Sequence[dict], width: int, height: int) -> np.ndarray:
    """RemoteResults -> [H, W, 4] uint8 (image.RGBA.Pix layout); raises on a result that carries an error"""
    by_id = {c["id"]: c for c in chunks}
    img = np.zeros((height, width, 4), dtype=np.uint8)
    for res in results:
        if res.get("error"):
            raise RuntimeError("chunk %s: %s" % (res.get("chunk_id"), res["error"]))
        c = by_id[res["chunk_id"]]
        img[c["start_y"]:c["end_y"], c["start_x"]:c["end_x"]] = decode_pixels(res, c)
    return img


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="serve /render and /status (the reference's chunk-farm protocol) from this GPU")
    ap.add_argument("--port", type=int, default=8080)
    ap.add_argument("--host", default="127.0.0.1", help="interface to listen on (0.0.0.0 to serve a dispatcher on another host)")
    ap.add_argument("--scene-root", default=".", help="directory the `scene` paths of a request are resolved in")
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--max-depth", type=int, default=50)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    node = ChunkNode(GpuChunkRenderer(a.samples, a.max_depth, a.seed, 0, a.device, a.scene_root), a.port, a.host)
    print("chunk node %s listening on %s:%d" % (node.node_id, a.host, node.port), flush=True)
    node.server.serve_forever()


if __name__ == "__main__":
    main()
