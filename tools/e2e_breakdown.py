"""Where the end-to-end time of one C1-view frame goes (host buffers in, host frame out)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench as B
G = importlib.import_module("concurrent-raytracer-go_b200")
kind, W, H, spp, depth, options, desc = B.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c1_view"]
flat = B.Workload(kind, options).flat(G)
r = G.NewParallelRenderer(1)
r.SetSamples(spp); r.SetMaxDepth(depth); r.SetSeed(20240601)
host = torch.zeros(W * H * 4, dtype=torch.uint8).pin_memory().numpy().reshape(H, W, 4)
for _ in range(10):
    r.UploadScene(flat); r.Render(flat, W, H, out=host)
N = 50
t_up = t_r = 0.0
acc = {"kernel_ms": 0, "cull_ms": 0, "trace_ms": 0, "resolve_ms": 0, "total_ms": 0, "upload_ms": 0, "bvh_build_ms": 0}
for _ in range(N):
    t0 = time.perf_counter(); r.UploadScene(flat); t1 = time.perf_counter(); r.Render(flat, W, H, out=host); t2 = time.perf_counter()
    t_up += t1 - t0; t_r += t2 - t1
    for k in acc: acc[k] += getattr(r.lastStats, k)
print("python wall: UploadScene %.1f us, Render %.1f us, sum %.1f us" % (1e6 * t_up / N, 1e6 * t_r / N, 1e6 * (t_up + t_r) / N))
print("library: " + ", ".join("%s %.1f us" % (k, 1e3 * v / N) for k, v in acc.items()))
r.close()
