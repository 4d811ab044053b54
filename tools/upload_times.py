"""Wall time of UploadScene (host flatten + BVH build + H2D) for a bench workload, repeated.  usage: upload_times.py <workload> [reps]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B
G = importlib.import_module("concurrent-raytracer-go_b200")
name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kind, W, H, spp, depth, options, desc = B.WORKLOADS[name]
flat = B.Workload(kind, options).flat(G)
r = G.NewParallelRenderer(1)
for i in range(reps):
    t0 = time.perf_counter()
    r.UploadScene(flat)
    print(name, "UploadScene %.1f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
