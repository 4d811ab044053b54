#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 render hot path (contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c1_view|c1_ref|c2_view|c2_faithful|c3|c4|c5|c5_spp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full frame of the workload.  Default workload = the configuration BASELINE.json's metric is
quoted on: demo-assets/sphere_reflections_light.json, 800x600, 100 spp, max_depth 50, soft shadows, jitter —
in its C1-view variant (camera mirrored to z=+8: with the committed camera the scene is behind the viewer and
the frame is black; `--workload c1_ref` runs that literal case).  Metric: the reference's own
rays_per_second = width*height*samples / render_time (README.md:60-61, cmd/benchmark/main.go:125-127), in Mrays/s.

  value : frames rendered with the scene resident in HBM and the RGBA8 frame left in HBM (device-timed).
  e2e   : the same frame through the public API with HOST buffers: scene description uploaded (host BVH
          build + H2D) and the RGBA8 frame copied back (D2H), every step, inside the timed region.
  N > 1 : one process per GPU; 32x32 tiles statically interleaved (tile_id % N == rank); each rank renders its
          tile-major slab, one NCCL all-gather of the slabs over NVLink, rank 0 un-swizzles.  Strong scaling
          (the frame is fixed).
  --impl reference : the reference's algorithm on the host CPU cores — the float64 C++ oracle port in reference
          mode (linear-scan hitWorld, 32x32 tiles, one thread per core, sequential RNG).  The Go original cannot
          be built: there is no Go toolchain in this image.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "rays_per_second (primary samples/s = width*height*samples/render_time)"
UNIT = "Mrays/s"
NOMINAL_FP32_TFLOPS = 74.4  # 148 SM x 128 lanes x 2 x 1.965 GHz

WORKLOADS = {
    # name: (scene builder, width, height, spp, depth, options, description)
    "c1_view": ("c1_view", 800, 600, 100, 50, 0, "sphere_reflections_light 800x600 100spp depth50 soft-shadows jitter, camera z mirrored to +8 (C1-view)"),
    "c1_ref": ("c1_ref", 800, 600, 100, 50, 0, "sphere_reflections_light 800x600 100spp depth50, committed camera (scene behind viewer: black frame)"),
    "c2_view": ("c2_view", 1200, 900, 100, 50, 1, "final_silver_prism_purple_cube 1200x900 100spp depth50, camera z mirrored to +25, prism extension on"),
    # the same scene as the reference's own loader sees it: triangularPrism objects are unknown to it and skipped (scene/scene.go:69-83)
    "c2_faithful": ("c2_view", 1200, 900, 100, 50, 0, "final_silver_prism_purple_cube 1200x900 100spp depth50, camera z mirrored to +25, prisms skipped as the reference's loader does"),
    "c3": ("c3", 800, 600, 1, 8, 0, "two_red_cubes 800x600 1spp depth8 no-jitter hard-shadows (deterministic correctness config)"),
    "c4": ("c4", 1920, 1080, 64, 16, 0, "synthetic 100k random spheres 1920x1080 64spp depth16 3 lights"),
    "c5": ("c5", 3840, 2160, 256, 32, 2, "synthetic 1M-primitive sphere/box scene 3840x2160 256spp depth32 fog on"),
    # the same 4K frame at 32 spp: 8x cheaper stand-in for scaling runs on a GPU-minute budget (less work per GPU,
    # so its parallel efficiency is a lower bound for the 256-spp frame)
    "c5_spp32": ("c5", 3840, 2160, 32, 32, 2, "synthetic 1M-primitive sphere/box scene 3840x2160 32spp depth32 fog on"),
}


def config_of(wl, world, link):
    """the `config` object of a bench line: identical keys and values for the GPU arm and the reference arm"""
    kind, W, H, spp, depth, options, desc = wl
    return {"workload": desc, "width": W, "height": H, "samples": spp, "max_depth": depth, "camera_mode": "reference",
            "anti_aliasing": kind != "c3", "soft_shadows": kind != "c3", "recursive_reflections": True,
            "prism_extension": bool(options & 1), "fog_extension": bool(options & 2)}


class Workload:
    """Scene of a workload for both arms: .flat(G) -> gort FlatScene (the gort_scene_desc a Go host would
    pass after Flatten()); .oracle(O) -> the oracle's scene built independently from the same description."""

    def __init__(self, kind, options):
        import common as Cm
        self.kind, self.options = kind, options
        self.dict, self.arrays = None, None
        if kind == "c1_view":
            self.dict = Cm.c1_view()
        elif kind == "c1_ref":
            self.dict = Cm.load_scene_dict("sphere_reflections_light.json")
        elif kind == "c2_view":
            self.dict = Cm.c2_view()
        elif kind == "c3":
            self.dict = Cm.c3()
        elif kind in ("c4", "c5"):
            import synth
            self.arrays = synth.c4_arrays() if kind == "c4" else synth.c5_arrays()
        else:
            raise SystemExit("unknown workload " + kind)

    def flat(self, G):
        if self.dict is not None:
            return G.HostScene(json.dumps(self.dict), self.options).to_flat()
        import synth
        return synth.to_gort(self.arrays)

    def oracle(self, O):
        if self.dict is not None:
            return O.Scene(self.dict, prisms=bool(self.options & 1), fog=bool(self.options & 2))
        import synth
        return synth.to_oracle(self.arrays)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args, wl):
    """--impl reference: the oracle port of the reference's CPU path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    kind, W, H, spp, depth, options, desc = wl
    cores = len(os.sched_getaffinity(0))
    scene = Workload(kind, options).oracle(O)
    jitter, soft = (kind != "c3"), (kind != "c3")
    # bounded sample: the full frame when the linear scan can finish it in seconds, else a centred crop at reduced spp
    crop, s_spp, sample = None, spp, "full frame, all %d samples/pixel" % spp
    n_prims = scene.counts()["spheres"] + scene.counts()["triangles"]
    if n_prims > 1000:
        cw = ch = 64 if n_prims <= 200_000 else 24
        s_spp = min(spp, 2)
        crop = ((W - cw) // 2, (H - ch) // 2, (W + cw) // 2, (H + ch) // 2)
        sample = "centred %dx%d crop at %d spp, linear scan over %d primitives; rays/s from the crop's own sample count" % (cw, ch, s_spp, n_prims)

    def step():
        t0 = time.perf_counter()
        _, _, cnt = scene.render(W, H, samples=s_spp, max_depth=depth, jitter=jitter, soft_shadows=soft, rng_mode=O.RNG_MT,
                                 seed=int(t0 * 1e6) & 0xffff, threads=cores, crop=crop)
        return time.perf_counter() - t0, cnt["samples"]

    warm, steps = args.warmup, args.steps
    if n_prims > 1000:  # one bounded sample is ~30 s of CPU work: the driver's K and W are meant for sub-second frames
        warm, steps = min(warm, 1), min(steps, 2)
    for _ in range(warm):
        step()
    times, samples = [], 0
    for _ in range(steps):
        dt, n = step()
        times.append(dt)
        samples += n
    total = sum(times)
    value = samples / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(wl, 1, None),
        "pixels_per_second": value * 1e6 / s_spp,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C++ -O2 float64 transcription of the Go renderer (no Go toolchain here); GOMAXPROCS n/a, threads = %d" % cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gort", choices=["gort", "reference"])
    ap.add_argument("--workload", default="c1_view", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scale-c5", action="store_true", help="skip the 4K 1 M-primitive sub-record of the default line")
    ap.add_argument("--gather", default="link", choices=["link", "nccl"], help="N > 1: how the frame is assembled on rank 0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "gort" else args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import numpy as np
    import torch
    import torch.distributed as dist

    kind, W, H, spp, depth, options, desc = wl
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun: python -m torch.distributed.run --nproc-per-node %d bench.py --gpus %d" % (args.gpus, args.gpus, args.gpus))
        raise SystemExit("--gpus (%d) != WORLD_SIZE (%d)" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libgort has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    G = importlib.import_module("concurrent-raytracer-go_b200")
    work = Workload(kind, options)
    flat = work.flat(G)  # the gort_scene_desc a Go host passes after Flatten()
    r = G.NewParallelRenderer(1, devices=[local_rank])
    r.SetSamples(spp); r.SetMaxDepth(depth); r.SetSeed(20240601)
    r.SetAntiAliasing(kind != "c3"); r.SetSoftShadows(kind != "c3")
    r.SetShard(rank, world)
    stream = torch.cuda.Stream(device=dev)  # every kernel, copy and collective of the bench runs on this stream
    torch.cuda.set_stream(stream)
    r.set_stream(stream.cuda_stream)
    r.UploadScene(flat)

    # N > 1: assemble the frame over NVLink peer memory (frame link: every rank's resolve kernel stores its tiles
    # straight into rank 0's frame, flag handshake, no collective); --gather nccl keeps the slab all-gather
    link = None
    if world > 1 and args.gather == "link":
        try:
            ht = torch.zeros(64, dtype=torch.uint8, device=dev)
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            if rank == 0:
                link, handle = r.LinkCreate(W, H, world)
                ht.copy_(torch.frombuffer(bytearray(handle), dtype=torch.uint8))
            dist.broadcast(ht, 0)
            if rank != 0:
                try:
                    link = r.LinkOpen(bytes(ht.cpu().numpy().tobytes()), W, H, world, rank)
                except G.GortError as e:
                    sys.stderr.write("rank %d: frame link unavailable (%s): falling back to the NCCL gather\n" % (rank, e))
                    ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if link is not None:
                    r.LinkClose(link)
                link = None
        except G.GortError as e:
            raise SystemExit("frame link: %s" % e)
    slab_bytes = G.shard_slab_bytes(W, H, world)
    slab = torch.zeros(slab_bytes, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(slab_bytes * world, dtype=torch.uint8, device=dev) if world > 1 else None
    frame = torch.zeros(W * H * 4, dtype=torch.uint8, device=dev)
    host_frame = torch.zeros(W * H * 4, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    launches_per_frame = [3]  # cull + trace + resolve; replaced by gort_stats.kernel_launches of the measured frame below

    def step_device():
        """one frame, everything resident in HBM; returns number of kernels launched"""
        if world == 1:
            r.RenderDevice(W, H, frame.data_ptr())
            return launches_per_frame[0]  # libgort kernels (memsets and the L2 flush are not counted)
        if link is not None:
            r.RenderLinked(W, H, link)
            # rank 0 owns the frame: + the wait for the peers' arrivals (its release rides in the cull pass; a peer launches
            # a wait and a signal around its resolve pass).  GORT_LINK_UNFUSED=1: + a release kernel
            return launches_per_frame[0] + (2 if os.environ.get("GORT_LINK_UNFUSED") else 1)
        r.RenderShardDevice(W, H, slab.data_ptr())
        dist.all_gather_into_tensor(gathered, slab)
        if rank == 0:
            r.UnswizzleDevice(gathered.data_ptr(), world, W, H, frame.data_ptr())
            return launches_per_frame[0] + 1  # + unswizzle (the all-gather is NCCL's)
        return launches_per_frame[0]

    def step_e2e():
        """public API with host buffers: scene upload (host flatten + BVH + H2D) and frame D2H inside the step"""
        r.UploadScene(flat)
        if world == 1:
            img = r.Render(flat, W, H, out=host_frame.numpy().reshape(H, W, 4))
            return img
        if link is not None:
            r.RenderLinked(W, H, link)
            if rank == 0:
                r.LinkRead(link, W, H, out=host_frame.numpy().reshape(H, W, 4))
            else:
                torch.cuda.synchronize()
            return host_frame
        r.RenderShardDevice(W, H, slab.data_ptr())
        dist.all_gather_into_tensor(gathered, slab)
        if rank == 0:
            r.UnswizzleDevice(gathered.data_ptr(), world, W, H, frame.data_ptr())
            host_frame.copy_(frame, non_blocking=True)
        torch.cuda.synchronize()
        return host_frame

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up --------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    if world == 1:
        launches_per_frame[0] = int(r.RenderDevice(W, H, frame.data_ptr(), want_stats=True).kernel_launches)
    else:
        launches_per_frame[0] = int(r.RenderShardDevice(W, H, slab.data_ptr(), want_stats=True).kernel_launches)
        barrier()

    # ---- timed: K steps, device events per step, L2 flushed between steps (outside the events) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        ev[i][0].record(stream)
        launches += step_device()
        ev[i][1].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())

    # ---- the trace kernel alone (roofline), CUDA events inside the library on the launching stream ----
    trace_ms = []
    for i in range(min(args.steps, 10)):
        flush.fill_(i)
        st = r.RenderShardDevice(W, H, slab.data_ptr(), want_stats=True) if world > 1 else r.RenderDevice(W, H, frame.data_ptr(), want_stats=True)
        trace_ms.append(st.trace_ms)
    barrier()

    # ---- e2e: host buffers, copies inside the timed region ---------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - e0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- N > 1: the link-assembled frame must be the frame the NCCL slab gather produces, bit for bit ----
    frame_check = None
    if link is not None:
        r.RenderLinked(W, H, link)
        if rank == 0:
            r.LinkRead(link, W, H, out=host_frame.numpy().reshape(H, W, 4))
        r.RenderShardDevice(W, H, slab.data_ptr())
        dist.all_gather_into_tensor(gathered, slab)
        if rank == 0:
            r.UnswizzleDevice(gathered.data_ptr(), world, W, H, frame.data_ptr())
            torch.cuda.synchronize()
            frame_check = "identical" if bool((frame.cpu() == host_frame).all()) else "MISMATCH"
        barrier()

    # ---- algorithmic FLOPs of one frame (device counters, separate untimed launch) -----------------
    r.SetCollectStats(True)
    st = r.RenderShardDevice(W, H, slab.data_ptr(), want_stats=True) if world > 1 else r.RenderDevice(W, H, frame.data_ptr(), want_stats=True)
    r.SetCollectStats(False)
    flops = st.algorithmic_flops
    segs = st.closest_queries + st.shadow_queries
    fl = torch.tensor([flops, float(segs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(fl)
    flops, segs = float(fl[0].item()), float(fl[1].item())
    peak_tflops, _ = r.MeasureFp32Peak()

    if rank == 0:
        rays = W * H * spp
        ms_per_step = dev_ms / args.steps
        value = rays / (ms_per_step * 1e-3) / 1e6
        e2e_value = rays * args.steps / e2e_s / 1e6
        tr = sum(trace_ms) / len(trace_ms)
        achieved = (st.algorithmic_flops / (tr * 1e-3)) / 1e12
        # measured FLOPs of the same kernel(s): ncu smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on of the committed
        # capture (tools/gpu_round.sh writes profiles/flops.json: fadd + fmul + 2 ffma per launch)
        measured = None
        try:
            # a whole frame's count: only comparable with the kernel time when one GPU renders the whole frame
            measured = json.load(open(os.path.join(ROOT, "profiles", "flops.json"))).get(args.workload) if world == 1 else None
        except (OSError, ValueError):
            pass
        h2d = int(st.bvh_bytes + flat.desc.n_materials * 64 + flat.desc.n_lights * 32)
        traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of trace_kernel from the committed ncu capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        except (OSError, ValueError):
            pass
        # the HBM side of the roofline, to show it is not the bound: measured DRAM traffic of the kernel over its time
        hbm = None
        try:
            peak_gbs = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json"
        except (OSError, ValueError, KeyError):
            peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        if traffic:
            gbs = traffic / (tr * 1e-3) / 1e9
            hbm = {"achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs, "peak_source": peak_src}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, world, link),
            "run": {"l2": "flushed between timed steps (256 MiB fill outside the events)",
                    "sharding": ("tile_id %% %d == rank, " % world + ("tiles stored into rank 0's frame over NVLink peer memory (frame link)" if link is not None else "NCCL all-gather of RGBA8 slabs")) if world > 1 else "single GPU",
                    "rng": "philox4x32-10 seed 20240601",
                    "render_path": {0: "parameter-bank scan (per-warp-queue kernel)", 1: "per-warp-queue BVH kernel", 2: "global-queue wavefront pipeline"}[st.render_path]},
            "pixels_per_second": value * 1e6 / spp,
            "ray_segments_per_second": segs / (ms_per_step * 1e-3),
            "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": W * H * 4,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                         "traffic": traffic,
                         "kernel": "trace_kernel" if st.render_path != 2 else "wavefront pipeline (pool_trace / pool_cone / scatter / shade kernels of one frame)",
                         "scope": "whole frame" if world == 1 else "rank 0's tiles (1/%d of the frame): counters and kernel time of that rank" % world,
                         "kernel_ms": tr, "cull_ms": st.cull_ms, "algorithmic_flops_per_launch": st.algorithmic_flops,
                         "flops_model": "SURVEY 8d per-operation costs x device counters of the same frame; ray generation only for the "
                                        "%d of %d primary samples that were generated (the rest sit in pixel blocks the cull pass proved empty); no tone-map term" % (st.primary_generated, st.primary_rays),
                         "measured_flops_per_launch": measured,
                         "frac_measured": (measured / (tr * 1e-3) / 1e12 / peak_tflops) if measured else None,
                         "peak_source": "measured live: dependent-FFMA microbenchmark (MEASURED_PEAKS.json has no fp32 entry; nominal %.1f)" % NOMINAL_FP32_TFLOPS,
                         "note": "divergent traversal + shading: bounded by FP32/INT issue and node-fetch latency, not HBM (scene fits L1/L2)",
                         "hbm": hbm},
            "culled_sample_fraction": 1.0 - st.primary_generated / max(1, st.primary_rays),
        }
        if frame_check is not None:
            line["run"]["frame_link_vs_nccl_gather"] = frame_check
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl, work)
    if world > 1:
        dist.barrier()
        if link is not None:
            torch.cuda.synchronize()
            r.LinkClose(link)
    r.close()
    c5 = None
    if args.workload == "c1_view" and not args.no_scale_c5:
        c5 = scale_c5(args, G, torch, dist, world, rank, local_rank, dev, stream)
    if rank == 0:
        if c5 is not None:
            line["scale_c5"] = c5
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def scale_c5(args, G, torch, dist, world, rank, local_rank, dev, stream):
    """The north-star's multi-GPU configuration next to the headline line: the 4K synthetic 1 M-primitive scene (C5, fog on,
    depth 32) at 32 spp — a frame of ~1.5 s on one GPU, where tile sharding is throughput- and not latency-bound.
    ms/frame device-resident (max over ranks, NCCL all-gather of the RGBA8 slabs + un-swizzle on rank 0), end to end with
    host buffers (scene upload incl. BVH build every step + frame D2H), and the same frame on ONE GPU in the same run
    (rank 0, the other ranks idle) for `efficiency_vs_n1` = t(1) / (N t(N))."""
    kind, W, H, spp, depth, options, desc = WORKLOADS["c5_spp32"]
    flat = Workload(kind, options).flat(G)
    r = G.NewParallelRenderer(1, devices=[local_rank])
    r.SetSamples(spp); r.SetMaxDepth(depth); r.SetSeed(20240602); r.SetShard(rank, world)
    r.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    r.UploadScene(flat)
    upload_s = time.perf_counter() - t0
    slab_bytes = G.shard_slab_bytes(W, H, world)
    slab = torch.zeros(slab_bytes, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(slab_bytes * world, dtype=torch.uint8, device=dev) if world > 1 else None
    frame = torch.zeros(W * H * 4, dtype=torch.uint8, device=dev)
    host_frame = torch.zeros(W * H * 4, dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if world == 1:
            r.RenderDevice(W, H, frame.data_ptr())
            return
        r.RenderShardDevice(W, H, slab.data_ptr())
        dist.all_gather_into_tensor(gathered, slab)
        if rank == 0:
            r.UnswizzleDevice(gathered.data_ptr(), world, W, H, frame.data_ptr())

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step()
    barrier()
    steps = 2
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record(stream)
        step()
        b.record(stream)
    barrier()
    ms = reduce_max(sum(a.elapsed_time(b) for a, b in ev) / steps)
    # end to end: scene upload (host flatten + BVH build + H2D) and frame D2H inside the timed region
    barrier()
    e0 = time.perf_counter()
    for _ in range(steps):
        r.UploadScene(flat)
        step()
        if rank == 0:
            host_frame.copy_(frame, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_ms = reduce_max(1e3 * (time.perf_counter() - e0) / steps)
    st = r.RenderShardDevice(W, H, slab.data_ptr(), want_stats=True) if world > 1 else r.RenderDevice(W, H, frame.data_ptr(), want_stats=True)
    n1_ms = ms
    if world > 1:
        barrier()
        if rank == 0:  # the same frame on one GPU, same process, same run
            r.SetShard(0, 1)
            r.RenderDevice(W, H, frame.data_ptr())
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            r.RenderDevice(W, H, frame.data_ptr())
            b.record(stream)
            torch.cuda.synchronize()
            n1_ms = a.elapsed_time(b)
        barrier()
        n1_ms = reduce_max(n1_ms if rank == 0 else 0.0)
    r.close()
    rays = W * H * spp
    return {"workload": desc, "width": W, "height": H, "samples": spp, "max_depth": depth, "n_gpus": world,
            "ms_per_frame": ms, "value": rays / (ms * 1e-3) / 1e6, "unit": UNIT,
            "e2e_ms_per_frame": e2e_ms, "scene_upload_s": upload_s, "bvh_build_ms": st.bvh_build_ms,
            "n1_ms_per_frame_same_run": n1_ms, "efficiency_vs_n1": n1_ms / (world * ms),
            "render_path": int(st.render_path), "gather": "NCCL all-gather of RGBA8 slabs" if world > 1 else "none"}


def cpu_baseline(wl, work):
    """The oracle port timed on this box's host cores (bounded sample), rank 0 at N=1 only."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    kind, W, H, spp, depth, options, desc = wl
    cores = len(os.sched_getaffinity(0))
    scene = work.oracle(O)
    n_prims = scene.counts()["spheres"] + scene.counts()["triangles"]
    crop, s_spp, sample = None, spp, "full frame %dx%d at %d spp" % (W, H, spp)
    if n_prims > 1000:
        # the linear scan costs samples x primitives: keep the sample to ~30 s of CPU work whatever the scene size
        s_spp = min(spp, 2)
        side = 64 if n_prims <= 200_000 else 24
        crop = ((W - side) // 2, (H - side) // 2, (W + side) // 2, (H + side) // 2)
        sample = "centred %dx%d crop at %d spp (linear scan over %d primitives)" % (side, side, s_spp, n_prims)
    # repeat the sample until ~10 s of CPU work have been timed (at least 2 passes, the first is a warm-up)
    scene.render(W, H, samples=s_spp, max_depth=depth, jitter=(kind != "c3"), soft_shadows=(kind != "c3"),
                 rng_mode=O.RNG_MT, seed=0, threads=cores, crop=crop)
    total, samples, passes = 0.0, 0, 0
    while total < 10.0 and passes < 200:
        t0 = time.perf_counter()
        _, _, cnt = scene.render(W, H, samples=s_spp, max_depth=depth, jitter=(kind != "c3"), soft_shadows=(kind != "c3"),
                                 rng_mode=O.RNG_MT, seed=1 + passes, threads=cores, crop=crop)
        total += time.perf_counter() - t0
        samples += cnt["samples"]
        passes += 1
    out = {"value": samples / total / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%s, %d passes" % (sample, passes), "seconds": total,
           "note": "C++ -O2 float64 transcription of the Go renderer, one thread per host core (no Go toolchain in this image)"}
    if n_prims > 1000:
        # the fairer CPU number for large scenes: the same float64 port with hitWorld over the oracle's own BVH
        # (identical hits to the linear scan, tests/test_oracle_accel.py) on a centred 256x256 crop at the same 2 spp
        crop2 = ((W - 256) // 2, (H - 256) // 2, (W + 256) // 2, (H + 256) // 2)
        scene.render(W, H, samples=s_spp, max_depth=depth, rng_mode=O.RNG_MT, seed=0, threads=cores, crop=crop2, use_accel=True)
        t0 = time.perf_counter()
        _, _, cnt = scene.render(W, H, samples=s_spp, max_depth=depth, rng_mode=O.RNG_MT, seed=1, threads=cores, crop=crop2, use_accel=True)
        dt = time.perf_counter() - t0
        out["with_bvh"] = {"value": cnt["samples"] / dt / 1e6, "unit": UNIT, "sample": "centred 256x256 crop at %d spp, oracle-side BVH" % s_spp, "seconds": dt}
    return out


if __name__ == "__main__":
    sys.exit(main())
