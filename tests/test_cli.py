"""CLI mirrors of the reference's cmd/raytracer (main.go:14-69) and cmd/benchmark (main.go:290-328) over the C ABI."""
import json
import os
import subprocess

import numpy as np
import pytest

import common as Cm

BIN = os.path.join(Cm.ROOT, "concurrent-raytracer-go_b200", "bin")


def run(args, **kw):
    return subprocess.run(args, capture_output=True, text=True, timeout=300, **kw)


def test_raytracer_usage_and_argument_errors():
    """Same messages and exit code as cmd/raytracer/main.go:18-36."""
    r = run([os.path.join(BIN, "raytracer")])
    assert r.returncode == 1 and "Usage: raytracer <scene_file> <output_file> <width> <height>" in r.stdout
    assert "Example: raytracer scene.json output.png 800 600" in r.stdout
    r = run([os.path.join(BIN, "raytracer"), "s.json", "o.png", "abc", "600"])
    assert r.returncode == 1 and "Invalid width: abc" in r.stdout
    r = run([os.path.join(BIN, "raytracer"), "s.json", "o.png", "800", "6x"])
    assert r.returncode == 1 and "Invalid height: 6x" in r.stdout


def test_benchmark_rejects_unknown_flags():
    r = run([os.path.join(BIN, "benchmark"), "-nonsense", "1"])
    assert r.returncode == 2 and "flag provided but not defined: -nonsense" in r.stderr


@pytest.mark.gpu
def test_raytracer_cli_renders_png_and_benchmark_json(gort, tmp_path):
    from PIL import Image
    scene = tmp_path / "c1_view.json"
    scene.write_text(json.dumps(Cm.c1_view()))
    out = tmp_path / "out" / "frame"  # no extension: ".png" is appended (main.go:52-55)
    os.makedirs(tmp_path / "out")
    readme = tmp_path / "readme.json"
    r = run([os.path.join(BIN, "raytracer"), "-samples", "8", "-seed", "5", "-readme-json", str(readme), str(scene), str(out), "320", "240"])
    assert r.returncode == 0, r.stdout + r.stderr
    for msg in ("Loading scene from:", "Rendering at 320x240 resolution...", "Render completed in", "Saving to:", "Benchmark data saved"):
        assert msg in r.stdout
    img = np.array(Image.open(str(out) + ".png").convert("RGBA"))
    assert img.shape == (240, 320, 4)
    # identical bytes to the library call the Python mirror makes
    rr = gort.NewParallelRenderer(1)
    rr.SetSamples(8); rr.SetMaxDepth(50); rr.SetSeed(5)
    ref = rr.Render(gort.SceneFromDict(Cm.c1_view()), 320, 240)
    rr.close()
    assert np.array_equal(img, ref)
    bd = json.load(open(tmp_path / "out" / "benchmark_data.json"))
    assert set(bd) == {"scene_name", "resolution", "render_time_seconds", "samples", "max_depth", "num_workers", "objects", "lights", "timestamp", "features"}
    assert bd["resolution"] == "320x240" and bd["samples"] == 8 and bd["max_depth"] == 50 and bd["objects"] == 5 and bd["lights"] == 2
    rj = json.load(open(readme))  # README.md:50-71 schema
    assert list(rj) == sorted(rj) and {"rays_per_second", "pixels_per_second", "render_time", "worker_count", "bvh_build_time", "setup_time"} <= set(rj)
    assert rj["rays_per_second"] == pytest.approx(rj["pixels_per_second"] * 8, rel=1e-3) and rj["render_time"].endswith("s")


@pytest.mark.gpu
def test_benchmark_cli_sweeps_and_reports(tmp_path):
    out = tmp_path / "bench.json"
    r = run([os.path.join(BIN, "benchmark"), "-width", "200", "-height", "150", "-workers", "1", "-samples", "2,4", "-max-depth", "5", "-duration", "50ms",
             "-output", str(out)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "BENCHMARK SUMMARY" in r.stdout and r.stdout.count("Completed: 1 workers") == 2
    rep = json.load(open(out))
    assert set(rep) == {"summary", "results", "config", "timestamp", "system_info"}  # main.go:166-172
    assert len(rep["results"]) == 2
    for res in rep["results"]:
        assert {"config", "worker_count", "samples", "max_depth", "scene", "duration", "rays_per_second", "pixels_per_second", "memory_usage", "cpu_usage",
                "speedup", "efficiency"} <= set(res)  # BenchmarkResult tags main.go:33-46
        assert res["rays_per_second"] == pytest.approx(res["pixels_per_second"] * res["samples"], rel=1e-3)
        assert res["duration"] > 0 and res["frames"] >= 1
    assert rep["summary"]["total_benchmarks"] == 2


@pytest.mark.gpu
def test_raytracer_scene_settings_extension(gort, tmp_path):
    """-scene-settings uses the scene's own "renderer" block (samples 200, maxDepth 20 in the shipped cube scene)."""
    scene = os.path.join(Cm.SCENES, "final_silver_prism_purple_cube_.json")
    out = tmp_path / "o.png"
    r = run([os.path.join(BIN, "raytracer"), "-scene-settings", "-seed", "1", scene, str(out), "64", "48"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Samples per pixel: 200" in r.stdout and "Max depth: 20" in r.stdout
    rr = gort.NewParallelRenderer(1)
    rr.UploadScene(gort.LoadFromFile(scene))
    assert rr.SceneRenderHints() == {"samples": 200, "maxDepth": 20, "antiAliasing": True, "recursiveReflections": True, "softShadows": True}
    rr.close()
