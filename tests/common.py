"""Shared helpers for the parity tests: scene variants of SURVEY §8d and image metrics."""
import copy
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "scenes")


def load_scene_dict(name):
    with open(os.path.join(SCENES, name)) as f:
        return json.load(f)


def c1_view():
    """C1-view: sphere_reflections_light with the camera mirrored to z=+8 so geometry is in frame
    under the reference camera (which ignores lookAt and looks down -Z)."""
    d = load_scene_dict("sphere_reflections_light.json")
    d["camera"]["position"][2] = 8
    return d


def c2_view():
    """C2-view: final_silver_prism_purple_cube with the camera mirrored to z=+25."""
    d = load_scene_dict("final_silver_prism_purple_cube_.json")
    d["camera"]["position"][2] = 25
    return d


def c3():
    return load_scene_dict("two_red_cubes_scene.json")


def within_one(a, b):
    """fraction of pixels whose RGB channels all differ by <= 1 (8-bit)."""
    d = np.abs(a[..., :3].astype(np.int32) - b[..., :3].astype(np.int32)).max(axis=-1)
    return float((d <= 1).mean())


def mae(a, b):
    return float(np.abs(a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)).mean())


def psnr(a, b):
    mse = float(((a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)) ** 2).mean())
    if mse == 0:
        return float("inf")
    return 10.0 * np.log10(255.0 ** 2 / mse)


def random_sphere_scene(n, seed, extent=10.0, cam_z=30.0, lights=3):
    """Synthetic mixed-material sphere scene as a reference-format JSON dict (C4-style)."""
    rng = np.random.default_rng(seed)
    objs = []
    for i in range(n):
        pos = rng.uniform(-extent, extent, 3)
        r = float(rng.uniform(0.1, 0.5) * extent / 10.0)
        k = i % 3
        if k == 0:
            mat = {"type": "metal", "color": rng.uniform(0.2, 1.0, 3).tolist(), "roughness": float(rng.uniform(0, 0.3))}
        elif k == 1:
            mat = {"type": "glass", "color": rng.uniform(0.5, 1.0, 3).tolist(), "refractionIndex": 1.5}
        else:
            mat = {"type": "dielectric", "refractionIndex": 1.5}
        objs.append({"type": "sphere", "position": pos.tolist(), "radius": r, "material": mat})
    lts = [{"type": "point", "position": [4 * extent, 6 * extent, 4 * extent], "color": [1, 1, 1], "intensity": 20.0 * extent * extent},
           {"type": "point", "position": [-4 * extent, 6 * extent, 4 * extent], "color": [1, 0.9, 0.8], "intensity": 20.0 * extent * extent},
           {"type": "area", "position": [0, 8 * extent, 0], "color": [0.8, 0.9, 1], "intensity": 30.0 * extent * extent}][:lights]
    return {"camera": {"position": [0, 0, cam_z], "lookAt": [0, 0, 0], "up": [0, 1, 0], "fov": 40, "aspectRatio": 16 / 9},
            "objects": objs, "lights": lts}
