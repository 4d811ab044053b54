"""Synthetic large scenes of SURVEY §8d (C4, C5) as flat arrays (no JSON round trip) + small JSON-dict
versions the linear-scan oracle can check.  Generator: numpy PCG64 with the fixed seeds below."""
import importlib

import numpy as np

C4_SEED, C5_SEED = 20240601, 20240602


def _materials(rng, n):
    """material by index mod 3: metal(roughness U[0,0.3], colour U[0.2,1]^3) / glass(1.5, colour U[0.5,1]^3) / dielectric(1.5)"""
    mats = []
    rough = rng.uniform(0, 0.3, n)
    cm = rng.uniform(0.2, 1.0, (n, 3))
    cg = rng.uniform(0.5, 1.0, (n, 3))
    for i in range(n):
        k = i % 3
        if k == 0:
            mats.append({"type": 1, "color": cm[i], "roughness": float(rough[i]), "metallic": 1.0, "specular": 1.0, "ior": 1.5})
        elif k == 1:
            mats.append({"type": 4, "color": cg[i], "roughness": 0.0, "metallic": 0.0, "specular": 1.0, "ior": 1.5})
        else:
            mats.append({"type": 5, "color": (1, 1, 1), "roughness": 0.0, "metallic": 0.0, "specular": 1.0, "ior": 1.5})
    return mats


def _cube_tris(pos, size):
    h = size / 2.0
    sg = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)
    v = pos + sg * h
    faces = [(0, 1, 2, 3), (1, 5, 6, 2), (5, 4, 7, 6), (4, 0, 3, 7), (3, 2, 6, 7), (4, 5, 1, 0)]  # scene.go:164-171
    out = []
    for f in faces:
        out.append(np.concatenate([v[f[0]], v[f[1]], v[f[2]]]))
        out.append(np.concatenate([v[f[0]], v[f[2]], v[f[3]]]))
    return out


def scene_arrays(n_spheres, n_cubes, extent, seed, cam_z, radius=(0.1, 0.5), lights=None, fog=None, aspect=16 / 9):
    """-> dict(camera, materials, spheres, triangles, lights, fog): feeds both gort.FlatScene and oracle.Scene.from_flat"""
    rng = np.random.default_rng(seed)
    n_obj = n_spheres + n_cubes
    mats = _materials(rng, n_obj)
    centers = rng.uniform(-extent, extent, (n_obj, 3))
    # keep the camera clear of geometry
    cam = np.array([0.0, 0.0, cam_z])
    near = np.linalg.norm(centers - cam, axis=1) < 3.0
    centers[near] += np.array([0, 0, -10.0])
    radii = rng.uniform(radius[0], radius[1], n_obj)
    is_cube = np.zeros(n_obj, dtype=bool)
    if n_cubes:
        is_cube[rng.choice(n_obj, n_cubes, replace=False)] = True
    spheres, tris, order = [], [], 0
    for i in range(n_obj):
        if is_cube[i]:
            for t in _cube_tris(centers[i], np.full(3, 2 * radii[i])):
                tris.append((t, i, order))
                order += 1
        else:
            spheres.append((centers[i], float(radii[i]), i, order))
            order += 1
    if lights is None:
        lights = [((40.0, 60.0, 40.0), (1, 1, 1), 2000.0), ((-40.0, 60.0, 40.0), (1, 1, 1), 2000.0), ((0.0, 80.0, 0.0), (1, 1, 1), 3000.0)]
    camera = {"position": cam, "lookAt": (0, 0, 0), "up": (0, 1, 0), "fov": 40.0, "aspectRatio": aspect}
    return {"camera": camera, "materials": mats, "spheres": spheres, "triangles": tris, "lights": lights, "fog": fog}


def to_gort(a):
    G = importlib.import_module("concurrent-raytracer-go_b200")
    return G.FlatScene(a["camera"], a["materials"], a["spheres"], a["triangles"], a["lights"], fog=a["fog"])


def to_oracle(a):
    import oracle as O
    return O.Scene.from_flat(a["camera"], a["materials"], a["spheres"], a["triangles"], a["lights"], fog=a["fog"])


def c4_arrays():
    """C4: 100 000 spheres, centres U[-50,50]^3, radius U[0.1,0.5], 2 point + 1 'area' (=point, F10) light, camera (0,0,120)."""
    return scene_arrays(100_000, 0, 50.0, C4_SEED, 120.0)


def c5_arrays():
    """C5: 1 000 000 primitives = 750 000 spheres + 20 833 cubes (12 triangles each), box U[-200,200]^3, fog on."""
    lights = [((160.0, 240.0, 160.0), (1, 1, 1), 32000.0), ((-160.0, 240.0, 160.0), (1, 1, 1), 32000.0), ((0.0, 320.0, 0.0), (1, 1, 1), 48000.0)]
    return scene_arrays(750_000, 20_833, 200.0, C5_SEED, 480.0, radius=(0.4, 2.0), lights=lights,
                      fog={"density": 0.002, "color": (0.25, 0.25, 0.25)})
