"""ctypes wrapper around oracle/_build/liboracle.so — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It parses the reference's scene JSON with Python's json module
(independently of the product's C++ loader) and feeds the objects to the float64 oracle, which
applies the reference's material defaults and cube expansion itself
(internal/scene/scene.go:104-190).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

MAT_TYPES = {
    "lambertian": 0,
    "metal": 1,
    "shiny": 2,
    "perfectmirror": 3,
    "glass": 4,
    "dielectric": 5,
    "diffuselight": 6,
}
CAMERA_REFERENCE, CAMERA_LOOKAT = 0, 1
RNG_MT, RNG_PHILOX = 0, 1
STREAM_JITTER, STREAM_SCATTER, STREAM_SHADOW = 0, 1, 2


class OrcMaterial(C.Structure):
    _fields_ = [
        ("type", C.c_int),
        ("has_color", C.c_int), ("color", C.c_double * 3),
        ("has_roughness", C.c_int), ("roughness", C.c_double),
        ("has_metallic", C.c_int), ("metallic", C.c_double),
        ("has_specular", C.c_int), ("specular", C.c_double),
        ("has_ior", C.c_int), ("ior", C.c_double),
    ]


class OrcParams(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("samples", C.c_int), ("max_depth", C.c_int),
        ("jitter", C.c_int), ("recursive_reflections", C.c_int), ("soft_shadows", C.c_int),
        ("camera_mode", C.c_int), ("rng_mode", C.c_int),
        ("seed", C.c_ulonglong),
        ("threads", C.c_int),
        ("x0", C.c_int), ("y0", C.c_int), ("x1", C.c_int), ("y1", C.c_int),
        ("use_accel", C.c_int),
    ]


class OrcCounters(C.Structure):
    _fields_ = [(n, C.c_ulonglong) for n in
                ("samples", "hit_world", "sphere_tests", "tri_tests", "shadow_rays", "scatters")]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.orc_scene_new.restype = C.c_void_p
        L.orc_scene_free.argtypes = [C.c_void_p]
        L.orc_scene_set_camera.argtypes = [C.c_void_p, dp, dp, dp, C.c_double, C.c_double]
        L.orc_scene_add_sphere.argtypes = [C.c_void_p, dp, C.c_double, C.POINTER(OrcMaterial)]
        L.orc_scene_add_cube.argtypes = [C.c_void_p, dp, dp, C.POINTER(OrcMaterial)]
        L.orc_scene_add_prism.argtypes = [C.c_void_p, dp, C.POINTER(OrcMaterial)]
        L.orc_scene_add_mesh.argtypes = [C.c_void_p, dp, C.c_int, C.POINTER(OrcMaterial)]
        L.orc_scene_add_light.argtypes = [C.c_void_p, dp, dp, C.c_double]
        L.orc_scene_set_fog.argtypes = [C.c_void_p, C.c_int, C.c_double, dp]
        L.orc_scene_set_sky.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_sky_color.argtypes = [C.c_void_p, dp, dp]
        ip = C.POINTER(C.c_int)
        L.orc_scene_counts.argtypes = [C.c_void_p, ip, ip, ip, ip]
        L.orc_scene_get_triangle.argtypes = [C.c_void_p, C.c_int, dp, ip]
        L.orc_scene_get_material.argtypes = [C.c_void_p, C.c_int, ip, dp]
        L.orc_render.argtypes = [C.c_void_p, C.POINTER(OrcParams), C.c_void_p, C.c_void_p, C.POINTER(OrcCounters)]
        for name in ("orc_vec_add", "orc_vec_sub", "orc_vec_mul", "orc_vec_cross", "orc_vec_reflect"):
            getattr(L, name).argtypes = [dp, dp, dp]
        L.orc_vec_dot.argtypes = [dp, dp]
        L.orc_vec_dot.restype = C.c_double
        L.orc_vec_length.argtypes = [dp]
        L.orc_vec_length.restype = C.c_double
        L.orc_vec_normalize.argtypes = [dp, dp]
        L.orc_vec_refract.argtypes = [dp, dp, C.c_double, dp]
        L.orc_vec_clamp.argtypes = [dp, C.c_double, C.c_double, dp]
        L.orc_vec_to_rgb.argtypes = [dp, C.POINTER(C.c_ubyte)]
        L.orc_reflectance.argtypes = [C.c_double, C.c_double]
        L.orc_reflectance.restype = C.c_double
        L.orc_schlick.argtypes = [C.c_double, C.c_double]
        L.orc_schlick.restype = C.c_double
        L.orc_tone_map.argtypes = [dp, dp]
        L.orc_tone_map_rgb.argtypes = [dp, C.POINTER(C.c_ubyte)]
        L.orc_sphere_hit.argtypes = [dp, C.c_double, dp, dp, C.c_double, C.c_double, dp]
        L.orc_triangle_hit.argtypes = [dp, dp, dp, C.c_double, C.c_double, dp]
        L.orc_scatter.argtypes = [C.POINTER(OrcMaterial), dp, dp, dp, dp, C.c_int, C.c_ulonglong,
                                  C.c_uint, C.c_uint, C.c_uint, dp]
        up = C.POINTER(C.c_uint)
        L.orc_philox4x32_10.argtypes = [up, up, up]
        L.orc_in_unit_sphere.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, dp]
        L.orc_in_unit_sphere_half.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_int, dp]
        L.orc_get_ray.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, dp]
        L.orc_hit_world.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_double, C.c_int, dp]
        _lib = L
    return _lib


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _vec3(v, default=(0.0, 0.0, 0.0)):
    """Vec3.UnmarshalJSON (internal/math/vector.go:176-193): array of 3 or {X,Y,Z} object."""
    if v is None:
        return tuple(default)
    if isinstance(v, dict):
        return (float(v.get("X", 0.0)), float(v.get("Y", 0.0)), float(v.get("Z", 0.0)))
    if len(v) != 3:
        raise ValueError("expected 3 elements for Vec3, got %d" % len(v))
    return (float(v[0]), float(v[1]), float(v[2]))


def make_material(md: dict) -> OrcMaterial:
    m = OrcMaterial()
    t = md["type"]  # reference: unchecked type assertion -> panic when absent (scene.go:105)
    m.type = MAT_TYPES.get(t, 0)  # default branch -> lambertian (scene.go:144-146)
    if "color" in md:
        m.has_color = 1
        m.color[:] = _vec3(md["color"])
    for key, has, fld in (("roughness", "has_roughness", "roughness"), ("metallic", "has_metallic", "metallic"),
                          ("specular", "has_specular", "specular"), ("refractionIndex", "has_ior", "ior")):
        if key in md:
            setattr(m, has, 1)
            setattr(m, fld, float(md[key]))
    return m


# NewDefaultAtmosphere / NewWhiteAtmosphere / NewSunsetAtmosphere / NewNightAtmosphere (atmosphere/atmosphere.go:28-98), fields in
# declaration order: SkyColorTop SkyColorBottom SunDirection SunColor SunIntensity SunSize RayleighScattering MieScattering
# AtmosphericDepth FogDensity FogColor HazeIntensity TimeOfDay
SKY_PRESETS = {
    "default": [0.6, 0.8, 1.0, 0.9, 0.95, 1.0, 0.0, 0.8, -0.6, 1.0, 0.98, 0.95, 1.2, 0.015, 0.6, 0.8, 1.0, 1.0, 0.98, 0.95, 0.3, 0.0, 0.9, 0.92, 0.95, 0.05, 0.6],
    "white": [0.98, 0.98, 1.0, 0.92, 0.92, 0.95, 0.0, 0.8, -0.6, 1.0, 0.99, 0.97, 0.8, 0.012, 0.9, 0.9, 0.95, 0.95, 0.95, 0.98, 0.2, 0.0, 0.95, 0.95, 0.98, 0.02, 0.6],
    "sunset": [1.0, 0.4, 0.2, 1.0, 0.8, 0.6, 0.0, 0.3, -0.9, 1.0, 0.6, 0.3, 1.2, 0.03, 1.0, 0.4, 0.2, 1.0, 0.8, 0.6, 0.8, 0.1, 1.0, 0.8, 0.6, 0.3, 0.8],
    "night": [0.1, 0.1, 0.3, 0.2, 0.2, 0.4, 0.0, -0.7, -0.7, 0.8, 0.8, 1.0, 0.3, 0.005, 0.1, 0.1, 0.3, 0.8, 0.8, 1.0, 0.2, 0.0, 0.1, 0.1, 0.2, 0.0, 0.0],
}
_SKY_FIELDS = (("skyColorTop", 0, 3), ("skyColorBottom", 3, 3), ("sunDirection", 6, 3), ("sunColor", 9, 3), ("sunIntensity", 12, 1),
               ("sunSize", 13, 1), ("rayleighScattering", 14, 3), ("mieScattering", 17, 3), ("atmosphericDepth", 20, 1),
               ("fogDensity", 21, 1), ("fogColor", 22, 3), ("hazeIntensity", 25, 1), ("timeOfDay", 26, 1))


def sky_params(block: dict):
    """the scene JSON's "sky" block (extension) -> the 27 AtmosphereConfig values: a preset, then per-field overrides"""
    q = list(SKY_PRESETS[block.get("preset", "default")])
    for name, off, n in _SKY_FIELDS:
        if name in block:
            v = block[name]
            q[off:off + n] = [float(x) for x in v] if n == 3 else [float(v)]
    return q


class Scene:
    """Oracle-side scene built from a parsed scene JSON dict (scene.go:12-39 schema)."""

    def __init__(self, desc: dict, prisms: bool = False, fog: bool = False, sky: bool = False):
        L = lib()
        self.h = C.c_void_p(L.orc_scene_new())
        cam = desc.get("camera", {})
        L.orc_scene_set_camera(self.h, _d3(_vec3(cam.get("position"))), _d3(_vec3(cam.get("lookAt"))),
                               _d3(_vec3(cam.get("up"), (0, 0, 0))), float(cam.get("fov", 0.0)),
                               float(cam.get("aspectRatio", 0.0)))
        for obj in desc.get("objects", []) or []:
            t = obj.get("type", "")
            if t == "sphere":
                m = make_material(obj["material"])
                L.orc_scene_add_sphere(self.h, _d3(_vec3(obj.get("position"))), float(obj.get("radius", 0.0)), C.byref(m))
            elif t == "cube":
                m = make_material(obj["material"])
                L.orc_scene_add_cube(self.h, _d3(_vec3(obj.get("position"))), _d3(_vec3(obj.get("size"))), C.byref(m))
            elif t == "triangularPrism" and prisms:
                m = make_material(obj["material"])
                flat = [float(c) for v in obj["vertices"] for c in _vec3(v)]
                L.orc_scene_add_prism(self.h, (C.c_double * 18)(*flat), C.byref(m))
            else:
                continue  # "Unknown object type" -> skipped (scene.go:80-82)
        for lt in desc.get("lights", []) or []:
            L.orc_scene_add_light(self.h, _d3(_vec3(lt.get("position"))), _d3(_vec3(lt.get("color"))),
                                  float(lt.get("intensity", 0.0)))
        fg = desc.get("fog") or {}
        if fog and fg.get("enabled"):
            L.orc_scene_set_fog(self.h, 1, float(fg.get("density", 0.0)), _d3(_vec3(fg.get("color"))))
        sk = desc.get("sky") or {}
        if sky and sk.get("enabled"):
            L.orc_scene_set_sky(self.h, 1, (C.c_double * 27)(*sky_params(sk)))

    def sky_color(self, direction):
        out = (C.c_double * 3)()
        lib().orc_sky_color(self.h, _d3(direction), out)
        return [out[0], out[1], out[2]]

    @classmethod
    def from_flat(cls, camera: dict, materials, spheres, triangles, lights, fog=None) -> "Scene":
        """Mirror of a flat gort_scene_desc (tests/synth.py): materials are post-constructor dicts with an
        integer type; spheres (center, radius, material, order); triangles (v9, material, order) — runs of
        triangles sharing a material become one Mesh, in scan order."""
        self = cls({"camera": camera, "objects": [], "lights": []})
        L = lib()

        def mat(md):
            m = OrcMaterial()
            m.type = int(md["type"])
            m.has_color, m.has_roughness, m.has_metallic, m.has_specular, m.has_ior = 1, 1, 1, 1, 1
            m.color[:] = [float(x) for x in md["color"]]
            m.roughness, m.metallic, m.specular, m.ior = float(md["roughness"]), float(md["metallic"]), float(md["specular"]), float(md["ior"])
            return m

        items = [(s[3], "s", s) for s in spheres] + [(t[2], "t", t) for t in triangles]
        items.sort(key=lambda it: it[0])
        i = 0
        while i < len(items):
            _, kind, it = items[i]
            if kind == "s":
                m = mat(materials[it[2]])
                L.orc_scene_add_sphere(self.h, _d3(it[0]), float(it[1]), C.byref(m))
                i += 1
            else:
                j = i
                verts = []
                while j < len(items) and items[j][1] == "t" and items[j][2][1] == it[1]:
                    verts.extend(float(x) for x in items[j][2][0])
                    j += 1
                m = mat(materials[it[1]])
                L.orc_scene_add_mesh(self.h, (C.c_double * len(verts))(*verts), (j - i), C.byref(m))
                i = j
        for lt in lights:
            L.orc_scene_add_light(self.h, _d3(lt[0]), _d3(lt[1]), float(lt[2]))
        if fog:
            L.orc_scene_set_fog(self.h, 1, float(fog["density"]), _d3(fog["color"]))
        return self

    @classmethod
    def from_file(cls, path: str, **kw) -> "Scene":
        with open(path) as f:
            return cls(json.load(f), **kw)

    def __del__(self):
        try:
            lib().orc_scene_free(self.h)
        except Exception:
            pass

    def counts(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib().orc_scene_counts(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return {"spheres": a.value, "triangles": b.value, "hittables": c.value, "lights": d.value}

    def triangle(self, i):
        out = (C.c_double * 12)()
        mat = C.c_int()
        lib().orc_scene_get_triangle(self.h, i, out, C.byref(mat))
        return np.array(out[:]), mat.value

    def material(self, i):
        out = (C.c_double * 7)()
        t = C.c_int()
        lib().orc_scene_get_material(self.h, i, C.byref(t), out)
        return t.value, np.array(out[:])

    def render(self, width, height, samples=100, max_depth=50, jitter=True, recursive_reflections=True,
               soft_shadows=True, camera_mode=CAMERA_REFERENCE, rng_mode=RNG_MT, seed=0, threads=None,
               crop=None, use_accel=False, want_radiance=False):
        """Returns (rgba uint8 [H,W,4], radiance float64 [H,W,3] | None, counters dict)."""
        p = OrcParams()
        p.width, p.height, p.samples, p.max_depth = width, height, samples, max_depth
        p.jitter, p.recursive_reflections, p.soft_shadows = int(jitter), int(recursive_reflections), int(soft_shadows)
        p.camera_mode, p.rng_mode, p.seed = camera_mode, rng_mode, seed
        p.threads = threads or len(os.sched_getaffinity(0))
        if crop:
            p.x0, p.y0, p.x1, p.y1 = crop
        p.use_accel = int(use_accel)
        rgba = np.zeros((height, width, 4), dtype=np.uint8)
        rad = np.zeros((height, width, 3), dtype=np.float64) if want_radiance else None
        cnt = OrcCounters()
        lib().orc_render(self.h, C.byref(p), rgba.ctypes.data_as(C.c_void_p),
                         rad.ctypes.data_as(C.c_void_p) if rad is not None else None, C.byref(cnt))
        return rgba, rad, {n: getattr(cnt, n) for n, _ in OrcCounters._fields_}

    def get_ray(self, u, v, camera_mode=CAMERA_REFERENCE):
        out = (C.c_double * 6)()
        lib().orc_get_ray(self.h, camera_mode, u, v, out)
        return np.array(out[:3]), np.array(out[3:])

    def hit_world(self, ro, rd, tmin=0.001, tmax=float("inf"), use_accel=False):
        out = (C.c_double * 10)()
        ok = lib().orc_hit_world(self.h, _d3(ro), _d3(rd), tmin, tmax, int(use_accel), out)
        if not ok:
            return None
        o = np.array(out[:])
        return {"t": o[0], "point": o[1:4], "normal": o[4:7], "front_face": bool(o[7]), "material": int(o[8]), "prim": int(o[9])}


# ---- unit-level helpers (known-answer tests) --------------------------------
def _v3call(name, *vecs):
    out = (C.c_double * 3)()
    getattr(lib(), name)(*[_d3(v) for v in vecs], out)
    return np.array(out[:])


def vec_add(a, b): return _v3call("orc_vec_add", a, b)
def vec_sub(a, b): return _v3call("orc_vec_sub", a, b)
def vec_mul(a, b): return _v3call("orc_vec_mul", a, b)
def vec_cross(a, b): return _v3call("orc_vec_cross", a, b)
def vec_reflect(a, n): return _v3call("orc_vec_reflect", a, n)
def vec_normalize(a): return _v3call("orc_vec_normalize", a)
def vec_dot(a, b): return lib().orc_vec_dot(_d3(a), _d3(b))
def vec_length(a): return lib().orc_vec_length(_d3(a))


def vec_refract(a, n, eta):
    out = (C.c_double * 3)()
    lib().orc_vec_refract(_d3(a), _d3(n), eta, out)
    return np.array(out[:])


def vec_clamp(a, lo, hi):
    out = (C.c_double * 3)()
    lib().orc_vec_clamp(_d3(a), lo, hi, out)
    return np.array(out[:])


def vec_to_rgb(a):
    out = (C.c_ubyte * 3)()
    lib().orc_vec_to_rgb(_d3(a), out)
    return tuple(out[:])


def tone_map(c):
    out = (C.c_double * 3)()
    lib().orc_tone_map(_d3(c), out)
    return np.array(out[:])


def tone_map_rgb(c):
    out = (C.c_ubyte * 3)()
    lib().orc_tone_map_rgb(_d3(c), out)
    return tuple(out[:])


def reflectance(cosine, ref_idx): return lib().orc_reflectance(cosine, ref_idx)
def schlick(ior, cos_theta): return lib().orc_schlick(ior, cos_theta)


def _hit(out):
    o = np.array(out[:])
    return {"t": o[0], "point": o[1:4], "normal": o[4:7], "front_face": bool(o[7])}


def sphere_hit(center, radius, ro, rd, tmin=0.001, tmax=float("inf")):
    out = (C.c_double * 8)()
    ok = lib().orc_sphere_hit(_d3(center), radius, _d3(ro), _d3(rd), tmin, tmax, out)
    return _hit(out) if ok else None


def triangle_hit(v0, v1, v2, ro, rd, tmin=0.001, tmax=float("inf")):
    out = (C.c_double * 8)()
    v9 = (C.c_double * 9)(*[float(x) for v in (v0, v1, v2) for x in v])
    ok = lib().orc_triangle_hit(v9, _d3(ro), _d3(rd), tmin, tmax, out)
    return _hit(out) if ok else None


def scatter(material: dict, ro, rd, point, normal, front_face, seed=0, pixel=0, sample=0, bounce=0):
    m = make_material(material)
    out = (C.c_double * 9)()
    ok = lib().orc_scatter(C.byref(m), _d3(ro), _d3(rd), _d3(point), _d3(normal), int(front_face), seed, pixel, sample, bounce, out)
    if not ok:
        return None
    o = np.array(out[:])
    return {"origin": o[:3], "direction": o[3:6], "attenuation": o[6:9]}


def philox4x32_10(ctr, key):
    out = (C.c_uint * 4)()
    lib().orc_philox4x32_10((C.c_uint * 4)(*ctr), (C.c_uint * 2)(*key), out)
    return tuple(out[:])


def in_unit_sphere_half(seed, pixel, sample, bounce, stream, seq_base, half):
    out = (C.c_double * 3)()
    lib().orc_in_unit_sphere_half(seed, pixel, sample, bounce, stream, seq_base, half, out)
    return np.array(out[:])


def in_unit_sphere(seed, pixel, sample, bounce, stream, seq_base):
    out = (C.c_double * 3)()
    lib().orc_in_unit_sphere(seed, pixel, sample, bounce, stream, seq_base, out)
    return np.array(out[:])
