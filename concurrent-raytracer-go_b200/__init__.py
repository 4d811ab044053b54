"""Host-side mirror (Python/ctypes) of the reference's renderer surface over libgort.so.

The product is the C-ABI library `lib/libgort.so` (include/gort.h): hand-written sm_100a CUDA for
the per-pixel render hot path of JoshElkind/concurrent-raytracer-go.  This module only binds it
and mirrors the reference's Go API names so tests read like the reference's call sites
(/root/reference cmd/raytracer/main.go:38-65):

    scene    = LoadFromFile(path)                      # scene.LoadFromFile    scene/scene.go:45
    renderer = NewParallelRenderer(num_workers)        # renderer.go:54
    renderer.SetSamples(100); renderer.SetMaxDepth(50) # settings.go:3-25
    img      = renderer.Render(scene, width, height)   # renderer.go:67 -> [H, W, 4] uint8 (image.RGBA.Pix)

There is NO CPU fallback: importing works anywhere, but creating a renderer without the built
library or without a CUDA device raises GortError.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GORT_LIB") or os.path.join(_HERE, "lib", "libgort.so")  # GORT_LIB: A/B builds of the same ABI

ABI_VERSION = 3
TILE = 32
CAMERA_REFERENCE, CAMERA_LOOKAT = 0, 1
LOAD_PRISMS, LOAD_FOG, LOAD_SKY = 1, 2, 4

MAT_TYPES = {"lambertian": 0, "metal": 1, "shiny": 2, "perfectmirror": 3, "glass": 4, "dielectric": 5, "diffuselight": 6}

EXPORTED_SYMBOLS = [
    "gort_abi_version", "gort_device_count", "gort_create", "gort_destroy", "gort_last_error", "gort_set_stream",
    "gort_scene_upload", "gort_scene_load_json", "gort_scene_load_file", "gort_scene_counts", "gort_scene_get_triangle",
    "gort_scene_get_material", "gort_render", "gort_render_device", "gort_shard_slab_bytes", "gort_render_shard_device",
    "gort_unswizzle_device", "gort_read_radiance", "gort_trace_rays", "gort_measure_fp32_peak",
    "gort_host_scene_parse", "gort_host_scene_from_desc", "gort_host_scene_free", "gort_host_scene_counts", "gort_host_scene_get_sphere",
    "gort_host_scene_get_triangle", "gort_host_scene_get_material", "gort_host_scene_get_light", "gort_host_scene_get_camera",
    "gort_host_scene_bvh_validate",
    "gort_link_create", "gort_link_open", "gort_link_close", "gort_link_frame", "gort_render_linked",
    "gort_link_read", "gort_link_open_local", "gort_scene_render_hints", "gort_host_alloc", "gort_host_free",
]


class GortError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libgort error %d: %s" % (code, msg))
        self.code = code


class SceneDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("reserved0", C.c_uint32),
        ("cam_position", C.c_double * 3), ("cam_look_at", C.c_double * 3), ("cam_up", C.c_double * 3),
        ("cam_fov", C.c_double), ("cam_aspect", C.c_double),
        ("n_materials", C.c_int32), ("reserved1", C.c_int32),
        ("mat_type", C.POINTER(C.c_int32)), ("mat_color", C.POINTER(C.c_double)), ("mat_roughness", C.POINTER(C.c_double)),
        ("mat_metallic", C.POINTER(C.c_double)), ("mat_specular", C.POINTER(C.c_double)), ("mat_ior", C.POINTER(C.c_double)),
        ("n_spheres", C.c_int32), ("reserved2", C.c_int32),
        ("sphere_center", C.POINTER(C.c_double)), ("sphere_radius", C.POINTER(C.c_double)),
        ("sphere_material", C.POINTER(C.c_int32)), ("sphere_order", C.POINTER(C.c_int32)),
        ("n_triangles", C.c_int32), ("reserved3", C.c_int32),
        ("tri_vertices", C.POINTER(C.c_double)), ("tri_material", C.POINTER(C.c_int32)), ("tri_order", C.POINTER(C.c_int32)),
        ("n_lights", C.c_int32), ("reserved4", C.c_int32),
        ("light_position", C.POINTER(C.c_double)), ("light_color", C.POINTER(C.c_double)), ("light_intensity", C.POINTER(C.c_double)),
        ("fog_enabled", C.c_int32), ("reserved5", C.c_int32),
        ("fog_density", C.c_double), ("fog_color", C.c_double * 3),
        ("sky_enabled", C.c_int32), ("reserved6", C.c_int32), ("sky_params", C.c_double * 27),
    ]


class RenderParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("width", C.c_int32), ("height", C.c_int32), ("samples", C.c_int32), ("max_depth", C.c_int32),
        ("anti_aliasing", C.c_int32), ("recursive_reflections", C.c_int32), ("soft_shadows", C.c_int32),
        ("camera_mode", C.c_int32), ("shard_rank", C.c_int32), ("shard_count", C.c_int32), ("collect_stats", C.c_int32),
        ("seed", C.c_uint64),
        ("crop_x0", C.c_int32), ("crop_y0", C.c_int32), ("crop_x1", C.c_int32), ("crop_y1", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_double), ("trace_ms", C.c_double), ("resolve_ms", C.c_double), ("total_ms", C.c_double),
        ("upload_ms", C.c_double), ("bvh_build_ms", C.c_double), ("device_ms", C.c_double * 8),
        ("n_devices", C.c_int32), ("n_tiles", C.c_int32),
        ("primary_rays", C.c_uint64), ("bvh_nodes", C.c_uint64), ("bvh_bytes", C.c_uint64),
        ("closest_queries", C.c_uint64), ("shadow_queries", C.c_uint64), ("nodes_visited", C.c_uint64),
        ("sphere_tests", C.c_uint64), ("sphere_hits", C.c_uint64), ("tri_tests", C.c_uint64), ("tri_hits", C.c_uint64),
        ("tri_rejects", C.c_uint64 * 4),
        ("shaded_hits", C.c_uint64), ("rng_blocks", C.c_uint64), ("light_evals", C.c_uint64), ("soft_shadow_rays", C.c_uint64),
        ("diffuse_evals", C.c_uint64), ("specular_evals", C.c_uint64),
        ("paths_depth_ge5", C.c_uint64), ("paths_depth_ge20", C.c_uint64), ("paths_depth_max", C.c_uint64),
        ("cone_tests", C.c_uint64),
        ("walk_lane_visits", C.c_uint64 * 4), ("walk_warp_visits", C.c_uint64 * 4),
        ("algorithmic_flops", C.c_double),
        ("soft_pairs_skipped", C.c_uint64),
        ("pairs_backfacing", C.c_uint64),
        ("cull_ms", C.c_double), ("primary_generated", C.c_uint64), ("render_path", C.c_int32), ("kernel_launches", C.c_int32),
    ]

    def as_dict(self) -> dict:
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


_lib = None


def load_library() -> C.CDLL:
    """dlopen lib/libgort.so; raises GortError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GortError(-3, "libgort.so not built at %s — run __graft_entry__.build() / make -C %s" % (LIB_PATH, _HERE))
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.gort_abi_version.restype = C.c_int
    L.gort_device_count.restype = C.c_int
    L.gort_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.gort_destroy.argtypes = [vp]
    L.gort_destroy.restype = None
    L.gort_last_error.argtypes = [vp]
    L.gort_last_error.restype = C.c_char_p
    L.gort_set_stream.argtypes = [vp, vp]
    L.gort_scene_upload.argtypes = [vp, C.POINTER(SceneDesc)]
    L.gort_scene_load_json.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_uint32]
    L.gort_scene_load_file.argtypes = [vp, C.c_char_p, C.c_uint32]
    L.gort_scene_counts.argtypes = [vp, ip, ip, ip, ip, ip]
    L.gort_scene_get_triangle.argtypes = [vp, C.c_int32, dp, ip]
    L.gort_scene_get_material.argtypes = [vp, C.c_int32, ip, dp]
    L.gort_render.argtypes = [vp, C.POINTER(RenderParams), vp, C.c_size_t, C.POINTER(Stats)]
    L.gort_render_device.argtypes = [vp, C.POINTER(RenderParams), vp, C.c_size_t, C.POINTER(Stats)]
    L.gort_shard_slab_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.gort_shard_slab_bytes.restype = C.c_size_t
    L.gort_render_shard_device.argtypes = [vp, C.POINTER(RenderParams), vp, C.c_size_t, C.POINTER(Stats)]
    L.gort_unswizzle_device.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, C.c_size_t]
    L.gort_read_radiance.argtypes = [vp, dp, C.c_size_t]
    L.gort_trace_rays.argtypes = [vp, C.c_int32, dp, dp, C.c_double, C.c_double, C.c_int32, dp, ip]
    L.gort_measure_fp32_peak.argtypes = [vp, dp, dp]
    L.gort_host_scene_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.gort_host_scene_from_desc.argtypes = [C.POINTER(SceneDesc), C.POINTER(vp), C.c_char_p, C.c_size_t]
    L.gort_host_scene_free.argtypes = [vp]
    L.gort_host_scene_free.restype = None
    L.gort_host_scene_counts.argtypes = [vp, ip]
    L.gort_host_scene_get_sphere.argtypes = [vp, C.c_int32, dp, ip, ip]
    L.gort_host_scene_get_triangle.argtypes = [vp, C.c_int32, dp, ip, ip]
    L.gort_host_scene_get_material.argtypes = [vp, C.c_int32, ip, dp]
    L.gort_host_scene_get_light.argtypes = [vp, C.c_int32, dp]
    L.gort_host_scene_get_camera.argtypes = [vp, dp]
    L.gort_host_scene_bvh_validate.argtypes = [vp, C.POINTER(C.c_int64), C.c_char_p, C.c_size_t]
    L.gort_link_create.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.POINTER(vp)]
    L.gort_link_open.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.gort_link_close.argtypes = [vp, vp]
    L.gort_link_close.restype = None
    L.gort_link_frame.argtypes = [vp]
    L.gort_link_frame.restype = vp
    L.gort_render_linked.argtypes = [vp, C.POINTER(RenderParams), vp, C.POINTER(Stats)]
    L.gort_link_read.argtypes = [vp, vp, vp, C.c_size_t]
    L.gort_scene_render_hints.argtypes = [vp, ip]
    L.gort_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.gort_host_free.argtypes = [vp]
    L.gort_host_free.restype = None
    L.gort_link_open_local.argtypes = [vp, vp, C.c_int32, C.POINTER(vp)]
    _lib = L
    return L


# --------------------------------------------------------------------------------------------
# scene.Scene mirror (internal/scene/scene.go:12-39)
# --------------------------------------------------------------------------------------------
class Scene:
    """Holds the scene JSON text (what scene.LoadFromFile reads); parsing, createMaterial defaults and
    cube expansion happen in libgort's C++ loader (csrc/host_scene.cpp)."""

    def __init__(self, json_text: str, options: int = 0, name: str = "demo_scene"):
        self.json_text = json_text
        self.options = options
        self.name = name  # GetSceneName() is the constant "demo_scene" (scene.go:100-102)

    def GetSceneName(self) -> str:
        return "demo_scene"


def LoadFromFile(filename: str, options: int = 0) -> Scene:
    """scene.LoadFromFile (scene.go:45-57)."""
    try:
        with open(filename) as f:
            text = f.read()
    except OSError as e:
        raise GortError(-7, "error reading file: %s" % e)
    return Scene(text, options, name=os.path.splitext(os.path.basename(filename))[0])


def SceneFromDict(desc: dict, options: int = 0) -> Scene:
    return Scene(json.dumps(desc), options)


class HostScene:
    """Host-only parse of the reference's scene JSON by libgort's loader (no CUDA needed): the
    flattened GetHittables()/GetLights() view in the reference's scan order."""

    def __init__(self, json_text: Optional[str], options: int = 0):
        L = load_library()
        self._L = L
        self._h = C.c_void_p()
        if json_text is None:  # from_desc fills it
            return
        err = C.create_string_buffer(512)
        raw = json_text.encode()
        rc = L.gort_host_scene_parse(raw, len(raw), options, C.byref(self._h), err, len(err))
        if rc != 0:
            raise GortError(rc, err.value.decode())

    @classmethod
    def from_desc(cls, flat: "FlatScene", into: Optional["HostScene"] = None) -> "HostScene":
        """The host side of UploadScene(FlatScene): validation + copy of a gort_scene_desc, no CUDA.  `into`: rebuild that
        host scene in place (its arrays are reused)."""
        hs = into if into is not None else cls(None)
        err = C.create_string_buffer(512)
        rc = hs._L.gort_host_scene_from_desc(C.byref(flat.desc), C.byref(hs._h), err, len(err))
        if rc != 0:
            raise GortError(rc, err.value.decode())
        return hs

    def __del__(self):
        try:
            if self._h.value:
                self._L.gort_host_scene_free(self._h)
        except Exception:
            pass

    def counts(self) -> dict:
        c = (C.c_int32 * 5)()
        self._L.gort_host_scene_counts(self._h, c)
        return dict(zip(("spheres", "triangles", "materials", "lights", "hittables"), list(c)))

    def sphere(self, i):
        o = (C.c_double * 4)()
        m, k = C.c_int32(), C.c_int32()
        if self._L.gort_host_scene_get_sphere(self._h, i, o, C.byref(m), C.byref(k)) != 0:
            raise IndexError(i)
        return np.array(o[:3]), o[3], m.value, k.value

    def triangle(self, i):
        o = (C.c_double * 9)()
        m, k = C.c_int32(), C.c_int32()
        if self._L.gort_host_scene_get_triangle(self._h, i, o, C.byref(m), C.byref(k)) != 0:
            raise IndexError(i)
        return np.array(o[:]), m.value, k.value

    def material(self, i):
        o = (C.c_double * 7)()
        t = C.c_int32()
        if self._L.gort_host_scene_get_material(self._h, i, C.byref(t), o) != 0:
            raise IndexError(i)
        return t.value, np.array(o[:])

    def light(self, i):
        o = (C.c_double * 7)()
        if self._L.gort_host_scene_get_light(self._h, i, o) != 0:
            raise IndexError(i)
        return np.array(o[:])

    def camera(self):
        o = (C.c_double * 11)()
        self._L.gort_host_scene_get_camera(self._h, o)
        return np.array(o[:])

    def bvh_validate(self) -> dict:
        info = (C.c_int64 * 4)()
        err = C.create_string_buffer(256)
        rc = self._L.gort_host_scene_bvh_validate(self._h, info, err, len(err))
        if rc != 0:
            raise GortError(rc, err.value.decode())
        return dict(zip(("nodes", "max_depth", "leaves", "bytes"), list(info)))

    def to_flat(self) -> "FlatScene":
        """Re-express the parsed scene as a gort_scene_desc (what a Go host's Flatten() would pass)."""
        c = self.counts()
        cam = self.camera()
        mats = []
        for i in range(c["materials"]):
            t, v = self.material(i)
            mats.append({"type": t, "color": v[:3], "roughness": v[3], "metallic": v[4], "specular": v[5], "ior": v[6]})
        sph = [(s[0], s[1], s[2], s[3]) for s in (self.sphere(i) for i in range(c["spheres"]))]
        tri = [self.triangle(i) for i in range(c["triangles"])]
        lts = [(l[:3], l[3:6], l[6]) for l in (self.light(i) for i in range(c["lights"]))]
        return FlatScene({"position": cam[:3], "lookAt": cam[3:6], "up": cam[6:9], "fov": cam[9], "aspectRatio": cam[10]},
                         mats, sph, tri, lts)


class FlatScene:
    """A gort_scene_desc built from Python arrays (what a Go host would pass after Flatten())."""

    def __init__(self, camera: dict, materials: Sequence[dict], spheres=(), triangles=(), lights=(), fog=None, sky=None):
        """materials: dicts with type (int), color, roughness, metallic, specular, ior (post-constructor values).
        spheres: (center3, radius, material, order); triangles: (v9, material, order); lights: (pos3, color3, intensity)."""
        self._keep = []
        d = SceneDesc()
        d.abi_version = ABI_VERSION
        d.cam_position[:] = camera.get("position", (0, 0, 0))
        d.cam_look_at[:] = camera.get("lookAt", (0, 0, 0))
        d.cam_up[:] = camera.get("up", (0, 1, 0))
        d.cam_fov = camera.get("fov", 60.0)
        d.cam_aspect = camera.get("aspectRatio", 1.0)

        def arr(values, ctype, nptype):
            a = np.ascontiguousarray(np.array(values, dtype=nptype).reshape(-1))
            self._keep.append(a)
            return a.ctypes.data_as(C.POINTER(ctype))

        d.n_materials = len(materials)
        if materials:
            d.mat_type = arr([m["type"] for m in materials], C.c_int32, np.int32)
            d.mat_color = arr([m.get("color", (1, 1, 1)) for m in materials], C.c_double, np.float64)
            d.mat_roughness = arr([m.get("roughness", 0.0) for m in materials], C.c_double, np.float64)
            d.mat_metallic = arr([m.get("metallic", 0.0) for m in materials], C.c_double, np.float64)
            d.mat_specular = arr([m.get("specular", 0.0) for m in materials], C.c_double, np.float64)
            d.mat_ior = arr([m.get("ior", 1.5) for m in materials], C.c_double, np.float64)
        d.n_spheres = len(spheres)
        if len(spheres):
            d.sphere_center = arr([s[0] for s in spheres], C.c_double, np.float64)
            d.sphere_radius = arr([s[1] for s in spheres], C.c_double, np.float64)
            d.sphere_material = arr([s[2] for s in spheres], C.c_int32, np.int32)
            d.sphere_order = arr([s[3] for s in spheres], C.c_int32, np.int32)
        d.n_triangles = len(triangles)
        if len(triangles):
            d.tri_vertices = arr([t[0] for t in triangles], C.c_double, np.float64)
            d.tri_material = arr([t[1] for t in triangles], C.c_int32, np.int32)
            d.tri_order = arr([t[2] for t in triangles], C.c_int32, np.int32)
        d.n_lights = len(lights)
        if len(lights):
            d.light_position = arr([l[0] for l in lights], C.c_double, np.float64)
            d.light_color = arr([l[1] for l in lights], C.c_double, np.float64)
            d.light_intensity = arr([l[2] for l in lights], C.c_double, np.float64)
        if fog:
            d.fog_enabled = 1
            d.fog_density = fog["density"]
            d.fog_color[:] = fog["color"]
        if sky:  # the 27 AtmosphereConfig values (atmosphere/atmosphere.go:8-26), extension
            d.sky_enabled = 1
            d.sky_params[:] = [float(x) for x in sky]
        self.desc = d


# --------------------------------------------------------------------------------------------
# renderer.ParallelRenderer mirror (internal/renderer/renderer.go:20-126; settings.go:3-37)
# --------------------------------------------------------------------------------------------
class ParallelRenderer:
    def __init__(self, numWorkers: int = 1, devices: Optional[Sequence[int]] = None):
        """numWorkers keeps the reference's meaning of "how many workers" = how many GPUs of this
        process render the frame (worker_count in the benchmark JSON).  `devices` overrides ids."""
        L = load_library()
        self._L = L
        if L.gort_abi_version() != ABI_VERSION:
            raise GortError(-1, "libgort ABI version mismatch")
        ids = list(devices) if devices is not None else list(range(max(1, int(numWorkers))))
        self.numWorkers = len(ids)
        self._ctx = C.c_void_p()
        arr = (C.c_int * len(ids))(*ids)
        rc = L.gort_create(arr, len(ids), C.byref(self._ctx))
        if rc != 0:
            raise GortError(rc, (L.gort_last_error(None) or b"").decode())
        # NewParallelRenderer defaults (renderer.go:54-65)
        self.maxDepth = 50
        self.samples = 100
        self.antiAliasing = True
        self.recursiveReflections = True
        self.softShadows = True
        self.depthOfField = False
        # additive knobs (not in the reference)
        self.cameraMode = CAMERA_REFERENCE
        self.seed = 0
        self.shardRank, self.shardCount = 0, 1
        self.collectStats = False
        self.crop = (0, 0, 0, 0)
        self._scene_token = None
        self.lastStats: Optional[Stats] = None
        self._bench_raw = None
        self._params_cache = None

    # ---- settings.go:3-25 ----
    def SetSamples(self, samples: int): self.samples = int(samples)
    def SetMaxDepth(self, maxDepth: int): self.maxDepth = int(maxDepth)
    def SetAntiAliasing(self, antiAliasing: bool): self.antiAliasing = bool(antiAliasing)
    def SetRecursiveReflections(self, v: bool): self.recursiveReflections = bool(v)
    def SetSoftShadows(self, v: bool): self.softShadows = bool(v)
    def SetDepthOfField(self, v: bool): self.depthOfField = bool(v)
    # ---- additive ----
    def SetCameraMode(self, mode: int): self.cameraMode = int(mode)
    def SetSeed(self, seed: int): self.seed = int(seed)
    def SetShard(self, rank: int, count: int): self.shardRank, self.shardCount = int(rank), int(count)
    def SetCollectStats(self, v: bool): self.collectStats = bool(v)

    def SetCrop(self, x0: int = 0, y0: int = 0, x1: int = 0, y1: int = 0):
        """region render (RenderChunk): only [x0, x1) x [y0, y1) is traced, the rest of the frame is black; () = whole frame"""
        self.crop = (int(x0), int(y0), int(x1), int(y1))

    def GetStats(self) -> dict:  # settings.go:27-37
        return {"workers": self.numWorkers, "samples": self.samples, "maxDepth": self.maxDepth,
                "antiAliasing": self.antiAliasing, "recursiveReflections": self.recursiveReflections,
                "softShadows": self.softShadows, "depthOfField": self.depthOfField}

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.gort_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise GortError(rc, (self._L.gort_last_error(self._ctx) or b"").decode())

    def set_stream(self, cuda_stream_ptr: int):
        """Launch on a caller-owned stream.  A torch default stream has handle 0, which the C ABI reads as
        "restore the ctx's own stream", so it is passed as cudaStreamLegacy (0x1)."""
        self._check(self._L.gort_set_stream(self._ctx, C.c_void_p(cuda_stream_ptr if cuda_stream_ptr else 1)))

    # ---- scene upload (GetHittables/GetLights feeding Render, renderer.go:72-74) ----
    def UploadScene(self, scene) -> None:
        if isinstance(scene, FlatScene):
            self._check(self._L.gort_scene_upload(self._ctx, C.byref(scene.desc)))
        elif isinstance(scene, Scene):
            raw = scene.json_text.encode()
            self._check(self._L.gort_scene_load_json(self._ctx, raw, len(raw), scene.options))
        else:
            raise TypeError("scene must be a Scene or FlatScene")
        self._scene_token = scene

    def _params(self, width: int, height: int) -> RenderParams:
        key = (width, height, self.samples, self.maxDepth, self.antiAliasing, self.recursiveReflections, self.softShadows,
               self.cameraMode, self.shardRank, self.shardCount, self.collectStats, self.seed, self.crop)
        if self._params_cache is not None and self._params_cache[0] == key:
            return self._params_cache[1]
        p = RenderParams()
        p.abi_version = ABI_VERSION
        p.width, p.height, p.samples, p.max_depth = width, height, self.samples, self.maxDepth
        p.anti_aliasing = int(self.antiAliasing)
        p.recursive_reflections = int(self.recursiveReflections)
        p.soft_shadows = int(self.softShadows)
        p.camera_mode = self.cameraMode
        p.shard_rank, p.shard_count = self.shardRank, self.shardCount
        p.collect_stats = int(self.collectStats)
        p.seed = self.seed
        p.crop_x0, p.crop_y0, p.crop_x1, p.crop_y1 = self.crop
        self._params_cache = (key, p)
        return p

    def Render(self, scene, width: int, height: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """ParallelRenderer.Render (renderer.go:67-126): blocking; returns image.RGBA.Pix as [H, W, 4] uint8.
        A scene object is uploaded when it is first seen and is treated as immutable afterwards (the reference's Scene is
        not mutated during Render either): after changing a FlatScene in place, call UploadScene(scene) again."""
        start = time.time()
        if scene is not self._scene_token:
            self.UploadScene(scene)
        img = out if out is not None else np.zeros((height, width, 4), dtype=np.uint8)
        assert img.dtype == np.uint8 and img.flags["C_CONTIGUOUS"] and img.size == width * height * 4
        st = Stats()
        p = self._params(width, height)
        self._check(self._L.gort_render(self._ctx, C.byref(p), img.ctypes.data_as(C.c_void_p), img.size, C.byref(st)))
        self.lastStats = st
        # BenchmarkData (renderer.go:31-42,103-117) is recorded per frame; the dict itself is only built when somebody
        # reads it (benchmarkData property): the formatting costs more than the cgo-sized call it sits next to
        self._bench_raw = (width, height, time.time() - start, time.time(), self.samples, self.maxDepth)
        return img

    @property
    def benchmarkData(self) -> dict:
        """BenchmarkData of the last Render (renderer.go:31-42,103-117); {} before the first frame."""
        if self._bench_raw is None:
            return {}
        width, height, seconds, when, samples, depth = self._bench_raw
        counts = self.SceneCounts()
        return {
            "scene_name": "demo_scene", "resolution": "%dx%d" % (width, height), "render_time_seconds": seconds,
            "samples": samples, "max_depth": depth, "num_workers": self.numWorkers,
            "objects": counts["hittables"], "lights": counts["lights"],
            "timestamp": time.strftime("%Y-%m-%dT%H:%M:%S%z", time.localtime(when)),
            "features": ["Improved metallic reflections with Fresnel effect",
                         "Shiny materials with configurable roughness and specular",
                         "Enhanced light source reflections", "Better specular highlights for metallic surfaces"],
        }

    def RenderDevice(self, width: int, height: int, d_rgba_ptr: int, want_stats: bool = False) -> Optional[Stats]:
        """Frame into device memory (row-major RGBA8 at d_rgba_ptr); asynchronous unless want_stats."""
        p = self._params(width, height)
        st = Stats() if want_stats else None
        self._check(self._L.gort_render_device(self._ctx, C.byref(p), C.c_void_p(d_rgba_ptr), width * height * 4,
                                               C.byref(st) if st is not None else None))
        if st is not None:
            self.lastStats = st
        return st

    def RenderShardDevice(self, width: int, height: int, d_slab_ptr: int, want_stats: bool = False) -> Optional[Stats]:
        p = self._params(width, height)
        nbytes = self._L.gort_shard_slab_bytes(width, height, self.shardCount)
        st = Stats() if want_stats else None
        self._check(self._L.gort_render_shard_device(self._ctx, C.byref(p), C.c_void_p(d_slab_ptr), nbytes,
                                                     C.byref(st) if st is not None else None))
        if st is not None:
            self.lastStats = st
        return st

    # ---- frame link: one process per GPU, tiles stored straight into the owner's frame over NVLink peer memory ----
    def LinkCreate(self, width: int, height: int, n_ranks: int):
        """owner rank: -> (link, 64-byte handle to send to the peers)"""
        buf = C.create_string_buffer(64)
        link = C.c_void_p()
        self._check(self._L.gort_link_create(self._ctx, width, height, n_ranks, buf, C.byref(link)))
        return link, bytes(buf.raw)

    def LinkOpen(self, handle: bytes, width: int, height: int, n_ranks: int, rank: int):
        link = C.c_void_p()
        self._check(self._L.gort_link_open(self._ctx, handle, width, height, n_ranks, rank, C.byref(link)))
        return link

    def LinkOpenLocal(self, owner_link, rank: int):
        link = C.c_void_p()
        self._check(self._L.gort_link_open_local(self._ctx, owner_link, rank, C.byref(link)))
        return link

    def LinkRead(self, link, width: int, height: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        img = out if out is not None else np.zeros((height, width, 4), dtype=np.uint8)
        self._check(self._L.gort_link_read(self._ctx, link, img.ctypes.data_as(C.c_void_p), img.size))
        return img

    def LinkClose(self, link) -> None:
        self._L.gort_link_close(self._ctx, link)

    def LinkFrame(self, link) -> int:
        return int(self._L.gort_link_frame(link))

    def RenderLinked(self, width: int, height: int, link, want_stats: bool = False) -> Optional[Stats]:
        p = self._params(width, height)
        st = Stats() if want_stats else None
        self._check(self._L.gort_render_linked(self._ctx, C.byref(p), link, C.byref(st) if st is not None else None))
        if st is not None:
            self.lastStats = st
        return st

    def UnswizzleDevice(self, d_slabs_ptr: int, shard_count: int, width: int, height: int, d_rgba_ptr: int):
        self._check(self._L.gort_unswizzle_device(self._ctx, C.c_void_p(d_slabs_ptr), shard_count, width, height,
                                                  C.c_void_p(d_rgba_ptr), width * height * 4))

    def ReadRadiance(self, width: int, height: int) -> np.ndarray:
        out = np.zeros((height, width, 3), dtype=np.float64)
        self._check(self._L.gort_read_radiance(self._ctx, out.ctypes.data_as(C.POINTER(C.c_double)), out.nbytes))
        return out

    def SceneRenderHints(self) -> dict:
        """The uploaded scene's own "renderer" block (extension; None where absent)."""
        h = (C.c_int32 * 5)()
        self._check(self._L.gort_scene_render_hints(self._ctx, h))
        keys = ("samples", "maxDepth", "antiAliasing", "recursiveReflections", "softShadows")
        return {k: (None if v < 0 else (int(v) if i < 2 else bool(v))) for i, (k, v) in enumerate(zip(keys, h))}

    def SceneCounts(self) -> dict:
        v = [C.c_int32() for _ in range(5)]
        self._check(self._L.gort_scene_counts(self._ctx, *[C.byref(x) for x in v]))
        return dict(zip(("spheres", "triangles", "materials", "lights", "hittables"), [x.value for x in v]))

    def SceneTriangle(self, i: int):
        v9 = (C.c_double * 9)()
        mat = C.c_int32()
        self._check(self._L.gort_scene_get_triangle(self._ctx, i, v9, C.byref(mat)))
        return np.array(v9[:]), mat.value

    def SceneMaterial(self, i: int):
        out = (C.c_double * 7)()
        t = C.c_int32()
        self._check(self._L.gort_scene_get_material(self._ctx, i, C.byref(t), out))
        return t.value, np.array(out[:])

    def TraceRays(self, origins, directions, t_min=0.001, t_max=float("inf"), any_hit=False):
        o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
        n = o.shape[0]
        t = np.zeros(n, dtype=np.float64)
        order = np.zeros(n, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        self._check(self._L.gort_trace_rays(self._ctx, n, o.ctypes.data_as(dp), d.ctypes.data_as(dp), t_min, t_max, int(any_hit),
                                            t.ctypes.data_as(dp), order.ctypes.data_as(C.POINTER(C.c_int32))))
        return t, order

    def MeasureFp32Peak(self):
        tf, ms = C.c_double(), C.c_double()
        self._check(self._L.gort_measure_fp32_peak(self._ctx, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    # ---- output (renderer.go:438-451,473-485) ----
    def SaveImage(self, img: np.ndarray, filename: str) -> None:
        from PIL import Image
        d = os.path.dirname(filename)
        if d:
            os.makedirs(d, exist_ok=True)
        Image.fromarray(img, "RGBA").save(filename)

    def SaveBenchmarkData(self, outputPath: str) -> None:
        d = os.path.dirname(outputPath)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(outputPath, "w") as f:
            json.dump(self.benchmarkData, f, indent=2)


def NewParallelRenderer(numWorkers: int = 1, devices: Optional[Sequence[int]] = None) -> ParallelRenderer:
    """renderer.NewParallelRenderer (renderer.go:54-65)."""
    return ParallelRenderer(numWorkers, devices)


class HostFrame:
    """Page-locked [H, W, 4] uint8 frame (gort_host_alloc): Render(..., out=frame.array) needs no staging copy."""

    def __init__(self, width: int, height: int):
        self._L = load_library()
        p = C.c_void_p()
        rc = self._L.gort_host_alloc(width * height * 4, C.byref(p))
        if rc != 0:
            raise GortError(rc, (self._L.gort_last_error(None) or b"").decode())
        self._p = p
        self.array = np.ctypeslib.as_array((C.c_uint8 * (width * height * 4)).from_address(p.value)).reshape(height, width, 4)

    def close(self):
        if self._p:
            self.array = None
            self._L.gort_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_slab_bytes(width: int, height: int, shard_count: int) -> int:
    return int(load_library().gort_shard_slab_bytes(width, height, shard_count))


def tiles_of_shard(width: int, height: int, rank: int, count: int):
    """Static interleaved tile assignment: global row-major 32x32 tile ids owned by `rank`."""
    tx, ty = (width + TILE - 1) // TILE, (height + TILE - 1) // TILE
    return list(range(rank, tx * ty, count))


def unswizzle_host(slabs: np.ndarray, shard_count: int, width: int, height: int) -> np.ndarray:
    """numpy restatement of gort_unswizzle_device (host logic shared with the gloo tests):
    slabs [shard_count, tiles_per_shard, 32, 32, 4] -> frame [H, W, 4]."""
    tx, ty = (width + TILE - 1) // TILE, (height + TILE - 1) // TILE
    n_tiles = tx * ty
    per = (n_tiles + shard_count - 1) // shard_count
    s = np.asarray(slabs, dtype=np.uint8).reshape(shard_count, per, TILE, TILE, 4)
    out = np.zeros((height, width, 4), dtype=np.uint8)
    for t in range(n_tiles):
        x0, y0 = (t % tx) * TILE, (t // tx) * TILE
        w, h = min(TILE, width - x0), min(TILE, height - y0)
        out[y0:y0 + h, x0:x0 + w] = s[t % shard_count, t // shard_count, :h, :w]
    return out
