/*
 * gort.h — C ABI of libgort.so: the B200-native (sm_100a) replacement for the per-pixel render
 * hot path of JoshElkind/concurrent-raytracer-go.
 *
 * The reference is pure Go and has no FFI of its own; the seam this ABI replaces is
 * the body of
 *     func (r *ParallelRenderer) Render(scene *scene.Scene, width, height int) *image.RGBA
 *                                                    (internal/renderer/renderer.go:67-126)
 * together with the state it reads: the ParallelRenderer fields/setters
 * (renderer.go:20-29,54-65; settings.go:3-25), scene.GetHittables/GetLights/Camera
 * (internal/scene/scene.go:12-39,59-98) and the collector's toneMap + ToRGB + img.Set
 * (renderer.go:92-97,348-367; internal/math/vector.go:106-109).
 * One cgo call per frame; the Go signatures stay, only Render's body changes.
 *
 * Conventions: plain pointers and sizes only; the caller owns every host pointer; the library
 * copies what it needs before returning and retains no caller pointer (cgo pointer rules).
 * Every function returns 0 (GORT_OK) or a negative gort_status; gort_last_error() gives the text.
 * A gort_ctx is externally synchronised (one call at a time per ctx — the reference's Render is
 * not re-entrant either: renderer.go:103-117).  There is no CPU fallback: without a usable CUDA
 * device gort_create fails with GORT_ERR_NO_DEVICE.
 */
#ifndef GORT_H
#define GORT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GORT_ABI_VERSION 3u /* 3: sky extension in gort_scene_desc; gort_stats: cull_ms, primary_generated, render_path, kernel_launches; gort_render_params: crop */
#define GORT_MAX_DEVICES 8
#define GORT_TILE 32 /* createRenderTasks tileSize, renderer.go:401 */

typedef enum gort_status {
    GORT_OK = 0,
    GORT_ERR_INVALID = -1,   /* bad argument / ABI version / size */
    GORT_ERR_CUDA = -2,      /* CUDA runtime error (text in gort_last_error) */
    GORT_ERR_NO_DEVICE = -3, /* no CUDA device: there is no CPU fallback */
    GORT_ERR_NO_SCENE = -4,  /* gort_render before gort_scene_upload */
    GORT_ERR_NCCL = -5,
    GORT_ERR_PARSE = -6, /* scene JSON malformed, or would panic in the reference (scene.go:105-146) */
    GORT_ERR_IO = -7
} gort_status;

/* material.Material implementations reachable from createMaterial (scene.go:104-148) */
typedef enum gort_material_type {
    GORT_MAT_LAMBERTIAN = 0,    /* material.go:18-55 */
    GORT_MAT_METAL = 1,         /* material.go:57-149 */
    GORT_MAT_SHINY = 2,         /* material.go:151-225 */
    GORT_MAT_PERFECTMIRROR = 3, /* advanced_materials.go:111-171 */
    GORT_MAT_GLASS = 4,         /* advanced_materials.go:9-66 */
    GORT_MAT_DIELECTRIC = 5,    /* material.go:227-280 */
    GORT_MAT_DIFFUSELIGHT = 6   /* material.go:288-318 */
} gort_material_type;

typedef enum gort_camera_mode {
    GORT_CAMERA_REFERENCE = 0, /* getRay, renderer.go:377-390: position+aspect only, looks down -Z, v=0 is row 0 */
    GORT_CAMERA_LOOKAT = 1     /* extension: honours lookAt/up/fov (scene.go:18-24 fields), upright image */
} gort_camera_mode;

/*
 * Flat scene description = what scene.GetHittables()/GetLights()/Camera hand to Render, after
 * createMaterial's defaults (scene.go:104-148) and createCube's 12-triangle expansion
 * (scene.go:150-190) have been applied by the host.  Struct-of-arrays, float64 like the reference.
 * `*_order` is the primitive's position in the reference's linear scan (hitWorld renderer.go:337,
 * Mesh.Hit scene.go:200): it decides exact-tie winners (last wins).
 */
typedef struct gort_scene_desc {
    uint32_t abi_version; /* GORT_ABI_VERSION */
    uint32_t reserved0;

    /* scene.Camera (scene.go:18-24) */
    double cam_position[3];
    double cam_look_at[3];
    double cam_up[3];
    double cam_fov;
    double cam_aspect;

    /* materials, one per object (post-constructor values: roughness/metallic/specular already min(x,1)) */
    int32_t n_materials;
    int32_t reserved1;
    const int32_t* mat_type;     /* gort_material_type [n] */
    const double* mat_color;     /* Albedo / Color / Emit [3n] */
    const double* mat_roughness; /* [n] */
    const double* mat_metallic;  /* [n]  (GetMetallic: PerfectMirror -> 1 is applied by the library) */
    const double* mat_specular;  /* [n] */
    const double* mat_ior;       /* [n]  RefractionIndex (glass, dielectric); Metal/Shiny 1.5, PerfectMirror 2.0 */

    /* geometry.Sphere (sphere.go:8-20) */
    int32_t n_spheres;
    int32_t reserved2;
    const double* sphere_center;     /* [3n] */
    const double* sphere_radius;     /* [n] */
    const int32_t* sphere_material;  /* [n] */
    const int32_t* sphere_order;     /* [n] */

    /* geometry.Triangle (triangle.go:7-20); flat-shaded: Normals = face normal */
    int32_t n_triangles;
    int32_t reserved3;
    const double* tri_vertices;   /* [9n] v0 v1 v2 */
    const int32_t* tri_material;  /* [n] */
    const int32_t* tri_order;     /* [n] */

    /* scene.Light (scene.go:34-39); Type is never read by the renderer */
    int32_t n_lights;
    int32_t reserved4;
    const double* light_position;  /* [3n] */
    const double* light_color;     /* [3n] */
    const double* light_intensity; /* [n] */

    /* extension: exponential fog on the primary-hit distance; 0 = reference behaviour */
    int32_t fog_enabled;
    int32_t reserved5;
    double fog_density;
    double fog_color[3];

    /* extension (second half): sky gradient as the colour of a ray that leaves the scene, after
     * AtmosphereConfig.GetSkyColor (internal/atmosphere/atmosphere.go:100-135; the reference never calls it: a miss is
     * black, renderer.go:171-173).  0 = reference behaviour.  sky_params = the AtmosphereConfig fields in declaration order
     * (atmosphere.go:8-26): SkyColorTop[3] SkyColorBottom[3] SunDirection[3] SunColor[3] SunIntensity SunSize
     * RayleighScattering[3] MieScattering[3] AtmosphericDepth FogDensity FogColor[3] HazeIntensity TimeOfDay. */
    int32_t sky_enabled;
    int32_t reserved6;
    double sky_params[27];
} gort_scene_desc;

/* ParallelRenderer fields (renderer.go:20-29) + the additive knobs of this implementation. */
typedef struct gort_render_params {
    uint32_t abi_version; /* GORT_ABI_VERSION */
    int32_t width, height;
    int32_t samples;               /* renderer.go:58 default 100 */
    int32_t max_depth;             /* renderer.go:57 default 50 */
    int32_t anti_aliasing;         /* 1 = jitter each sample (what the reference always does, renderer.go:155-156);
                                      0 = declared extension: sub-pixel offset (0.5,0.5) */
    int32_t recursive_reflections; /* renderer.go:60 */
    int32_t soft_shadows;          /* renderer.go:61 */
    int32_t camera_mode;           /* gort_camera_mode */
    int32_t shard_rank;            /* multi-process sharding: this process renders tiles with    */
    int32_t shard_count;           /*   tile_id % shard_count == shard_rank (0/1 = whole frame)  */
    int32_t collect_stats;         /* 1 = also count ray segments / node visits / primitive tests (slower kernel variant) */
    uint64_t seed;                 /* Philox key; the image is a pure function of (scene, params, seed) */
    /* Region render (the reference's RenderChunk, internal/distributed/distributed_renderer.go:29-39): only the pixel blocks that
     * touch [crop_x0, crop_x1) x [crop_y0, crop_y1) are traced; every other pixel of the frame is written black.  Pixels inside
     * the region are bit-identical to the full frame's.  crop_x1 <= crop_x0 (all zero): the whole frame. */
    int32_t crop_x0, crop_y0, crop_x1, crop_y1;
} gort_render_params;

typedef struct gort_stats {
    double kernel_ms;    /* device time of the frame (cull + trace + resolve), max over the ctx's devices.  Per-warp-queue path
                            (render_path 0/1): from %globaltimer stamps the kernels write — first cull thread to last resolve CTA —
                            so that asking for stats does not put event records between the kernels; wavefront pipeline
                            (render_path 2) and n_devices > 1: CUDA events on the stream */
    double trace_ms;     /* the trace kernel(s) alone (device 0) */
    double resolve_ms;   /* tone-map/quantise/pack kernel (+ gather/unswizzle when n_devices > 1) */
    double total_ms;     /* host wall clock of the whole call (launch + D2H where applicable) */
    double upload_ms;    /* last gort_scene_upload: host flatten + H2D */
    double bvh_build_ms; /* last gort_scene_upload: host BVH build */
    double device_ms[GORT_MAX_DEVICES];
    int32_t n_devices;
    int32_t n_tiles;     /* tiles rendered by this call */
    uint64_t primary_rays;  /* width*height*samples of the rendered tiles (the reference's "rays") */
    uint64_t bvh_nodes, bvh_bytes;
    /* filled when collect_stats = 1 (device-side counters of the same frame) */
    uint64_t closest_queries; /* hitWorld calls with closest-hit semantics */
    uint64_t shadow_queries;  /* hitWorld calls used as a boolean (hard + soft shadow rays) */
    uint64_t nodes_visited;   /* BVH inner nodes fetched (2 slab tests each) */
    uint64_t sphere_tests, sphere_hits;
    uint64_t tri_tests, tri_hits;
    uint64_t tri_rejects[4];  /* rejected at |a|<1e-6, u, v, t (triangle.go:42-65) */
    uint64_t shaded_hits;     /* calculateDirectLighting + Scatter evaluations */
    uint64_t rng_blocks;      /* Philox4x32-10 blocks generated */
    uint64_t light_evals;     /* (hit, light) pairs that cast the hard shadow ray */
    uint64_t soft_shadow_rays;
    uint64_t diffuse_evals;   /* (hit, light) pairs with shadowFactor > 0 */
    uint64_t specular_evals;  /* ... of which metallic > 0.5 (Blinn-Phong term) */
    uint64_t paths_depth_ge5, paths_depth_ge20, paths_depth_max; /* samples whose path reached that depth */
    uint64_t cone_tests;      /* cone-vs-box / cone-vs-primitive tests of the soft-shadow candidate pass */
    /* SIMT use of the BVH walk per call site {FILL, EXTEND, SHADE hard shadows, SHADE soft shadows without
     * candidates}: node visits summed over lanes / 32 x the longest lane of every warp-level call */
    uint64_t walk_lane_visits[4], walk_warp_visits[4];
    double algorithmic_flops; /* SURVEY §8d per-operation costs applied to the counters above */
    uint64_t soft_pairs_skipped; /* lit (hit, light) pairs whose shadow cone held no primitive: shadowFactor 16/16
                                  * (renderer.go:326-328) without casting the 16 rays */
    uint64_t pairs_backfacing;   /* (hit, light) pairs with hit.Normal . lightDir <= 0: cosTheta = 0 (renderer.go:259) zeroes the
                                  * diffuse and specular terms, so no shadow ray is cast for them */
    double cull_ms;              /* the beam-cull pass (device 0); trace_ms no longer includes it */
    uint64_t primary_generated;  /* collect_stats: primary rays actually generated (samples of the pixel blocks the cull pass
                                  * kept); primary_rays - primary_generated samples were resolved as misses without a ray */
    int32_t render_path;         /* 0 parameter-bank scan (tiny sphere scenes), 1 per-warp-queue BVH kernel, 2 global-queue
                                  * wavefront pipeline (large scenes) */
    int32_t kernel_launches;     /* kernels of libgort launched for this frame on device 0 */
} gort_stats;

typedef struct gort_ctx gort_ctx;

/* ---- lifecycle ------------------------------------------------------------------------- */
int gort_abi_version(void);
int gort_device_count(void); /* >= 0, or a negative gort_status */
/* device_ids == NULL -> devices 0..n_devices-1.  n_devices > 1 = single-process multi-GPU
 * (tiles interleaved over the devices, gathered to device_ids[0]). */
int gort_create(const int* device_ids, int n_devices, gort_ctx** out);
void gort_destroy(gort_ctx* ctx);
const char* gort_last_error(const gort_ctx* ctx); /* ctx may be NULL: error of a failed gort_create */
/* Launch on a caller-owned CUDA stream (cudaStream_t as void*) of device_ids[0]; NULL restores the
 * ctx's own (non-blocking) stream; name the legacy default stream with cudaStreamLegacy (0x1).
 * Lets a host framework time/order the work with its own events. */
int gort_set_stream(gort_ctx* ctx, void* cuda_stream);

/* ---- scene (replaces scene.GetHittables()/GetLights() feeding Render, renderer.go:72-74) -- */
int gort_scene_upload(gort_ctx* ctx, const gort_scene_desc* desc);
/* Host mirror of scene.LoadFromFile + GetHittables (scene.go:45-90,104-190) for hosts that are not
 * the Go program: parses the reference's scene JSON, applies its defaults, expands cubes, uploads.
 * options: bit0 = load "triangularPrism" objects (extension; reference skips them, scene.go:80-82)
 *          bit1 = honour the "fog" block (extension; reference ignores it)
 *          bit2 = honour a "sky" block (extension): {"enabled": true, "preset": "default"|"white"|"sunset"|"night", and any of
 *                 skyColorTop skyColorBottom sunDirection sunColor sunIntensity sunSize rayleighScattering mieScattering
 *                 atmosphericDepth fogDensity fogColor hazeIntensity timeOfDay} -> gort_scene_desc::sky_params */
int gort_scene_load_json(gort_ctx* ctx, const char* json_text, size_t json_len, uint32_t options);
int gort_scene_load_file(gort_ctx* ctx, const char* path, uint32_t options);
/* Introspection of the uploaded scene (flattened, reference scan order). */
int gort_scene_counts(const gort_ctx* ctx, int32_t* n_spheres, int32_t* n_triangles, int32_t* n_materials,
                      int32_t* n_lights, int32_t* n_hittables);
int gort_scene_get_triangle(const gort_ctx* ctx, int32_t order_index_among_triangles, double* v9, int32_t* material);
int gort_scene_get_material(const gort_ctx* ctx, int32_t index, int32_t* type, double* color3_rough_metal_spec_ior7);
/* Extension: the loaded scene's own "renderer" block (README.md:285-291; the reference's loader drops it, scene.go:12-16):
 * hints5 = {samples, maxDepth, antiAliasing, recursiveReflections, softShadows}, -1 where the key is absent (and for
 * scenes uploaded as gort_scene_desc).  The library never applies them; a host may copy them into gort_render_params. */
int gort_scene_render_hints(const gort_ctx* ctx, int32_t* hints5);

/* ---- render (replaces the body of ParallelRenderer.Render, renderer.go:67-126) ------------ */
/* rgba_out: caller-owned host buffer of width*height*4 bytes, row-major, row y=0 first ==
 * image.RGBA.Pix with Stride = 4*width (renderer.go:70,96).  With shard_count > 1 only the
 * shard's tiles are written; other bytes are left untouched. */
int gort_render(gort_ctx* ctx, const gort_render_params* params, uint8_t* rgba_out, size_t rgba_bytes,
                gort_stats* stats_out);
/* Page-locked host memory for frames.  gort_render into such a buffer needs no staging copy: the resolve kernel stores
 * the pixels straight into it over PCIe, the culled (black) regions already while the frame is still being traced.
 * (A Go host wraps it as img.Pix: unsafe.Slice((*byte)(p), n) — see INTEGRATION.md.)  Any other host pointer works
 * too and costs one extra memcpy of the frame. */
int gort_host_alloc(size_t bytes, void** out);
void gort_host_free(void* p);
/* Same frame, but the row-major RGBA8 image stays in device memory (device_ids[0]); d_rgba is a
 * device pointer with width*height*4 bytes.  Stream-ordered on the ctx stream; returns without waiting for the frame unless
 * stats_out != NULL — except for large scenes (render_path 2), whose bounce loop is driven from the host and has finished all
 * but the resolve pass when the call returns. */
int gort_render_device(gort_ctx* ctx, const gort_render_params* params, void* d_rgba, size_t rgba_bytes,
                       gort_stats* stats_out);
/* Multi-process sharding (one process per GPU): render only this shard's tiles into a TILE-MAJOR
 * slab in device memory: slab tile j = global tile (shard_rank + j*shard_count), each tile
 * GORT_TILE*GORT_TILE*4 bytes, pixel (lx,ly) at ((ly*GORT_TILE+lx)*4).  slab_bytes must be
 * gort_shard_slab_bytes().  The slabs of all ranks, concatenated rank-major (what an NCCL gather
 * produces), are turned into the row-major frame by gort_unswizzle_device. */
size_t gort_shard_slab_bytes(int32_t width, int32_t height, int32_t shard_count);
int gort_render_shard_device(gort_ctx* ctx, const gort_render_params* params, void* d_slab, size_t slab_bytes,
                             gort_stats* stats_out);
int gort_unswizzle_device(gort_ctx* ctx, const void* d_slabs, int32_t shard_count, int32_t width, int32_t height,
                          void* d_rgba, size_t rgba_bytes);
/* ---- frame link: one process per GPU, frame assembled over NVLink peer memory (no collective) ----
 * The reference's distributed stub ships pixel rectangles as JSON over HTTP (internal/distributed/
 * distributed_renderer.go:29-61); here the owner rank allocates the row-major RGBA8 frame, every other rank maps
 * it (CUDA IPC) and its resolve kernel stores its tiles straight into it; two counters in the owner's memory
 * order completion and reuse.  Every rank calls gort_render_linked once per frame, with the same params.
 *   owner:  gort_link_create -> send `handle` to the peers -> per frame gort_render_linked; when it returns,
 *           the work that waits for all ranks' tiles is enqueued on the ctx stream, and gort_link_frame() may be
 *           read by anything enqueued after it.  The next gort_render_linked releases the frame for reuse.
 *   peer:   gort_link_open(handle) -> per frame gort_render_linked. */
#define GORT_LINK_HANDLE_BYTES 64
typedef struct gort_link gort_link;
int gort_link_create(gort_ctx* ctx, int32_t width, int32_t height, int32_t n_ranks, uint8_t* handle_out /* [64] */, gort_link** out);
int gort_link_open(gort_ctx* ctx, const uint8_t* handle /* [64] */, int32_t width, int32_t height, int32_t n_ranks, int32_t rank,
                   gort_link** out);
void gort_link_close(gort_ctx* ctx, gort_link* link);
void* gort_link_frame(const gort_link* link); /* device pointer of the frame in this process (owner: its own memory) */
int gort_render_linked(gort_ctx* ctx, const gort_render_params* params, gort_link* link, gort_stats* stats_out);
/* owner: copy the assembled frame to host memory (stream-ordered after gort_render_linked, then synchronised) */
int gort_link_read(gort_ctx* ctx, gort_link* link, uint8_t* rgba_out, size_t rgba_bytes);
/* test hook: a peer link onto an owner link of the SAME process (CUDA IPC cannot open a handle in the process that
 * made it); lets a single-GPU box exercise the protocol with the ranks run one after the other, peers first.  An owner
 * with local peers never enqueues a device-side wait: gort_render_linked checks on the host that every peer's tiles of
 * the frame have arrived and returns GORT_ERR_INVALID if the owner was rendered first. */
int gort_link_open_local(gort_ctx* ctx, gort_link* owner, int32_t rank, gort_link** out);

/* Sample-averaged linear radiance (before tone-map) of the last render on this ctx, float64 RGB
 * [height][width][3] on the host; tiles not owned by the shard are left untouched.  Test hook. */
int gort_read_radiance(gort_ctx* ctx, double* radiance_out, size_t bytes);

/* ---- host-only scene model (no CUDA needed): what the loader + BVH builder produce ---------- */
typedef struct gort_host_scene gort_host_scene;
/* Parse the reference's scene JSON exactly like gort_scene_load_json but keep the result on the host.
 * errbuf (may be NULL) receives the message on failure. */
int gort_host_scene_parse(const char* json_text, size_t json_len, uint32_t options, gort_host_scene** out, char* errbuf,
                          size_t errbuf_len);
/* The host side of gort_scene_upload (cgo: scene.Flatten() -> gort_scene_desc): the same validation and copy,
 * kept on the host.  *scene_inout NULL: a new host scene is made; non-NULL: that one is rebuilt in place, its arrays reused
 * (what gort_scene_upload does with the context's spare scene).  GORT_ERR_INVALID + errbuf on a bad description (the scene
 * passed in is then unspecified but still valid to free or rebuild). */
int gort_host_scene_from_desc(const gort_scene_desc* desc, gort_host_scene** scene_inout, char* errbuf, size_t errbuf_len);
void gort_host_scene_free(gort_host_scene* scene);
/* counts5 = {spheres, triangles, materials, lights, hittables} */
int gort_host_scene_counts(const gort_host_scene* scene, int32_t* counts5);
int gort_host_scene_get_sphere(const gort_host_scene* scene, int32_t i, double* center3_radius4, int32_t* material, int32_t* order);
int gort_host_scene_get_triangle(const gort_host_scene* scene, int32_t i, double* v9, int32_t* material, int32_t* order);
int gort_host_scene_get_material(const gort_host_scene* scene, int32_t i, int32_t* type, double* color3_rough_metal_spec_ior7);
int gort_host_scene_get_light(const gort_host_scene* scene, int32_t i, double* pos3_color3_intensity7);
int gort_host_scene_get_camera(const gort_host_scene* scene, double* pos3_lookat3_up3_fov_aspect11);
/* Build the BVH on the host and check its invariants (every primitive in exactly one leaf, leaves
 * homogeneous and <= 4 primitives, every child box encloses its subtree, depth within the traversal
 * stack).  Returns GORT_OK or GORT_ERR_INVALID; fills info4 = {inner nodes, max depth, leaves, bytes}. */
int gort_host_scene_bvh_validate(const gort_host_scene* scene, int64_t* info4, char* errbuf, size_t errbuf_len);

/* ---- test / measurement hooks ------------------------------------------------------------ */
/* hitWorld on the GPU BVH for n rays (host arrays, float64 in, float64 out): out_t[n] (<0 = miss),
 * out_order[n] = reference scan-order index of the primitive hit.  Contract: identical closest hit
 * to the linear scan (renderer.go:333-346) up to fp32 rounding. */
int gort_trace_rays(gort_ctx* ctx, int32_t n, const double* origins3, const double* directions3, double t_min,
                    double t_max, int32_t any_hit, double* out_t, int32_t* out_order);
/* Dependent-FFMA-chain microbenchmark on device_ids[0]: measured FP32 issue peak in TFLOP/s
 * (FMA = 2 flops) — the roofline denominator MEASURED_PEAKS.json lacks. */
int gort_measure_fp32_peak(gort_ctx* ctx, double* tflops_out, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* GORT_H */
