// kernels.h — launch interface between the host context (capi.cu) and the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace gort {

constexpr int kTile = 32;                 // createRenderTasks tileSize (renderer.go:401)
constexpr int kTilePixels = kTile * kTile;
constexpr int kSmallMax = 12;             // sphere-only scenes up to this size use the linear scan from the parameter bank
constexpr int kSmallLights = 4;           // ... if they also have at most this many lights
constexpr int kAccumFracBits = 30;        // fixed-point radiance accumulators: value * 2^30 in int64
constexpr float kSampleClamp = 65536.0f;  // |per-sample radiance| clamp before fixed-point conversion

enum Stream : uint32_t { kStreamJitter = 0, kStreamScatter = 1, kStreamShadow = 2 };

// counters collected by the STATS kernel variant (indices into TraceParams::stats)
enum StatIndex {
    kStatClosest = 0,      // closest-hit hitWorld queries
    kStatShadow = 1,       // boolean hitWorld queries (hard + soft shadow rays)
    kStatNodes = 2,        // BVH inner nodes visited (two slab tests each)
    kStatSphereTests = 3,
    kStatSphereHits = 4,   // sphere tests that produced an accepted root
    kStatTriTests = 5,
    kStatTriRejA = 6,      // rejected at |a| < 1e-6   (triangle.go:42-44)
    kStatTriRejU = 7,      // rejected at the u test   (triangle.go:50-52)
    kStatTriRejV = 8,      // rejected at the v test   (triangle.go:57-59)
    kStatTriRejT = 9,      // rejected at the t range  (triangle.go:63-65)
    kStatTriHits = 10,
    kStatShaded = 11,      // calculateDirectLighting + Scatter evaluations
    kStatRngBlocks = 12,   // Philox4x32-10 blocks
    kStatLightEvals = 13,  // (hit, light) pairs that cast a hard shadow ray
    kStatSoftRays = 14,    // soft shadow rays
    kStatDiffuse = 15,     // (hit, light) pairs with shadowFactor > 0
    kStatSpec = 16,        // ... that also evaluate the Blinn-Phong term
    kStatDepth5 = 17,      // samples whose path reached depth >= 5 / >= 20 / max_depth
    kStatDepth20 = 18,
    kStatDepthMax = 19,
    kStatConeTests = 20,   // cone-vs-box / cone-vs-primitive tests of the soft-shadow candidate pass
    // SIMT use of the BVH walk per call site (FILL, EXTEND, SHADE-B, SHADE-C overflow): node visits summed over
    // lanes, and 32 x the longest lane of each warp-level call; their ratio is the lane utilisation
    kStatWalkLane0 = 21, kStatWalkWarp0 = 25,
    kStatSoftSkipped = 29,  // lit (hit, light) pairs whose shadow cone is empty: factor 16/16 without casting the 16 rays
    kStatBackfacing = 30,   // (hit, light) pairs with hit.Normal . lightDir <= 0: cosTheta = 0 zeroes both lighting terms, no shadow rays
    kStatPairSetups = 31,   // (hit, light) pairs whose light direction / distance set-up ran and that were not back-facing
    kStatPrimary = 32,      // primary rays actually generated (samples of the pixel blocks the cull pass kept)
    kStatCount = 33
};

struct DevCamera {
    float ox, oy, oz;     // ray origin
    float llx, lly, llz;  // lower-left corner minus origin
    float hx, hy, hz;     // horizontal span (u = 0..1)
    float vx, vy, vz;     // vertical span   (v = 0..1; v = (y + jitter)/H)
};

struct SceneView {
    const float4* nodes;
    const float4* spheres;
    const int2* sphere_meta;
    const float4* tris;
    const float4* mats;    // 4 float4 per material (see capi.cu pack_material)
    const float4* lights;  // 2 float4 per light: (pos.xyz, intensity) (color.rgb, 0)
    int n_nodes;
    int n_lights;
    int n_spheres, n_tris;
    // quantised copy of the nodes behind the fp32 ones (nodes + 4 n_nodes, 32 bytes each; bvh.h): grid origin and cell size
    float qox, qoy, qoz, qcx, qcy, qcz;
};

struct TraceParams {
    SceneView scene;
    DevCamera cam;
    int width, height, samples, max_depth;
    float inv_w, inv_h;
    int jitter, recursive, soft;
    int tiles_x, tiles_y;
    int shard_rank, shard_count, n_local_tiles;
    int crop_x0, crop_y0, crop_x1, crop_y1;  // region render: blocks outside are not traced (crop_x1 <= crop_x0: whole frame)
    uint32_t target_units;           // desired number of work units (dynamic balance), see trace_kernel
    const uint32_t* active_list;     // pixel blocks kept by the cull pass: (local tile << 5) | block;
                                     // [0, n_deep) from the front, n_norm more from the back of [0, 32*n_local_tiles)
    const unsigned int* active_count;  // {n_deep, n_norm}
    uint8_t* block_active;       // [32 * n_local_tiles] written by the cull pass for EVERY block: 1 = kept (its accumulators are
                                 // zeroed there), 0 = culled (accumulators stale: resolve writes black without reading them)
    unsigned long long* accum;   // [n_local_tiles][1024][3] int64 fixed point
    unsigned int* work_counter;  // zeroed before launch
    unsigned long long* stats;   // [kStatCount] or nullptr
    uint8_t* early_out;               // nullptr, or the row-major RGBA8 frame: one warp in eight CTAs of the trace kernel first writes black
                                      // into the pixel blocks the cull pass dropped (a page-locked host frame fills over PCIe under the trace)
    unsigned long long* stamps;       // nullptr, or %globaltimer stamps of the frame: [0] cull pass starts, [1] trace kernel starts (cull done),
                                      // [2] resolve starts (trace done), [3] resolve ends — the kernel times of gort_stats without event records
    unsigned long long* debug_times;  // nullptr, or [1 + 2*n_warps]: kernel start, then per warp (end of units, end of drain) in ns
    uint32_t rk[20];             // Philox4x32-10 round keys: rk[2r] = key0 + r*W0, rk[2r+1] = key1 + r*W1
    int fog_enabled;
    float fog_density, fog_r, fog_g, fog_b;
    // sky extension: a ray that leaves the scene returns AtmosphereConfig.GetSkyColor(direction) (atmosphere.go:100-135) instead
    // of black; sky[] = the 27 config values of gort_scene_desc::sky_params
    int sky_enabled;
    float sky[27];
    // wavefront pipeline: primitives per unit volume of the world box x the volume factor of a shadow cone (pi/3 tan^2(asin 0.1)).
    // A lit pair whose cone is expected to hold far more primitives than a candidate list takes skips its cone walk and sends
    // its 16 rays through the BVH at once (same answers either way; 0 = always walk the cone)
    float cone_skip;
    // Upper bound on |emitted + w * direct| x |throughput growth| of any further bounce per unit of
    // throughput (host, float64, deliberately loose): used by the exact dead-path test in trace_kernel
    // (a path is dropped once every add it could still make rounds to zero in the fixed-point accumulator).
    // 0 disables it.
    float dead_bound;
    int urgent_depth;  // > 0: a warp whose newest survivors reached this depth extends them before refilling (tail latency)
    int no_cone_cull;  // test switch (GORT_NO_CONE_CULL): no exact culls — back-facing pairs cast their shadow rays, soft-shadow
                       // rays test every primitive / walk the BVH themselves
    // tiny sphere-only scene, in the reference's scan order (small_n == 0: use the BVH)
    int small_n;
    int small_mat[kSmallMax];
    float4 small_sph[kSmallMax];
    uint16_t small_inside[kSmallMax];  // per sphere: the spheres a ray can hit while it is inside it (itself + overlapping ones)
    float4 small_mats[kSmallMax][4];      // the scene's materials (same packing as SceneView::mats)
    float4 small_lights[kSmallLights][2];  // the scene's lights (same packing as SceneView::lights)
};

struct ResolveParams {
    const uint8_t* block_active;  // see TraceParams
    const unsigned long long* accum;
    int n_local_tiles, shard_rank, shard_count;
    int tiles_x, width, height, samples;
    uint8_t* out;      // row-major frame (slab_mode 0) or tile-major slab (slab_mode 1)
    int slab_mode;
    int part;          // 0 every pixel; 1 only the culled blocks (black; needs nothing but the cull pass); 2 only the kept blocks
    unsigned long long* stamps;  // see TraceParams; [3] = the last CTA of this pass is done
    unsigned int* done_count;    // with stamps: device word, zero between launches
};

// frame-link work folded into the cull pass and per-frame counter upkeep: its first thread zeroes the counter bank of the
// NEXT frame (four words) and, on the owner of a frame link, publishes "frame store_value consumed" (system scope)
struct CullExtras {
    unsigned int* zero_bank = nullptr;
    unsigned int* store_flag = nullptr;
    unsigned int store_value = 0;
};

// All launchers enqueue on `stream` and return the launch error (no synchronisation).
// `dependent` (trace, resolve): programmatic dependent launch — the kernel directly before it on the stream lets this one's CTAs
// become resident while it drains (cull -> trace -> resolve of a frame), and this one waits for it with griddepcontrol.wait.
// Only valid when the previous operation on the stream IS that kernel (no event record, copy or memset in between).
cudaError_t launch_cull(const TraceParams& p, uint32_t* active_list, unsigned int* active_count, cudaStream_t stream, const CullExtras& x = CullExtras());
cudaError_t launch_trace(const TraceParams& p, bool stats, int sm_count, cudaStream_t stream, bool dependent = false);
// max_blocks > 0: a small grid-stride launch (co-resident with the persistent trace grid)
cudaError_t launch_resolve(const ResolveParams& p, cudaStream_t stream, int max_blocks = 0, bool dependent = false);
// Frame link (one process per GPU, the owner's frame mapped into every peer): system-scope flag handshake in
// the owner's memory.  signal: fence + atomicAdd(flag, 1); store: flag = value; wait: spin until flag >= target
// (gives up after 20 s and sets *timed_out instead of hanging the GPU when a rank has died).
cudaError_t launch_link_signal(unsigned int* flag, cudaStream_t stream, bool dependent = false);
cudaError_t launch_link_store(unsigned int* flag, unsigned int value, cudaStream_t stream);
cudaError_t launch_link_wait(const unsigned int* flag, unsigned int target, unsigned int* timed_out, cudaStream_t stream, bool dependent = false);
cudaError_t launch_unswizzle(const uint8_t* slabs, int shard_count, int width, int height, uint8_t* rgba, cudaStream_t stream);
cudaError_t launch_trace_rays(const SceneView& scene, int n, const float* origins, const float* dirs, float tmin, float tmax,
                              int any_hit, float* out_t, int* out_order, cudaStream_t stream);
cudaError_t launch_ffma_peak(float* sink, int iters, int blocks, int threads, cudaStream_t stream);
int trace_kernel_regs(bool stats);

}  // namespace gort
