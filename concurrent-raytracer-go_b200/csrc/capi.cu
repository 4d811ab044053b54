// capi.cu — the C ABI of libgort.so (include/gort.h): context, scene upload, frame render.
// Replaces the body of ParallelRenderer.Render (/root/reference internal/renderer/renderer.go:67-126).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gort.h"
#include "bvh.h"
#include "host_scene.h"
#include "kernels.h"
#include "lbvh.h"
#include "stream.h"

using namespace gort;

namespace {

thread_local std::string g_create_error;

struct DeviceState {
    int dev = -1;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    // scene (replicated on every device)
    float4* d_nodes = nullptr;
    float4* d_spheres = nullptr;
    int2* d_meta = nullptr;
    float4* d_tris = nullptr;
    float4* d_mats = nullptr;
    float4* d_lights = nullptr;
    size_t cap_nodes = 0, cap_spheres = 0, cap_meta = 0, cap_tris = 0, cap_mats = 0, cap_lights = 0;
    // small scenes (<= kBlobMax bytes in all): one pinned staging buffer, one device blob, ONE async copy per upload
    uint8_t* h_stage = nullptr;
    uint8_t* d_blob = nullptr;
    size_t stage_bytes = 0, blob_bytes = 0;
    cudaEvent_t ev_upload = nullptr;
    bool blob_in_use = false;  // the scene pointers currently point into d_blob
    // per-frame buffers
    unsigned long long* d_accum = nullptr;
    size_t accum_tiles = 0;
    // two banks of {work counter, deep active blocks, other active blocks, -}: a frame uses one and its cull pass clears the
    // other for the next frame; behind them the resolve kernel's count of finished CTAs (frame link)
    unsigned int* d_counter = nullptr;
    int bank = 0;                       // the bank the LAST enqueued frame used
    unsigned int* counters() const { return d_counter + 4 * bank; }
    uint32_t* d_active = nullptr;       // active pixel-block list
    size_t active_bytes = 0;
    unsigned long long* d_stats = nullptr;
    unsigned long long* d_debug = nullptr;  // GORT_DEBUG_TIMES
    uint8_t* d_out = nullptr;  // frame (row-major) or slab (tile-major)
    size_t out_bytes = 0;
    uint8_t* d_gather = nullptr;  // lead device only: rank-major slabs of all devices
    size_t gather_bytes = 0;
    uint8_t* h_pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // zero-copy host frames: the culled (black) blocks are written by a small kernel on aux_stream right after the
    // cull pass, under the trace kernel; only the kept blocks wait for the end
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_cull = nullptr, ev_aux = nullptr;
    uint64_t wide_serial = 0;     // scene_serial whose tree the 4-wide nodes behind d_nodes were collapsed from
    uint8_t* d_build = nullptr;   // device BVH build: primitive arrays in scene order + scratch
    size_t build_bytes = 0;
    cudaEvent_t ev_tc = nullptr;  // end of the cull pass (timing): trace_ms starts here
    // per-warp-queue path: the kernels stamp %globaltimer instead (TraceParams::stamps), so that a frame whose caller wants
    // gort_stats is the same chain of dependent launches as one whose caller does not
    unsigned long long* h_stamps = nullptr;  // page-locked host memory the kernels store into (no copy to read them back)
    bool stamped = false;         // the last enqueued frame was timed by stamps, not by ev_tc / ev[1]
    // global-queue wavefront pipeline (stream.cu): path / record / pair queues of one batch, counter ring, readback
    uint8_t* d_stream = nullptr;
    size_t stream_bytes = 0;
    float4* d_top = nullptr;           // first four levels of the 4-wide tree as one block (lbvh.h: wide_top_block)
    uint8_t* d_sort = nullptr;         // sorted mode: (key, entry) double buffers + radix-sort scratch
    size_t sort_bytes = 0;
    unsigned int* d_ctl = nullptr;     // 2 x kCtlWords + the frame's primary cursor (64 bit)
    unsigned int* h_count = nullptr;   // pinned: {active deep, active other, first 6 counters of the iteration}
    cudaEvent_t ev_count = nullptr;
    int last_path = 0, last_launches = 0;
};

// tuning knobs that are set for a whole process are read once (a getenv is ~0.3 us; a C1 frame is 170 us and asked a dozen).
// The switches the tests flip between frames (GORT_PATH, GORT_BVH, GORT_NO_CONE_CULL, GORT_NO_ZERO_COPY, ...) stay live.
#define GORT_ENV_ONCE(name) ([]() -> const char* { static const char* const v = getenv(name); return v; }())

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// GORT_HOST_TIMES=1: where the HOST time of the upload and render calls goes (accumulated per phase, printed by gort_destroy)
enum { kLapDesc, kLapBvh, kLapPack, kLapH2D, kLapPre, kLapEnqueue, kLapSync, kLapStats, kLapEnqSetup, kLapEnqLaunch, kLapCount };
const char* const kLapNames[kLapCount] = {"upload: desc -> host scene", "upload: BVH build", "upload: pack + stage", "upload: H2D enqueue",
                                          "render: validate + pointer attributes", "render: enqueue", "render: wait for the GPU", "render: stats",
                                          "  enqueue: parameters", "  enqueue: launches"};
double g_lap_us[kLapCount];
unsigned long long g_lap_n[kLapCount];
const bool g_laps = getenv("GORT_HOST_TIMES") != nullptr;
struct Lap {
    double t;
    Lap() : t(g_laps ? now_ms() : 0.0) {}
    void mark(int k) {
        if (!g_laps) return;
        const double n = now_ms();
        if (g_lap_n[k]++ >= 8) g_lap_us[k] += (n - t) * 1e3;  // the first calls pay for module load and allocations
        t = n;
    }
};

}  // namespace

struct gort_ctx {
    std::vector<DeviceState> devs;
    cudaStream_t user_stream = nullptr;
    bool use_user_stream = false;
    HostScene scene;
    HostScene scene_spare;       // the scene before the current one: gort_scene_upload builds into it and swaps
    FlatBvh bvh;
    bool has_scene = false;
    bool peer_direct = false;  // multi-device ctx: all devices can store into the lead device's memory
    std::string err;
    double upload_ms = 0, bvh_ms = 0;
    size_t bvh_bytes = 0;     // node + primitive arrays on the device
    uint64_t scene_serial = 0;   // bumped by every scene upload (the 4-wide collapse of the tree is derived per upload, on first use)
    bool bvh_on_device = false;  // built by lbvh.cu: ctx->bvh holds the counts and the grid only
    uint8_t* h_build = nullptr;  // page-locked staging of the device builder's input (primitives, materials, lights), kept across uploads
    size_t h_build_bytes = 0;
    // last render (for gort_read_radiance)
    int last_w = 0, last_h = 0, last_samples = 0, last_rank = 0, last_count = 1;
    // the call wants gort_stats: timing events are recorded between the kernels of the frame.  Without them the frame is
    // cull -> trace -> resolve back to back, chained by programmatic dependent launches.
    bool timing = false;
    std::vector<int> last_local_tiles;  // per device
};

struct gort_link {
    bool owner = false;
    int32_t width = 0, height = 0, n_ranks = 1, rank = 0;
    size_t frame_bytes = 0;
    uint8_t* base = nullptr;       // owner: cudaMalloc; peer: cudaIpcOpenMemHandle mapping of the owner's allocation
    unsigned int* ctrl = nullptr;  // {arrived, consumed, timed_out} behind the frame
    unsigned int frame_no = 0;
    bool local_alias = false;      // test hook: peer link sharing the owner's pointer inside one process
    bool has_local_peers = false;  // owner: some peers are local aliases (ranks run one after the other on this GPU)
};

namespace {

int fail(gort_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                                       \
    do {                                                                                                          \
        cudaError_t e__ = (expr);                                                                                 \
        if (e__ != cudaSuccess)                                                                                   \
            return fail(ctx, GORT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorName(e__) + " (" + cudaGetErrorString(e__) + ")"); \
    } while (0)

cudaStream_t stream_of(gort_ctx* ctx, int i) {
    if (i == 0 && ctx->use_user_stream) return ctx->user_stream;
    return ctx->devs[i].own_stream;
}

template <typename T>
int ensure(gort_ctx* ctx, T*& ptr, size_t& have, size_t want_bytes) {
    if (have >= want_bytes && ptr) return GORT_OK;
    if (ptr) CUDA_TRY(ctx, cudaFree(ptr));
    ptr = nullptr;
    have = 0;
    CUDA_TRY(ctx, cudaMalloc(&ptr, std::max<size_t>(want_bytes, 16)));
    have = want_bytes;
    return GORT_OK;
}

void free_scene(DeviceState& d) {
    cudaSetDevice(d.dev);
    if (d.blob_in_use) {
        d.d_nodes = d.d_spheres = d.d_tris = d.d_mats = d.d_lights = nullptr;
        d.d_meta = nullptr;
        d.blob_in_use = false;
    }
    cudaFree(d.d_blob);
    d.d_blob = nullptr; d.blob_bytes = 0;
    if (d.h_stage) cudaFreeHost(d.h_stage);
    d.h_stage = nullptr; d.stage_bytes = 0;
    if (d.ev_upload) cudaEventDestroy(d.ev_upload);
    d.ev_upload = nullptr;
    cudaFree(d.d_nodes); cudaFree(d.d_spheres); cudaFree(d.d_meta); cudaFree(d.d_tris); cudaFree(d.d_mats); cudaFree(d.d_lights);
    d.d_nodes = d.d_spheres = d.d_tris = d.d_mats = d.d_lights = nullptr;
    d.d_meta = nullptr;
    d.cap_nodes = d.cap_spheres = d.cap_meta = d.cap_tris = d.cap_mats = d.cap_lights = 0;
}

float as_float(int32_t i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}

// Per-material constants the kernel needs, with every threshold ladder of the reference resolved
// in float64 on the host: ambient/diffuse/specular ladders (renderer.go:236-243,262-273,282-287),
// reflection/direct weights (renderer.go:193-223), Fresnel parameters (material.go:88,103,117;
// material.go:180; advanced_materials.go:137,147).
void pack_material(const HostMaterial& m, F4 out[4]) {
    const bool metal_like = (m.type == GORT_MAT_METAL || m.type == GORT_MAT_SHINY || m.type == GORT_MAT_PERFECTMIRROR);
    const double metallic = metal_like ? (m.type == GORT_MAT_PERFECTMIRROR ? 1.0 : m.metallic) : 0.0;  // GetMetallic
    double color[3] = {m.color[0], m.color[1], m.color[2]};
    if (m.type == GORT_MAT_DIELECTRIC) color[0] = color[1] = color[2] = 1.0;  // GetAlbedo / attenuation material.go:236,266
    double ambient = 0.1;
    if (metallic > 0.9) ambient = 0.05;
    else if (metallic > 0.7) ambient = 0.07;
    else if (metallic > 0.5) ambient = 0.08;
    double kd = 0.25;
    if (metallic > 0.95) kd = 0.05;
    else if (metallic > 0.9) kd = 0.08;
    else if (metallic > 0.8) kd = 0.12;
    else if (metallic > 0.7) kd = 0.15;
    else if (metallic > 0.5) kd = 0.2;
    double wr = 1.0, wd = 1.0;
    if (metallic > 0.95) { wr = 0.85; wd = 0.15; }
    else if (metallic > 0.9) { wr = 0.8; wd = 0.2; }
    else if (metallic > 0.8) { wr = 0.75; wd = 0.25; }
    else if (metallic > 0.7) { wr = 0.7; wd = 0.3; }
    else if (metallic > 0.5) { wr = 0.6; wd = 0.4; }
    else if (metallic > 0.2) { wr = 0.4; wd = 0.6; }
    double spec_power = 0.0;  // 0 => no specular term (metallic <= 0.5)
    if (metallic > 0.5) {
        spec_power = 32.0;
        if (metallic > 0.9) spec_power = 64.0;
        else if (metallic > 0.8) spec_power = 48.0;
    }
    const double ior = m.ior;
    const double f0 = std::pow((ior - 1.0) / (ior + 1.0), 2.0);
    double fs = 0.0, mf = -1.0;
    if (m.type == GORT_MAT_METAL) {
        fs = 0.6 + m.metallic * 0.4;
        if (m.metallic > 0.8) mf = 0.4 + m.metallic * 0.5;
    } else if (m.type == GORT_MAT_SHINY) {
        fs = 0.4 + m.specular * 0.4;
    } else if (m.type == GORT_MAT_PERFECTMIRROR) {
        fs = 0.9;
    } else if (m.type == GORT_MAT_GLASS || m.type == GORT_MAT_DIELECTRIC) {
        fs = 1.0 / ior;  // refractionRatio on the front face (advanced_materials.go:25-29, material.go:239-243)
    }
    out[0] = F4{as_float(m.type), (float)color[0], (float)color[1], (float)color[2]};
    out[1] = F4{(float)m.roughness, (float)metallic, (float)m.specular, (float)ior};
    out[2] = F4{(float)ambient, (float)kd, (float)wr, (float)wd};
    out[3] = F4{(float)spec_power, (float)f0, (float)fs, (float)mf};
}

constexpr double kCoordLimit = 1e30;  // |coordinate| a scene may hold (fp32 device arithmetic squares distances)

// Large scenes: the BVH is built on the device (lbvh.cu).  The host packs the primitives in scene order (parallel), one H2D
// copy per array, and the builder writes the node / primitive arrays the kernels read.  Returns GORT_OK, an error, or
// 1 = "use the host builder" (tree deeper than the traversal stack: many coincident centroids).
int upload_scene_device_bvh(gort_ctx* ctx) {
    const double t0 = now_ms();
    const HostScene& hs = ctx->scene;
    const size_t nS = hs.spheres.size(), nT = hs.tris.size(), n = nS + nT;
    const bool times = GORT_ENV_ONCE("GORT_BVH_TIMES") != nullptr;
    auto lap = [&](const char* what) {
        if (times) fprintf(stderr, "[gort lbvh] %-28s %8.2f ms\n", what, now_ms() - t0);
    };
    // everything the device needs is packed straight into page-locked memory (kept across uploads): no staging copy, and the
    // H2D copies run at PCIe speed
    auto a256 = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t nM = hs.mats.size(), nL = hs.lights.size();
    const size_t off_sph = 0, off_meta = off_sph + a256(nS * 16), off_tri = off_meta + a256(nS * 8), in_bytes = off_tri + a256(nT * 64);
    const size_t off_mats = in_bytes, off_lights = off_mats + a256(nM * 64), stage_bytes = off_lights + a256(nL * 32);
    if (ctx->h_build_bytes < stage_bytes) {
        if (ctx->h_build) CUDA_TRY(ctx, cudaFreeHost(ctx->h_build));
        ctx->h_build = nullptr; ctx->h_build_bytes = 0;
        CUDA_TRY(ctx, cudaHostAlloc(&ctx->h_build, stage_bytes, cudaHostAllocPortable));
        ctx->h_build_bytes = stage_bytes;
    }
    F4* sph = reinterpret_cast<F4*>(ctx->h_build + off_sph);
    I2* meta = reinterpret_cast<I2*>(ctx->h_build + off_meta);
    F4* tri = reinterpret_cast<F4*>(ctx->h_build + off_tri);
    F4* mats = reinterpret_cast<F4*>(ctx->h_build + off_mats);
    F4* lights = reinterpret_cast<F4*>(ctx->h_build + off_lights);
    lap("staging buffer");
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<double> lo(hw * 3, 1e300), hi(hw * 3, -1e300);
    auto work = [&](unsigned t) {
        // (bounds in locals, stored once: the threads' slots of lo / hi share cache lines)
        double l[3] = {1e300, 1e300, 1e300}, h[3] = {-1e300, -1e300, -1e300};
        for (size_t i = nS * t / hw; i < nS * (t + 1) / hw; i++) {
            pack_sphere(hs.spheres[i], sph[i], meta[i]);
            const double r = std::fabs(hs.spheres[i].r);
            for (int a = 0; a < 3; a++) { l[a] = std::min(l[a], hs.spheres[i].c[a] - r); h[a] = std::max(h[a], hs.spheres[i].c[a] + r); }
        }
        for (size_t i = nT * t / hw; i < nT * (t + 1) / hw; i++) {
            pack_triangle(hs.tris[i], &tri[4 * i]);
            for (int k = 0; k < 3; k++)
                for (int a = 0; a < 3; a++) { l[a] = std::min(l[a], hs.tris[i].v[k][a]); h[a] = std::max(h[a], hs.tris[i].v[k][a]); }
        }
        for (size_t i = nM * t / hw; i < nM * (t + 1) / hw; i++) pack_material(hs.mats[i], &mats[4 * i]);
        for (int a = 0; a < 3; a++) { lo[t * 3 + a] = l[a]; hi[t * 3 + a] = h[a]; }
    };
    {
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < hw; t++) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
    }
    for (size_t i = 0; i < nL; i++) {
        const HostLight& l = hs.lights[i];
        lights[2 * i] = F4{(float)l.pos[0], (float)l.pos[1], (float)l.pos[2], (float)l.intensity};
        lights[2 * i + 1] = F4{(float)l.color[0], (float)l.color[1], (float)l.color[2], 0.f};
    }
    lap("primitives + materials packed");
    double wlo[3] = {1e300, 1e300, 1e300}, whi[3] = {-1e300, -1e300, -1e300};
    for (unsigned t = 0; t < hw; t++)
        for (int a = 0; a < 3; a++) { wlo[a] = std::min(wlo[a], lo[t * 3 + a]); whi[a] = std::max(whi[a], hi[t * 3 + a]); }
    LbvhIn in;
    memset(&in, 0, sizeof(in));
    in.n_spheres = (uint32_t)nS; in.n_tris = (uint32_t)nT;
    double extent = 0;
    for (int a = 0; a < 3; a++) extent = std::max(extent, std::max(std::fabs(wlo[a]), std::fabs(whi[a])));
    if (!(extent < kCoordLimit)) {
        ctx->has_scene = false;
        return fail(ctx, GORT_ERR_INVALID, "scene coordinates exceed the range of the fp32 device path (|x| must stay below 1e30)");
    }
    const double pad = 4e-7 * std::max(1.0, extent);  // as bvh.cpp
    in.pad = (float)pad;
    FlatBvh& b = ctx->bvh;
    b = FlatBvh();
    for (int a = 0; a < 3; a++) {
        in.world_lo[a] = (float)wlo[a]; in.world_hi[a] = (float)whi[a];
        const double l = wlo[a] - 4.0 * pad, h = whi[a] + 4.0 * pad;  // (the device pads in fp32: a little more room than the host grid)
        const double qc = std::max((h - l) / 65520.0, 1e-30);
        b.qcell[a] = (float)qc;
        b.qorigin[a] = (float)(l - 8.0 * qc);
        in.qcell[a] = b.qcell[a]; in.qorigin[a] = b.qorigin[a];
    }
    b.n_nodes = (int32_t)(n - 1);
    int max_depth = 0;
    for (size_t i = 0; i < ctx->devs.size(); i++) {
        DeviceState& d = ctx->devs[i];
        CUDA_TRY(ctx, cudaSetDevice(d.dev));
        cudaStream_t st = stream_of(ctx, (int)i);
        if (d.blob_in_use) {
            d.d_nodes = d.d_spheres = d.d_tris = d.d_mats = d.d_lights = nullptr;
            d.d_meta = nullptr;
            d.cap_nodes = d.cap_spheres = d.cap_meta = d.cap_tris = d.cap_mats = d.cap_lights = 0;
            d.blob_in_use = false;
        }
        const size_t scratch = lbvh_scratch_bytes((uint32_t)n);
        if (int rc = ensure(ctx, d.d_build, d.build_bytes, in_bytes + scratch)) return rc;
        float4* d_sph_in = (float4*)(d.d_build + off_sph);
        int2* d_meta_in = (int2*)(d.d_build + off_meta);
        float4* d_tri_in = (float4*)(d.d_build + off_tri);
        CUDA_TRY(ctx, cudaMemcpyAsync(d.d_build, ctx->h_build, in_bytes, cudaMemcpyHostToDevice, st));  // the three primitive arrays at once
        if (int rc = ensure(ctx, d.d_nodes, d.cap_nodes, (n - 1) * 10 * sizeof(F4))) return rc;
        if (int rc = ensure(ctx, d.d_spheres, d.cap_spheres, nS * sizeof(F4))) return rc;
        if (int rc = ensure(ctx, d.d_meta, d.cap_meta, nS * sizeof(I2))) return rc;
        if (int rc = ensure(ctx, d.d_tris, d.cap_tris, nT * 4 * sizeof(F4))) return rc;
        if (int rc = ensure(ctx, d.d_mats, d.cap_mats, nM * 4 * sizeof(F4))) return rc;
        if (int rc = ensure(ctx, d.d_lights, d.cap_lights, nL * 2 * sizeof(F4))) return rc;
        if (nM) CUDA_TRY(ctx, cudaMemcpyAsync(d.d_mats, mats, nM * 4 * sizeof(F4), cudaMemcpyHostToDevice, st));
        if (nL) CUDA_TRY(ctx, cudaMemcpyAsync(d.d_lights, lights, nL * 2 * sizeof(F4), cudaMemcpyHostToDevice, st));
        in.spheres = d_sph_in; in.sphere_meta = d_meta_in; in.tris = d_tri_in;
        LbvhOut out;
        out.nodes = d.d_nodes; out.spheres = d.d_spheres; out.sphere_meta = d.d_meta; out.tris = d.d_tris;
        int depth_i = 0;
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        lap("buffers + H2D");
        CUDA_TRY(ctx, lbvh_build(in, out, d.d_build + in_bytes, scratch, &depth_i, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        lap("device build");
        max_depth = std::max(max_depth, depth_i);
    }
    if (times) fprintf(stderr, "[gort lbvh] %zu primitives, %zu inner nodes, depth %d%s\n", n, n - 1, max_depth, max_depth > 62 ? " -> host builder" : "");
    if (max_depth > 62) return 1;
    b.max_depth = max_depth;
    ctx->bvh_on_device = true;
    ctx->scene_serial++;
    ctx->bvh_bytes = (n - 1) * 6 * sizeof(F4) + nS * (sizeof(F4) + sizeof(I2)) + nT * 4 * sizeof(F4);
    ctx->has_scene = true;
    ctx->bvh_ms = now_ms() - t0;  // pack + H2D + device build
    b.build_ms = ctx->bvh_ms;
    ctx->upload_ms = 0;
    return GORT_OK;
}

int upload_scene(gort_ctx* ctx) {
    const double t0 = now_ms();
    // limits of the device code: a leaf's first primitive is a 26-bit index per primitive type; the walks keep a 64-entry stack
    if (ctx->scene.spheres.size() > (size_t)kLeafStartMask + 1 || ctx->scene.tris.size() > (size_t)kLeafStartMask + 1) {
        ctx->has_scene = false;
        return fail(ctx, GORT_ERR_INVALID, "scene too large: at most 2^26 spheres and 2^26 triangles");
    }
    {
        // which builder: binned SAH on the host (better trees, ~0.45 s per million primitives) or LBVH on the device (a few ms)
        const size_t n_prims = ctx->scene.spheres.size() + ctx->scene.tris.size();
        const char* which = getenv("GORT_BVH");
        const char* mn = GORT_ENV_ONCE("GORT_BVH_DEVICE_MIN");
        const size_t device_min = mn ? (size_t)atoll(mn) : (size_t)200000;
        const bool on_device = n_prims >= 2 && (which ? !strcmp(which, "device") : n_prims >= device_min);
        ctx->bvh_on_device = false;
        if (on_device) {
            const int rc = upload_scene_device_bvh(ctx);
            if (rc != 1) return rc;
        }
    }
    Lap lap;
    build_bvh(ctx->scene, ctx->bvh);
    lap.mark(kLapBvh);
    ctx->bvh_ms = ctx->bvh.build_ms;
    ctx->bvh_bytes = ctx->bvh.nodes.size() * sizeof(F4) + ctx->bvh.spheres.size() * sizeof(F4) + ctx->bvh.sphere_meta.size() * sizeof(I2) +
                     ctx->bvh.tris.size() * sizeof(F4);
    if (ctx->bvh.max_depth > 62) {
        ctx->has_scene = false;
        return fail(ctx, GORT_ERR_INVALID, "BVH deeper than the traversal stack (62 levels): degenerate geometry");
    }
    for (int a = 0; a < 3; a++) {
        // the reference computes in float64; the device path is fp32: a scene whose box leaves that range is refused, not mis-rendered
        const double o = ctx->bvh.qorigin[a], c = ctx->bvh.qcell[a];
        if (!(std::isfinite(o) && std::isfinite(c) && std::fabs(o) < kCoordLimit && c * 65536.0 < kCoordLimit)) {
            ctx->has_scene = false;
            return fail(ctx, GORT_ERR_INVALID, "scene coordinates exceed the range of the fp32 device path (|x| must stay below 1e30)");
        }
    }
    const HostScene& hs = ctx->scene;
    std::vector<F4> mats(hs.mats.size() * 4);
    for (size_t i = 0; i < hs.mats.size(); i++) pack_material(hs.mats[i], &mats[4 * i]);
    std::vector<F4> lights(hs.lights.size() * 2);
    for (size_t i = 0; i < hs.lights.size(); i++) {
        const HostLight& l = hs.lights[i];
        lights[2 * i] = F4{(float)l.pos[0], (float)l.pos[1], (float)l.pos[2], (float)l.intensity};
        lights[2 * i + 1] = F4{(float)l.color[0], (float)l.color[1], (float)l.color[2], 0.f};
    }
    const FlatBvh& b = ctx->bvh;
    const void* src[6] = {b.nodes.data(), b.spheres.data(), b.sphere_meta.data(), b.tris.data(), mats.data(), lights.data()};
    const size_t nodes_cap = (size_t)b.n_nodes * 10 * sizeof(F4);  // 4 fp32 + 2 quantised float4 per node from the host, room for the 4-wide collapse
    const size_t len[6] = {b.nodes.size() * sizeof(F4), b.spheres.size() * sizeof(F4), b.sphere_meta.size() * sizeof(I2),
                           b.tris.size() * sizeof(F4), mats.size() * sizeof(F4), lights.size() * sizeof(F4)};
    size_t off[6], total = 0;
    for (int k = 0; k < 6; k++) {
        off[k] = total;
        total += (std::max<size_t>(k == 0 ? nodes_cap : len[k], 16) + 255) / 256 * 256;
    }
    constexpr size_t kBlobMax = 1u << 20;
    for (size_t i = 0; i < ctx->devs.size(); i++) {
        DeviceState& d = ctx->devs[i];
        CUDA_TRY(ctx, cudaSetDevice(d.dev));
        cudaStream_t st = stream_of(ctx, (int)i);
        if (total <= kBlobMax) {
            // A frame of the README scenes is sub-millisecond, so the upload must not cost several API round trips:
            // the arrays are packed into pinned staging memory and go down in one cudaMemcpyAsync, with no host
            // synchronisation (the staging buffer is reused only after the previous copy's event has completed).
            if (!d.ev_upload) CUDA_TRY(ctx, cudaEventCreateWithFlags(&d.ev_upload, cudaEventDisableTiming));
            if (d.stage_bytes < total) {
                if (d.h_stage) { CUDA_TRY(ctx, cudaEventSynchronize(d.ev_upload)); CUDA_TRY(ctx, cudaFreeHost(d.h_stage)); }
                d.h_stage = nullptr; d.stage_bytes = 0;
                CUDA_TRY(ctx, cudaMallocHost(&d.h_stage, kBlobMax));
                d.stage_bytes = kBlobMax;
            }
            if (int rc = ensure(ctx, d.d_blob, d.blob_bytes, kBlobMax)) return rc;
            if (!d.blob_in_use) {  // switching from separately allocated arrays
                cudaFree(d.d_nodes); cudaFree(d.d_spheres); cudaFree(d.d_meta); cudaFree(d.d_tris); cudaFree(d.d_mats); cudaFree(d.d_lights);
                d.cap_nodes = d.cap_spheres = d.cap_meta = d.cap_tris = d.cap_mats = d.cap_lights = 0;
                d.blob_in_use = true;
            }
            CUDA_TRY(ctx, cudaEventSynchronize(d.ev_upload));
            for (int k = 0; k < 6; k++)
                if (len[k]) memcpy(d.h_stage + off[k], src[k], len[k]);
            lap.mark(kLapPack);
            CUDA_TRY(ctx, cudaMemcpyAsync(d.d_blob, d.h_stage, total, cudaMemcpyHostToDevice, st));
            CUDA_TRY(ctx, cudaEventRecord(d.ev_upload, st));
            lap.mark(kLapH2D);
            d.d_nodes = reinterpret_cast<float4*>(d.d_blob + off[0]);
            d.d_spheres = reinterpret_cast<float4*>(d.d_blob + off[1]);
            d.d_meta = reinterpret_cast<int2*>(d.d_blob + off[2]);
            d.d_tris = reinterpret_cast<float4*>(d.d_blob + off[3]);
            d.d_mats = reinterpret_cast<float4*>(d.d_blob + off[4]);
            d.d_lights = reinterpret_cast<float4*>(d.d_blob + off[5]);
            continue;
        }
        if (d.blob_in_use) {  // back to separately allocated arrays: the pointers into the blob are not ours to free
            d.d_nodes = d.d_spheres = d.d_tris = d.d_mats = d.d_lights = nullptr;
            d.d_meta = nullptr;
            d.cap_nodes = d.cap_spheres = d.cap_meta = d.cap_tris = d.cap_mats = d.cap_lights = 0;
            d.blob_in_use = false;
        }
        // device buffers are kept across uploads and only grow (a re-upload per frame costs no cudaMalloc)
        auto up = [&](auto*& dst, size_t& cap, const void* src, size_t bytes) -> int {
            if (int rc = ensure(ctx, dst, cap, bytes)) return rc;
            if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
            return GORT_OK;
        };
        if (int rc = ensure(ctx, d.d_nodes, d.cap_nodes, nodes_cap)) return rc;
        if (int rc = up(d.d_nodes, d.cap_nodes, b.nodes.data(), b.nodes.size() * sizeof(F4))) return rc;
        if (int rc = up(d.d_spheres, d.cap_spheres, b.spheres.data(), b.spheres.size() * sizeof(F4))) return rc;
        if (int rc = up(d.d_meta, d.cap_meta, b.sphere_meta.data(), b.sphere_meta.size() * sizeof(I2))) return rc;
        if (int rc = up(d.d_tris, d.cap_tris, b.tris.data(), b.tris.size() * sizeof(F4))) return rc;
        if (int rc = up(d.d_mats, d.cap_mats, mats.data(), mats.size() * sizeof(F4))) return rc;
        if (int rc = up(d.d_lights, d.cap_lights, lights.data(), lights.size() * sizeof(F4))) return rc;
        CUDA_TRY(ctx, cudaStreamSynchronize(st));  // host vectors go out of scope: the copies must be done
    }
    ctx->has_scene = true;
    ctx->scene_serial++;
    ctx->upload_ms = now_ms() - t0 - ctx->bvh_ms;
    return GORT_OK;
}

DevCamera make_camera(const HostScene& s, int mode) {
    DevCamera c;
    c.ox = (float)s.cam_pos[0]; c.oy = (float)s.cam_pos[1]; c.oz = (float)s.cam_pos[2];
    if (mode == GORT_CAMERA_LOOKAT) {
        // extension: classic look-at pinhole (vertical fov in degrees), upright image
        const double kPi = 3.14159265358979323846;
        const double hh = std::tan(s.cam_fov * kPi / 180.0 / 2.0);
        const double vh = 2.0 * hh, vw = vh * s.cam_aspect;
        auto norm = [](double* v) {
            double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            if (l == 0) { v[0] = v[1] = v[2] = 0; return; }
            v[0] /= l; v[1] /= l; v[2] /= l;
        };
        double w[3] = {s.cam_pos[0] - s.cam_look_at[0], s.cam_pos[1] - s.cam_look_at[1], s.cam_pos[2] - s.cam_look_at[2]};
        norm(w);
        double u[3] = {s.cam_up[1] * w[2] - s.cam_up[2] * w[1], s.cam_up[2] * w[0] - s.cam_up[0] * w[2], s.cam_up[0] * w[1] - s.cam_up[1] * w[0]};
        norm(u);
        double v[3] = {w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0]};
        double H[3], V[3], ll[3];
        for (int a = 0; a < 3; a++) {
            H[a] = u[a] * vw;
            V[a] = v[a] * vh;
            ll[a] = -H[a] / 2 - V[a] / 2 - w[a];
        }
        // dir = ll + u*H + (1-v)*V  (row 0 = top of the view)
        c.llx = (float)(ll[0] + V[0]); c.lly = (float)(ll[1] + V[1]); c.llz = (float)(ll[2] + V[2]);
        c.hx = (float)H[0]; c.hy = (float)H[1]; c.hz = (float)H[2];
        c.vx = (float)-V[0]; c.vy = (float)-V[1]; c.vz = (float)-V[2];
        return c;
    }
    // getRay (renderer.go:377-390): viewport height 2, width 2*aspect, focal length 1, looking down -Z
    const double vw = 2.0 * s.cam_aspect;
    c.llx = (float)(-vw / 2.0); c.lly = -1.0f; c.llz = -1.0f;
    c.hx = (float)vw; c.hy = 0.f; c.hz = 0.f;
    c.vx = 0.f; c.vy = 2.0f; c.vz = 0.f;
    return c;
}

// Loose float64 upper bound on |traceRay(...)| per unit of throughput, for the exact dead-path test:
//   per bounce   emitted <= E, direct <= ambient + sum_l |I_l| max(1,|color_l|) / tMin^2 * (kd*albedo + 3)
//                (lights closer than 0.001 are skipped, renderer.go:252; cos, shadow factor, metallic <= 1)
//   attenuation  |w_r * atten| <= F = max(1, max albedo, max |Schlick|), |Schlick| <= max(1, f0 + (1-f0)(|D|max-1)^5)
//                with |D|max = the longest (unnormalised) primary direction — reflection keeps |D|
//   radiance     R <= max_depth * (E + D) * F^max_depth
float dead_path_bound(const HostScene& s, const DevCamera& c, int max_depth) {
    double emax = 0, amax = 1.0;
    for (const HostMaterial& m : s.mats) {
        const double cm = std::max(std::fabs(m.color[0]), std::max(std::fabs(m.color[1]), std::fabs(m.color[2])));
        if (m.type == GORT_MAT_DIFFUSELIGHT) emax = std::max(emax, cm);
        amax = std::max(amax, cm);
    }
    double dsum = 0.1;
    for (const HostLight& l : s.lights) {
        const double cl = std::max(1.0, std::max(std::fabs(l.color[0]), std::max(std::fabs(l.color[1]), std::fabs(l.color[2]))));
        dsum += std::fabs(l.intensity) * cl * 1e6 * (0.25 * amax + 3.0);
    }
    double dmax = 1.0;
    for (int i = 0; i < 4; i++) {
        const double u = (i & 1) ? 1.0 : 0.0, v = (i & 2) ? 1.0 : 0.0;
        const double x = c.llx + u * c.hx + v * c.vx, y = c.lly + u * c.hy + v * c.vy, z = c.llz + u * c.hz + v * c.vz;
        dmax = std::max(dmax, std::sqrt(x * x + y * y + z * z) * 1.001);
    }
    const double fres = std::max(1.0, 1.0 + std::pow(dmax - 1.0, 5.0));
    const double F = std::max(amax, fres) * 1.0001;
    const double bound = (double)std::max(1, max_depth) * (emax + dsum) * std::pow(F, (double)std::max(1, max_depth));
    if (!(bound < 1e30)) return 0.f;  // no useful bound: never cut
    return (float)bound;
}

int local_tile_count(int n_tiles, int rank, int count) { return rank < n_tiles ? (n_tiles - rank + count - 1) / count : 0; }

int validate(gort_ctx* ctx, const gort_render_params* p) {
    if (!ctx) return GORT_ERR_INVALID;
    if (!p) return fail(ctx, GORT_ERR_INVALID, "params is NULL");
    if (p->abi_version != GORT_ABI_VERSION) return fail(ctx, GORT_ERR_INVALID, "gort_render_params.abi_version mismatch");
    if (!ctx->has_scene) return fail(ctx, GORT_ERR_NO_SCENE, "no scene uploaded");
    if (p->width <= 0 || p->height <= 0) return fail(ctx, GORT_ERR_INVALID, "width/height must be positive");
    if ((int64_t)p->width * p->height > (int64_t)1 << 28) return fail(ctx, GORT_ERR_INVALID, "image too large");
    if (p->width > 65535 || p->height > 65535) return fail(ctx, GORT_ERR_INVALID, "width/height must be <= 65535");
    if (p->samples <= 0 || p->samples > 65535) return fail(ctx, GORT_ERR_INVALID, "samples must be in 1..65535");
    // the per-warp-queue kernel numbers its work units (8x4 pixel block, sample batch) in 32 bits
    if ((int64_t)p->width * p->height * (int64_t)p->samples > ((int64_t)1 << 36))
        return fail(ctx, GORT_ERR_INVALID, "width*height*samples must be <= 2^36 per call");
    if (p->max_depth < 0 || p->max_depth > 65535) return fail(ctx, GORT_ERR_INVALID, "max_depth must be in 0..65535");
    const int sc = p->shard_count <= 0 ? 1 : p->shard_count;
    if (p->shard_rank < 0 || p->shard_rank >= sc) return fail(ctx, GORT_ERR_INVALID, "shard_rank out of range");
    if (p->camera_mode != GORT_CAMERA_REFERENCE && p->camera_mode != GORT_CAMERA_LOOKAT) return fail(ctx, GORT_ERR_INVALID, "camera_mode");
    if (p->crop_x1 > p->crop_x0 && (p->crop_x0 < 0 || p->crop_y0 < 0 || p->crop_x1 > p->width || p->crop_y1 > p->height || p->crop_y1 <= p->crop_y0))
        return fail(ctx, GORT_ERR_INVALID, "crop rectangle outside the frame");
    return GORT_OK;
}


// Which kernel family renders this scene (gort_stats::render_path): tiny sphere scenes scan the parameter bank, small BVH
// scenes run the per-warp-queue kernel, large ones the global-queue wavefront pipeline (stream.h says why).
// GORT_PATH=queue|stream overrides the size rule (tests render the same scene through both).
enum RenderPath { kPathSmall = 0, kPathQueue = 1, kPathStream = 2 };

int choose_path(const gort_ctx* ctx, const gort_render_params* p) {
    const HostScene& hs = ctx->scene;
    const size_t n_prims = hs.spheres.size() + hs.tris.size();
    const bool small_ok = hs.tris.empty() && !hs.spheres.empty() && (int)hs.spheres.size() <= kSmallMax && (int)hs.mats.size() <= kSmallMax &&
                          (int)hs.lights.size() <= kSmallLights && !hs.sky_enabled;  // (the sky extension lives in the BVH kernels)
    const bool stream_ok = ctx->bvh.n_nodes > 0 && (int)hs.lights.size() <= kStreamLightChunk * kStreamMaxChunks;
    const char* force = getenv("GORT_PATH");
    if (force && !strcmp(force, "stream") && stream_ok) return kPathStream;
    if (force && !strcmp(force, "queue")) return small_ok ? kPathSmall : kPathQueue;
    if (small_ok) return kPathSmall;
    // The pipeline pays ~10 launches per bounce whatever the frame holds: it wins once a frame has enough rays to fill them.
    // Measured crossover (profiles/README.md): C4 (depth 16) ~8 M primary samples per frame, C5 (depth 32) ~12 M.
    const char* mp = GORT_ENV_ONCE("GORT_STREAM_MIN_PRIMS");
    const char* ms = GORT_ENV_ONCE("GORT_STREAM_MIN_SAMPLES");
    const size_t min_prims = mp ? (size_t)atoll(mp) : 4096;
    const int64_t min_samples = ms ? (int64_t)atoll(ms) : (int64_t)12 << 20;
    const int64_t frame_samples = (int64_t)p->width * p->height * p->samples;
    return (stream_ok && n_prims >= min_prims && frame_samples >= min_samples) ? kPathStream : kPathQueue;
}

// The wavefront pipeline for one device's share of the frame: batches of samples, one bounce at a time (stream.h).
// Called after the cull pass has been enqueued on `st`.  Synchronises with the device once per bounce to learn whether
// any path is still alive (the stages of the bounce are already enqueued by then, so the GPU does not idle).
int run_stream(gort_ctx* ctx, DeviceState& d, const TraceParams& tp, const gort_render_params* p, cudaStream_t st) {
    const bool stats = p->collect_stats != 0;
    const int geom = (tp.scene.n_spheres > 0 ? 1 : 0) | (tp.scene.n_tris > 0 ? 2 : 0);
    CUDA_TRY(ctx, cudaMemcpyAsync(d.h_count, d.counters() + 1, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const uint32_t n_deep = d.h_count[0], n_active = d.h_count[0] + d.h_count[1];
    if (n_active == 0) return GORT_OK;
    if (stream_wants_wide_nodes() && d.wide_serial != ctx->scene_serial) {
        // the traversal kernel walks the 4-wide collapse of the tree: derived on the device, once per uploaded scene
        const size_t need = bvh_collapse_scratch_bytes(ctx->bvh.n_nodes);
        if (int rc = ensure(ctx, d.d_build, d.build_bytes, need)) return rc;
        CUDA_TRY(ctx, bvh_collapse_wide(d.d_nodes, ctx->bvh.n_nodes, d.d_nodes + 6 * (size_t)ctx->bvh.n_nodes, ctx->bvh.qorigin, ctx->bvh.qcell, d.d_build,
                                        d.build_bytes, st));
        // ... and the top of it as one block, for the traversal kernel's shared-memory staging (GORT_TOP)
        size_t top_have = d.d_top ? kTopNodes * 64 : 0;
        if (int rc = ensure(ctx, d.d_top, top_have, (size_t)kTopNodes * 64)) return rc;
        CUDA_TRY(ctx, wide_top_block(d.d_nodes + 6 * (size_t)ctx->bvh.n_nodes, d.d_top, st));
        d.wide_serial = ctx->scene_serial;
    }
    const uint64_t per_sample = (uint64_t)n_active * 32u;
    const uint64_t prim_total = per_sample * (uint64_t)p->samples;
    // path slots: every launch works on (up to) this many paths; the queue is topped up with new primary rays each iteration
    const char* be = GORT_ENV_ONCE("GORT_STREAM_BATCH");
    const uint64_t want = be ? std::max<uint64_t>(1024, (uint64_t)atoll(be)) : (uint64_t)16 << 20;
    const uint64_t cap = std::min<uint64_t>(std::min(want, prim_total), (uint64_t)1 << 26);

    // carve the queues out of one allocation (grows only)
    const size_t slot_bytes = stream_bytes_per_slot();
    if (int rc = ensure(ctx, d.d_stream, d.stream_bytes, (size_t)cap * slot_bytes + 8192)) return rc;
    StreamView v;
    memset(&v, 0, sizeof(v));
    {
        uint8_t* q = d.d_stream;
        auto take = [&](size_t bytes_per_slot) { uint8_t* r = q; q += ((size_t)cap * bytes_per_slot + 255) / 256 * 256; return r; };
        for (int b = 0; b < 2; b++) {
            v.qa[b] = (float4*)take(16); v.qb[b] = (float4*)take(16); v.qc[b] = (float4*)take(16); v.qd[b] = (uint2*)take(8);
        }
        v.ra = (float4*)take(16); v.rb = (float4*)take(16); v.rc = (float4*)take(16); v.racc = (float4*)take(16); v.rd = (uint2*)take(8);
        v.lit = (uint8_t*)take(kStreamLightChunk);
        v.cnt = (unsigned int*)take(4 * kStreamLightChunk);
        v.hard_list = (uint32_t*)take(4 * kStreamLightChunk);
        v.walk_list = (uint32_t*)take(4 * kStreamLightChunk);
        v.lit_list = (uint32_t*)take(4 * kStreamLightChunk);
        v.cand_recs = (uint4*)take(32 * kStreamLightChunk);
        if ((size_t)(q - d.d_stream) > d.stream_bytes) return fail(ctx, GORT_ERR_INVALID, "stream buffer carve-out overflow");
    }
    // Sorted mode (GORT_SORT=0..3): the entries of an iteration's queue are scattered in Morton order of their hit points.
    // Queues below kSortMin entries are not worth the sort's launches.
    const char* so = getenv("GORT_SORT");
    const int sort_mode = so ? atoi(so) : kStreamSortDefault;  // 1: Morton order; 2: coarser cells; 3: live entries first only
    const bool sorted = sort_mode >= 1 && sort_mode <= 3;
    const int sort_begin_bit = sort_mode == 2 ? 8 : (sort_mode == 3 ? 23 : 0);
    constexpr uint32_t kSortMin = 1u << 15;
    if (sorted)
        if (int rc = ensure(ctx, d.d_sort, d.sort_bytes, stream_sort_bytes((uint32_t)cap))) return rc;
    // GORT_TOP=0|1: the traversal kernel stages the top of the wide tree in shared memory
    const char* te = getenv("GORT_TOP");
    v.top = (stream_wants_wide_nodes() && (te ? atoi(te) != 0 : kStreamTopDefault)) ? d.d_top : nullptr;
    v.top_plain = te && atoi(te) == 2;
    v.cap = (uint32_t)cap;
    v.n_active = n_active; v.n_deep = n_deep;
    v.prim_total = prim_total;
    v.prim_cursor = reinterpret_cast<unsigned long long*>(d.d_ctl + 2 * kCtlWords);
    const int n_lights = tp.scene.n_lights;
    const int n_chunks = std::max(1, (n_lights + kStreamLightChunk - 1) / kStreamLightChunk);  // no lights: one pass adds the ambient term

    // iteration i: scatter the current queue (paths of any depth) -> shade records + survivors in the next queue; top the
    // next queue up with new primary rays; trace both; shade the records.  Ring of two counter blocks.
    CUDA_TRY(ctx, cudaMemsetAsync(d.d_ctl, 0, (2 * kCtlWords + 2) * sizeof(unsigned int), st));
    uint64_t generated = 0;
    uint32_t n_cur = 0;  // entries of the current queue (= the previous iteration's kCtlNextTotal, read back below)
    for (int it = 0;; it++) {
        v.cur = it & 1;
        v.ctl = d.d_ctl + (size_t)(it & 1) * kCtlWords;
        v.ctl_prev = d.d_ctl + (size_t)((it & 1) ^ 1) * kCtlWords;
        v.chunk = 0;
        v.order = nullptr;
        if (it > 0) {
            CUDA_TRY(ctx, cudaMemsetAsync(v.ctl, 0, kCtlWords * sizeof(unsigned int), st));
            if (sorted && n_cur >= kSortMin) {
                CUDA_TRY(ctx, stream_launch_sort(tp, v, n_cur, sort_begin_bit, d.d_sort, d.sort_bytes, st));
                d.last_launches++;  // the key pass (the radix sort's passes are CUB's)
            }
            CUDA_TRY(ctx, stream_launch_scatter(tp, v, geom, stats, d.sm_count, st));
            d.last_launches++;
        }
        CUDA_TRY(ctx, stream_launch_plan(v, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(d.h_count + 2, v.ctl, 6 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaEventRecord(d.ev_count, st));
        if (it > 0) CUDA_TRY(ctx, stream_launch_trace_ext(tp, v, geom, stats, d.sm_count, st));
        if (generated < prim_total) CUDA_TRY(ctx, stream_launch_primary(tp, v, geom, stats, d.sm_count, st));
        d.last_launches += 1 + (it > 0 ? 1 : 0) + (generated < prim_total ? 1 : 0);
        if (it > 0) {
            for (int c = 0; c < n_chunks; c++) {
                v.chunk = c; v.l0 = c * kStreamLightChunk; v.lc = std::max(0, std::min(kStreamLightChunk, n_lights - v.l0));
                v.last_chunk = c == n_chunks - 1;
                CUDA_TRY(ctx, stream_launch_shade_chunk(tp, v, geom, stats, d.sm_count, st));
                d.last_launches += tp.soft ? 6 : 3;
            }
        }
        CUDA_TRY(ctx, cudaEventSynchronize(d.ev_count));
        generated += d.h_count[2 + kCtlNew];
        n_cur = d.h_count[2 + kCtlNextTotal];
        if (d.h_count[2 + kCtlNextTotal] == 0 && generated >= prim_total) break;  // nothing left to scatter, nothing left to generate
    }
    return GORT_OK;
}

// Enqueue one device's share of the frame: zero accumulators, trace, resolve into d.d_out.
// eff_rank/eff_count: the tiles this device owns.  slab_mode: tile-major slab vs row-major frame.
// out_override: write the resolved pixels there instead of d.d_out (device pointer on this device).
// Optional stream-ordered hooks around the resolve of one device's share (frame link)
struct ResolveHooks {
    bool early_black = false;                 // resolve the culled blocks on the aux stream while the trace kernel runs
    const unsigned int* wait_flag = nullptr;  // before resolve: wait until *wait_flag >= wait_target
    unsigned int wait_target = 0;
    unsigned int* timed_out = nullptr;        // set by a wait that gave up
    unsigned int* signal_flag = nullptr;      // after resolve: fence.sys + atomicAdd(*signal_flag, 1)
    unsigned int* store_flag = nullptr;       // first thing in the frame: *store_flag = store_value (system scope)
    unsigned int store_value = 0;
};

int enqueue_device(gort_ctx* ctx, int di, const gort_render_params* p, int eff_rank, int eff_count, int slab_mode, uint8_t* out_override,
                   size_t slab_bytes, const ResolveHooks* hooks = nullptr) {
    Lap lap;
    DeviceState& d = ctx->devs[di];
    CUDA_TRY(ctx, cudaSetDevice(d.dev));
    cudaStream_t st = stream_of(ctx, di);
    const int tiles_x = (p->width + kTile - 1) / kTile, tiles_y = (p->height + kTile - 1) / kTile;
    const int n_tiles = tiles_x * tiles_y;
    const int n_local = local_tile_count(n_tiles, eff_rank, eff_count);
    ctx->last_local_tiles[di] = n_local;

    size_t accum_have = d.accum_tiles * kTilePixels * 3 * sizeof(unsigned long long);
    const size_t accum_want = (size_t)n_local * kTilePixels * 3 * sizeof(unsigned long long);
    if (int rc = ensure(ctx, d.d_accum, accum_have, accum_want)) return rc;
    d.accum_tiles = accum_have / (kTilePixels * 3 * sizeof(unsigned long long));

    uint8_t* out = out_override;
    if (!out) {
        const size_t want = slab_mode ? slab_bytes : (size_t)p->width * p->height * 4;
        if (int rc = ensure(ctx, d.d_out, d.out_bytes, want)) return rc;
        out = d.d_out;
    }

    const int path = choose_path(ctx, p);
    // gort_stats wanted: the wavefront pipeline (host round trips between its launches anyway) is timed by events between
    // the kernels, the per-warp-queue frame by stamps the kernels write
    const bool events = ctx->timing && (path == kPathStream || n_local == 0);
    const bool timing = events;
    d.stamped = ctx->timing && !events;
    // the owner's release of a frame link rides in the cull pass (GORT_LINK_UNFUSED=1: a one-thread kernel in front of it)
    static const bool link_unfused = getenv("GORT_LINK_UNFUSED") != nullptr;
    static const bool no_pdl = getenv("GORT_NO_PDL") != nullptr;
    if (events || (ctx->timing && ctx->devs.size() > 1)) CUDA_TRY(ctx, cudaEventRecord(d.ev[0], st));
    // the accumulators are cleared block by block in the cull pass (kept blocks only), not wholesale; the counters come
    // cleared from the previous frame's cull pass (two banks)
    if (hooks && hooks->store_flag && (link_unfused || n_local == 0)) CUDA_TRY(ctx, launch_link_store(hooks->store_flag, hooks->store_value, st));
    CullExtras cx;
    if (n_local > 0) {
        cx.zero_bank = d.counters();  // the last frame's
        d.bank ^= 1;
        if (hooks && hooks->store_flag && !link_unfused) { cx.store_flag = hooks->store_flag; cx.store_value = hooks->store_value; }
    }
    unsigned int* const counters = d.counters();
    // active-block list (uint32 per block) followed by the kept/culled byte of every block
    if (int rc = ensure(ctx, d.d_active, d.active_bytes, (size_t)n_local * 32 * (sizeof(uint32_t) + 1))) return rc;
    if (p->collect_stats) CUDA_TRY(ctx, cudaMemsetAsync(d.d_stats, 0, kStatCount * sizeof(unsigned long long), st));
    if (slab_mode && slab_bytes > (size_t)n_local * kTilePixels * 4)
        CUDA_TRY(ctx, cudaMemsetAsync(out + (size_t)n_local * kTilePixels * 4, 0, slab_bytes - (size_t)n_local * kTilePixels * 4, st));

    TraceParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.scene.nodes = d.d_nodes; tp.scene.spheres = d.d_spheres; tp.scene.sphere_meta = d.d_meta; tp.scene.tris = d.d_tris;
    tp.scene.mats = d.d_mats; tp.scene.lights = d.d_lights;
    tp.scene.n_nodes = ctx->bvh.n_nodes; tp.scene.n_lights = (int)ctx->scene.lights.size();
    tp.scene.n_spheres = (int)ctx->scene.spheres.size(); tp.scene.n_tris = (int)ctx->scene.tris.size();
    tp.scene.qox = ctx->bvh.qorigin[0]; tp.scene.qoy = ctx->bvh.qorigin[1]; tp.scene.qoz = ctx->bvh.qorigin[2];
    tp.scene.qcx = ctx->bvh.qcell[0]; tp.scene.qcy = ctx->bvh.qcell[1]; tp.scene.qcz = ctx->bvh.qcell[2];
    tp.cam = make_camera(ctx->scene, p->camera_mode);
    tp.width = p->width; tp.height = p->height; tp.samples = p->samples; tp.max_depth = p->max_depth;
    tp.inv_w = 1.0f / (float)p->width; tp.inv_h = 1.0f / (float)p->height;
    // tiny sphere-only scenes: the spheres ride in the kernel parameters, in scan order
    tp.small_n = 0;
    d.last_path = path;
    d.last_launches = 0;
    if (path == kPathSmall) {
        std::vector<const HostSphere*> sorted;
        for (const HostSphere& hs : ctx->scene.spheres) sorted.push_back(&hs);
        std::sort(sorted.begin(), sorted.end(), [](const HostSphere* a, const HostSphere* b) { return a->order < b->order; });
        tp.small_n = (int)sorted.size();
        for (int i = 0; i < tp.small_n; i++) {
            tp.small_sph[i] = make_float4((float)sorted[i]->c[0], (float)sorted[i]->c[1], (float)sorted[i]->c[2], (float)sorted[i]->r);
            tp.small_mat[i] = sorted[i]->mat;
        }
        // small_inside[i]: the spheres a ray can meet while it is INSIDE sphere i — i itself and every sphere whose
        // surface comes within reach of i's volume (centre distance <= r_i + r_j, with a margin).  A non-overlapping
        // sphere lies wholly outside i, so it can only be hit beyond i's exit point: never the closest hit.
        for (int i = 0; i < tp.small_n; i++) {
            uint32_t m = 1u << i;
            for (int j = 0; j < tp.small_n; j++) {
                if (j == i) continue;
                double d2 = 0;
                for (int a = 0; a < 3; a++) d2 += (sorted[i]->c[a] - sorted[j]->c[a]) * (sorted[i]->c[a] - sorted[j]->c[a]);
                const double reach = (std::fabs(sorted[i]->r) + std::fabs(sorted[j]->r)) * (1.0 + 1e-4) + 1e-5;
                if (std::sqrt(d2) <= reach) m |= 1u << j;
            }
            tp.small_inside[i] = (uint16_t)m;
        }
        for (size_t i = 0; i < ctx->scene.mats.size(); i++) {
            F4 m[4];
            pack_material(ctx->scene.mats[i], m);
            for (int k = 0; k < 4; k++) tp.small_mats[i][k] = make_float4(m[k].x, m[k].y, m[k].z, m[k].w);
        }
        for (size_t i = 0; i < ctx->scene.lights.size(); i++) {
            const HostLight& l = ctx->scene.lights[i];
            tp.small_lights[i][0] = make_float4((float)l.pos[0], (float)l.pos[1], (float)l.pos[2], (float)l.intensity);
            tp.small_lights[i][1] = make_float4((float)l.color[0], (float)l.color[1], (float)l.color[2], 0.f);
        }
    }
    tp.jitter = p->anti_aliasing ? 1 : 0; tp.recursive = p->recursive_reflections ? 1 : 0; tp.soft = p->soft_shadows ? 1 : 0;
    tp.tiles_x = tiles_x; tp.tiles_y = tiles_y;
    tp.shard_rank = eff_rank; tp.shard_count = eff_count; tp.n_local_tiles = n_local;
    tp.crop_x0 = p->crop_x0; tp.crop_y0 = p->crop_y0; tp.crop_x1 = p->crop_x1; tp.crop_y1 = p->crop_y1;
    // work units = (sample batch, active 8x4 block), sized on the device: aim for >= 16 per resident warp
    {
        // work units per resident warp.  Measured 4 / 8 / 16 / 32 / 64: C2-view 0.777 / 0.709 / 0.667 / 0.667 / 0.665 ms
        // (finer units shorten the end of the frame), C4 and C5 flat; C1-view is at one sample per unit either way.
        const char* upw = GORT_ENV_ONCE("GORT_UNITS_PER_WARP");
        const int k = upw ? std::max(1, atoi(upw)) : 16;
        tp.target_units = (uint32_t)k * (uint32_t)d.sm_count * 32u;
    }
    tp.active_list = d.d_active; tp.active_count = counters + 1;
    tp.block_active = reinterpret_cast<uint8_t*>(d.d_active + (size_t)n_local * 32);

    tp.debug_times = d.d_debug;
    tp.stamps = nullptr;
    if (d.stamped) {
        void* dp = nullptr;
        CUDA_TRY(ctx, cudaHostGetDevicePointer(&dp, d.h_stamps, 0));
        tp.stamps = static_cast<unsigned long long*>(dp);
    }
    tp.accum = d.d_accum; tp.work_counter = counters; tp.stats = p->collect_stats ? d.d_stats : nullptr;
    const uint32_t k0 = (uint32_t)p->seed, k1 = (uint32_t)(p->seed >> 32);
    for (int r = 0; r < 10; r++) {
        tp.rk[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u;
        tp.rk[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
    tp.dead_bound = GORT_ENV_ONCE("GORT_NO_DEAD_PATH") ? 0.f : dead_path_bound(ctx->scene, tp.cam, p->max_depth);
    tp.no_cone_cull = getenv("GORT_NO_CONE_CULL") ? 1 : 0;
    {
        const char* ud = GORT_ENV_ONCE("GORT_URGENT_DEPTH");
        tp.urgent_depth = ud ? atoi(ud) : 3;
    }
    tp.fog_enabled = ctx->scene.fog_enabled;
    tp.fog_density = (float)ctx->scene.fog_density;
    tp.fog_r = (float)ctx->scene.fog_color[0]; tp.fog_g = (float)ctx->scene.fog_color[1]; tp.fog_b = (float)ctx->scene.fog_color[2];
    {
        const double ex = 65520.0 * ctx->bvh.qcell[0], ey = 65520.0 * ctx->bvh.qcell[1], ez = 65520.0 * ctx->bvh.qcell[2];
        const double vol = ex * ey * ez;
        const double n_prims = (double)(ctx->scene.spheres.size() + ctx->scene.tris.size());
        tp.cone_skip = (vol > 0 && !GORT_ENV_ONCE("GORT_NO_CONE_SKIP")) ? (float)(n_prims / vol * 0.010578) : 0.f;
    }
    tp.sky_enabled = ctx->scene.sky_enabled;
    for (int k = 0; k < 27; k++) tp.sky[k] = (float)ctx->scene.sky_params[k];
    lap.mark(kLapEnqSetup);
    CUDA_TRY(ctx, launch_cull(tp, d.d_active, counters + 1, st, cx));
    if (timing) CUDA_TRY(ctx, cudaEventRecord(d.ev_tc, st));
    ResolveParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.block_active = tp.block_active;
    rp.accum = d.d_accum; rp.n_local_tiles = n_local; rp.shard_rank = eff_rank; rp.shard_count = eff_count;
    rp.tiles_x = tiles_x; rp.width = p->width; rp.height = p->height; rp.samples = p->samples;
    rp.out = out; rp.slab_mode = slab_mode; rp.part = 0;
    // the culled blocks of a host frame: written by the trace kernel itself (per-warp-queue path: the frame stays one chain of
    // dependent launches), or by a small resolve pass on the aux stream beside the wavefront pipeline's kernels
    static const bool early_aux = getenv("GORT_EARLY_AUX") != nullptr;
    // (for a frame in device memory the trace kernel writing the culled blocks changed nothing: 0.1746 vs 0.1756 ms)
    const bool early_any = hooks && hooks->early_black && n_local > 0 && !slab_mode;
    const bool early_in_trace = early_any && path != kPathStream && !early_aux;
    const bool early = early_any && !early_in_trace;
    if (early_in_trace) { tp.early_out = out; rp.part = 2; }
    if (early) {
        if (!d.aux_stream) {
            CUDA_TRY(ctx, cudaStreamCreateWithFlags(&d.aux_stream, cudaStreamNonBlocking));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&d.ev_cull, cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&d.ev_aux, cudaEventDisableTiming));
        }
        CUDA_TRY(ctx, cudaEventRecord(d.ev_cull, st));
        CUDA_TRY(ctx, cudaStreamWaitEvent(d.aux_stream, d.ev_cull, 0));
        rp.part = 1;
        CUDA_TRY(ctx, launch_resolve(rp, d.aux_stream, d.sm_count));  // 128-thread blocks: they fit beside the resident trace CTAs
        CUDA_TRY(ctx, cudaEventRecord(d.ev_aux, d.aux_stream));
        rp.part = 2;
    }
    // programmatic dependent launches: only between kernels that follow each other directly on the stream
    bool chained = false;
    if (path == kPathStream) {
        if (int rc = run_stream(ctx, d, tp, p, st)) return rc;
    } else {
        chained = !no_pdl && n_local > 0;
        CUDA_TRY(ctx, launch_trace(tp, p->collect_stats != 0, d.sm_count, st, chained && !timing && !early));
        d.last_launches++;
    }
    d.last_launches += 2 + (early ? 1 : 0);  // cull + resolve (+ the early pass over the culled blocks)
    if (timing) CUDA_TRY(ctx, cudaEventRecord(d.ev[1], st));

    // frame link, peer side: wait for the owner's release before the first store into its frame, tell it when the last one is
    // out.  One-thread kernels around the resolve pass: folded into it (every CTA polling the owner's word, a system-scope fence
    // per CTA before the last one signals) the two-GPU frame was 3 us slower (profiles/README.md).
    // They join the frame's chain of dependent launches: each is resident, parked in griddepcontrol.wait, before its turn comes.
    static const bool link_chain = getenv("GORT_LINK_NO_CHAIN") == nullptr;
    const bool chain = chained && !timing;
    const bool waits = hooks && hooks->wait_flag;
    if (waits) CUDA_TRY(ctx, launch_link_wait(hooks->wait_flag, hooks->wait_target, hooks->timed_out, st, chain && link_chain));
    rp.stamps = tp.stamps; rp.done_count = d.d_counter + 8;
    CUDA_TRY(ctx, launch_resolve(rp, st, 0, chain && (!waits || link_chain)));
    if (hooks && hooks->signal_flag) CUDA_TRY(ctx, launch_link_signal(hooks->signal_flag, st, chain && link_chain));
    if (early) CUDA_TRY(ctx, cudaStreamWaitEvent(st, d.ev_aux, 0));
    CUDA_TRY(ctx, cudaEventRecord(d.ev[2], st));
    lap.mark(kLapEnqLaunch);
    return GORT_OK;
}

int collect_stats(gort_ctx* ctx, const gort_render_params* p, gort_stats* s, double t_start_ms, int n_devs_used) {
    if (!s) return GORT_OK;
    memset(s, 0, sizeof(*s));
    s->n_devices = n_devs_used;
    s->upload_ms = ctx->upload_ms;
    s->bvh_build_ms = ctx->bvh_ms;
    s->bvh_nodes = (uint64_t)ctx->bvh.n_nodes;
    s->bvh_bytes = ctx->bvh_bytes;
    unsigned long long tot[kStatCount] = {0};
    int64_t pixels = 0;
    const int tiles_x = (p->width + kTile - 1) / kTile;
    for (int i = 0; i < n_devs_used; i++) {
        DeviceState& d = ctx->devs[i];
        CUDA_TRY(ctx, cudaSetDevice(d.dev));
        CUDA_TRY(ctx, cudaEventSynchronize(d.ev[2]));
        float c = 0, a = 0, b = 0;
        if (d.stamped) {
            // cull pass = first cull thread -> first trace thread past its wait for the cull pass; trace = that -> first resolve
            // thread past its wait for the trace kernel; resolve = that -> its last CTA
            const volatile unsigned long long* t = d.h_stamps;  // the frame is complete (ev[2]): the stores have landed
            c = (float)((double)(t[1] - t[0]) * 1e-6);
            a = (float)((double)(t[2] - t[1]) * 1e-6);
            b = (float)((double)(t[3] - t[2]) * 1e-6);
        } else {
            CUDA_TRY(ctx, cudaEventElapsedTime(&c, d.ev[0], d.ev_tc));  // memsets + cull pass
            CUDA_TRY(ctx, cudaEventElapsedTime(&a, d.ev_tc, d.ev[1]));  // the trace kernel(s) alone
            CUDA_TRY(ctx, cudaEventElapsedTime(&b, d.ev[1], d.ev[2]));
        }
        s->device_ms[i] = c + a + b;
        if (i == 0) {
            s->cull_ms = c; s->trace_ms = a; s->resolve_ms = b;
            s->render_path = d.last_path; s->kernel_launches = d.last_launches;
        }
        s->kernel_ms = std::max(s->kernel_ms, (double)(c + a + b));
        s->n_tiles += ctx->last_local_tiles[i];
        if (p->collect_stats) {
            unsigned long long h[kStatCount];
            CUDA_TRY(ctx, cudaMemcpy(h, d.d_stats, sizeof(h), cudaMemcpyDeviceToHost));
            for (int k = 0; k < kStatCount; k++) tot[k] += h[k];
        }
    }
    // pixels actually covered by the rendered tiles
    {
        const int eff_count = ctx->last_count * n_devs_used;
        const int tiles_y = (p->height + kTile - 1) / kTile;
        for (int i = 0; i < n_devs_used; i++) {
            const int eff_rank = ctx->last_rank + i * ctx->last_count;
            for (int t = eff_rank; t < tiles_x * tiles_y; t += eff_count) {
                const int tx = t % tiles_x, ty = t / tiles_x;
                pixels += (int64_t)std::min(kTile, p->width - tx * kTile) * std::min(kTile, p->height - ty * kTile);
            }
        }
    }
    s->primary_rays = (uint64_t)pixels * (uint64_t)p->samples;
    if (p->collect_stats) {
        s->closest_queries = tot[kStatClosest]; s->shadow_queries = tot[kStatShadow]; s->nodes_visited = tot[kStatNodes];
        s->sphere_tests = tot[kStatSphereTests]; s->sphere_hits = tot[kStatSphereHits];
        s->tri_tests = tot[kStatTriTests]; s->tri_hits = tot[kStatTriHits];
        s->tri_rejects[0] = tot[kStatTriRejA]; s->tri_rejects[1] = tot[kStatTriRejU]; s->tri_rejects[2] = tot[kStatTriRejV]; s->tri_rejects[3] = tot[kStatTriRejT];
        s->shaded_hits = tot[kStatShaded]; s->rng_blocks = tot[kStatRngBlocks]; s->light_evals = tot[kStatLightEvals];
        s->soft_pairs_skipped = tot[kStatSoftSkipped]; s->pairs_backfacing = tot[kStatBackfacing];
        s->soft_shadow_rays = tot[kStatSoftRays]; s->diffuse_evals = tot[kStatDiffuse]; s->specular_evals = tot[kStatSpec];
        s->paths_depth_ge5 = tot[kStatDepth5]; s->paths_depth_ge20 = tot[kStatDepth20]; s->paths_depth_max = tot[kStatDepthMax];
        s->cone_tests = tot[kStatConeTests];
        for (int k = 0; k < 4; k++) { s->walk_lane_visits[k] = tot[kStatWalkLane0 + k]; s->walk_warp_visits[k] = tot[kStatWalkWarp0 + k]; }
        // SURVEY §8d operation costs (FMA = 2 flops): ray generation 12, AABB slab 24 (two per node),
        // sphere 23 miss / 47 hit, triangle 20/30/46/52 staged rejects / 92 accept, 30 per (hit, light)
        // set-up, 36 per soft-shadow direction, 50 per diffuse term, 45 per specular term, ~70 per
        // scatter; 30 per cone test of the soft-shadow candidate pass (this implementation's own pruning
        // work, like the BVH slabs).
        // Only work the trace kernels execute is counted: ray generation for the samples actually generated (the
        // blocks the cull pass drops never produce a ray), no tone-map term (that is resolve_kernel's work).
        s->primary_generated = tot[kStatPrimary];
        s->algorithmic_flops = 12.0 * (double)s->primary_generated + 48.0 * (double)s->nodes_visited +
                               23.0 * (double)(s->sphere_tests - s->sphere_hits) + 47.0 * (double)s->sphere_hits +
                               20.0 * (double)s->tri_rejects[0] + 30.0 * (double)s->tri_rejects[1] + 46.0 * (double)s->tri_rejects[2] +
                               52.0 * (double)s->tri_rejects[3] + 92.0 * (double)s->tri_hits + 30.0 * (double)tot[kStatPairSetups] +
                               36.0 * (double)s->soft_shadow_rays + 50.0 * (double)s->diffuse_evals + 45.0 * (double)s->specular_evals +
                               70.0 * (double)s->shaded_hits + 30.0 * (double)s->cone_tests;
    }
    s->total_ms = now_ms() - t_start_ms;
    return GORT_OK;
}

}  // namespace

extern "C" {

int gort_abi_version(void) { return (int)GORT_ABI_VERSION; }

int gort_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? 0 : GORT_ERR_CUDA;
    }
    return n;
}

int gort_create(const int* device_ids, int n_devices, gort_ctx** out) {
    if (!out) return fail(nullptr, GORT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_devices <= 0 || n_devices > GORT_MAX_DEVICES) return fail(nullptr, GORT_ERR_INVALID, "n_devices must be in 1..8");
    const int avail = gort_device_count();
    if (avail <= 0) return fail(nullptr, GORT_ERR_NO_DEVICE, "no CUDA device available (libgort has no CPU fallback): " + g_create_error);
    gort_ctx* ctx = new gort_ctx();
    ctx->devs.resize(n_devices);
    ctx->last_local_tiles.assign(n_devices, 0);
    auto bail = [&](int code, const std::string& msg) {
        g_create_error = msg;
        gort_destroy(ctx);
        return code;
    };
    for (int i = 0; i < n_devices; i++) {
        DeviceState& d = ctx->devs[i];
        d.dev = device_ids ? device_ids[i] : i;
        if (d.dev < 0 || d.dev >= avail) return bail(GORT_ERR_INVALID, "device id out of range");
        cudaError_t e = cudaSetDevice(d.dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.dev);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.own_stream, cudaStreamNonBlocking);
        for (int k = 0; k < 4 && e == cudaSuccess; k++) e = cudaEventCreate(&d.ev[k]);
        if (e == cudaSuccess) e = cudaEventCreate(&d.ev_tc);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_count, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&d.d_ctl, (2 * kCtlWords + 4) * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMallocHost(&d.h_count, 64);
        if (e == cudaSuccess) e = cudaHostAlloc(&d.h_stamps, 4 * sizeof(unsigned long long), cudaHostAllocPortable | cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaMalloc(&d.d_counter, 12 * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMemset(d.d_counter, 0, 12 * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMalloc(&d.d_stats, kStatCount * sizeof(unsigned long long));
        if (e == cudaSuccess && getenv("GORT_DEBUG_TIMES")) e = cudaMalloc(&d.d_debug, (1 + 16 * 148 * 64) * sizeof(unsigned long long));
        if (e != cudaSuccess) return bail(GORT_ERR_CUDA, std::string("device init: ") + cudaGetErrorString(e));
    }
    if (n_devices > 1) {  // NVLink peer access: the other devices store their tiles straight into the lead's frame
        ctx->peer_direct = true;
        for (int i = 1; i < n_devices; i++) {
            int can = 0, can_back = 0;
            cudaDeviceCanAccessPeer(&can, ctx->devs[0].dev, ctx->devs[i].dev);
            cudaDeviceCanAccessPeer(&can_back, ctx->devs[i].dev, ctx->devs[0].dev);
            if (can) {
                cudaSetDevice(ctx->devs[0].dev);
                cudaDeviceEnablePeerAccess(ctx->devs[i].dev, 0);
            }
            if (can_back) {
                cudaSetDevice(ctx->devs[i].dev);
                cudaError_t e = cudaDeviceEnablePeerAccess(ctx->devs[0].dev, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can_back = 0;
            }
            cudaGetLastError();
            if (!can_back) ctx->peer_direct = false;  // fall back to slabs + peer copies + un-swizzle
        }
        if (getenv("GORT_NO_PEER_DIRECT")) ctx->peer_direct = false;
    }
    *out = ctx;
    return GORT_OK;
}

void gort_destroy(gort_ctx* ctx) {
    if (!ctx) return;
    if (g_laps) {
        for (int k = 0; k < kLapCount; k++)
            if (g_lap_n[k] > 8) fprintf(stderr, "[gort host] %-40s %9.2f us x %llu\n", kLapNames[k], g_lap_us[k] / (double)(g_lap_n[k] - 8), g_lap_n[k] - 8);
        memset(g_lap_us, 0, sizeof(g_lap_us)); memset(g_lap_n, 0, sizeof(g_lap_n));
    }
    for (DeviceState& d : ctx->devs) {
        if (d.dev < 0) continue;
        cudaSetDevice(d.dev);
        if (d.own_stream) cudaStreamSynchronize(d.own_stream);
        free_scene(d);
        cudaFree(d.d_debug); cudaFree(d.d_accum); cudaFree(d.d_active); cudaFree(d.d_counter); cudaFree(d.d_stats); cudaFree(d.d_out); cudaFree(d.d_gather);
        if (d.h_pinned) cudaFreeHost(d.h_pinned);
        if (d.h_stamps) cudaFreeHost(d.h_stamps);
        for (auto& e : d.ev)
            if (e) cudaEventDestroy(e);
        if (d.aux_stream) { cudaStreamSynchronize(d.aux_stream); cudaStreamDestroy(d.aux_stream); }
        if (d.ev_cull) cudaEventDestroy(d.ev_cull);
        if (d.ev_aux) cudaEventDestroy(d.ev_aux);
        if (d.ev_tc) cudaEventDestroy(d.ev_tc);
        if (d.ev_count) cudaEventDestroy(d.ev_count);
        cudaFree(d.d_stream); cudaFree(d.d_sort); cudaFree(d.d_top); cudaFree(d.d_ctl); cudaFree(d.d_build);
        if (d.h_count) cudaFreeHost(d.h_count);
        if (d.own_stream) cudaStreamDestroy(d.own_stream);
    }
    if (ctx->h_build) cudaFreeHost(ctx->h_build);
    delete ctx;
}

const char* gort_last_error(const gort_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gort_set_stream(gort_ctx* ctx, void* cuda_stream) {
    if (!ctx) return GORT_ERR_INVALID;
    for (DeviceState& d : ctx->devs)  // an asynchronous scene upload on the old stream must land before the new stream reads it
        if (d.ev_upload) {
            cudaSetDevice(d.dev);
            cudaEventSynchronize(d.ev_upload);
        }
    ctx->user_stream = (cudaStream_t)cuda_stream;
    ctx->use_user_stream = cuda_stream != nullptr;
    return GORT_OK;
}

int gort_scene_upload(gort_ctx* ctx, const gort_scene_desc* desc) {
    if (!ctx) return GORT_ERR_INVALID;
    if (!desc) return fail(ctx, GORT_ERR_INVALID, "desc is NULL");
    Lap lap;
    // built in the spare scene and swapped in: a failed upload leaves the current scene alone, and the arrays of the scene
    // before last are reused (re-uploading a 1 M-primitive scene neither maps nor unmaps its 90 MB)
    std::string err = scene_from_desc(*desc, ctx->scene_spare);
    if (!err.empty()) return fail(ctx, GORT_ERR_INVALID, err);
    std::swap(ctx->scene, ctx->scene_spare);
    lap.mark(kLapDesc);
    return upload_scene(ctx);
}

int gort_scene_load_json(gort_ctx* ctx, const char* json_text, size_t json_len, uint32_t options) {
    if (!ctx) return GORT_ERR_INVALID;
    if (!json_text) return fail(ctx, GORT_ERR_INVALID, "json_text is NULL");
    HostScene hs;
    std::string err = scene_from_json(json_text, json_len, options, hs);
    if (!err.empty()) return fail(ctx, GORT_ERR_PARSE, err);
    ctx->scene = std::move(hs);
    return upload_scene(ctx);
}

int gort_scene_load_file(gort_ctx* ctx, const char* path, uint32_t options) {
    if (!ctx) return GORT_ERR_INVALID;
    if (!path) return fail(ctx, GORT_ERR_INVALID, "path is NULL");
    std::ifstream f(path, std::ios::binary);
    if (!f) return fail(ctx, GORT_ERR_IO, std::string("error reading file: ") + path);  // scene.go:46-49
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    return gort_scene_load_json(ctx, text.data(), text.size(), options);
}

int gort_scene_counts(const gort_ctx* ctx, int32_t* n_spheres, int32_t* n_triangles, int32_t* n_materials, int32_t* n_lights,
                      int32_t* n_hittables) {
    if (!ctx || !ctx->has_scene) return GORT_ERR_NO_SCENE;
    if (n_spheres) *n_spheres = (int32_t)ctx->scene.spheres.size();
    if (n_triangles) *n_triangles = (int32_t)ctx->scene.tris.size();
    if (n_materials) *n_materials = (int32_t)ctx->scene.mats.size();
    if (n_lights) *n_lights = (int32_t)ctx->scene.lights.size();
    if (n_hittables) *n_hittables = ctx->scene.n_hittables;
    return GORT_OK;
}

int gort_scene_get_triangle(const gort_ctx* ctx, int32_t i, double* v9, int32_t* material) {
    if (!ctx || !ctx->has_scene) return GORT_ERR_NO_SCENE;
    if (i < 0 || i >= (int32_t)ctx->scene.tris.size() || !v9) return GORT_ERR_INVALID;
    memcpy(v9, ctx->scene.tris[i].v, 9 * sizeof(double));
    if (material) *material = ctx->scene.tris[i].mat;
    return GORT_OK;
}

int gort_scene_get_material(const gort_ctx* ctx, int32_t i, int32_t* type, double* out7) {
    if (!ctx || !ctx->has_scene) return GORT_ERR_NO_SCENE;
    if (i < 0 || i >= (int32_t)ctx->scene.mats.size() || !out7) return GORT_ERR_INVALID;
    const HostMaterial& m = ctx->scene.mats[i];
    if (type) *type = m.type;
    out7[0] = m.color[0]; out7[1] = m.color[1]; out7[2] = m.color[2];
    out7[3] = m.roughness; out7[4] = m.metallic; out7[5] = m.specular; out7[6] = m.ior;
    return GORT_OK;
}

int gort_scene_render_hints(const gort_ctx* ctx, int32_t* hints5) {
    if (!ctx || !ctx->has_scene) return GORT_ERR_NO_SCENE;
    if (!hints5) return GORT_ERR_INVALID;
    memcpy(hints5, ctx->scene.render_hints, 5 * sizeof(int32_t));
    return GORT_OK;
}

size_t gort_shard_slab_bytes(int32_t width, int32_t height, int32_t shard_count) {
    if (width <= 0 || height <= 0) return 0;
    if (shard_count <= 0) shard_count = 1;
    const int64_t n_tiles = (int64_t)((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    return (size_t)((n_tiles + shard_count - 1) / shard_count) * kTilePixels * 4;
}

int gort_render_shard_device(gort_ctx* ctx, const gort_render_params* p, void* d_slab, size_t slab_bytes, gort_stats* stats_out) {
    const double t0 = now_ms();
    if (int rc = validate(ctx, p)) return rc;
    const int sc = p->shard_count <= 0 ? 1 : p->shard_count;
    if (!d_slab || slab_bytes != gort_shard_slab_bytes(p->width, p->height, sc)) return fail(ctx, GORT_ERR_INVALID, "slab pointer/size mismatch");
    ctx->last_w = p->width; ctx->last_h = p->height; ctx->last_samples = p->samples; ctx->last_rank = p->shard_rank; ctx->last_count = sc;
    ctx->timing = stats_out != nullptr;
    if (int rc = enqueue_device(ctx, 0, p, p->shard_rank, sc, 1, (uint8_t*)d_slab, slab_bytes)) return rc;
    if (stats_out) return collect_stats(ctx, p, stats_out, t0, 1);
    return GORT_OK;
}

int gort_unswizzle_device(gort_ctx* ctx, const void* d_slabs, int32_t shard_count, int32_t width, int32_t height, void* d_rgba,
                          size_t rgba_bytes) {
    if (!ctx) return GORT_ERR_INVALID;
    if (!d_slabs || !d_rgba || shard_count <= 0 || width <= 0 || height <= 0 || rgba_bytes != (size_t)width * height * 4)
        return fail(ctx, GORT_ERR_INVALID, "gort_unswizzle_device: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    CUDA_TRY(ctx, launch_unswizzle((const uint8_t*)d_slabs, shard_count, width, height, (uint8_t*)d_rgba, stream_of(ctx, 0)));
    return GORT_OK;
}

// Frame into device memory of the lead device (row-major).  n_devices > 1: every device renders its
// interleaved tiles into a slab, slabs are gathered to the lead device over NVLink and unswizzled.
static int render_frame_device(gort_ctx* ctx, const gort_render_params* p, uint8_t* d_rgba, double t0, gort_stats* stats_out, bool sync,
                               const ResolveHooks* hooks = nullptr) {
    const int nd = (int)ctx->devs.size();
    const int sc = p->shard_count <= 0 ? 1 : p->shard_count;
    ctx->last_w = p->width; ctx->last_h = p->height; ctx->last_samples = p->samples; ctx->last_rank = p->shard_rank; ctx->last_count = sc;
    if (nd == 1) {
        if (int rc = enqueue_device(ctx, 0, p, p->shard_rank, sc, 0, d_rgba, 0, hooks)) return rc;
    } else {
        if (sc != 1) return fail(ctx, GORT_ERR_INVALID, "a multi-device ctx renders whole frames (shard_count must be 1)");
        const size_t slab = gort_shard_slab_bytes(p->width, p->height, nd);
        DeviceState& lead = ctx->devs[0];
        if (ctx->peer_direct) {
            // every device resolves its interleaved tiles into the lead's row-major frame (peer stores over NVLink)
            for (int i = 0; i < nd; i++)
                if (int rc = enqueue_device(ctx, i, p, i, nd, 0, d_rgba, 0)) return rc;
            CUDA_TRY(ctx, cudaSetDevice(lead.dev));
            cudaStream_t st0 = stream_of(ctx, 0);
            for (int i = 1; i < nd; i++) CUDA_TRY(ctx, cudaStreamWaitEvent(st0, ctx->devs[i].ev[2], 0));
            CUDA_TRY(ctx, cudaEventRecord(lead.ev[3], st0));
        } else {
        CUDA_TRY(ctx, cudaSetDevice(lead.dev));
        if (int rc = ensure(ctx, lead.d_gather, lead.gather_bytes, slab * nd)) return rc;
        for (int i = 0; i < nd; i++) {
            uint8_t* dst = (i == 0) ? lead.d_gather : nullptr;  // lead device resolves straight into the gather buffer
            if (int rc = enqueue_device(ctx, i, p, i, nd, 1, dst, slab)) return rc;
        }
        CUDA_TRY(ctx, cudaSetDevice(lead.dev));
        cudaStream_t st0 = stream_of(ctx, 0);
        for (int i = 1; i < nd; i++) {
            CUDA_TRY(ctx, cudaStreamWaitEvent(st0, ctx->devs[i].ev[2], 0));
            CUDA_TRY(ctx, cudaMemcpyPeerAsync(lead.d_gather + slab * i, lead.dev, ctx->devs[i].d_out, ctx->devs[i].dev, slab, st0));
        }
        CUDA_TRY(ctx, launch_unswizzle(lead.d_gather, nd, p->width, p->height, d_rgba, st0));
        CUDA_TRY(ctx, cudaEventRecord(lead.ev[3], st0));
        }
    }
    if (stats_out || sync) {
        CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
        CUDA_TRY(ctx, cudaStreamSynchronize(stream_of(ctx, 0)));
    }
    if (stats_out) {
        if (int rc = collect_stats(ctx, p, stats_out, t0, nd)) return rc;
        if (nd > 1) {
            float g = 0;
            CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
            CUDA_TRY(ctx, cudaEventElapsedTime(&g, ctx->devs[0].ev[0], ctx->devs[0].ev[3]));
            stats_out->kernel_ms = std::max(stats_out->kernel_ms, (double)g);
            stats_out->resolve_ms = g - stats_out->trace_ms - stats_out->cull_ms;
        }
    }
    return GORT_OK;
}

int gort_render_device(gort_ctx* ctx, const gort_render_params* p, void* d_rgba, size_t rgba_bytes, gort_stats* stats_out) {
    const double t0 = now_ms();
    if (int rc = validate(ctx, p)) return rc;
    if (!d_rgba || rgba_bytes != (size_t)p->width * p->height * 4) return fail(ctx, GORT_ERR_INVALID, "d_rgba pointer/size mismatch");
    ctx->timing = stats_out != nullptr;
    return render_frame_device(ctx, p, (uint8_t*)d_rgba, t0, stats_out, false);
}

int gort_render(gort_ctx* ctx, const gort_render_params* p, uint8_t* rgba_out, size_t rgba_bytes, gort_stats* stats_out) {
    const double t0 = now_ms();
    if (int rc = validate(ctx, p)) return rc;
    const size_t frame_bytes = (size_t)p->width * p->height * 4;
    if (!rgba_out || rgba_bytes != frame_bytes) return fail(ctx, GORT_ERR_INVALID, "rgba_out pointer/size mismatch (want width*height*4)");
    ctx->timing = stats_out != nullptr;
    Lap lap;
    DeviceState& lead = ctx->devs[0];
    CUDA_TRY(ctx, cudaSetDevice(lead.dev));
    const int sc = p->shard_count <= 0 ? 1 : p->shard_count;
    cudaStream_t st0 = stream_of(ctx, 0);
    if (sc == 1 && ctx->devs.size() == 1 && !getenv("GORT_NO_ZERO_COPY")) {
        // Page-locked caller memory is mapped into the device's address space (UVA): resolve_kernel then stores the RGBA8
        // pixels straight into the caller's buffer over PCIe — 128 contiguous bytes per warp — instead of into HBM followed
        // by a device-to-host copy: one launch and one DMA set-up less, and the transfer overlaps the tone mapping.
        // Pageable caller memory (a Go slice, a numpy array) cannot be mapped: the same stores then go to the ctx's own
        // page-locked frame and one host memcpy moves it into the caller's buffer.
        cudaPointerAttributes pa0;
        uint8_t* target = rgba_out;
        bool staged = false;
        if (!(cudaPointerGetAttributes(&pa0, rgba_out) == cudaSuccess && pa0.type == cudaMemoryTypeHost && pa0.devicePointer)) {
            cudaGetLastError();
            if (lead.pinned_bytes < frame_bytes) {
                if (lead.h_pinned) CUDA_TRY(ctx, cudaFreeHost(lead.h_pinned));
                lead.h_pinned = nullptr; lead.pinned_bytes = 0;
                CUDA_TRY(ctx, cudaMallocHost(&lead.h_pinned, frame_bytes));
                lead.pinned_bytes = frame_bytes;
            }
            target = lead.h_pinned;
            staged = true;
        }
        if (cudaPointerGetAttributes(&pa0, target) == cudaSuccess && pa0.type == cudaMemoryTypeHost && pa0.devicePointer) {
            ResolveHooks hk;
            hk.early_black = !GORT_ENV_ONCE("GORT_NO_EARLY_BLACK");
            lap.mark(kLapPre);
            if (int rc = render_frame_device(ctx, p, (uint8_t*)pa0.devicePointer, t0, nullptr, false, &hk)) return rc;
            lap.mark(kLapEnqueue);
            CUDA_TRY(ctx, cudaStreamSynchronize(st0));
            lap.mark(kLapSync);
            if (staged) memcpy(rgba_out, lead.h_pinned, frame_bytes);
            if (stats_out) {
                if (int rc = collect_stats(ctx, p, stats_out, t0, 1)) return rc;
                stats_out->total_ms = now_ms() - t0;
            }
            lap.mark(kLapStats);
            return GORT_OK;
        }
        cudaGetLastError();
    }
    if (sc == 1) {
        if (lead.pinned_bytes < frame_bytes) {
            if (lead.h_pinned) CUDA_TRY(ctx, cudaFreeHost(lead.h_pinned));
            lead.h_pinned = nullptr; lead.pinned_bytes = 0;
            CUDA_TRY(ctx, cudaMallocHost(&lead.h_pinned, frame_bytes));
            lead.pinned_bytes = frame_bytes;
        }
        if (ctx->devs.size() == 1) {
            if (int rc = render_frame_device(ctx, p, nullptr, t0, nullptr, false)) return rc;  // into lead.d_out
        } else {
            if (int rc = ensure(ctx, lead.d_out, lead.out_bytes, std::max(frame_bytes, gort_shard_slab_bytes(p->width, p->height, (int)ctx->devs.size())))) return rc;
            // lead.d_out doubles as the frame; its own slab goes straight into d_gather
            if (int rc = render_frame_device(ctx, p, lead.d_out, t0, nullptr, false)) return rc;
        }
        CUDA_TRY(ctx, cudaSetDevice(lead.dev));
        // page-locked caller memory (cudaHostRegister / cudaMallocHost) takes the DMA directly;
        // pageable memory (a Go slice) goes through the ctx's pinned staging buffer
        cudaPointerAttributes pa;
        const bool caller_pinned = cudaPointerGetAttributes(&pa, rgba_out) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (caller_pinned) {
            CUDA_TRY(ctx, cudaMemcpyAsync(rgba_out, lead.d_out, frame_bytes, cudaMemcpyDeviceToHost, st0));
            CUDA_TRY(ctx, cudaStreamSynchronize(st0));
        } else {
            CUDA_TRY(ctx, cudaMemcpyAsync(lead.h_pinned, lead.d_out, frame_bytes, cudaMemcpyDeviceToHost, st0));
            CUDA_TRY(ctx, cudaStreamSynchronize(st0));
            memcpy(rgba_out, lead.h_pinned, frame_bytes);
        }
    } else {
        // process-level shard with a host destination: render the slab, copy back, write own tiles only
        if (ctx->devs.size() != 1) return fail(ctx, GORT_ERR_INVALID, "a multi-device ctx renders whole frames (shard_count must be 1)");
        const size_t slab = gort_shard_slab_bytes(p->width, p->height, sc);
        if (lead.pinned_bytes < slab) {
            if (lead.h_pinned) CUDA_TRY(ctx, cudaFreeHost(lead.h_pinned));
            lead.h_pinned = nullptr; lead.pinned_bytes = 0;
            CUDA_TRY(ctx, cudaMallocHost(&lead.h_pinned, slab));
            lead.pinned_bytes = slab;
        }
        ctx->last_w = p->width; ctx->last_h = p->height; ctx->last_samples = p->samples; ctx->last_rank = p->shard_rank; ctx->last_count = sc;
        if (int rc = enqueue_device(ctx, 0, p, p->shard_rank, sc, 1, nullptr, slab)) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(lead.h_pinned, lead.d_out, slab, cudaMemcpyDeviceToHost, st0));
        CUDA_TRY(ctx, cudaStreamSynchronize(st0));
        const int tiles_x = (p->width + kTile - 1) / kTile, tiles_y = (p->height + kTile - 1) / kTile;
        int j = 0;
        for (int t = p->shard_rank; t < tiles_x * tiles_y; t += sc, j++) {
            const int tx = t % tiles_x, ty = t / tiles_x;
            const int w = std::min(kTile, p->width - tx * kTile), h = std::min(kTile, p->height - ty * kTile);
            for (int ly = 0; ly < h; ly++)
                memcpy(rgba_out + ((size_t)(ty * kTile + ly) * p->width + tx * kTile) * 4, lead.h_pinned + ((size_t)j * kTilePixels + ly * kTile) * 4, (size_t)w * 4);
        }
    }
    if (stats_out) {
        if (int rc = collect_stats(ctx, p, stats_out, t0, (int)ctx->devs.size())) return rc;
        if (ctx->devs.size() > 1) {
            float g = 0;
            CUDA_TRY(ctx, cudaSetDevice(lead.dev));
            CUDA_TRY(ctx, cudaEventElapsedTime(&g, lead.ev[0], lead.ev[3]));
            stats_out->kernel_ms = std::max(stats_out->kernel_ms, (double)g);
            stats_out->resolve_ms = g - stats_out->trace_ms - stats_out->cull_ms;
        }
        stats_out->total_ms = now_ms() - t0;
    }
    return GORT_OK;
}

int gort_host_alloc(size_t bytes, void** out) {
    if (!out || bytes == 0) return GORT_ERR_INVALID;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        g_create_error = std::string("gort_host_alloc: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? GORT_ERR_NO_DEVICE : GORT_ERR_CUDA;
    }
    return GORT_OK;
}

void gort_host_free(void* p) {
    if (p) cudaFreeHost(p);
    cudaGetLastError();
}

int gort_link_create(gort_ctx* ctx, int32_t width, int32_t height, int32_t n_ranks, uint8_t* handle_out, gort_link** out) {
    if (!ctx || !out || !handle_out) return GORT_ERR_INVALID;
    *out = nullptr;
    if (width <= 0 || height <= 0 || n_ranks < 1 || n_ranks > 64) return fail(ctx, GORT_ERR_INVALID, "gort_link_create: bad arguments");
    if (ctx->devs.size() != 1) return fail(ctx, GORT_ERR_INVALID, "a frame link joins single-device contexts (one process per GPU)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    gort_link* l = new gort_link();
    l->owner = true; l->width = width; l->height = height; l->n_ranks = n_ranks; l->rank = 0;
    l->frame_bytes = (size_t)width * height * 4;
    const size_t total = ((l->frame_bytes + 255) / 256) * 256 + 256;  // frame + control block {arrived, consumed}
    cudaError_t e = cudaMalloc(&l->base, total);
    if (e == cudaSuccess) e = cudaMemset(l->base, 0, total);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, l->base);
    if (e != cudaSuccess) {
        if (l->base) cudaFree(l->base);
        delete l;
        cudaGetLastError();
        return fail(ctx, GORT_ERR_CUDA, std::string("gort_link_create: ") + cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == GORT_LINK_HANDLE_BYTES, "IPC handle size");
    memcpy(handle_out, &h, sizeof(h));
    l->ctrl = reinterpret_cast<unsigned int*>(l->base + ((l->frame_bytes + 255) / 256) * 256);
    *out = l;
    return GORT_OK;
}

int gort_link_open(gort_ctx* ctx, const uint8_t* handle, int32_t width, int32_t height, int32_t n_ranks, int32_t rank, gort_link** out) {
    if (!ctx || !out || !handle) return GORT_ERR_INVALID;
    *out = nullptr;
    if (width <= 0 || height <= 0 || n_ranks < 2 || rank < 1 || rank >= n_ranks) return fail(ctx, GORT_ERR_INVALID, "gort_link_open: bad arguments");
    if (ctx->devs.size() != 1) return fail(ctx, GORT_ERR_INVALID, "a frame link joins single-device contexts (one process per GPU)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, GORT_ERR_CUDA, std::string("gort_link_open: ") + cudaGetErrorString(e));
    }
    gort_link* l = new gort_link();
    l->owner = false; l->width = width; l->height = height; l->n_ranks = n_ranks; l->rank = rank;
    l->frame_bytes = (size_t)width * height * 4;
    l->base = (uint8_t*)base;
    l->ctrl = reinterpret_cast<unsigned int*>(l->base + ((l->frame_bytes + 255) / 256) * 256);
    *out = l;
    return GORT_OK;
}

int gort_link_open_local(gort_ctx* ctx, gort_link* owner, int32_t rank, gort_link** out) {
    if (!ctx || !owner || !out || !owner->owner || rank < 1 || rank >= owner->n_ranks) return GORT_ERR_INVALID;
    gort_link* l = new gort_link(*owner);
    l->owner = false; l->local_alias = true; l->rank = rank; l->frame_no = 0; l->has_local_peers = false;
    owner->has_local_peers = true;
    *out = l;
    return GORT_OK;
}

int gort_link_read(gort_ctx* ctx, gort_link* l, uint8_t* rgba_out, size_t rgba_bytes) {
    if (!ctx || !l || !rgba_out) return GORT_ERR_INVALID;
    if (!l->owner || rgba_bytes != l->frame_bytes) return fail(ctx, GORT_ERR_INVALID, "gort_link_read: owner link and width*height*4 bytes expected");
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    cudaStream_t st = stream_of(ctx, 0);
    CUDA_TRY(ctx, cudaMemcpyAsync(rgba_out, l->base, rgba_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    unsigned int timed_out = 0;
    CUDA_TRY(ctx, cudaMemcpy(&timed_out, l->ctrl + 2, sizeof(timed_out), cudaMemcpyDeviceToHost));
    if (timed_out) {
        CUDA_TRY(ctx, cudaMemset(l->ctrl + 2, 0, sizeof(unsigned int)));  // reported once: later frames start clean
        return fail(ctx, GORT_ERR_CUDA, "frame link: a rank did not deliver its tiles within 20 s (frame incomplete)");
    }
    return GORT_OK;
}

void gort_link_close(gort_ctx* ctx, gort_link* l) {
    if (!l) return;
    if (ctx && !ctx->devs.empty()) {
        cudaSetDevice(ctx->devs[0].dev);
        cudaStreamSynchronize(stream_of(ctx, 0));
    }
    if (l->base && !l->local_alias) {
        if (l->owner) cudaFree(l->base);
        else cudaIpcCloseMemHandle(l->base);
    }
    cudaGetLastError();
    delete l;
}

void* gort_link_frame(const gort_link* l) { return l ? (void*)l->base : nullptr; }

int gort_render_linked(gort_ctx* ctx, const gort_render_params* p, gort_link* l, gort_stats* stats_out) {
    const double t0 = now_ms();
    if (int rc = validate(ctx, p)) return rc;
    if (!l) return fail(ctx, GORT_ERR_INVALID, "link is NULL");
    if (p->width != l->width || p->height != l->height) return fail(ctx, GORT_ERR_INVALID, "gort_render_linked: size differs from the link's frame");
    if (ctx->devs.size() != 1) return fail(ctx, GORT_ERR_INVALID, "a frame link joins single-device contexts (one process per GPU)");
    const unsigned int k = ++l->frame_no;  // 1, 2, ...
    ctx->last_w = p->width; ctx->last_h = p->height; ctx->last_samples = p->samples; ctx->last_rank = l->rank; ctx->last_count = l->n_ranks;
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
    cudaStream_t st = stream_of(ctx, 0);
    ResolveHooks hk;
    ctx->timing = stats_out != nullptr;
    if (l->owner) {
        // everything this stream did with frame k-1 is done when the frame's first kernel runs: the peers may overwrite it
        hk.store_flag = l->ctrl + 1; hk.store_value = k - 1;
        if (l->has_local_peers) {
            // Ranks that share this GPU cannot run concurrently with a kernel that waits for them: the peers must have rendered
            // (and finished) frame k already.  Checked on the host instead of enqueuing a wait that could only time out.
            unsigned int arrived = 0;
            CUDA_TRY(ctx, cudaDeviceSynchronize());
            CUDA_TRY(ctx, cudaMemcpy(&arrived, l->ctrl + 0, sizeof(arrived), cudaMemcpyDeviceToHost));
            if ((int)(arrived - k * (unsigned int)(l->n_ranks - 1)) < 0) {
                l->frame_no--;
                return fail(ctx, GORT_ERR_INVALID, "frame link on one GPU: render every peer rank's frame before the owner's");
            }
        }
        if (int rc = enqueue_device(ctx, 0, p, 0, l->n_ranks, 0, l->base, 0, &hk)) return rc;
        if (l->n_ranks > 1 && !l->has_local_peers) CUDA_TRY(ctx, launch_link_wait(l->ctrl + 0, k * (unsigned int)(l->n_ranks - 1), l->ctrl + 2, st));
    } else {
        // (a local alias renders before its owner and on the same GPU: there is no one to wait for)
        if (!l->local_alias) { hk.wait_flag = l->ctrl + 1; hk.wait_target = k - 1; hk.timed_out = l->ctrl + 2; }
        hk.signal_flag = l->ctrl + 0;
        if (int rc = enqueue_device(ctx, 0, p, l->rank, l->n_ranks, 0, l->base, 0, &hk)) return rc;
    }
    if (stats_out) {
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        return collect_stats(ctx, p, stats_out, t0, 1);
    }
    return GORT_OK;
}

int gort_read_radiance(gort_ctx* ctx, double* out, size_t bytes) {
    if (!ctx) return GORT_ERR_INVALID;
    if (ctx->last_w <= 0) return fail(ctx, GORT_ERR_INVALID, "no frame rendered yet");
    const int W = ctx->last_w, H = ctx->last_h;
    if (!out || bytes != (size_t)W * H * 3 * sizeof(double)) return fail(ctx, GORT_ERR_INVALID, "radiance buffer size mismatch");
    const int nd = (int)ctx->devs.size();
    const int tiles_x = (W + kTile - 1) / kTile;
    const double inv = 1.0 / ((double)(1u << kAccumFracBits) * (double)ctx->last_samples);
    for (int i = 0; i < nd; i++) {
        DeviceState& d = ctx->devs[i];
        const int n_local = ctx->last_local_tiles[i];
        if (n_local == 0) continue;
        CUDA_TRY(ctx, cudaSetDevice(d.dev));
        CUDA_TRY(ctx, cudaStreamSynchronize(stream_of(ctx, i)));
        std::vector<long long> h((size_t)n_local * kTilePixels * 3);
        CUDA_TRY(ctx, cudaMemcpy(h.data(), d.d_accum, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        std::vector<uint8_t> kept((size_t)n_local * 32);  // culled blocks never had their accumulators cleared: they are black
        CUDA_TRY(ctx, cudaMemcpy(kept.data(), reinterpret_cast<uint8_t*>(d.d_active + (size_t)n_local * 32), kept.size(), cudaMemcpyDeviceToHost));
        const int eff_count = ctx->last_count * nd, eff_rank = ctx->last_rank + i * ctx->last_count;
        for (int lt = 0; lt < n_local; lt++) {
            const int gt = eff_rank + lt * eff_count;
            const int tx = gt % tiles_x, ty = gt / tiles_x;
            for (int p = 0; p < kTilePixels; p++) {
                const int x = tx * kTile + (p & (kTile - 1)), y = ty * kTile + p / kTile;
                if (x >= W || y >= H) continue;
                const bool k = kept[(size_t)lt * 32 + ((p / kTile) >> 2) * 4 + ((p & (kTile - 1)) >> 3)] != 0;
                for (int c = 0; c < 3; c++) out[((size_t)y * W + x) * 3 + c] = k ? (double)h[((size_t)lt * kTilePixels + p) * 3 + c] * inv : 0.0;
            }
        }
    }
    return GORT_OK;
}

int gort_trace_rays(gort_ctx* ctx, int32_t n, const double* origins3, const double* directions3, double t_min, double t_max,
                    int32_t any_hit, double* out_t, int32_t* out_order) {
    if (!ctx) return GORT_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GORT_ERR_NO_SCENE, "no scene uploaded");
    if (n < 0 || (n > 0 && (!origins3 || !directions3 || !out_t || !out_order))) return fail(ctx, GORT_ERR_INVALID, "gort_trace_rays: bad arguments");
    if (n == 0) return GORT_OK;
    DeviceState& d = ctx->devs[0];
    CUDA_TRY(ctx, cudaSetDevice(d.dev));
    cudaStream_t st = stream_of(ctx, 0);
    std::vector<float> ho(3 * (size_t)n), hd(3 * (size_t)n);
    for (size_t i = 0; i < 3 * (size_t)n; i++) {
        ho[i] = (float)origins3[i];
        hd[i] = (float)directions3[i];
    }
    float *d_o = nullptr, *d_d = nullptr, *d_t = nullptr;
    int* d_ord = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_o, ho.size() * 4));
    CUDA_TRY(ctx, cudaMalloc(&d_d, hd.size() * 4));
    CUDA_TRY(ctx, cudaMalloc(&d_t, (size_t)n * 4));
    CUDA_TRY(ctx, cudaMalloc(&d_ord, (size_t)n * 4));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_o, ho.data(), ho.size() * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_d, hd.data(), hd.size() * 4, cudaMemcpyHostToDevice, st));
    SceneView sv;
    sv.nodes = d.d_nodes; sv.spheres = d.d_spheres; sv.sphere_meta = d.d_meta; sv.tris = d.d_tris; sv.mats = d.d_mats; sv.lights = d.d_lights;
    sv.n_nodes = ctx->bvh.n_nodes; sv.n_lights = (int)ctx->scene.lights.size();
    sv.n_spheres = (int)ctx->scene.spheres.size(); sv.n_tris = (int)ctx->scene.tris.size();
    sv.qox = sv.qoy = sv.qoz = 0.f; sv.qcx = sv.qcy = sv.qcz = 1.f;  // (the test hook walks the fp32 nodes)
    const float tmax_f = std::isinf(t_max) ? INFINITY : (float)t_max;
    CUDA_TRY(ctx, launch_trace_rays(sv, n, d_o, d_d, (float)t_min, tmax_f, any_hit, d_t, d_ord, st));
    std::vector<float> ht(n);
    CUDA_TRY(ctx, cudaMemcpyAsync(ht.data(), d_t, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_order, d_ord, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++) out_t[i] = ht[i];
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_t); cudaFree(d_ord);
    return GORT_OK;
}

int gort_measure_fp32_peak(gort_ctx* ctx, double* tflops_out, double* ms_out) {
    if (!ctx || !tflops_out) return GORT_ERR_INVALID;
    DeviceState& d = ctx->devs[0];
    CUDA_TRY(ctx, cudaSetDevice(d.dev));
    cudaStream_t st = stream_of(ctx, 0);
    float* sink = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&sink, 16));
    const int blocks = d.sm_count * 8, threads = 256, iters = 8192;
    CUDA_TRY(ctx, launch_ffma_peak(sink, 256, blocks, threads, st));  // warm-up
    double best = 1e30;
    for (int rep = 0; rep < 3; rep++) {
        CUDA_TRY(ctx, cudaEventRecord(d.ev[0], st));
        CUDA_TRY(ctx, launch_ffma_peak(sink, iters, blocks, threads, st));
        CUDA_TRY(ctx, cudaEventRecord(d.ev[1], st));
        CUDA_TRY(ctx, cudaEventSynchronize(d.ev[1]));
        float ms = 0;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
        best = std::min(best, (double)ms);
    }
    cudaFree(sink);
    const double flops = (double)blocks * threads * (double)iters * 64.0 * 2.0;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return GORT_OK;
}

}  // extern "C"
