// kernels.cu — sm_100a kernels of libgort: the per-pixel render hot path of
// concurrent-raytracer-go (/root/reference internal/renderer/renderer.go:150-390) as an iterative
// wavefront executed by persistent warps.
//
//   trace_kernel   : persistent warps pull (tile, 8x4 pixel block, sample batch) work units from an
//                    atomic counter and run three decoupled stages over two per-warp queues in shared
//                    memory (see the comment above the kernel): FILL (primary rays), EXTEND (Scatter +
//                    scattered ray; the recursion of traceRay as a throughput chain) and SHADE
//                    (calculateDirectLighting: hard shadow rays, cone-culled soft-shadow rays, shading
//                    arithmetic).  Every shaded hit adds T * (emitted + w * direct) to per-pixel int64
//                    fixed-point accumulators (order independent => the image is bit-reproducible for
//                    any schedule / GPU count).
//   resolve_kernel : toneMap + ToRGB + img.Set (renderer.go:92-97,348-367) in float64 from the exact
//                    accumulator, packed as RGBA8 (row-major frame or tile-major shard slab).
//
// Arithmetic is fp32 (the reference is float64); RNG is counter-based Philox4x32-10 keyed on
// (pixel, sample, bounce, purpose) so any work-to-lane assignment draws the same numbers.
#include "device.cuh"

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <vector>

namespace gort {

// ---------------------------------------------------------------------------------------------
// per-warp state in shared memory: two queues (structure of arrays, one column per entry)
//   PQ  paths that hit something and wait for Material.Scatter + the scattered ray   (EXTEND)
//   SQ  hit records that wait for calculateDirectLighting                             (SHADE)
// and the scratch of the shade round.
// ---------------------------------------------------------------------------------------------
constexpr int kWarpsPerCta = 4;
// Queue capacity.  Shared memory is what limits the resident warps (64 registers per thread would allow 32
// warps per SM): 48 entries instead of 64 brings the tiny-scene kernel from 6 to 8 CTAs per SM.  Invariants:
// FILL (pushes <= 32 paths) only runs with pqn <= kQueueCap - 32; EXTEND pops n <= 32 paths and pushes n
// records, and only runs with sqn + n <= kQueueCap (SHADE makes room first).
constexpr int kQueueCap = 48;
constexpr int kLightChunkMax = 4;  // lights handled per pass of a shade round
enum PField { PF_PX, PF_PY, PF_PZ, PF_DX, PF_DY, PF_DZ, PF_PRIM, PF_TR, PF_TG, PF_TB, PF_PIXG, PF_PIXL, PF_SD, PF_FOG, PF_COUNT };
enum SField { SF_PX, SF_PY, SF_PZ, SF_NX, SF_NY, SF_NZ, SF_TR, SF_TG, SF_TB, SF_MAT, SF_PIXG, SF_PIXL, SF_SD, SF_FOG, SF_COUNT };
// *_SD = sample | depth << 16;  *_FOG = fog factor of the path's primary hit (extension)

template <bool SMALL>
struct WarpShared {
    static constexpr int kLightChunk = SMALL ? kSmallLights : kLightChunkMax;
    uint32_t pq[PF_COUNT][kQueueCap];
    uint32_t sq[SF_COUNT][kQueueCap];
    uint8_t lit[kLightChunk][32];        // hard shadow ray unoccluded
    uint8_t cnt[kLightChunk][32];        // unoccluded soft shadow rays (of 16)
    uint16_t pairs[kLightChunk * 32];    // lit (light, item) pairs of the chunk: (light << 8) | item
    uint16_t cmask[SMALL ? kLightChunk : 1][32];  // tiny scenes: spheres the pair's shadow cone can reach
    uint8_t ncand[SMALL ? 4 : 32];                // BVH scenes: candidates of the 32 pairs in flight (or kCandOverflow)
    uint8_t sel[SMALL ? 4 : 32];                  // BVH scenes: slots of the pairs in flight that have candidates
    uint32_t cand[SMALL ? 1 : 32][kMaxCand];
};

__device__ __forceinline__ float qf(const uint32_t (*Q)[kQueueCap], int f, int slot) { return __uint_as_float(Q[f][slot]); }


// ---------------------------------------------------------------------------------------------
// the trace kernel
//
// traceRay (renderer.go:165-227) returns  emitted + w_d*direct + w_r*atten*traceRay(scattered): affine in
// the recursive term.  Unrolled, a sample's radiance is  sum_k T_k * (emitted_k + w_k*direct_k)  with the
// throughput T_k = prod_{i<k} w_r,i*atten_i, and T_k depends only on the chain of Scatter calls — not on
// any lighting.  So each warp runs two decoupled stages over its queues:
//   FILL    32 primary rays (tracePixel/getRay + hitWorld); hits enter PQ, misses are black and done;
//   EXTEND  <= 32 paths of PQ: hit record + Material.Scatter + the scattered ray's hitWorld.  Every path
//           leaves one hit record (point, normal, material, T_k) in SQ; survivors return to PQ.  A
//           50-bounce glass path is a chain of these short rounds and never waits for a shadow ray;
//   SHADE   32 records of SQ: one packed batch of hard shadow rays for every (record, light), then the 16
//           soft-shadow rays of every lit pair (a quarter warp per pair, two rays per lane, candidates
//           pre-culled with the pair's cone), then calculateDirectLighting's arithmetic and ONE
//           fixed-point add of T_k * (...) to the pixel.
// ---------------------------------------------------------------------------------------------
// resident CTAs per SM the register allocation must allow: 8 (64 registers) for the tiny-scene kernel; 7 (72) for
// the BVH kernel.  Its walk is bound by per-warp latency (node fetch + dependent slab arithmetic): measured on the
// 100 k / 1 M primitive scenes and the 40-triangle scene, 7 CTAs beat 6 by 5 / 8 / 4 % although the walk then spills
// ~70 bytes, and 8 is no better than 7.  Shared memory (27 KB per CTA: 6 candidates per pair, 4 lights per pass)
// is sized so that 7-8 CTAs fit.
// SKY (BVH variants only): the sky extension's two call sites are compiled in only for scenes that enable it — the kernel is
// bound by its code footprint (the same sites, present but never executed, cost C2-view 8 %)
// TraceParams::early_out: black (alpha 255) for every pixel of the blocks the cull pass dropped — what resolve_kernel writes for
// them (part 1), but from inside the trace kernel: the frame of a sparse scene is mostly such blocks, and a frame in page-locked
// host memory takes them over PCIe while the kept blocks are still being traced.  One warp per participating CTA; a 32-pixel tile
// row is one 128-byte store.
static __device__ __noinline__ void fill_culled_rows(const TraceParams& P, int first, int stride) {
    const int lane = threadIdx.x & 31;
    uchar4* out = reinterpret_cast<uchar4*>(P.early_out);
    for (int seg = first; seg < P.n_local_tiles * kTile; seg += stride) {
        const int lt = seg / kTile, ly = seg % kTile;
        if (P.block_active[lt * 32 + (ly >> 2) * 4 + (lane >> 3)]) continue;
        const int gt = P.shard_rank + lt * P.shard_count;
        const int x = (gt % P.tiles_x) * kTile + lane, y = (gt / P.tiles_x) * kTile + ly;
        if (x < P.width && y < P.height) out[(size_t)y * P.width + x] = make_uchar4(0, 0, 0, 255);
    }
}

template <bool STATS, bool SMALL, int GEOM, bool SKY = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, SMALL ? 8 : 7) trace_kernel(const __grid_constant__ TraceParams P) {
    __shared__ WarpShared<SMALL> wsh[kWarpsPerCta];
    const int lane = threadIdx.x & 31;
    WarpShared<SMALL>& W = wsh[threadIdx.x >> 5];
    uint32_t(*PQ)[kQueueCap] = W.pq;
    uint32_t(*SQ)[kQueueCap] = W.sq;
    const SceneView& S = P.scene;
    Stats st;
    if (STATS) {
#pragma unroll
        for (int i = 0; i < kStatCount; i++) st.v[i] = 0;
    }

    // Programmatic dependent launch (launch_trace(dependent)): this grid's CTAs may have become resident while the cull pass
    // was still running, and resolve_kernel's CTAs may take the slots this grid's CTAs leave.  Nothing of the cull pass is
    // read before the wait; a launch without the attribute passes straight through both instructions.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (P.stamps && blockIdx.x == 0 && threadIdx.x == 0)
        asm volatile("{ .reg .u64 t; mov.u64 t, %%globaltimer; st.global.u64 [%0], t; }" ::"l"(P.stamps + 1) : "memory");
    if (P.early_out && (blockIdx.x & 7u) == 0 && threadIdx.x < 32) fill_culled_rows(P, (int)(blockIdx.x >> 3), (int)((gridDim.x + 7u) >> 3));
    // Work units are sized on the device from the number of pixel blocks the cull pass kept:
    // enough units for dynamic balance (target_units), at most 16 samples each.
    const uint32_t n_deep = P.active_count[0], n_norm = P.active_count[1];
    const uint32_t n_active = n_deep + n_norm;
    if (n_active == 0) return;
    int spu;
    {
        const uint32_t want_batches = (P.target_units + n_active - 1) / n_active;
        spu = max(1, min(16, P.samples / (int)max(1u, want_batches)));
    }
    const uint32_t n_batches = (uint32_t)((P.samples + spu - 1) / spu);
    const uint32_t n_units = n_active * n_batches, deep_units = n_deep * n_batches;

    // per-warp timeline, compiled in only with -DGORT_DEBUG (lib/libgort_dbg.so; tools/debug_times.py)
    unsigned long long t_units_done = 0, dbg_rounds = 0, dbg_paths = 0, dbg_shades = 0, dbg_t_ext = 0, dbg_t_shade = 0, dbg_t0 = 0;
    // anatomy of the EXTEND rounds after the units ran out (cycles): queue read + hit record + materials / Scatter / hitWorld / rest
    unsigned long long dbg_sec[4] = {0, 0, 0, 0}, dbg_tp = 0;
#define GORT_DBG_SECTION(k)                                   \
    if (kDbg && P.debug_times && !more_units) {               \
        const unsigned long long t__ = clock64();             \
        dbg_sec[k] += t__ - dbg_tp;                            \
        dbg_tp = t__;                                          \
    }
    if (kDbg && P.debug_times && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(P.debug_times, t);
    }

    int pqn = 0, sqn = 0;    // warp-uniform queue sizes
    bool more_units = true;  // warp-uniform
    bool urgent = false;     // warp-uniform: PQ's top holds a path at depth >= urgent_depth
    int s_cur = 0, s_end = 0;
    // this lane's pixel in the current work unit: (y << 16) | x, local accumulator index
    uint32_t pix_xy = 0, pixl = 0;
    bool lane_valid = false, jit_valid = false;
    uint32_t jit_z = 0, jit_w = 0;
    const unsigned lt_mask = (1u << lane) - 1u;
    constexpr int kLightChunk = WarpShared<SMALL>::kLightChunk;

    for (;;) {
        const bool can_fill = (s_cur < s_end) || more_units;
        int action;  // 0 FILL, 1 EXTEND, 2 SHADE
        if (sqn >= 32) action = 2;
        else if (pqn > kQueueCap - 32 || (pqn > 0 && (urgent || !can_fill)))
            action = (sqn + min(32, pqn) > kQueueCap) ? 2 : 1;  // EXTEND, after SHADE has made room for its records
        else if (can_fill) action = 0;
        else if (sqn > 0) action = 2;
        else break;

        if (action == 0) {
            // ================= FILL: primary rays (tracePixel renderer.go:150-163, getRay :377-390) =========
            if (s_cur >= s_end) {
                uint32_t u = 0;
                if (lane == 0) u = atomicAdd(P.work_counter, 1u);
                u = __shfl_sync(FULL_MASK, u, 0);
                if (u >= n_units) {
                    more_units = false;
                    if (kDbg && P.debug_times) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_units_done));
                    continue;
                }
                // unit = (sample batch, active 8x4 block), batch-major; all units of the blocks that can see
                // glass (deep paths) come first, the other blocks follow
                uint32_t batch, idx;
                if (u < deep_units) {
                    batch = u / n_deep;
                    idx = u - batch * n_deep;
                } else {
                    const uint32_t v = u - deep_units;
                    batch = v / n_norm;
                    idx = P.n_local_tiles * 32u - 1u - (v - batch * n_norm);  // the normal list grows down from the end
                }
                const uint32_t packed = __ldg(P.active_list + idx);
                const uint32_t ltile = packed >> 5, block = packed & 31u;
                const uint32_t gtile = (uint32_t)P.shard_rank + ltile * (uint32_t)P.shard_count;
                const uint32_t tx = gtile % (uint32_t)P.tiles_x, ty = gtile / (uint32_t)P.tiles_x;
                const uint32_t lx = ((block & 3u) << 3) + (lane & 7u), ly = ((block >> 2) << 2) + (lane >> 3);
                const uint32_t x = tx * kTile + lx, y = ty * kTile + ly;
                lane_valid = (x < (uint32_t)P.width) && (y < (uint32_t)P.height);
                pix_xy = (y << 16) | x;
                pixl = ltile * kTilePixels + ly * kTile + lx;
                s_cur = (int)batch * spu;
                s_end = min(s_cur + spu, P.samples);
                jit_valid = false;
            }
            bool hit = false;
            float t = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
            int prim = 0;
            const uint32_t x = pix_xy & 0xffffu, y = pix_xy >> 16;
            const uint32_t pixg = y * (uint32_t)P.width + x;
            if (lane_valid) {
                stat_add<STATS>(st, kStatPrimary);
                float ju = 0.5f, jv = 0.5f;
                if (P.jitter) {
                    // one Philox block serves two consecutive samples: (x,y) the even one, (z,w) the odd one
                    uint32_t jx, jy;
                    if ((s_cur & 1) && jit_valid) {
                        jx = jit_z; jy = jit_w;
                    } else {
                        const uint4 r = philox_at<SMALL>(P.rk, pixg, (uint32_t)s_cur >> 1, kStreamJitter, 0u);
                        stat_add<STATS>(st, kStatRngBlocks);
                        jit_z = r.z; jit_w = r.w;
                        jx = (s_cur & 1) ? r.z : r.x;
                        jy = (s_cur & 1) ? r.w : r.y;
                    }
                    jit_valid = !(s_cur & 1);
                    ju = (float)(jx >> 8) * (1.0f / 16777216.0f);
                    jv = (float)(jy >> 8) * (1.0f / 16777216.0f);
                }
                const float u = ((float)x + ju) * P.inv_w, v = ((float)y + jv) * P.inv_h;
                dx = fmaf(v, P.cam.vx, fmaf(u, P.cam.hx, P.cam.llx));
                dy = fmaf(v, P.cam.vy, fmaf(u, P.cam.hy, P.cam.lly));
                dz = fmaf(v, P.cam.vz, fmaf(u, P.cam.hz, P.cam.llz));
                // traceRay depth 0 (renderer.go:166-173); max_depth <= 0 returns black before any hit test
                if (P.max_depth > 0)
                    {
                        const unsigned int nv0 = STATS ? st.v[kStatNodes] : 0u;
                        hit = query<STATS, SMALL, GEOM>(P, P.cam.ox, P.cam.oy, P.cam.oz, dx, dy, dz, 0.001f, FLT_MAX * 2.0f, false, t, prim, st);
                        walk_account<STATS>(st, nv0, 0);
                    }
            }
            if (SKY && lane_valid && !hit && P.max_depth > 0) {
                // sky extension: a primary ray that leaves the scene sees the sky (the reference returns black, renderer.go:171-173)
                const float3 c = sky_color(P.sky, dx, dy, dz);
                add_radiance(P, pixl, c.x, c.y, c.z);
            }
            const unsigned hm = __ballot_sync(FULL_MASK, hit);
            if (hit) {
                const int slot = pqn + __popc(hm & lt_mask);
                PQ[PF_PX][slot] = __float_as_uint(fmaf(t, dx, P.cam.ox));  // rec.Point = ray.At(t) (sphere.go:43, triangle.go:69)
                PQ[PF_PY][slot] = __float_as_uint(fmaf(t, dy, P.cam.oy));
                PQ[PF_PZ][slot] = __float_as_uint(fmaf(t, dz, P.cam.oz));
                PQ[PF_DX][slot] = __float_as_uint(dx); PQ[PF_DY][slot] = __float_as_uint(dy); PQ[PF_DZ][slot] = __float_as_uint(dz);
                PQ[PF_PRIM][slot] = (uint32_t)prim;
                PQ[PF_TR][slot] = PQ[PF_TG][slot] = PQ[PF_TB][slot] = __float_as_uint(1.0f);
                PQ[PF_PIXG][slot] = pixg; PQ[PF_PIXL][slot] = pixl; PQ[PF_SD][slot] = (uint32_t)s_cur;
                float fog = 0.f;
                if (P.fog_enabled) {  // extension: exponential fog on the primary-hit distance
                    const float dist = t * sqrt_fast(dot3(dx, dy, dz, dx, dy, dz));
                    fog = 1.0f - expf(-P.fog_density * dist);
                }
                PQ[PF_FOG][slot] = __float_as_uint(fog);
            }
            pqn += __popc(hm);
            s_cur++;
            __syncwarp();
            continue;
        }

        if (action == 1) {
            // ================= EXTEND: hit record + Material.Scatter + scattered ray (lane = path) ============
            const int n = min(32, pqn);
            const int base = pqn - n;
            const bool act = lane < n;
            const int slot = base + (act ? lane : 0);
            if (kDbg && P.debug_times && !more_units) {
                dbg_rounds++;
                dbg_paths += n;
                dbg_t0 = clock64();
                dbg_tp = dbg_t0;
            }
            bool survive = false;
            float px = 0.f, py = 0.f, pz = 0.f, sx = 0.f, sy = 0.f, sz = 0.f, tr = 0.f, tg = 0.f, tb = 0.f, fog = 0.f;
            uint32_t pixg = 0, pixl2 = 0, sd = 0;
            int prim2 = 0;
            if (act) {
                stat_add<STATS>(st, kStatShaded);
                px = qf(PQ, PF_PX, slot); py = qf(PQ, PF_PY, slot); pz = qf(PQ, PF_PZ, slot);
                const float dx = qf(PQ, PF_DX, slot), dy = qf(PQ, PF_DY, slot), dz = qf(PQ, PF_DZ, slot);
                const int prim = (int)PQ[PF_PRIM][slot];
                tr = qf(PQ, PF_TR, slot); tg = qf(PQ, PF_TG, slot); tb = qf(PQ, PF_TB, slot);
                pixg = PQ[PF_PIXG][slot]; pixl2 = PQ[PF_PIXL][slot]; sd = PQ[PF_SD][slot];
                fog = qf(PQ, PF_FOG, slot);
                uint4 rnd_early = make_uint4(0u, 0u, 0u, 0u);
                if (SMALL) rnd_early = philox(P.rk, pixg, sd & 0xffffu, ((sd >> 16) << 8) | kStreamScatter, 0u);
                // hit record (sphere.go:42-50, triangle.go:69-73)
                float nx, ny, nz;
                int mat;
                if (SMALL) {
                    const float4 s = P.small_sph[prim];
                    const float inv_r = rcp_fast(s.w);
                    nx = (px - s.x) * inv_r; ny = (py - s.y) * inv_r; nz = (pz - s.z) * inv_r;
                    mat = P.small_mat[prim];
                } else if (GEOM == 1 || (GEOM == 3 && prim >= 0)) {
                    const float4 s = ldg4(S.spheres + prim);
                    const float inv_r = rcp_fast(s.w);
                    nx = (px - s.x) * inv_r; ny = (py - s.y) * inv_r; nz = (pz - s.z) * inv_r;
                    mat = __ldg(&S.sphere_meta[prim]).x;
                } else {
                    const float4* tp = S.tris + 4 * (size_t)(prim & 0x7fffffff);
                    mat = __float_as_int(ldg4(tp).w);
                    const float4 nn = ldg4(tp + 3);
                    nx = nn.x; ny = nn.y; nz = nn.z;
                }
                const float ddn0 = dot3(dx, dy, dz, nx, ny, nz);
                const bool front = ddn0 < 0.f;
                if (!front) { nx = -nx; ny = -ny; nz = -nz; }
                const float ddn = front ? ddn0 : -ddn0;  // ray.Direction . normal (<= 0)

                // the hit record goes to the shade queue with the throughput it is seen through
                {
                    const int ss = sqn + lane;
                    SQ[SF_PX][ss] = __float_as_uint(px); SQ[SF_PY][ss] = __float_as_uint(py); SQ[SF_PZ][ss] = __float_as_uint(pz);
                    SQ[SF_NX][ss] = __float_as_uint(nx); SQ[SF_NY][ss] = __float_as_uint(ny); SQ[SF_NZ][ss] = __float_as_uint(nz);
                    SQ[SF_TR][ss] = __float_as_uint(tr); SQ[SF_TG][ss] = __float_as_uint(tg); SQ[SF_TB][ss] = __float_as_uint(tb);
                    SQ[SF_MAT][ss] = (uint32_t)mat; SQ[SF_PIXG][ss] = pixg; SQ[SF_PIXL][ss] = pixl2; SQ[SF_SD][ss] = sd;
                    SQ[SF_FOG][ss] = __float_as_uint(fog);
                }

                const float4 m0 = mat4<SMALL>(P, mat, 0), m1 = mat4<SMALL>(P, mat, 1), m2 = mat4<SMALL>(P, mat, 2), m3 = mat4<SMALL>(P, mat, 3);
                const int mtype = __float_as_int(m0.x);
                GORT_DBG_SECTION(0)
                const uint32_t depth = sd >> 16, sample = sd & 0xffffu;
                const uint32_t bs = (depth << 8) | kStreamScatter;
                bool scattered = true;
                float ar = 0.f, ag = 0.f, ab = 0.f;
                // Scatter draws at most one Philox block per hit, counter (pixel, sample, bounce|scatter, 0):
                // Lambertian and rough Metal/Shiny/Mirror turn it into a ball point, Glass/Dielectric use word 0
                const bool rough = (mtype == 2) ? (m1.x > 0.f) : (m1.x > 0.001f);
                uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
                const bool draws = mtype == 0 || (mtype <= 3 && rough) || mtype == 4 || mtype == 5;
                if (SMALL) {
                    // drawn for every hit, up front: the block depends on nothing but the path's counters, so its ten
                    // dependent rounds overlap the queue reads, the hit record and the material fetch instead of following them
                    // (the tail of a frame is single glass paths, one dependent chain per round; profiles/README.md)
                    rnd = rnd_early;
                } else if (draws) {
                    rnd = philox_at<SMALL>(P.rk, pixg, sample, bs, 0u);
                }
                if (draws) stat_add<STATS>(st, kStatRngBlocks);
                if (mtype == 0) {  // Lambertian (material.go:26-35)
                    float bx, by, bz;
                    ball_from_block(rnd, bx, by, bz);
                    sx = nx + bx; sy = ny + by; sz = nz + bz;
                    if (fabsf(sx) < 1e-8f && fabsf(sy) < 1e-8f && fabsf(sz) < 1e-8f) { sx = nx; sy = ny; sz = nz; }
                    normalize3(sx, sy, sz);
                    ar = m0.y; ag = m0.z; ab = m0.w;
                } else if (mtype <= 3) {  // Metal / Shiny / PerfectMirror (material.go:75-113,169-189; advanced_materials.go:125-144)
                    sx = fmaf(-2.0f * ddn, nx, dx); sy = fmaf(-2.0f * ddn, ny, dy); sz = fmaf(-2.0f * ddn, nz, dz);  // Reflect vector.go:77
                    if (rough) {
                        float bx, by, bz;
                        ball_from_block(rnd, bx, by, bz);
                        sx = fmaf(m1.x, bx, sx); sy = fmaf(m1.x, by, sy); sz = fmaf(m1.x, bz, sz);
                        normalize3(sx, sy, sz);
                    }
                    const float cosT = fabsf(ddn);  // ray direction is NOT normalised here (material.go:85)
                    const float fres = fmaf(1.0f - m3.y, pow5(1.0f - cosT), m3.y);
                    const float fs = m3.z;
                    ar = fmaf(m0.y, 1.0f - fs, fres * fs); ag = fmaf(m0.z, 1.0f - fs, fres * fs); ab = fmaf(m0.w, 1.0f - fs, fres * fs);
                    if (mtype == 1) {
                        ar = fmaxf(0.f, fminf(1.f, ar)); ag = fmaxf(0.f, fminf(1.f, ag)); ab = fmaxf(0.f, fminf(1.f, ab));
                        if (m3.w >= 0.f) {  // metallic > 0.8 (material.go:102-109)
                            const float mf = m3.w;
                            ar = fmaf(ar, 1.0f - mf, fres * mf); ag = fmaf(ag, 1.0f - mf, fres * mf); ab = fmaf(ab, 1.0f - mf, fres * mf);
                        }
                    } else if (mtype == 2) {
                        ar = fminf(1.f, ar); ag = fminf(1.f, ag); ab = fminf(1.f, ab);
                    }
                } else if (mtype <= 5) {  // Glass / Dielectric (advanced_materials.go:21-46; material.go:235-260)
                    ar = m0.y; ag = m0.z; ab = m0.w;  // Glass colour; Dielectric packed as (1,1,1)
                    const float ratio = front ? m3.z : m1.w;  // 1/ior precomputed in float64 on the host
                    float ux = dx, uy = dy, uz = dz;
                    normalize3(ux, uy, uz);
                    const float udn = dot3(ux, uy, uz, nx, ny, nz);
                    const float cosT = fminf(-udn, 1.0f);
                    const float sinT = sqrt_fast(fmaf(-cosT, cosT, 1.0f));
                    bool reflect = ratio * sinT > 1.0f;  // cannotRefract
                    if (!reflect) {
                        const float r0 = m3.y;  // ((1-x)/(1+x))^2 is the same for x = ior and x = 1/ior
                        const float refl = fmaf(1.0f - r0, pow5(1.0f - cosT), r0);  // reflectance material.go:282-286
                        reflect = refl > (float)(rnd.x >> 8) * (1.0f / 16777216.0f);
                    }
                    if (reflect) {
                        sx = fmaf(-2.0f * udn, nx, ux); sy = fmaf(-2.0f * udn, ny, uy); sz = fmaf(-2.0f * udn, nz, uz);
                    } else {
                        // Vec3.Refract (vector.go:81-96) with v = unit direction, normal against the ray
                        float cn = udn, eta = ratio, rnx = nx, rny = ny, rnz = nz;
                        if (cn > 0.f) { rnx = -nx; rny = -ny; rnz = -nz; eta = rcp_fast(eta); cn = -cn; }
                        // explicit roundings: left to the compiler, the two kernels that share this code contracted different products
                        // into fused multiply-adds and drifted apart by an ulp per refraction
                        const float sin2 = __fmul_rn(eta * eta, fmaf(-cn, cn, 1.0f));
                        if (sin2 > 1.0f) {
                            const float d2 = dot3(ux, uy, uz, rnx, rny, rnz);
                            sx = fmaf(-2.0f * d2, rnx, ux); sy = fmaf(-2.0f * d2, rny, uy); sz = fmaf(-2.0f * d2, rnz, uz);
                        } else {
                            const float k = fmaf(eta, cn, sqrt_fast(__fsub_rn(1.0f, sin2)));
                            sx = fmaf(eta, ux, -k * rnx); sy = fmaf(eta, uy, -k * rny); sz = fmaf(eta, uz, -k * rnz);
                        }
                    }
                } else {  // DiffuseLight (material.go:296-298): no scatter
                    scattered = false;
                }
                // traceRay(scattered, depth+1) is black at once when depth+1 >= maxDepth or when
                // recursiveReflections is off (renderer.go:166-168,186-189): no ray needed
                bool cont = scattered && P.recursive && (int)(depth + 1) < P.max_depth;
                const float wr = m2.z;
                tr *= ar * wr; tg *= ag * wr; tb *= ab * wr;
                // Exact dead-path test.  Everything the remaining bounces can add reaches the pixel as
                // fixed-point adds of T * c with |c| <= dead_bound.  If T * dead_bound < 2^-31 in every
                // channel, each add rounds to zero: tracing on cannot change the accumulator.  (Rays trapped
                // inside a rough-metal sphere otherwise bounce to max_depth with throughput ~0.03^k.)
                const float db = P.dead_bound;
                if (db > 0.f && fabsf(tr) * db < 4.6566e-10f && fabsf(tg) * db < 4.6566e-10f && fabsf(tb) * db < 4.6566e-10f) cont = false;
                GORT_DBG_SECTION(1)
                if (cont) {
                    float t2;
                    {
                        const unsigned int nv0 = STATS ? st.v[kStatNodes] : 0u;
                        if (SMALL) {
                            // A scattered ray that heads INTO the sphere it starts on (refraction, internal reflection) stays inside
                            // it up to its exit point: only that sphere and the ones overlapping it can be the closest hit
                            // (TraceParams::small_inside).  The 50-bounce chains of a glass sphere scan one sphere instead of all.
                            // The ray leaves the sphere again at t = -2 g / |s|^2: four times tMin at least, so that the reference
                            // accepts that root too (sphere.go:35-40) and nothing beyond it can be the closest hit.
                            const float4 s0 = P.small_sph[prim];
                            const float g = dot3(sx, sy, sz, px - s0.x, py - s0.y, pz - s0.z);
                            const float s2 = dot3(sx, sy, sz, sx, sy, sz);
                            const bool inward = g < -0.002f * s2 && !P.no_cone_cull;
                            const uint32_t mask = inward ? (uint32_t)P.small_inside[prim] : ((1u << P.small_n) - 1u);
                            survive = small_query<STATS, true>(P, mask, px, py, pz, sx, sy, sz, 0.001f, FLT_MAX * 2.0f, false, t2, prim2, st);
                        } else {
                            survive = query<STATS, SMALL, GEOM>(P, px, py, pz, sx, sy, sz, 0.001f, FLT_MAX * 2.0f, false, t2, prim2, st);
                        }
                        walk_account<STATS>(st, nv0, 1);
                    }
                    if (survive) {
                        px = fmaf(t2, sx, px); py = fmaf(t2, sy, py); pz = fmaf(t2, sz, pz);
                        sd += 0x10000u;
                    } else if (SKY) {
                        // sky extension: traceRay(scattered) = sky colour; it reaches the pixel through the updated throughput
                        // (and the primary hit's fog factor)
                        const float3 c = sky_color(P.sky, sx, sy, sz);
                        const float k = P.fog_enabled ? 1.0f - fog : 1.0f;
                        add_radiance(P, pixl2, tr * c.x * k, tg * c.y * k, tb * c.z * k);
                    }
                }
                GORT_DBG_SECTION(2)
                if (STATS && !survive) {
                    const uint32_t dd = depth + (scattered ? 1u : 0u);
                    if (dd >= 5) stat_add<STATS>(st, kStatDepth5);
                    if (dd >= 20) stat_add<STATS>(st, kStatDepth20);
                    if ((int)dd >= P.max_depth) stat_add<STATS>(st, kStatDepthMax);
                }
            }
            __syncwarp();  // every lane has read its slot before the survivors are compacted over the popped region
            const unsigned hm = __ballot_sync(FULL_MASK, survive);
            if (survive) {
                const int d = base + __popc(hm & lt_mask);
                PQ[PF_PX][d] = __float_as_uint(px); PQ[PF_PY][d] = __float_as_uint(py); PQ[PF_PZ][d] = __float_as_uint(pz);
                PQ[PF_DX][d] = __float_as_uint(sx); PQ[PF_DY][d] = __float_as_uint(sy); PQ[PF_DZ][d] = __float_as_uint(sz);
                PQ[PF_PRIM][d] = (uint32_t)prim2;
                PQ[PF_TR][d] = __float_as_uint(tr); PQ[PF_TG][d] = __float_as_uint(tg); PQ[PF_TB][d] = __float_as_uint(tb);
                PQ[PF_PIXG][d] = pixg; PQ[PF_PIXL][d] = pixl2; PQ[PF_SD][d] = sd; PQ[PF_FOG][d] = __float_as_uint(fog);
            }
            pqn = base + __popc(hm);
            sqn += n;
            urgent = P.urgent_depth > 0 && __any_sync(FULL_MASK, survive && (int)(sd >> 16) >= P.urgent_depth);
            __syncwarp();
            GORT_DBG_SECTION(3)
            if (kDbg && P.debug_times && !more_units) dbg_t_ext += clock64() - dbg_t0;
            continue;
        }

        // ================= SHADE: calculateDirectLighting for the top n <= 32 records of SQ =================
        const int n = min(32, sqn);
        const int base = sqn - n;
        sqn = base;
        const bool act = lane < n;
        const int slot = base + (act ? lane : 0);
        const float inv_n = 1.0f / (float)n;
        if (kDbg && P.debug_times && !more_units) {
            dbg_shades++;
            dbg_t0 = clock64();
        }
        // lane = record: material constants and the running total (starts at the ambient term, renderer.go:236-246)
        float dr = 0.f, dg = 0.f, db = 0.f;
        float kar = 0.f, kag = 0.f, kab = 0.f, spec_pow = 0.f, spec_w = 0.f;
        if (act) {
            const int mat = (int)SQ[SF_MAT][slot];
            const float4 m0 = mat4<SMALL>(P, mat, 0), m2 = mat4<SMALL>(P, mat, 2);
            const bool is_light = __float_as_int(m0.x) == 6;
            // GetAlbedo: DiffuseLight -> 0 (material.go:304); Dielectric -> 1 (packed by the host)
            kar = is_light ? 0.f : m0.y * m2.y; kag = is_light ? 0.f : m0.z * m2.y; kab = is_light ? 0.f : m0.w * m2.y;
            dr = dg = db = m2.x;
            spec_pow = mat4<SMALL>(P, mat, 3).x;       // 0: metallic <= 0.5, no specular term
            spec_w = mat4<SMALL>(P, mat, 1).y * 3.0f;  // metallic * 3
        }
        const int nl = S.n_lights;
        for (int l0 = 0; l0 < nl; l0 += kLightChunk) {
            const int lc = min(kLightChunk, nl - l0);
            // ---- B: the hard shadow ray of every (record, light), light-major, 32 per step ----
            const int n_rays = n * lc;
            for (int r0 = 0; r0 < n_rays; r0 += 32) {
                const int r = r0 + lane;
                const bool valid = r < n_rays;
                const int li = valid ? (int)(((float)r + 0.5f) * inv_n) : 0;
                const int j = valid ? r - li * n : 0;
                const int sj = base + j;
                const float ox = qf(SQ, SF_PX, sj), oy = qf(SQ, SF_PY, sj), oz = qf(SQ, SF_PZ, sj);
                const float4 L0 = light4<SMALL>(P, l0 + li, 0);
                float dx = L0.x - ox, dy = L0.y - oy, dz = L0.z - oz;
                const float dist2 = dot3(dx, dy, dz, dx, dy, dz);
                const float inv_d = dist2 > 0.f ? rsqrt_fast(dist2) : 0.f;
                const float dist = dist2 * inv_d;  // lightDistance
                dx *= inv_d; dy *= inv_d; dz *= inv_d;
                // cosTheta = max(0, hit.Normal . lightDir) (renderer.go:259): at 0 the light's `intensity` is 0 and with it both
                // the diffuse and the specular term (renderer.go:260-287) — the shadow factor of such a pair is multiplied by
                // zero, so its 17 shadow rays are never cast.  Same expression as the shading step below: bit-identical image.
                const float ndl = dot3(qf(SQ, SF_NX, sj), qf(SQ, SF_NY, sj), qf(SQ, SF_NZ, sj), dx, dy, dz);
                const bool culls = !P.no_cone_cull;
                const bool go = valid && !(dist < 0.001f) && (ndl > 0.f || !culls);  // renderer.go:252-254
                if (STATS && valid && !(dist < 0.001f) && !go) stat_add<STATS>(st, kStatBackfacing);
                uint8_t lit_code = 0;
                if (go) {
                    stat_add<STATS>(st, kStatPairSetups);
                    float tt;
                    int pp;
                    if (SMALL && P.soft && culls) {
                        // Tiny scenes: which spheres can the pair's shadow cone (the hard ray is its axis) reach at all?
                        const float thr = tangent_threshold(ndl, ox, oy, oz);
                        const float nnx = qf(SQ, SF_NX, sj), nny = qf(SQ, SF_NY, sj), nnz = qf(SQ, SF_NZ, sj);
                        uint32_t cm = 0;
#pragma unroll 1
                        for (int si = 0; si < P.small_n; si++) {
                            stat_add<STATS>(st, kStatConeTests);
                            const float4 s = P.small_sph[si];
                            const float vx = s.x - ox, vy = s.y - oy, vz = s.z - oz;
                            if (dot3(nnx, nny, nnz, vx, vy, vz) + fabsf(s.w) < thr) continue;  // behind the tangent plane
                            if (cone_sphere_candidate(vx, vy, vz, fabsf(s.w), dx, dy, dz, dist)) cm |= 1u << si;
                        }
                        W.cmask[li][j] = (uint16_t)cm;
                        if (cm == 0) {
                            // Nothing in the cone: the hard ray and every one of the 16 jittered rays are unoccluded, so the
                            // pair needs neither ray tests nor random numbers (shadowFactor = 16/16, renderer.go:326-328)
                            lit_code = 2;
                            W.cnt[li][j] = 16;
                            stat_add<STATS>(st, kStatSoftSkipped);
                        } else {
                            stat_add<STATS>(st, kStatLightEvals);
                            lit_code = small_query<STATS, true>(P, cm, ox, oy, oz, dx, dy, dz, 0.001f, dist, true, tt, pp, st) ? 0 : 1;
                        }
                    } else {
                        stat_add<STATS>(st, kStatLightEvals);
                        const unsigned int nv0 = STATS ? st.v[kStatNodes] : 0u;
                        lit_code = query<STATS, SMALL, GEOM>(P, ox, oy, oz, dx, dy, dz, 0.001f, dist, true, tt, pp, st) ? 0 : 1;
                        walk_account<STATS>(st, nv0, 2);
                        if (SMALL && lit_code && P.soft) W.cmask[li][j] = (uint16_t)((1u << P.small_n) - 1u);  // culls off: every sphere
                    }
                }
                if (valid) W.lit[li][j] = lit_code;
            }
            __syncwarp();

            // ---- C: calculateSmartShadow's 16 jittered rays (renderer.go:311-328) for every lit pair of the chunk ----
            if (P.soft) {
                int np = 0;
                for (int li = 0; li < lc; li++) {
                    const bool bit = act && W.lit[li][lane] == 1;  // 2: already resolved (empty cone)
                    const unsigned m = __ballot_sync(FULL_MASK, bit);
                    if (bit) W.pairs[np + __popc(m & lt_mask)] = (uint16_t)((li << 8) | lane);
                    np += __popc(m);
                }
                __syncwarp();
                for (int p0 = 0; p0 < np; p0 += 32) {
                    const int pc = min(32, np - p0);
                    int pc_rays = pc;  // pairs of this batch that cast their 16 rays
                    if (!SMALL) {
                        // lane = pair: one cone walk collects the pair's candidate primitives
                        uint32_t nc_mine = 0;
                        if (lane < pc) {
                            const int pr = (int)W.pairs[p0 + lane];
                            const int sj = base + (pr & 31);
                            const float ox = qf(SQ, SF_PX, sj), oy = qf(SQ, SF_PY, sj), oz = qf(SQ, SF_PZ, sj);
                            const float4 L0 = light4<SMALL>(P, l0 + (pr >> 8), 0);
                            float ax = L0.x - ox, ay = L0.y - oy, az = L0.z - oz;
                            const float dist2 = dot3(ax, ay, az, ax, ay, az);
                            const float inv_d = rsqrt_fast(dist2);
                            ax *= inv_d; ay *= inv_d; az *= inv_d;
                            const float nnx = qf(SQ, SF_NX, sj), nny = qf(SQ, SF_NY, sj), nnz = qf(SQ, SF_NZ, sj);
                            const float thr = tangent_threshold(dot3(nnx, nny, nnz, ax, ay, az), ox, oy, oz);
                            nc_mine = P.no_cone_cull ? kCandOverflow
                                                     : cone_candidates<STATS, GEOM>(S, ox, oy, oz, ax, ay, az, dist2 * inv_d, nnx, nny, nnz, thr, W.cand[lane], st);
                            W.ncand[lane] = (uint8_t)nc_mine;
                            // an empty cone: all 16 rays are unoccluded whatever their jitter (shadowFactor = 16/16,
                            // renderer.go:326-328) — no random numbers, no ray tests
                            if (nc_mine == 0) {
                                W.cnt[pr >> 8][pr & 31] = 16;
                                stat_add<STATS>(st, kStatSoftSkipped);
                            }
                        }
                        // the pairs that do need their 16 rays, compacted: sel[k] = slot of the k-th of them
                        const unsigned need = __ballot_sync(FULL_MASK, lane < pc && nc_mine != 0);
                        if (lane < pc && nc_mine != 0) W.sel[__popc(need & lt_mask)] = (uint8_t)lane;
                        pc_rays = __popc(need);
                        __syncwarp();
                    }
                    // a quarter warp per pair; lane & 7 = k handles shadow samples 2k and 2k+1 (renderer.go:313)
                    for (int q0 = 0; q0 < pc_rays; q0 += 4) {
                        const int qk = q0 + (lane >> 3);
                        const bool valid = qk < pc_rays;
                        const int qi = SMALL ? qk : (valid ? (int)W.sel[qk] : 0);  // slot of the pair in this batch
                        const int pr = valid ? (int)W.pairs[p0 + qi] : 0;
                        const int li = pr >> 8, j = pr & 31;
                        const int sj = base + j;
                        bool unA = false, unB = false;
                        if (valid) {
                            const float ox = qf(SQ, SF_PX, sj), oy = qf(SQ, SF_PY, sj), oz = qf(SQ, SF_PZ, sj);
                            const float4 L0 = light4<SMALL>(P, l0 + li, 0);
                            float ax = L0.x - ox, ay = L0.y - oy, az = L0.z - oz;
                            const float dist2 = dot3(ax, ay, az, ax, ay, az);
                            const float inv_d = rsqrt_fast(dist2);
                            const float dist = dist2 * inv_d;
                            ax *= inv_d; ay *= inv_d; az *= inv_d;
                            const uint32_t sdw = SQ[SF_SD][sj];
                            const uint4 rb = philox_at<SMALL>(P.rk, SQ[SF_PIXG][sj], sdw & 0xffffu, ((sdw >> 16) << 8) | kStreamShadow,
                                                    ((uint32_t)(l0 + li) << 12) | ((uint32_t)(lane & 7) << 8));
                            stat_add<STATS>(st, kStatRngBlocks);
                            stat_add<STATS>(st, kStatSoftRays, 2);
                            float bx, by, bz;
                            ball_from_bits(rb.x, rb.y, bx, by, bz);
                            float dxa = fmaf(0.1f, bx, ax), dya = fmaf(0.1f, by, ay), dza = fmaf(0.1f, bz, az);
                            normalize3(dxa, dya, dza);
                            ball_from_bits(rb.z, rb.w, bx, by, bz);
                            float dxb = fmaf(0.1f, bx, ax), dyb = fmaf(0.1f, by, ay), dzb = fmaf(0.1f, bz, az);
                            normalize3(dxb, dyb, dzb);
                            bool occA = false, occB = false;
                            if (SMALL) {
                                uint32_t cm = W.cmask[li][j];
                                stat_add<STATS>(st, kStatShadow, 2);
                                while (cm) {
                                    const int si = __ffs(cm) - 1;
                                    cm &= cm - 1;
                                    stat_add<STATS>(st, kStatSphereTests, 2);
                                    const float4 s = P.small_sph[si];
                                    occA = occA || sphere_occludes_unit(s, ox, oy, oz, dxa, dya, dza, 0.001f, dist);
                                    occB = occB || sphere_occludes_unit(s, ox, oy, oz, dxb, dyb, dzb, 0.001f, dist);
                                }
                            } else {
                                const uint32_t nc = W.ncand[qi];
                                if (nc == kCandOverflow) {
                                    float tt;
                                    int pp;
                                    const unsigned int nv0 = STATS ? st.v[kStatNodes] : 0u;
                                    occA = traverse<STATS, GEOM>(S, ox, oy, oz, dxa, dya, dza, 0.001f, dist, true, tt, pp, st);
                                    occB = traverse<STATS, GEOM>(S, ox, oy, oz, dxb, dyb, dzb, 0.001f, dist, true, tt, pp, st);
                                    walk_account<STATS>(st, nv0, 3);
                                } else {
                                    stat_add<STATS>(st, kStatShadow, 2);
                                    for (uint32_t k = 0; k < nc; k++) {
                                        const uint32_t ref = W.cand[qi][k];
                                        if (GEOM == 2 || (GEOM == 3 && (ref & 0x80000000u))) {
                                            const float4* tp = S.tris + 4 * (size_t)(ref & 0x7fffffffu);
                                            const bool ha = tri_occludes(tp, ox, oy, oz, dxa, dya, dza, 0.001f, dist);
                                            const bool hb = tri_occludes(tp, ox, oy, oz, dxb, dyb, dzb, 0.001f, dist);
                                            if (STATS) {
                                                stat_add<STATS>(st, kStatTriTests, 2);
                                                stat_add<STATS>(st, ha ? kStatTriHits : kStatTriRejA);  // rejects counted at the cheapest stage
                                                stat_add<STATS>(st, hb ? kStatTriHits : kStatTriRejA);
                                            }
                                            occA = occA || ha;
                                            occB = occB || hb;
                                        } else {
                                            stat_add<STATS>(st, kStatSphereTests, 2);
                                            const float4 s = ldg4(S.spheres + ref);
                                            occA = occA || sphere_occludes_unit(s, ox, oy, oz, dxa, dya, dza, 0.001f, dist);
                                            occB = occB || sphere_occludes_unit(s, ox, oy, oz, dxb, dyb, dzb, 0.001f, dist);
                                        }
                                    }
                                }
                            }
                            unA = !occA;
                            unB = !occB;
                        }
                        const unsigned ua = __ballot_sync(FULL_MASK, unA), ub = __ballot_sync(FULL_MASK, unB);
                        if (valid && (lane & 7) == 0) {
                            const int sh = lane & 24;
                            W.cnt[li][j] = (uint8_t)(__popc((ua >> sh) & 0xFFu) + __popc((ub >> sh) & 0xFFu));
                        }
                    }
                    __syncwarp();
                }
            }

            // ---- D: calculateDirectLighting's arithmetic for the chunk (renderer.go:258-293), lane = record ----
            if (act) {
                const float px = qf(SQ, SF_PX, slot), py = qf(SQ, SF_PY, slot), pz = qf(SQ, SF_PZ, slot);
                const float nx = qf(SQ, SF_NX, slot), ny = qf(SQ, SF_NY, slot), nz = qf(SQ, SF_NZ, slot);
                for (int li = 0; li < lc; li++) {
                    if (!W.lit[li][lane]) continue;
                    const float factor = P.soft ? (float)W.cnt[li][lane] * (1.0f / 16.0f) : 1.0f;
                    if (!(factor > 0.0f)) continue;
                    stat_add<STATS>(st, kStatDiffuse);
                    const float4 L0 = light4<SMALL>(P, l0 + li, 0);
                    float ldx = L0.x - px, ldy = L0.y - py, ldz = L0.z - pz;
                    const float dist2 = dot3(ldx, ldy, ldz, ldx, ldy, ldz);
                    const float inv_d = rsqrt_fast(dist2);
                    ldx *= inv_d; ldy *= inv_d; ldz *= inv_d;
                    const float cosT = fmaxf(0.f, dot3(nx, ny, nz, ldx, ldy, ldz));
                    const float inten = cosT * L0.w * (inv_d * inv_d);
                    const float kdw = inten * factor;
                    dr = fmaf(kar, kdw, dr); dg = fmaf(kag, kdw, dg); db = fmaf(kab, kdw, db);
                    if (spec_pow > 0.f) {  // metallic > 0.5, resolved in float64 on the host
                        stat_add<STATS>(st, kStatSpec);
                        const float4 L1 = light4<SMALL>(P, l0 + li, 1);
                        float vx = -px, vy = -py, vz = -pz;  // viewDir toward the world origin (renderer.go:279)
                        normalize3(vx, vy, vz);
                        float hx = ldx + vx, hy = ldy + vy, hz = ldz + vz;
                        normalize3(hx, hy, hz);
                        const float nh = fmaxf(0.f, dot3(nx, ny, nz, hx, hy, hz));
                        const float x2 = nh * nh, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8, x32 = x16 * x16;
                        const float si = (spec_pow > 56.f) ? x32 * x32 : ((spec_pow > 40.f) ? x32 * x16 : x32);
                        const float sw = si * inten * factor * spec_w;
                        dr = fmaf(L1.x, sw, dr); dg = fmaf(L1.y, sw, dg); db = fmaf(L1.z, sw, db);
                    }
                }
            }
            __syncwarp();  // lit / cnt / cmask are reused by the next chunk
        }

        // ---- traceRay's weighting of this hit (renderer.go:177-226): T * (emitted + w * direct) into the pixel ----
        if (act) {
            const int mat = (int)SQ[SF_MAT][slot];
            const float4 m0 = mat4<SMALL>(P, mat, 0), m2 = mat4<SMALL>(P, mat, 2);
            const bool is_light = __float_as_int(m0.x) == 6;
            // DiffuseLight does not scatter: emitted + direct (renderer.go:182-184); everything else: emitted (0) + w_d * direct
            const float wd = is_light ? 1.0f : m2.w;
            const float er = is_light ? m0.y : 0.f, eg = is_light ? m0.z : 0.f, eb = is_light ? m0.w : 0.f;  // Emitted
            float r = qf(SQ, SF_TR, slot) * fmaf(dr, wd, er), g = qf(SQ, SF_TG, slot) * fmaf(dg, wd, eg), b = qf(SQ, SF_TB, slot) * fmaf(db, wd, eb);
            if (P.fog_enabled) {  // extension: final = (1-f) * radiance + f * fog colour, f from the primary hit
                const float f = qf(SQ, SF_FOG, slot);
                r *= 1.0f - f; g *= 1.0f - f; b *= 1.0f - f;
                if ((SQ[SF_SD][slot] >> 16) == 0) { r = fmaf(P.fog_r, f, r); g = fmaf(P.fog_g, f, g); b = fmaf(P.fog_b, f, b); }
            }
            add_radiance(P, SQ[SF_PIXL][slot], r, g, b);
        }
        __syncwarp();
        if (kDbg && P.debug_times && !more_units) dbg_t_shade += clock64() - dbg_t0;
    }

    if (kDbg && P.debug_times && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned int w = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
        unsigned long long* o = P.debug_times + 1 + 16 * (size_t)w;
        o[0] = t_units_done; o[1] = t; o[2] = dbg_rounds; o[3] = dbg_paths; o[4] = dbg_shades; o[5] = dbg_t_ext; o[6] = dbg_t_shade;
        o[7] = dbg_sec[0]; o[8] = dbg_sec[1]; o[9] = dbg_sec[2]; o[10] = dbg_sec[3];
    }
    if (STATS && P.stats) {
#pragma unroll
        for (int i = 0; i < kStatCount; i++) {
            unsigned long long v = st.v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
            if (lane == 0 && v) atomicAdd(P.stats + i, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cull pass: which 8x4 pixel blocks can see geometry at all?
//
// Every primary ray of a block (all samples, any jitter) has a direction that is an affine function
// of (u,v) over the block's rectangle, so it lies inside the pyramid spanned by the four corner
// directions.  The pyramid is walked down the BVH with the conservative plane/box test (a box is
// rejected only when it lies entirely behind one side plane); reaching a primitive whose own bound
// (sphere: centre/radius, triangle: vertices) survives marks the block active.  Blocks that are
// culled would only have produced misses (black, renderer.go:171-173), so the image is unchanged;
// in the reference's own scenes 97-100 % of the blocks are culled.
// ---------------------------------------------------------------------------------------------
struct Beam {
    float ox, oy, oz;
    float nx[4], ny[4], nz[4];  // inward side-plane normals
};

__device__ __forceinline__ bool beam_box_outside(const Beam& B, float lox, float hix, float loy, float hiy, float loz, float hiz) {
    const float ax = fmaxf(fabsf(lox - B.ox), fabsf(hix - B.ox)), ay = fmaxf(fabsf(loy - B.oy), fabsf(hiy - B.oy)),
                az = fmaxf(fabsf(loz - B.oz), fabsf(hiz - B.oz));
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float px = (B.nx[i] >= 0.f ? hix : lox) - B.ox, py = (B.ny[i] >= 0.f ? hiy : loy) - B.oy, pz = (B.nz[i] >= 0.f ? hiz : loz) - B.oz;
        const float d = dot3(B.nx[i], B.ny[i], B.nz[i], px, py, pz);
        const float tol = 4e-6f * (fabsf(B.nx[i]) * ax + fabsf(B.ny[i]) * ay + fabsf(B.nz[i]) * az);
        if (d < -tol) return true;
    }
    return false;
}

__global__ void __launch_bounds__(128) cull_kernel(const __grid_constant__ TraceParams P, uint32_t* __restrict__ active_list,
                                                    unsigned int* __restrict__ active_count, const CullExtras X) {
    // the trace kernel's CTAs may become resident while this pass runs (they wait for its end with griddepcontrol.wait)
    asm volatile("griddepcontrol.launch_dependents;");
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id == 0) {
        if (P.stamps) asm volatile("{ .reg .u64 t; mov.u64 t, %%globaltimer; st.global.u64 [%0], t; }" ::"l"(P.stamps) : "memory");
        // the other counter bank is the next frame's: cleared here instead of by a memset node in front of every frame
        if (X.zero_bank) *reinterpret_cast<uint4*>(X.zero_bank) = make_uint4(0u, 0u, 0u, 0u);
        // owner of a frame link: everything this stream did with the previous frame is done, the peers may overwrite it
        if (X.store_flag) {
            asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(X.store_flag), "r"(X.store_value) : "memory");
            __threadfence_system();
        }
    }
    if (id >= (uint32_t)P.n_local_tiles * 32u) return;
    const SceneView& S = P.scene;
    const uint32_t ltile = id >> 5, block = id & 31u;
    // The frame's accumulators are not cleared wholesale (12 MB for 800x600): a kept block clears its own 32 pixels
    // here, a culled block is marked and resolve_kernel writes black for it without reading them.
    P.block_active[id] = 0;
    if (P.max_depth <= 0 || (S.n_nodes == 0 && !P.sky_enabled)) return;
    const uint32_t gtile = (uint32_t)P.shard_rank + ltile * (uint32_t)P.shard_count;
    const uint32_t tx = gtile % (uint32_t)P.tiles_x, ty = gtile / (uint32_t)P.tiles_x;
    const uint32_t x0 = tx * kTile + ((block & 3u) << 3), y0 = ty * kTile + ((block >> 2) << 2);
    if (x0 >= (uint32_t)P.width || y0 >= (uint32_t)P.height) return;
    const uint32_t x1 = min(x0 + 8u, (uint32_t)P.width), y1 = min(y0 + 4u, (uint32_t)P.height);
    // region render (RenderChunk): a block that does not touch the region is not traced
    if (P.crop_x1 > P.crop_x0 && ((int)x1 <= P.crop_x0 || (int)x0 >= P.crop_x1 || (int)y1 <= P.crop_y0 || (int)y0 >= P.crop_y1)) return;
    // (u,v) rectangle of every sample of the block, widened by a rounding margin
    const float u0 = ((float)x0 - 1e-3f) / (float)P.width, u1 = ((float)x1 + 1e-3f) / (float)P.width;
    const float v0 = ((float)y0 - 1e-3f) / (float)P.height, v1 = ((float)y1 + 1e-3f) / (float)P.height;
    const DevCamera& C = P.cam;
    float cx[4], cy[4], cz[4];
    const float uu[4] = {u0, u1, u1, u0}, vv[4] = {v0, v0, v1, v1};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        cx[i] = fmaf(vv[i], C.vx, fmaf(uu[i], C.hx, C.llx));
        cy[i] = fmaf(vv[i], C.vy, fmaf(uu[i], C.hy, C.lly));
        cz[i] = fmaf(vv[i], C.vz, fmaf(uu[i], C.hz, C.llz));
    }
    const float mx = 0.25f * (cx[0] + cx[1] + cx[2] + cx[3]), my = 0.25f * (cy[0] + cy[1] + cy[2] + cy[3]), mz = 0.25f * (cz[0] + cz[1] + cz[2] + cz[3]);
    Beam B;
    B.ox = C.ox; B.oy = C.oy; B.oz = C.oz;
    bool degenerate = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int j = (i + 1) & 3;
        float nx = cy[i] * cz[j] - cz[i] * cy[j], ny = cz[i] * cx[j] - cx[i] * cz[j], nz = cx[i] * cy[j] - cy[i] * cx[j];
        const float sgn = dot3(nx, ny, nz, mx, my, mz);
        if (!(fabsf(sgn) > 0.f)) degenerate = true;  // collapsed pyramid (zero-area viewport): keep the block
        if (sgn < 0.f) { nx = -nx; ny = -ny; nz = -nz; }
        B.nx[i] = nx; B.ny[i] = ny; B.nz[i] = nz;
    }
    // active: some primitive may be visible; deep: one of them is glass/dielectric (paths can stay trapped
    // by total internal reflection up to max_depth) -> those blocks are scheduled first
    // (sky extension: a block that sees no geometry is not black — every block of the image is traced)
    bool active = degenerate || P.sky_enabled != 0, deep = false;
    int stack[64];
    int sp = 0;
    int node = 0;
    while (!deep && S.n_nodes > 0) {
        if (node >= 0) {
            const float4* np = S.nodes + 4 * (size_t)node;
            const float4 n0 = ldg4(np), n1 = ldg4(np + 1), n2 = ldg4(np + 2), n3 = ldg4(np + 3);
            const bool h0 = !beam_box_outside(B, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y);
            const bool h1 = !beam_box_outside(B, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w);
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                stack[sp++] = c1;
                node = c0;
            } else if (h0) {
                node = c0;
            } else if (h1) {
                node = c1;
            } else {
                if (sp == 0) break;
                node = stack[--sp];
            }
        } else {
            const uint32_t v = ~(uint32_t)node;
            const uint32_t start = v & 0x3FFFFFFu;
            const int cnt = (int)((v >> 26) & 15u) + 1;
            if (((v >> 30) & 1u) == 0) {
                for (int i = 0; i < cnt && !deep; i++) {
                    const float4 s = ldg4(S.spheres + start + i);
                    const float px = s.x - B.ox, py = s.y - B.oy, pz = s.z - B.oz;
                    const float r = fabsf(s.w) * 1.00001f + 1e-6f * (fabsf(px) + fabsf(py) + fabsf(pz));
                    bool out = false;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float d = dot3(B.nx[k], B.ny[k], B.nz[k], px, py, pz);
                        const float nl = sqrtf(dot3(B.nx[k], B.ny[k], B.nz[k], B.nx[k], B.ny[k], B.nz[k]));
                        if (d < -r * nl) out = true;
                    }
                    if (!out) {
                        active = true;
                        const int mt = __float_as_int(ldg4(S.mats + 4 * (size_t)__ldg(&S.sphere_meta[start + i]).x).x);
                        deep = (mt == 4 || mt == 5);
                    }
                }
            } else {
                for (int i = 0; i < cnt && !deep; i++) {
                    const float4* tp = S.tris + 4 * (size_t)(start + i);
                    const float4 a = ldg4(tp), e1 = ldg4(tp + 1), e2 = ldg4(tp + 2);
                    const float ax = a.x - B.ox, ay = a.y - B.oy, az = a.z - B.oz;
                    const float ext = 1e-5f * (fabsf(ax) + fabsf(ay) + fabsf(az) + fabsf(e1.x) + fabsf(e1.y) + fabsf(e1.z) + fabsf(e2.x) + fabsf(e2.y) + fabsf(e2.z));
                    bool out = false;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float nl = fabsf(B.nx[k]) + fabsf(B.ny[k]) + fabsf(B.nz[k]);
                        const float d0 = dot3(B.nx[k], B.ny[k], B.nz[k], ax, ay, az);
                        const float d1 = d0 + dot3(B.nx[k], B.ny[k], B.nz[k], e1.x, e1.y, e1.z);
                        const float d2 = d0 + dot3(B.nx[k], B.ny[k], B.nz[k], e2.x, e2.y, e2.z);
                        if (fmaxf(d0, fmaxf(d1, d2)) < -ext * nl) out = true;
                    }
                    if (!out) {
                        active = true;
                        const int mt = __float_as_int(ldg4(S.mats + 4 * (size_t)__float_as_int(a.w)).x);
                        deep = (mt == 4 || mt == 5);
                    }
                }
            }
            if (deep || sp == 0) break;
            node = stack[--sp];
        }
    }
    if (active || deep) {
        P.block_active[id] = 1;
        // 4 rows x 8 pixels x 3 channels of int64: 192 contiguous bytes per row
        const uint32_t lx0 = (block & 3u) << 3, ly0 = (block >> 2) << 2;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            ulonglong2* row = reinterpret_cast<ulonglong2*>(P.accum + 3 * ((size_t)ltile * kTilePixels + (ly0 + r) * kTile + lx0));
#pragma unroll
            for (int k = 0; k < 12; k++) row[k] = make_ulonglong2(0ull, 0ull);
        }
    }
    // one array, two cursors: deep blocks fill it from the front, the others from the back
    if (deep || degenerate) active_list[atomicAdd(active_count, 1u)] = id;
    else if (active) active_list[(uint32_t)P.n_local_tiles * 32u - 1u - atomicAdd(active_count + 1, 1u)] = id;
}

cudaError_t launch_cull(const TraceParams& p, uint32_t* active_list, unsigned int* active_count, cudaStream_t stream, const CullExtras& x) {
    const unsigned int n = (unsigned int)p.n_local_tiles * 32u;
    if (n == 0) return cudaSuccess;
    cull_kernel<<<(n + 127) / 128, 128, 0, stream>>>(p, active_list, active_count, x);
    return cudaGetLastError();
}

// launch with or without the programmatic-dependent-launch attribute (kernels.h)
template <typename... KArgs, typename... Args>
static cudaError_t launch_maybe_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, bool dependent, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = dependent ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}


template <bool STATS, bool SMALL, int GEOM, bool SKY = false>
static cudaError_t launch_trace_variant(const TraceParams& p, int sm_count, cudaStream_t stream, bool dependent) {
    // persistent grid: as many CTAs as fit on the chip at once (occupancy is register-bound)
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_kernel<STATS, SMALL, GEOM, SKY>, kWarpsPerCta * 32, 0);
        if (e != cudaSuccess) return e;
        ctas_per_sm = n > 0 ? n : 1;
    }
    if (kDbg && p.debug_times) {  // -DGORT_DEBUG build + GORT_DEBUG_TIMES=1: per-warp timeline of this launch on stderr
        const int nw = sm_count * ctas_per_sm * kWarpsPerCta;
        cudaMemsetAsync(p.debug_times, 0xff, 8, stream);
        cudaMemsetAsync(p.debug_times + 1, 0, (size_t)nw * 128, stream);
        trace_kernel<STATS, SMALL, GEOM, SKY><<<sm_count * ctas_per_sm, kWarpsPerCta * 32, 0, stream>>>(p);
        std::vector<unsigned long long> h(1 + 16 * (size_t)nw);
        cudaMemcpyAsync(h.data(), p.debug_times, h.size() * 8, cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        struct Rec { double units, end; unsigned long long rounds, paths, shades; double ext_us, shade_us; double sec[4]; };
        std::vector<Rec> recs;
        for (int w = 0; w < nw; w++) {
            const unsigned long long* o = &h[1 + 16 * (size_t)w];
            if (!o[1]) continue;
            recs.push_back(Rec{(double)(o[0] - h[0]) * 1e-3, (double)(o[1] - h[0]) * 1e-3, o[2], o[3], o[4], (double)o[5] / 1965.0, (double)o[6] / 1965.0,
                                {(double)o[7], (double)o[8], (double)o[9], (double)o[10]}});
        }
        std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.end < b.end; });
        auto at = [&](double q) -> const Rec& { return recs[(size_t)(q * (recs.size() - 1))]; };
        if (!recs.empty()) {
            fprintf(stderr, "[gort debug] warps %zu finish us p0 %.1f p10 %.1f p50 %.1f p90 %.1f p99 %.1f p100 %.1f\n", recs.size(), at(0).end, at(0.1).end,
                    at(0.5).end, at(0.9).end, at(0.99).end, at(1).end);
            for (size_t i = recs.size() > 6 ? recs.size() - 6 : 0; i < recs.size(); i++)
                fprintf(stderr, "[gort debug]   slow warp: units exhausted %.1f us, end %.1f us; after that %llu EXTEND rounds (%llu paths) %.1f us, %llu SHADE rounds %.1f us\n",
                        recs[i].units, recs[i].end, recs[i].rounds, recs[i].paths, recs[i].ext_us, recs[i].shades, recs[i].shade_us);
            // anatomy of those EXTEND rounds (cycles per round, summed over the slowest 32 warps)
            double sec[4] = {0, 0, 0, 0}, rounds = 0;
            for (size_t i = recs.size() > 32 ? recs.size() - 32 : 0; i < recs.size(); i++) {
                for (int k = 0; k < 4; k++) sec[k] += recs[i].sec[k];
                rounds += (double)recs[i].rounds;
            }
            if (rounds > 0)
                fprintf(stderr, "[gort debug]   cycles per drain-mode EXTEND round: queue read + hit record + materials %.0f, Scatter %.0f, hitWorld %.0f, compaction + bookkeeping %.0f\n",
                        sec[0] / rounds, sec[1] / rounds, sec[2] / rounds, sec[3] / rounds);
        }
        return cudaGetLastError();
    }
    return launch_maybe_dependent(trace_kernel<STATS, SMALL, GEOM, SKY>, dim3(sm_count * ctas_per_sm), dim3(kWarpsPerCta * 32), stream, dependent, p);
}

template <bool STATS, bool SKY>
static cudaError_t launch_trace_bvh(const TraceParams& p, int geom, int sm_count, cudaStream_t stream, bool dependent) {
    switch (geom) {
        case 1: return launch_trace_variant<STATS, false, 1, SKY>(p, sm_count, stream, dependent);
        case 2: return launch_trace_variant<STATS, false, 2, SKY>(p, sm_count, stream, dependent);
        default: return launch_trace_variant<STATS, false, 3, SKY>(p, sm_count, stream, dependent);
    }
}

cudaError_t launch_trace(const TraceParams& p, bool stats, int sm_count, cudaStream_t stream, bool dependent) {
    if (p.n_local_tiles == 0) return cudaSuccess;
    // kernel variant: tiny sphere scenes scan the parameter bank (an empty scene without sky is the degenerate case of that);
    // BVH scenes run the kernel specialised for what they hold
    const int geom = (p.scene.n_spheres > 0 ? 1 : 0) | (p.scene.n_tris > 0 ? 2 : 0);
    if (p.small_n > 0 || (geom == 0 && !p.sky_enabled))
        return stats ? launch_trace_variant<true, true, 1>(p, sm_count, stream, dependent) : launch_trace_variant<false, true, 1>(p, sm_count, stream, dependent);
    if (p.sky_enabled)
        return stats ? launch_trace_bvh<true, true>(p, geom, sm_count, stream, dependent) : launch_trace_bvh<false, true>(p, geom, sm_count, stream, dependent);
    return stats ? launch_trace_bvh<true, false>(p, geom, sm_count, stream, dependent) : launch_trace_bvh<false, false>(p, geom, sm_count, stream, dependent);
}

int trace_kernel_regs(bool stats) {
    cudaFuncAttributes a;
    cudaError_t e = stats ? cudaFuncGetAttributes(&a, trace_kernel<true, false, 3>) : cudaFuncGetAttributes(&a, trace_kernel<false, false, 3>);
    return e == cudaSuccess ? a.numRegs : -1;
}

// ---------------------------------------------------------------------------------------------
// resolve: collector of Render (renderer.go:92-97): toneMap (:348-367) + ToRGB (vector.go:106-109)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t tone_map_u8(long long fixed, double inv_scale_spp) {
    double c = (double)fixed * inv_scale_spp;      // color.DivScalar(samples) renderer.go:162
    c = 1.0 - exp(-c);                             // exposure 1.0
    c = pow(c, 1.0 / 2.2);                         // NaN for negative c, like math.Pow
    if (c != c) return 0;                          // declared: NaN -> 0
    c = fmax(0.0, fmin(1.0, c));
    return (uint8_t)(c * 255.0);                   // truncating conversion
}

__global__ void __launch_bounds__(256) resolve_kernel(const ResolveParams R) {
    // launched as a programmatic dependent of the trace kernel, the CTAs sit here while the last paths of the frame finish
    asm volatile("griddepcontrol.launch_dependents;");  // (a frame link's signal kernel)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (R.stamps && blockIdx.x == 0 && threadIdx.x == 0)
        asm volatile("{ .reg .u64 t; mov.u64 t, %%globaltimer; st.global.u64 [%0], t; }" ::"l"(R.stamps + 2) : "memory");
    const int n = R.n_local_tiles * kTilePixels;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int lt = i / kTilePixels, p = i % kTilePixels;
        const int lx = p & (kTile - 1), ly = p / kTile;
        const int gt = R.shard_rank + lt * R.shard_count;
        const int x = (gt % R.tiles_x) * kTile + lx, y = (gt / R.tiles_x) * kTile + ly;
        const bool inside = x < R.width && y < R.height;
        uchar4 px = make_uchar4(0, 0, 0, 0);
        const bool kept = R.block_active[lt * 32 + (ly >> 2) * 4 + (lx >> 3)] != 0;
        if ((R.part == 1 && kept) || (R.part == 2 && !kept)) continue;
        if (inside && !kept) px = make_uchar4(0, 0, 0, 255);  // culled block: every sample missed (renderer.go:171-173), toneMap(0) = 0
        if (inside && kept) {
            const double inv = 1.0 / ((double)(1u << kAccumFracBits) * (double)R.samples);
            const unsigned long long* a = R.accum + 3 * (size_t)i;
            px.x = tone_map_u8((long long)a[0], inv);
            px.y = tone_map_u8((long long)a[1], inv);
            px.z = tone_map_u8((long long)a[2], inv);
            px.w = 255;
        }
        uchar4* out = reinterpret_cast<uchar4*>(R.out);
        if (R.slab_mode) out[i] = px;
        else if (inside) out[(size_t)y * R.width + x] = px;
    }
    if (R.stamps) {
        // end of the frame for gort_stats: stamped by the last CTA to get here
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(R.done_count, 1u) == gridDim.x - 1) {
            *R.done_count = 0;  // for the next launch on this stream
            asm volatile("{ .reg .u64 t; mov.u64 t, %%globaltimer; st.global.u64 [%0], t; }" ::"l"(R.stamps + 3) : "memory");
        }
    }
}

cudaError_t launch_resolve(const ResolveParams& p, cudaStream_t stream, int max_blocks, bool dependent) {
    const int n = p.n_local_tiles * kTilePixels;
    if (n == 0) return cudaSuccess;
    if (max_blocks > 0) return launch_maybe_dependent(resolve_kernel, dim3(std::min(max_blocks, (n + 127) / 128)), dim3(128), stream, dependent, p);
    return launch_maybe_dependent(resolve_kernel, dim3((n + 255) / 256), dim3(256), stream, dependent, p);
}

// rank-major concatenation of tile-major shard slabs -> row-major frame
__global__ void __launch_bounds__(256) unswizzle_kernel(const uchar4* __restrict__ slabs, int shard_count, int tiles_per_shard,
                                                          int tiles_x, int n_tiles, int width, int height, uchar4* __restrict__ rgba) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tiles * kTilePixels) return;
    const int gt = i / kTilePixels, p = i % kTilePixels;
    const int x = (gt % tiles_x) * kTile + (p & (kTile - 1)), y = (gt / tiles_x) * kTile + p / kTile;
    if (x >= width || y >= height) return;
    const int rank = gt % shard_count, j = gt / shard_count;
    rgba[(size_t)y * width + x] = slabs[((size_t)rank * tiles_per_shard + j) * kTilePixels + p];
}

cudaError_t launch_unswizzle(const uint8_t* slabs, int shard_count, int width, int height, uint8_t* rgba, cudaStream_t stream) {
    const int tiles_x = (width + kTile - 1) / kTile, tiles_y = (height + kTile - 1) / kTile;
    const int n_tiles = tiles_x * tiles_y;
    const int tiles_per_shard = (n_tiles + shard_count - 1) / shard_count;
    const int n = n_tiles * kTilePixels;
    if (n == 0) return cudaSuccess;
    unswizzle_kernel<<<(n + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const uchar4*>(slabs), shard_count, tiles_per_shard, tiles_x,
                                                         n_tiles, width, height, reinterpret_cast<uchar4*>(rgba));
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// frame link: every rank's resolve_kernel stores its tiles straight into the owner's row-major frame through
// a peer mapping (NVLink), so the frame is assembled without a collective and without an un-swizzle pass.
// Completion and reuse are ordered by two counters in the owner's memory:
//   arrived  += 1 by every peer after its tiles of frame k are written (fence.sys first);
//              the owner waits for k * (ranks - 1) before anything that reads the frame
//   consumed  = k - 1 stored by the owner when it starts frame k (its stream has consumed frame k - 1 by then);
//              a peer waits for it before it overwrites the frame with its tiles of frame k
// Each kernel is one thread; a waiting kernel spins on a flag written from ANOTHER GPU (never on a kernel of
// its own GPU), with back-off.
// ---------------------------------------------------------------------------------------------
// (the signal and wait kernels can be launched as programmatic dependents of the kernel before them, like the frame's own)
__global__ void link_signal_kernel(unsigned int* flag) {
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the resolve pass is complete, its stores performed
    __threadfence_system();
    atomicAdd_system(flag, 1u);
}
__global__ void link_store_kernel(unsigned int* flag, unsigned int value) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
    __threadfence_system();
}
// spin (with back-off) until *flag >= target; a flag written from ANOTHER GPU.  20 s without it: a rank died — never hang the
// GPU, flag it and let the host report.
__device__ __forceinline__ void link_wait(const unsigned int* flag, unsigned int target, unsigned int* timed_out) {
    unsigned int ns = 100;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned int v;
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - target) >= 0) break;
        __nanosleep(ns);
        if (ns < 2000) ns *= 2;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 20000000000ull) {
            atomicExch_system(timed_out, 1u);
            break;
        }
    }
    __threadfence_system();
}

__global__ void link_wait_kernel(const unsigned int* flag, unsigned int target, unsigned int* timed_out) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    link_wait(flag, target, timed_out);
}
cudaError_t launch_link_signal(unsigned int* flag, cudaStream_t stream, bool dependent) {
    return launch_maybe_dependent(link_signal_kernel, dim3(1), dim3(1), stream, dependent, flag);
}
cudaError_t launch_link_store(unsigned int* flag, unsigned int value, cudaStream_t stream) {
    link_store_kernel<<<1, 1, 0, stream>>>(flag, value);
    return cudaGetLastError();
}
cudaError_t launch_link_wait(const unsigned int* flag, unsigned int target, unsigned int* timed_out, cudaStream_t stream, bool dependent) {
    return launch_maybe_dependent(link_wait_kernel, dim3(1), dim3(1), stream, dependent, flag, target, timed_out);
}

// ---------------------------------------------------------------------------------------------
// test hook: hitWorld for explicit rays
// ---------------------------------------------------------------------------------------------
__global__ void trace_rays_kernel(const SceneView S, int n, const float* __restrict__ o, const float* __restrict__ d, float tmin,
                                  float tmax, int any_hit, float* __restrict__ out_t, int* __restrict__ out_order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Stats st;
    float t = -1.f;
    int prim = 0;
    bool hit;
    if (any_hit) hit = traverse<false>(S, o[3 * i], o[3 * i + 1], o[3 * i + 2], d[3 * i], d[3 * i + 1], d[3 * i + 2], tmin, tmax, true, t, prim, st);
    else hit = traverse<false>(S, o[3 * i], o[3 * i + 1], o[3 * i + 2], d[3 * i], d[3 * i + 1], d[3 * i + 2], tmin, tmax, false, t, prim, st);
    if (any_hit) {
        out_t[i] = hit ? 1.f : -1.f;
        out_order[i] = -1;
    } else {
        out_t[i] = hit ? t : -1.f;
        out_order[i] = hit ? prim_order(S, prim) : -1;
    }
}

cudaError_t launch_trace_rays(const SceneView& scene, int n, const float* origins, const float* dirs, float tmin, float tmax,
                              int any_hit, float* out_t, int* out_order, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    trace_rays_kernel<<<(n + 127) / 128, 128, 0, stream>>>(scene, n, origins, dirs, tmin, tmax, any_hit, out_t, out_order);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// FP32 issue-rate microbenchmark: 8 independent FFMA chains per thread
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* sink, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456f) sink[0] = s;  // never true; keeps the chains alive
}

cudaError_t launch_ffma_peak(float* sink, int iters, int blocks, int threads, cudaStream_t stream) {
    ffma_peak_kernel<<<blocks, threads, 0, stream>>>(sink, iters);
    return cudaGetLastError();
}

}  // namespace gort
