// stream.cu — global-queue wavefront pipeline for large scenes (see stream.h for the stage list).
//
// Every stage is a plain grid over a queue in global memory whose length is read from a device counter, so the host
// enqueues a bounce without knowing how many paths survive.  Arithmetic, Philox counters and accumulator adds are
// those of kernels.cu (trace_kernel), stage by stage; the reference lines are cited there and again here.
#include "stream.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

#include "device.cuh"
#include "lbvh.h"

namespace gort {

constexpr uint32_t kDeadPrim = 0x7FFFFFFFu;  // qa.w of a path that ended (miss, no scatter, depth, dead-path cut)
#ifndef GORT_POOL_REFILL
#define GORT_POOL_REFILL 8
#endif
#ifndef GORT_POOL_CHUNK
#define GORT_POOL_CHUNK 128
#endif
#ifndef GORT_POOL_QNODES
#define GORT_POOL_QNODES 1
#endif
#ifndef GORT_POOL_WIDE
#define GORT_POOL_WIDE 1  // the traversal kernel walks the 4-wide collapse of the tree (0: the quantised binary nodes)
#endif
#ifndef GORT_POOL_DEFER
#define GORT_POOL_DEFER 1  // leaf tests deferred and run per warp in batches (4-wide walk only)
#endif
#ifndef GORT_LEAF_BATCH
#define GORT_LEAF_BATCH 8
#endif
#ifndef GORT_POOL_MINB
#define GORT_POOL_MINB 10
#endif
constexpr int kPoolRefill = GORT_POOL_REFILL;      // pool_trace refills as soon as this many lanes of a warp are idle
constexpr uint32_t kPoolChunk = GORT_POOL_CHUNK;   // rays a warp takes from the global cursor at a time
constexpr int kLC = kStreamLightChunk;
[[maybe_unused]] constexpr int kLeafBatch = GORT_LEAF_BATCH;     // deferred leaf tests run once this many lanes of a warp have a parked leaf
[[maybe_unused]] constexpr int kDoneNode = (int)0x80000000;      // "nothing left to walk" (not a valid leaf link: 2^26 - 1 primitives from start 2^26 - 1)
#ifndef GORT_CONE_BUDGET
#define GORT_CONE_BUDGET 16
#endif
constexpr int kConeBudget = GORT_CONE_BUDGET;  // node visits of one shadow-cone walk before the pair is sent to the walk list instead

enum PoolSrc { SRC_PRIMARY = 0, SRC_EXT = 1, SRC_HARD = 2, SRC_SOFT = 3 };

template <bool STATS>
__device__ __forceinline__ void stats_zero(Stats& st) {
    if (STATS) {
#pragma unroll
        for (int i = 0; i < kStatCount; i++) st.v[i] = 0;
    }
}

template <bool STATS>
__device__ __forceinline__ void stats_flush(const TraceParams& P, Stats& st) {
    if (STATS && P.stats) {
#pragma unroll
        for (int i = 0; i < kStatCount; i++) {
            unsigned long long v = st.v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(P.stats + i, v);
        }
    }
}

__device__ __forceinline__ void store_w(float4* p, uint32_t w) { reinterpret_cast<uint32_t*>(p)[3] = w; }

// direction and distance from a hit point to a light, exactly as the shade stage of trace_kernel forms them
// (lightDir / lightDistance, renderer.go:249-251)
__device__ __forceinline__ void light_dir(const float4 L0, float ox, float oy, float oz, float& dx, float& dy, float& dz, float& dist) {
    dx = L0.x - ox; dy = L0.y - oy; dz = L0.z - oz;
    const float dist2 = dot3(dx, dy, dz, dx, dy, dz);
    const float inv_d = dist2 > 0.f ? rsqrt_fast(dist2) : 0.f;
    dist = dist2 * inv_d;
    dx *= inv_d; dy *= inv_d; dz *= inv_d;
}

// ---------------------------------------------------------------------------------------------
// pool_trace: hitWorld (renderer.go:333-346) for a pool of rays, persistent warps with lane refill.
//
// A lane holds one ray and walks the BVH exactly like traverse() in device.cuh (same slab test, same near-first
// order, same leaf tests and tie rule).  Whenever kPoolRefill lanes of a warp have finished, they take the next
// rays of the pool (consecutive indices, so a warp keeps rays that were generated together), while the other
// lanes keep their walk state: the warp's instruction stream stays filled whatever the spread of walk lengths.
// ANY sources (shadow rays) stop at the first accepted primitive.
// ---------------------------------------------------------------------------------------------
// TOP (4-wide walk only): the first four wide levels of the tree (lbvh.h: wide_top_block, 5.4 KB) are staged in shared memory
// by one bulk copy (cp.async.bulk -> mbarrier) when the CTA starts, and visits of those nodes read them from there.
template <int SRC, bool STATS, int GEOM, bool TOP>
__global__ void __launch_bounds__(128, GORT_POOL_MINB) pool_trace_kernel(const __grid_constant__ TraceParams P, const __grid_constant__ StreamView V) {
    constexpr bool ANY = SRC == SRC_HARD || SRC == SRC_SOFT;
    [[maybe_unused]] const float4* top_s = nullptr;
    const SceneView& S = P.scene;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    Stats st;
    stats_zero<STATS>(st);

    uint32_t total;
    unsigned int* fetch;
    const int cb = kCtlChunk0 + 8 * V.chunk;
    if (SRC == SRC_PRIMARY) { total = V.ctl[kCtlNew]; fetch = V.ctl + kCtlFetchPrimary; }
    else if (SRC == SRC_EXT) { total = V.ctl[kCtlNext]; fetch = V.ctl + kCtlFetchExt; }
    else if (SRC == SRC_HARD) { total = V.ctl[cb + kCtlHard]; fetch = V.ctl + cb + kCtlFetchHard; }
    else { total = V.ctl[cb + kCtlWalk] * 16u; fetch = V.ctl + cb + kCtlFetchWalk; }
    if (total == 0) return;
    const int nxt = V.cur ^ 1;
    // primary rays of this iteration go behind the scattered rays in the next queue
    const uint32_t slot0 = SRC == SRC_PRIMARY ? V.ctl[kCtlNext] : 0u;
    const unsigned long long g0 = SRC == SRC_PRIMARY ? ((unsigned long long)V.ctl[kCtlPrimStart] | ((unsigned long long)V.ctl[kCtlPrimStart + 1] << 32)) : 0ull;
    // a pool smaller than one ray per lane of the grid is dealt out 32 rays at a time (the kernel then lasts one walk, not four)
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t chunk = total >= n_warps * kPoolChunk ? kPoolChunk : max(32u, ((total / n_warps) + 31u) & ~31u);

#if GORT_POOL_WIDE
    const float4* __restrict__ nodes = S.nodes + 6 * (size_t)S.n_nodes;  // the 4-wide collapse (lbvh.h): 64 bytes per wide node
#elif GORT_POOL_QNODES
    const float4* __restrict__ nodes = S.nodes + 4 * (size_t)S.n_nodes;  // the quantised copy: 32 bytes per node
#else
    const float4* __restrict__ nodes = S.nodes;
#endif
    const float4* __restrict__ spheres = GEOM != 2 ? S.spheres : nullptr;
    const float4* __restrict__ tris = GEOM != 1 ? S.tris : nullptr;

    int root = 0;  // where a walk starts: wide node 0, or its copy at the head of the staged block
    if constexpr (TOP) {
        __shared__ __align__(128) float4 top_block[kTopNodes * 4];
        __shared__ __align__(8) unsigned long long top_bar;
        top_s = top_block;
        constexpr uint32_t kTopBytes = kTopNodes * 4 * sizeof(float4);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&top_bar), dst = (uint32_t)__cvta_generic_to_shared(top_block);
        if (V.top_plain) {
            // (A/B switch: the same block copied by the CTA's threads, no bulk copy, no mbarrier)
            for (int k = threadIdx.x; k < kTopNodes * 4; k += blockDim.x) top_block[k] = __ldg(V.top + k);
            __syncthreads();
            root = kTopBase;
        }
        if (!V.top_plain && threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (!V.top_plain && threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTopBytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(V.top), "r"(kTopBytes), "r"(bar)
                         : "memory");
        }
        // every thread waits for phase 0 of the barrier (the copy's bytes); bounded, and a walk that never got its block
        // simply starts at the global root
        uint32_t ok = V.top_plain ? 1u : 0u;
        for (int tries = 0; tries < (1 << 20) && !ok; tries++)  // (2^16 failed tries took 12 ms on a B200: the bound is ~0.2 s)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar) : "memory");
        if (__syncthreads_and((int)ok)) root = kTopBase;
    }

    uint32_t wnext = 0, wend = 0;  // this warp's share of the pool: [wnext, wend)
    bool pool_done = false;        // warp-uniform
    bool have = false;             // this lane holds an unfinished ray
    uint32_t my = 0;               // its index in the pool
    RayQuery q;
    q.ox = q.oy = q.oz = q.dx = q.dy = q.dz = 0.f; q.a = 1.f; q.inv_a = 1.f; q.tmin = 0.001f; q.tbest = 0.f; q.best = 0; q.found = false;
    float idx = 0.f, idy = 0.f, idz = 0.f, oodx = 0.f, oody = 0.f, oodz = 0.f;
#if GORT_POOL_WIDE
    int stack[96];  // up to three pushes per wide level, (depth <= 62) / 2 levels
#else
    int stack[64];
#endif
    int sp = 0, node = 0;
#if GORT_POOL_WIDE && GORT_POOL_DEFER
    int pend = 0;  // a leaf this lane has reached but not tested yet (leaf links are negative; 0 = none)
#endif

    for (;;) {
        const unsigned act = __ballot_sync(FULL_MASK, have);
        if (!pool_done && __popc(act) <= 32 - kPoolRefill) {
            // ---------------- refill the idle lanes ----------------
            if (wnext >= wend) {
                uint32_t b = total;
                if (lane == 0 && *reinterpret_cast<volatile unsigned int*>(fetch) < total) b = atomicAdd(fetch, chunk);
                b = __shfl_sync(FULL_MASK, b, 0);
                if (b >= total) {
                    pool_done = true;
                } else {
                    wnext = b;
                    wend = min(b + chunk, total);
                }
            }
            if (!pool_done) {
                const unsigned idle = ~act;
                const uint32_t avail = wend - wnext;
                const uint32_t r = (uint32_t)__popc(idle & lt_mask);
                if (!have && r < avail) {
                    my = wnext + r;
                    bool valid = true;
                    float tmax = FLT_MAX * 2.0f;
                    if (SRC == SRC_PRIMARY) {
                        // tracePixel / getRay (renderer.go:150-163,377-390); pool index = ((sample, active block), lane of the 8x4 block)
                        const uint32_t per_s = V.n_active * 32u;
                        const unsigned long long g = g0 + my;  // index in the frame's primary sequence: sample-major over the active blocks
                        const uint32_t s = (uint32_t)(g / per_s), rr = (uint32_t)(g - (unsigned long long)s * per_s);
                        const uint32_t blk = rr >> 5, l = rr & 31u;
                        const uint32_t li = blk < V.n_deep ? blk : (uint32_t)P.n_local_tiles * 32u - 1u - (blk - V.n_deep);
                        const uint32_t packed = __ldg(P.active_list + li);
                        const uint32_t ltile = packed >> 5, block = packed & 31u;
                        const uint32_t gtile = (uint32_t)P.shard_rank + ltile * (uint32_t)P.shard_count;
                        const uint32_t tx = gtile % (uint32_t)P.tiles_x, ty = gtile / (uint32_t)P.tiles_x;
                        const uint32_t lx = ((block & 3u) << 3) + (l & 7u), ly = ((block >> 2) << 2) + (l >> 3);
                        const uint32_t x = tx * kTile + lx, y = ty * kTile + ly;
                        valid = (x < (uint32_t)P.width) && (y < (uint32_t)P.height);
                        if (valid) {
                            const uint32_t pixg = y * (uint32_t)P.width + x;
                            float ju = 0.5f, jv = 0.5f;
                            if (P.jitter) {
                                // one Philox block serves two consecutive samples: (x,y) the even one, (z,w) the odd one
                                const uint4 r4 = philox(P.rk, pixg, s >> 1, kStreamJitter, 0u);
                                stat_add<STATS>(st, kStatRngBlocks);
                                ju = (float)(((s & 1u) ? r4.z : r4.x) >> 8) * (1.0f / 16777216.0f);
                                jv = (float)(((s & 1u) ? r4.w : r4.y) >> 8) * (1.0f / 16777216.0f);
                            }
                            const float u = ((float)x + ju) * P.inv_w, v = ((float)y + jv) * P.inv_h;
                            q.dx = fmaf(v, P.cam.vx, fmaf(u, P.cam.hx, P.cam.llx));
                            q.dy = fmaf(v, P.cam.vy, fmaf(u, P.cam.hy, P.cam.lly));
                            q.dz = fmaf(v, P.cam.vz, fmaf(u, P.cam.hz, P.cam.llz));
                            q.ox = P.cam.ox; q.oy = P.cam.oy; q.oz = P.cam.oz;
                            V.qc[nxt][slot0 + my] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(s));
                            V.qd[nxt][slot0 + my] = make_uint2(pixg, ltile * kTilePixels + ly * kTile + lx);
                            stat_add<STATS>(st, kStatPrimary);
                        } else {
                            store_w(V.qa[nxt] + slot0 + my, kDeadPrim);
                        }
                    } else if (SRC == SRC_EXT) {
                        const float4 a = V.qa[nxt][my], b = V.qb[nxt][my];
                        q.ox = a.x; q.oy = a.y; q.oz = a.z;
                        q.dx = b.x; q.dy = b.y; q.dz = b.z;
                    } else {
                        const uint32_t e = SRC == SRC_HARD ? __ldg(V.hard_list + my) : __ldg(V.walk_list + (my >> 4));
                        const uint32_t rec = e & 0x3FFFFFFFu, li = e >> 30;
                        const float4 a = __ldg(V.ra + rec);
                        const float4 L0 = light4<false>(P, V.l0 + (int)li, 0);
                        q.ox = a.x; q.oy = a.y; q.oz = a.z;
                        if (SRC == SRC_HARD) {
                            light_dir(L0, a.x, a.y, a.z, q.dx, q.dy, q.dz, tmax);
                        } else {
                            // calculateSmartShadow's jittered ray k of the pair (renderer.go:313-320): the same Philox block
                            // and bits as the quarter-warp loop of soft_setup / trace_kernel (sample 2j from (x,y), 2j+1 from (z,w))
                            float ax = L0.x - a.x, ay = L0.y - a.y, az = L0.z - a.z;
                            const float dist2 = dot3(ax, ay, az, ax, ay, az);
                            const float inv_d = rsqrt_fast(dist2);
                            tmax = dist2 * inv_d;
                            ax *= inv_d; ay *= inv_d; az *= inv_d;
                            const uint32_t k = my & 15u;
                            const uint32_t sdw = __float_as_uint(__ldg(V.rb + rec).w);
                            const uint4 r4 = philox(P.rk, __ldg(V.rd + rec).x, sdw & 0xffffu, ((sdw >> 16) << 8) | kStreamShadow,
                                                    ((uint32_t)(V.l0 + (int)li) << 12) | ((k >> 1) << 8));
                            if (!(k & 1u)) stat_add<STATS>(st, kStatRngBlocks);
                            stat_add<STATS>(st, kStatSoftRays);
                            float bx, by, bz;
                            ball_from_bits((k & 1u) ? r4.z : r4.x, (k & 1u) ? r4.w : r4.y, bx, by, bz);
                            q.dx = fmaf(0.1f, bx, ax); q.dy = fmaf(0.1f, by, ay); q.dz = fmaf(0.1f, bz, az);
                            normalize3(q.dx, q.dy, q.dz);
                        }
                    }
                    if (valid) {
                        stat_add<STATS>(st, ANY ? kStatShadow : kStatClosest);
                        q.a = dot3(q.dx, q.dy, q.dz, q.dx, q.dy, q.dz);
                        q.inv_a = rcp_fast(q.a);
                        q.tbest = tmax; q.best = 0; q.found = false;
                        const float ooeps = 8.27180613e-25f;  // 2^-80
                        idx = rcp_fast(fabsf(q.dx) > ooeps ? q.dx : copysignf(ooeps, q.dx));
                        idy = rcp_fast(fabsf(q.dy) > ooeps ? q.dy : copysignf(ooeps, q.dy));
                        idz = rcp_fast(fabsf(q.dz) > ooeps ? q.dz : copysignf(ooeps, q.dz));
#if GORT_POOL_QNODES
                        // slab distance of grid plane k: (origin + k cell - o) / d = k (cell / d) + (origin - o) / d.  The node's
                        // 16-bit k becomes the float 2^23 + k by OR-ing it into the mantissa of 2^23 (no conversion instruction);
                        // the 2^23 is taken out again through the additive term.  That term's rounding moves the distance by at
                        // most half a cell: the builder rounds every box outward by two.
                        oodx = (q.ox - S.qox) * idx; oody = (q.oy - S.qoy) * idy; oodz = (q.oz - S.qoz) * idz;
                        idx *= S.qcx; idy *= S.qcy; idz *= S.qcz;
                        oodx = fmaf(8388608.0f, idx, oodx); oody = fmaf(8388608.0f, idy, oody); oodz = fmaf(8388608.0f, idz, oodz);
#else
                        oodx = q.ox * idx; oody = q.oy * idy; oodz = q.oz * idz;
#endif
                        sp = 0;
                        node = root;
                        have = true;
                    }
                }
                wnext += min((uint32_t)__popc(idle), avail);
                continue;
            }
        }
        if (act == 0) break;  // pool exhausted and every lane has finished

        if (STATS) {
            const unsigned vm = __ballot_sync(FULL_MASK, have && node >= 0);
            if (lane == 0 && vm) st.v[kStatWalkWarp0 + SRC] += 32u;
        }
        if (have) {
            [[maybe_unused]] bool fin = false;
#if GORT_POOL_WIDE
            if (node >= 0) {
                // one 4-wide node: two 256-bit loads, four slab tests (two binary visits' worth of boxes in one dependent fetch)
                stat_add<STATS>(st, kStatNodes, 2);
                stat_add<STATS>(st, kStatWalkLane0 + SRC);
                float4 A, B, C, D;
                if (TOP && node >= kTopBase) {
                    const float4* tn = top_s + 4 * (node - kTopBase);
                    A = tn[0]; B = tn[1]; C = tn[2]; D = tn[3];
                } else {
                    ldg8(nodes + 4 * (size_t)node, A, B);
                    ldg8(nodes + 4 * (size_t)node + 2, C, D);
                }
                const uint32_t magic = 0x4B000000u;  // 2^23
#define GORT_QLO(w) __uint_as_float((__float_as_uint(w) & 0xffffu) | magic)
#define GORT_QHI(w) __uint_as_float(__byte_perm(__float_as_uint(w), magic, 0x7632))
#define GORT_SLAB(wx, wy, wz, tn, tf)                                                                                        \
    {                                                                                                                        \
        const float lx = fmaf(GORT_QLO(wx), idx, -oodx), hx = fmaf(GORT_QHI(wx), idx, -oodx);                               \
        const float ly = fmaf(GORT_QLO(wy), idy, -oody), hy = fmaf(GORT_QHI(wy), idy, -oody);                               \
        const float lz = fmaf(GORT_QLO(wz), idz, -oodz), hz = fmaf(GORT_QHI(wz), idz, -oodz);                               \
        tn = fmaxf(fmaxf(fminf(lx, hx), fminf(ly, hy)), fmaxf(fminf(lz, hz), q.tmin));                                      \
        tf = fminf(fminf(fmaxf(lx, hx), fmaxf(ly, hy)), fminf(fmaxf(lz, hz), q.tbest));                                     \
    }
                float tn0, tf0, tn1, tf1, tn2, tf2, tn3, tf3;
                GORT_SLAB(A.x, A.y, A.z, tn0, tf0)
                GORT_SLAB(A.w, B.x, B.y, tn1, tf1)
                GORT_SLAB(B.z, B.w, C.x, tn2, tf2)
                GORT_SLAB(C.y, C.z, C.w, tn3, tf3)
#undef GORT_SLAB
#undef GORT_QLO
#undef GORT_QHI
                const int l0 = __float_as_int(D.x), l1 = __float_as_int(D.y), l2 = __float_as_int(D.z), l3 = __float_as_int(D.w);
                // (1 + 2^-22 widening of the far side keeps the fp32 slab test conservative)
                const bool h0 = tn0 <= tf0 * 1.0000002f && l0 != kWideNoChild, h1 = tn1 <= tf1 * 1.0000002f && l1 != kWideNoChild;
                const bool h2 = tn2 <= tf2 * 1.0000002f && l2 != kWideNoChild, h3 = tn3 <= tf3 * 1.0000002f && l3 != kWideNoChild;
                // the nearest hit child is walked next (closest-hit rays; shadow rays take the first), the others wait on the stack
                int next = kWideNoChild;
                float tnext = 0.f;
#define GORT_PUSH(l, tn) { stack[sp++] = l; }  // (keeping the waiting children nearest-on-top was measured: no gain)
#define GORT_TAKE(h, l, tn)                                              \
    if (h) {                                                             \
        if (next == kWideNoChild) { next = l; tnext = tn; }              \
        else if (!ANY && tn < tnext) { GORT_PUSH(next, tnext) next = l; tnext = tn; } \
        else GORT_PUSH(l, tn)                                            \
    }
                GORT_TAKE(h0, l0, tn0)
                GORT_TAKE(h1, l1, tn1)
                GORT_TAKE(h2, l2, tn2)
                GORT_TAKE(h3, l3, tn3)
#undef GORT_TAKE
#undef GORT_PUSH
                if (next != kWideNoChild) node = next;
#if GORT_POOL_DEFER
                else node = sp ? stack[--sp] : kDoneNode;
            }
            // Leaf tests are deferred: a lane that reaches a leaf parks it and walks on, and the warp tests its parked leaves
            // together — when kLeafBatch lanes have one, or when no lane can go on without it.  Tested on the spot, a leaf costs
            // the whole warp ~45 instructions for the two or three lanes that happen to be at one (a quarter of all issue slots
            // at 8 % lane efficiency).  The closest hit / the occlusion answer is the same; tbest shrinks a few visits later.
            if (node < 0 && node != kDoneNode && pend == 0) {
                pend = node;
                node = sp ? stack[--sp] : kDoneNode;
            }
        }
        {
            const unsigned pm = __ballot_sync(FULL_MASK, have && pend != 0), nm = __ballot_sync(FULL_MASK, have && node >= 0);
            if (pm != 0u && (__popc(pm) >= kLeafBatch || nm == 0u) && have && pend != 0) {
                const uint32_t v = ~(uint32_t)pend;
                const uint32_t start = v & 0x3FFFFFFu;
                const int cnt = (int)((v >> 26) & 15u) + 1;
                if (GEOM == 1) test_spheres<STATS, GEOM>(S, spheres, q, start, cnt, st);
                else if (GEOM == 2) test_tris<STATS, GEOM>(S, tris, q, start, cnt, st);
                else if (((v >> 30) & 1u) == 0) test_spheres<STATS, GEOM>(S, spheres, q, start, cnt, st);
                else test_tris<STATS, GEOM>(S, tris, q, start, cnt, st);
                pend = 0;
                if (ANY && q.found) { node = kDoneNode; sp = 0; }
            }
        }
        if (have) {
            bool fin = node == kDoneNode && pend == 0;
            if (false) {
#else
                else if (sp == 0) fin = true;
                else node = stack[--sp];
            } else {
#endif
#else
            if (node >= 0) {
                stat_add<STATS>(st, kStatNodes);
                stat_add<STATS>(st, kStatWalkLane0 + SRC);
#if GORT_POOL_QNODES
                float4 w0, w1;
                ldg8(nodes + 2 * (size_t)node, w0, w1);
                const uint32_t magic = 0x4B000000u;  // 2^23
#define GORT_QLO(w) __uint_as_float((__float_as_uint(w) & 0xffffu) | magic)
#define GORT_QHI(w) __uint_as_float(__byte_perm(__float_as_uint(w), magic, 0x7632))
                const float c0lox = fmaf(GORT_QLO(w0.x), idx, -oodx), c0hix = fmaf(GORT_QHI(w0.x), idx, -oodx);
                const float c0loy = fmaf(GORT_QLO(w0.y), idy, -oody), c0hiy = fmaf(GORT_QHI(w0.y), idy, -oody);
                const float c0loz = fmaf(GORT_QLO(w0.z), idz, -oodz), c0hiz = fmaf(GORT_QHI(w0.z), idz, -oodz);
                const float c1lox = fmaf(GORT_QLO(w1.x), idx, -oodx), c1hix = fmaf(GORT_QHI(w1.x), idx, -oodx);
                const float c1loy = fmaf(GORT_QLO(w1.y), idy, -oody), c1hiy = fmaf(GORT_QHI(w1.y), idy, -oody);
                const float c1loz = fmaf(GORT_QLO(w1.z), idz, -oodz), c1hiz = fmaf(GORT_QHI(w1.z), idz, -oodz);
#undef GORT_QLO
#undef GORT_QHI
                float4 n3;
                n3.x = w0.w; n3.y = w1.w;
#else
                const float4* np = nodes + 4 * (size_t)node;
                float4 n0, n1, n2, n3;
                ldg8(np, n0, n1);
                ldg8(np + 2, n2, n3);
                const float c0lox = fmaf(n0.x, idx, -oodx), c0hix = fmaf(n0.y, idx, -oodx);
                const float c0loy = fmaf(n0.z, idy, -oody), c0hiy = fmaf(n0.w, idy, -oody);
                const float c0loz = fmaf(n2.x, idz, -oodz), c0hiz = fmaf(n2.y, idz, -oodz);
                const float c1lox = fmaf(n1.x, idx, -oodx), c1hix = fmaf(n1.y, idx, -oodx);
                const float c1loy = fmaf(n1.z, idy, -oody), c1hiy = fmaf(n1.w, idy, -oody);
                const float c1loz = fmaf(n2.z, idz, -oodz), c1hiz = fmaf(n2.w, idz, -oodz);
#endif
                const float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), q.tmin));
                const float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), q.tbest));
                const float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), q.tmin));
                const float t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), q.tbest));
                // 1 + 2^-22 widening of the far side keeps the fp32 slab test conservative
                const bool h0 = t0n <= t0f * 1.0000002f;
                const bool h1 = t1n <= t1f * 1.0000002f;
                int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
                if (h0 && h1) {
                    if (t1n < t0n) {
                        const int tmp = c0;
                        c0 = c1;
                        c1 = tmp;
                    }
                    stack[sp++] = c1;
                    node = c0;
                } else if (h0) {
                    node = c0;
                } else if (h1) {
                    node = c1;
                } else {
                    if (sp == 0) fin = true;
                    else node = stack[--sp];
                }
            } else {
#endif
                const uint32_t v = ~(uint32_t)node;
                const uint32_t start = v & 0x3FFFFFFu;
                const int cnt = (int)((v >> 26) & 15u) + 1;
                if (GEOM == 1) test_spheres<STATS, GEOM>(S, spheres, q, start, cnt, st);
                else if (GEOM == 2) test_tris<STATS, GEOM>(S, tris, q, start, cnt, st);
                else if (((v >> 30) & 1u) == 0) test_spheres<STATS, GEOM>(S, spheres, q, start, cnt, st);
                else test_tris<STATS, GEOM>(S, tris, q, start, cnt, st);
                if ((q.found && ANY) || sp == 0) fin = true;
                else node = stack[--sp];
            }
            if (fin) {
                // ---------------- the ray's answer ----------------
                have = false;
                if (SRC == SRC_PRIMARY) {
                    if (q.found) {
                        // rec.Point = ray.At(t) (sphere.go:43, triangle.go:69); qb.w carries t (the fog factor needs it, see scatter)
                        V.qa[nxt][slot0 + my] = make_float4(fmaf(q.tbest, q.dx, q.ox), fmaf(q.tbest, q.dy, q.oy), fmaf(q.tbest, q.dz, q.oz), __int_as_float(q.best));
                        V.qb[nxt][slot0 + my] = make_float4(q.dx, q.dy, q.dz, q.tbest);
                    } else {
                        store_w(V.qa[nxt] + slot0 + my, kDeadPrim);  // traceRay: miss -> black (renderer.go:171-173)
                        if (P.sky_enabled) {  // sky extension: ... or the sky
                            const float3 c = sky_color(P.sky, q.dx, q.dy, q.dz);
                            add_radiance(P, V.qd[nxt][slot0 + my].y, c.x, c.y, c.z);
                        }
                    }
                } else if (SRC == SRC_EXT) {
                    if (q.found) {
                        V.qa[nxt][my] = make_float4(fmaf(q.tbest, q.dx, q.ox), fmaf(q.tbest, q.dy, q.oy), fmaf(q.tbest, q.dz, q.oz), __int_as_float(q.best));
                    } else {
                        store_w(V.qa[nxt] + my, kDeadPrim);
                        if (P.sky_enabled) {  // sky extension: traceRay(scattered) = sky colour, seen through the path's throughput
                            const float3 c = sky_color(P.sky, q.dx, q.dy, q.dz);
                            const float4 T = V.qc[nxt][my];
                            const float k = P.fog_enabled ? 1.0f - V.qb[nxt][my].w : 1.0f;
                            add_radiance(P, V.qd[nxt][my].y, T.x * c.x * k, T.y * c.y * k, T.z * c.z * k);
                        }
                        if (STATS) {
                            const uint32_t dd = __float_as_uint(V.qc[nxt][my].w) >> 16;  // already depth + 1
                            if (dd >= 5) stat_add<STATS>(st, kStatDepth5);
                            if (dd >= 20) stat_add<STATS>(st, kStatDepth20);
                            if ((int)dd >= P.max_depth) stat_add<STATS>(st, kStatDepthMax);
                        }
                    }
                } else if (SRC == SRC_HARD) {
                    if (!q.found) {
                        const uint32_t e = __ldg(V.hard_list + my);
                        V.lit[(size_t)(e & 0x3FFFFFFFu) * kLC + (e >> 30)] = 1;
                        if (P.soft) {  // a lit pair casts its 16 jittered rays (renderer.go:311): queue its cone walk
                            namespace cg = cooperative_groups;
                            cg::coalesced_group g = cg::coalesced_threads();
                            uint32_t lb = 0;
                            if (g.thread_rank() == 0) lb = atomicAdd(V.ctl + cb + kCtlLit, g.size());
                            V.lit_list[g.shfl(lb, 0) + g.thread_rank()] = e;
                        }
                    }
                } else {
                    if (!q.found) {
                        const uint32_t e = __ldg(V.walk_list + (my >> 4));
                        atomicAdd(V.cnt + (size_t)(e & 0x3FFFFFFFu) * kLC + (e >> 30), 1u);
                    }
                }
            }
        }
    }
    stats_flush<STATS>(P, st);
}

// ---------------------------------------------------------------------------------------------
// scatter: one lane per entry of the current queue (traceRay renderer.go:175-189 for a path that hit):
// hit record, the shade record for calculateDirectLighting, Material.Scatter, and the scattered ray's entry in the
// next queue (whose hit pool_trace<EXT> fills in).  Same code as the EXTEND stage of trace_kernel.
// ---------------------------------------------------------------------------------------------
template <bool STATS, int GEOM>
__global__ void __launch_bounds__(128) stream_scatter_kernel(const __grid_constant__ TraceParams P, const __grid_constant__ StreamView V) {
    const SceneView& S = P.scene;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    Stats st;
    stats_zero<STATS>(st);
    // sorted mode: the live entries, in Morton order of their hit points (stream_launch_sort)
    const bool sorted = V.order != nullptr;
    const uint32_t n_cur = sorted ? V.ctl[kCtlLive] : V.ctl_prev[kCtlNextTotal];
    const int cur = V.cur, nxt = cur ^ 1;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 32u; base < n_cur; base += n_warps * 32u) {
        uint32_t i = base + lane;
        float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
        bool alive = false;
        if (i < n_cur) {
            if (sorted) i = __ldg(V.order + i);
            A = V.qa[cur][i];
            alive = __float_as_uint(A.w) != kDeadPrim;
        }
        const unsigned am = __ballot_sync(FULL_MASK, alive);
        if (am == 0) continue;
        uint32_t rbase = 0;
        if (lane == 0) rbase = atomicAdd(V.ctl + kCtlRec, (unsigned int)__popc(am));
        rbase = __shfl_sync(FULL_MASK, rbase, 0);

        bool cont = false;
        float sx = 0.f, sy = 0.f, sz = 0.f, tr = 0.f, tg = 0.f, tb = 0.f, fog = 0.f;
        uint32_t sd = 0;
        uint2 D = make_uint2(0u, 0u);
        if (alive) {
            stat_add<STATS>(st, kStatShaded);
            const float4 B = V.qb[cur][i], C = V.qc[cur][i];
            D = V.qd[cur][i];
            const float px = A.x, py = A.y, pz = A.z;
            const float dx = B.x, dy = B.y, dz = B.z;
            const int prim = __float_as_int(A.w);
            tr = C.x; tg = C.y; tb = C.z;
            sd = __float_as_uint(C.w);
            const uint32_t depth = sd >> 16, sample = sd & 0xffffu;
            fog = B.w;
            if (depth == 0) {
                // primary entries carry t: exponential fog on the primary-hit distance (extension)
                fog = 0.f;
                if (P.fog_enabled) {
                    const float dist = B.w * sqrt_fast(dot3(dx, dy, dz, dx, dy, dz));
                    fog = 1.0f - expf(-P.fog_density * dist);
                }
            }
            // hit record (sphere.go:42-50, triangle.go:69-73)
            float nx, ny, nz;
            int mat;
            if (GEOM == 1 || (GEOM == 3 && prim >= 0)) {
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
                const uint32_t ss = rbase + (uint32_t)__popc(am & lt_mask);
                V.ra[ss] = make_float4(px, py, pz, __int_as_float(mat));
                V.rb[ss] = make_float4(nx, ny, nz, __uint_as_float(sd));
                V.rc[ss] = make_float4(tr, tg, tb, fog);
                V.rd[ss] = D;
            }

            const float4 m0 = mat4<false>(P, mat, 0), m1 = mat4<false>(P, mat, 1), m2 = mat4<false>(P, mat, 2), m3 = mat4<false>(P, mat, 3);
            const int mtype = __float_as_int(m0.x);
            const uint32_t bs = (depth << 8) | kStreamScatter;
            bool scattered = true;
            float ar = 0.f, ag = 0.f, ab = 0.f;
            // Scatter draws at most one Philox block per hit, counter (pixel, sample, bounce|scatter, 0):
            // Lambertian and rough Metal/Shiny/Mirror turn it into a ball point, Glass/Dielectric use word 0
            const bool rough = (mtype == 2) ? (m1.x > 0.f) : (m1.x > 0.001f);
            uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
            const bool draws = mtype == 0 || (mtype <= 3 && rough) || mtype == 4 || mtype == 5;
            if (draws) {
                rnd = philox(P.rk, D.x, sample, bs, 0u);
                stat_add<STATS>(st, kStatRngBlocks);
            }
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
            cont = scattered && P.recursive && (int)(depth + 1) < P.max_depth;
            const float wr = m2.z;
            tr *= ar * wr; tg *= ag * wr; tb *= ab * wr;
            // exact dead-path test (see trace_kernel)
            const float db = P.dead_bound;
            if (db > 0.f && fabsf(tr) * db < 4.6566e-10f && fabsf(tg) * db < 4.6566e-10f && fabsf(tb) * db < 4.6566e-10f) cont = false;
            if (STATS && !cont) {
                const uint32_t dd = depth + (scattered ? 1u : 0u);
                if (dd >= 5) stat_add<STATS>(st, kStatDepth5);
                if (dd >= 20) stat_add<STATS>(st, kStatDepth20);
                if ((int)dd >= P.max_depth) stat_add<STATS>(st, kStatDepthMax);
            }
        }
        const unsigned cm = __ballot_sync(FULL_MASK, cont);
        if (cm) {
            uint32_t nbase = 0;
            if (lane == 0) nbase = atomicAdd(V.ctl + kCtlNext, (unsigned int)__popc(cm));
            nbase = __shfl_sync(FULL_MASK, nbase, 0);
            if (cont) {
                const uint32_t d = nbase + (uint32_t)__popc(cm & lt_mask);
                V.qa[nxt][d] = make_float4(A.x, A.y, A.z, __uint_as_float(kDeadPrim));  // origin; pool_trace<EXT> writes the hit
                V.qb[nxt][d] = make_float4(sx, sy, sz, fog);
                V.qc[nxt][d] = make_float4(tr, tg, tb, __uint_as_float(sd + 0x10000u));
                V.qd[nxt][d] = D;
            }
        }
    }
    stats_flush<STATS>(P, st);
}

// ---------------------------------------------------------------------------------------------
// pair_setup: every (record, light of the chunk): lightDir / lightDistance (renderer.go:249-254), the back-facing cull
// (cosTheta = 0 zeroes both lighting terms, renderer.go:259-287) and the list of hard shadow rays to cast.
// Light-major within a warp's 32 records, so 32 consecutive rays of the list aim at the same light.
// ---------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(128) stream_pair_setup_kernel(const __grid_constant__ TraceParams P, const __grid_constant__ StreamView V) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    Stats st;
    stats_zero<STATS>(st);
    const uint32_t n_rec = V.ctl[kCtlRec];
    const int cb = kCtlChunk0 + 8 * V.chunk;
    const bool culls = !P.no_cone_cull;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 32u; base < n_rec; base += n_warps * 32u) {
        const uint32_t rec = base + lane;
        const bool valid = rec < n_rec;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), n = a;
        if (valid) { a = __ldg(V.ra + rec); n = __ldg(V.rb + rec); }
        for (int li = 0; li < V.lc; li++) {
            const float4 L0 = light4<false>(P, V.l0 + li, 0);
            float dx, dy, dz, dist;
            light_dir(L0, a.x, a.y, a.z, dx, dy, dz, dist);
            const float ndl = dot3(n.x, n.y, n.z, dx, dy, dz);
            const bool go = valid && !(dist < 0.001f) && (ndl > 0.f || !culls);  // renderer.go:252-254
            if (STATS && valid && !(dist < 0.001f) && !go) stat_add<STATS>(st, kStatBackfacing);
            if (go) {
                stat_add<STATS>(st, kStatPairSetups);
                stat_add<STATS>(st, kStatLightEvals);
            }
            if (valid) {
                V.lit[(size_t)rec * kLC + li] = 0;
                V.cnt[(size_t)rec * kLC + li] = 0u;
            }
            const unsigned gm = __ballot_sync(FULL_MASK, go);
            if (gm) {
                uint32_t hb = 0;
                if (lane == 0) hb = atomicAdd(V.ctl + cb + kCtlHard, (unsigned int)__popc(gm));
                hb = __shfl_sync(FULL_MASK, hb, 0);
                if (go) V.hard_list[hb + (uint32_t)__popc(gm & lt_mask)] = rec | ((uint32_t)li << 30);
            }
        }
    }
    stats_flush<STATS>(P, st);
}

// ---------------------------------------------------------------------------------------------
// pool_cone: the shadow-cone walk of every lit (record, light) pair (soft-shadow candidate culling, device.cuh), with the
// same lane refill as pool_trace — a cone walk ends after a handful of nodes (empty cone) or at the seventh candidate, so
// statically assigned lanes sat idle 28 of 32 (profiles/r2_ncu_pool_c4_v2.txt: 4.4 lanes per instruction).
// A finished pair goes one of three ways (calculateSmartShadow, renderer.go:311-328):
//   no candidate        all 16 jittered rays are unoccluded: cnt = 16, no rays
//   1..kMaxCand         a 32-byte record (pair, count, candidates) for soft_cand
//   more                the walk list: its 16 rays walk the BVH (pool_trace<SOFT>)
// ---------------------------------------------------------------------------------------------
template <bool STATS, int GEOM>
__global__ void __launch_bounds__(128, GORT_POOL_MINB) pool_cone_kernel(const __grid_constant__ TraceParams P, const __grid_constant__ StreamView V) {
    namespace cg = cooperative_groups;
    const SceneView& S = P.scene;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    Stats st;
    stats_zero<STATS>(st);
    const int cb = kCtlChunk0 + 8 * V.chunk;
    const uint32_t total = V.ctl[cb + kCtlLit];
    unsigned int* fetch = V.ctl + cb + kCtlFetchCone;
    if (total == 0) return;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t chunk = total >= n_warps * kPoolChunk ? kPoolChunk : max(32u, ((total / n_warps) + 31u) & ~31u);
    const float4* __restrict__ nodes = S.nodes;

    uint32_t wnext = 0, wend = 0;
    bool pool_done = false, have = false;
    uint32_t e = 0;  // the lane's pair: record | light-in-chunk << 30
    float ox = 0.f, oy = 0.f, oz = 0.f, ax = 0.f, ay = 0.f, az = 0.f, tmax = 0.f, nx = 0.f, ny = 0.f, nz = 0.f, thr = 0.f;
    uint32_t n = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
    int stack[64];
    int sp = 0, node = 0;
    int budget = 0;  // node visits this pair's cone walk may still make (see kConeBudget)

    for (;;) {
        const unsigned act = __ballot_sync(FULL_MASK, have);
        if (!pool_done && __popc(act) <= 32 - kPoolRefill) {
            if (wnext >= wend) {
                uint32_t b = total;
                if (lane == 0 && *reinterpret_cast<volatile unsigned int*>(fetch) < total) b = atomicAdd(fetch, chunk);
                b = __shfl_sync(FULL_MASK, b, 0);
                if (b >= total) {
                    pool_done = true;
                } else {
                    wnext = b;
                    wend = min(b + chunk, total);
                }
            }
            if (!pool_done) {
                const unsigned idle = ~act;
                const uint32_t avail = wend - wnext;
                const uint32_t r = (uint32_t)__popc(idle & lt_mask);
                if (!have && r < avail) {
                    e = V.lit_list[wnext + r];
                    const uint32_t rec = e & 0x3FFFFFFFu;
                    const float4 a = __ldg(V.ra + rec), nn = __ldg(V.rb + rec);
                    const float4 L0 = light4<false>(P, V.l0 + (int)(e >> 30), 0);
                    ox = a.x; oy = a.y; oz = a.z;
                    ax = L0.x - ox; ay = L0.y - oy; az = L0.z - oz;
                    const float dist2 = dot3(ax, ay, az, ax, ay, az);
                    const float inv_d = rsqrt_fast(dist2);
                    ax *= inv_d; ay *= inv_d; az *= inv_d;
                    tmax = dist2 * inv_d;
                    nx = nn.x; ny = nn.y; nz = nn.z;
                    thr = tangent_threshold(dot3(nx, ny, nz, ax, ay, az), ox, oy, oz);
                    bool skip = P.no_cone_cull != 0;  // test switch: every pair's rays walk the BVH
                    if (!skip && P.cone_skip > 0.f) {
                        // how many primitives does the cone hold, at the scene's mean density, up to where its axis leaves the world
                        // box (or reaches the light)?  Far more than a candidate list takes: do not walk the cone at all
                        const float wlx = fmaf(8.0f, S.qcx, S.qox), wly = fmaf(8.0f, S.qcy, S.qoy), wlz = fmaf(8.0f, S.qcz, S.qoz);
                        const float tx = ((ax > 0.f ? fmaf(65520.0f, S.qcx, wlx) : wlx) - ox) * rcp_fast(fabsf(ax) > 1e-12f ? ax : 1e-12f);
                        const float ty = ((ay > 0.f ? fmaf(65520.0f, S.qcy, wly) : wly) - oy) * rcp_fast(fabsf(ay) > 1e-12f ? ay : 1e-12f);
                        const float tz = ((az > 0.f ? fmaf(65520.0f, S.qcz, wlz) : wlz) - oz) * rcp_fast(fabsf(az) > 1e-12f ? az : 1e-12f);
                        const float L = fmaxf(0.f, fminf(fminf(tx, ty), fminf(tz, tmax)));
                        skip = P.cone_skip * L * L * L > 8.0f * (float)kMaxCand;
                    }
                    if (skip) {
                        cg::coalesced_group g = cg::coalesced_threads();
                        uint32_t wb = 0;
                        if (g.thread_rank() == 0) wb = atomicAdd(V.ctl + cb + kCtlWalk, g.size());
                        V.walk_list[g.shfl(wb, 0) + g.thread_rank()] = e;
                    } else {
                        n = 0; sp = 0; node = 0;
                        budget = kConeBudget;
                        have = true;
                    }
                }
                wnext += min((uint32_t)__popc(idle), avail);
                continue;
            }
        }
        if (act == 0) break;

        if (have) {
            bool fin = false, over = false;
            if (node >= 0 && --budget < 0) {
                // A cone walk is one lane's sequential work while the pair's 16 rays walk in parallel: a cone that has not been
                // settled within the budget goes to the walk list like an overflowing one (same answers, bounded latency)
                over = true;
                fin = true;
            } else if (node >= 0) {
                stat_add<STATS>(st, kStatConeTests, 2);
                const float4* np = nodes + 4 * (size_t)node;
                float4 n0, n1, n2, n3;
                ldg8(np, n0, n1);
                ldg8(np + 2, n2, n3);
                const bool h0 = cone_box_hit(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, ox, oy, oz, ax, ay, az, tmax, nx, ny, nz, thr);
                const bool h1 = cone_box_hit(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, ox, oy, oz, ax, ay, az, tmax, nx, ny, nz, thr);
                const int k0 = __float_as_int(n3.x), k1 = __float_as_int(n3.y);
                if (h0 && h1) {
                    stack[sp++] = k1;
                    node = k0;
                } else if (h0) {
                    node = k0;
                } else if (h1) {
                    node = k1;
                } else {
                    if (sp == 0) fin = true;
                    else node = stack[--sp];
                }
            } else {
                const uint32_t v = ~(uint32_t)node;
                const uint32_t start = v & 0x3FFFFFFu;
                const int cnt = (int)((v >> 26) & 15u) + 1;
                const bool is_tri = GEOM == 2 || (GEOM == 3 && ((v >> 30) & 1u) != 0);
                for (int i = 0; i < cnt && !over; i++) {
                    stat_add<STATS>(st, kStatConeTests);
                    if (cone_prim_keep<GEOM>(S, is_tri, start + i, ox, oy, oz, ax, ay, az, tmax, nx, ny, nz, thr)) {
                        const uint32_t ref = (start + i) | (is_tri ? 0x80000000u : 0u);
                        if (n >= (uint32_t)kMaxCand) over = true;
                        else if (n == 0) c0 = ref;
                        else if (n == 1) c1 = ref;
                        else if (n == 2) c2 = ref;
                        else if (n == 3) c3 = ref;
                        else if (n == 4) c4 = ref;
                        else c5 = ref;
                        if (!over) n++;
                    }
                }
                if (over || sp == 0) fin = true;
                else node = stack[--sp];
            }
            if (fin) {
                have = false;
                if (over) {
                    cg::coalesced_group g = cg::coalesced_threads();
                    uint32_t wb = 0;
                    if (g.thread_rank() == 0) wb = atomicAdd(V.ctl + cb + kCtlWalk, g.size());
                    V.walk_list[g.shfl(wb, 0) + g.thread_rank()] = e;
                } else if (n == 0) {
                    // an empty cone: all 16 rays are unoccluded whatever their jitter (shadowFactor = 16/16, renderer.go:326-328)
                    V.cnt[(size_t)(e & 0x3FFFFFFFu) * kLC + (e >> 30)] = 16u;
                    stat_add<STATS>(st, kStatSoftSkipped);
                } else {
                    cg::coalesced_group g = cg::coalesced_threads();
                    uint32_t wb = 0;
                    if (g.thread_rank() == 0) wb = atomicAdd(V.ctl + cb + kCtlCand, g.size());
                    const size_t slot = (size_t)g.shfl(wb, 0) + g.thread_rank();
                    V.cand_recs[2 * slot] = make_uint4(e, n, c0, c1);
                    V.cand_recs[2 * slot + 1] = make_uint4(c2, c3, c4, c5);
                }
            }
        }
    }
    stats_flush<STATS>(P, st);
}

// ---------------------------------------------------------------------------------------------
// soft_cand: the 16 jittered rays (renderer.go:311-328) of the pairs with 1..kMaxCand candidates, tested against the
// candidates only: a quarter warp per pair; lane & 7 = k handles shadow samples 2k and 2k+1 from one Philox block
// (the C phase of trace_kernel's shade stage).
// ---------------------------------------------------------------------------------------------
template <bool STATS, int GEOM>
__global__ void __launch_bounds__(128) stream_soft_cand_kernel(const __grid_constant__ TraceParams P, const __grid_constant__ StreamView V) {
    const SceneView& S = P.scene;
    const int lane = threadIdx.x & 31;
    Stats st;
    stats_zero<STATS>(st);
    const int cb = kCtlChunk0 + 8 * V.chunk;
    const uint32_t n_pairs = V.ctl[cb + kCtlCand];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 4u; base < n_pairs; base += n_warps * 4u) {
        const uint32_t qk = base + (uint32_t)(lane >> 3);
        const bool valid = qk < n_pairs;
        bool unA = false, unB = false;
        uint32_t rec = 0;
        int li = 0;
        if (valid) {
            const uint4 r0 = __ldg(V.cand_recs + 2 * (size_t)qk), r1 = __ldg(V.cand_recs + 2 * (size_t)qk + 1);
            rec = r0.x & 0x3FFFFFFFu;
            li = (int)(r0.x >> 30);
            const uint32_t nc = r0.y;
            const uint32_t cand[kMaxCand] = {r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            const float4 a = __ldg(V.ra + rec);
            const float ox = a.x, oy = a.y, oz = a.z;
            const float4 L0 = light4<false>(P, V.l0 + li, 0);
            float ax = L0.x - ox, ay = L0.y - oy, az = L0.z - oz;
            const float dist2 = dot3(ax, ay, az, ax, ay, az);
            const float inv_d = rsqrt_fast(dist2);
            const float dist = dist2 * inv_d;
            ax *= inv_d; ay *= inv_d; az *= inv_d;
            const uint32_t sdw = __float_as_uint(__ldg(V.rb + rec).w);
            const uint4 rb = philox(P.rk, __ldg(V.rd + rec).x, sdw & 0xffffu, ((sdw >> 16) << 8) | kStreamShadow,
                                    ((uint32_t)(V.l0 + li) << 12) | ((uint32_t)(lane & 7) << 8));
            stat_add<STATS>(st, kStatRngBlocks);
            stat_add<STATS>(st, kStatSoftRays, 2);
            stat_add<STATS>(st, kStatShadow, 2);
            float bx, by, bz;
            ball_from_bits(rb.x, rb.y, bx, by, bz);
            float dxa = fmaf(0.1f, bx, ax), dya = fmaf(0.1f, by, ay), dza = fmaf(0.1f, bz, az);
            normalize3(dxa, dya, dza);
            ball_from_bits(rb.z, rb.w, bx, by, bz);
            float dxb = fmaf(0.1f, bx, ax), dyb = fmaf(0.1f, by, ay), dzb = fmaf(0.1f, bz, az);
            normalize3(dxb, dyb, dzb);
            bool occA = false, occB = false;
#pragma unroll
            for (uint32_t k = 0; k < (uint32_t)kMaxCand; k++) {
                if (k < nc) {
                    const uint32_t ref = cand[k];
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
            unA = !occA;
            unB = !occB;
        }
        const unsigned ua = __ballot_sync(FULL_MASK, unA), ub = __ballot_sync(FULL_MASK, unB);
        if (valid && (lane & 7) == 0) {
            const int sh = lane & 24;
            V.cnt[(size_t)rec * kLC + li] = (unsigned int)(__popc((ua >> sh) & 0xFFu) + __popc((ub >> sh) & 0xFFu));
        }
    }
    stats_flush<STATS>(P, st);
}

// ---------------------------------------------------------------------------------------------
// shade_accum: calculateDirectLighting's arithmetic (renderer.go:236-293) for the lights of the chunk, and after the
// last chunk traceRay's weighting of the hit (renderer.go:177-226): one fixed-point add of T * (emitted + w * direct).
// The D phase and the epilogue of trace_kernel's shade stage, lane = record.
// ---------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(128) stream_shade_accum_kernel(const __grid_constant__ TraceParams P, const __grid_constant__ StreamView V) {
    Stats st;
    stats_zero<STATS>(st);
    const uint32_t n_rec = V.ctl[kCtlRec];
    for (uint32_t rec = blockIdx.x * blockDim.x + threadIdx.x; rec < n_rec; rec += gridDim.x * blockDim.x) {
        const float4 a = __ldg(V.ra + rec), nn = __ldg(V.rb + rec);
        const int mat = __float_as_int(a.w);
        const float4 m0 = mat4<false>(P, mat, 0), m2 = mat4<false>(P, mat, 2);
        const bool is_light = __float_as_int(m0.x) == 6;
        // GetAlbedo: DiffuseLight -> 0 (material.go:304); Dielectric -> 1 (packed by the host)
        const float kar = is_light ? 0.f : m0.y * m2.y, kag = is_light ? 0.f : m0.z * m2.y, kab = is_light ? 0.f : m0.w * m2.y;
        float dr = m2.x, dg = m2.x, db = m2.x;  // ambient (renderer.go:236-246)
        if (V.chunk > 0) {
            const float4 acc = V.racc[rec];
            dr = acc.x; dg = acc.y; db = acc.z;
        }
        const float spec_pow = mat4<false>(P, mat, 3).x;       // 0: metallic <= 0.5, no specular term
        const float spec_w = mat4<false>(P, mat, 1).y * 3.0f;  // metallic * 3
        const float px = a.x, py = a.y, pz = a.z, nx = nn.x, ny = nn.y, nz = nn.z;
        for (int li = 0; li < V.lc; li++) {
            if (!V.lit[(size_t)rec * kLC + li]) continue;
            const float factor = P.soft ? (float)V.cnt[(size_t)rec * kLC + li] * (1.0f / 16.0f) : 1.0f;
            if (!(factor > 0.0f)) continue;
            stat_add<STATS>(st, kStatDiffuse);
            const float4 L0 = light4<false>(P, V.l0 + li, 0);
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
                const float4 L1 = light4<false>(P, V.l0 + li, 1);
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
        if (!V.last_chunk) {
            V.racc[rec] = make_float4(dr, dg, db, 0.f);
            continue;
        }
        // DiffuseLight does not scatter: emitted + direct (renderer.go:182-184); everything else: emitted (0) + w_d * direct
        const float4 T = __ldg(V.rc + rec);
        const float wd = is_light ? 1.0f : m2.w;
        const float er = is_light ? m0.y : 0.f, eg = is_light ? m0.z : 0.f, eb = is_light ? m0.w : 0.f;  // Emitted
        float r = T.x * fmaf(dr, wd, er), g = T.y * fmaf(dg, wd, eg), b = T.z * fmaf(db, wd, eb);
        if (P.fog_enabled) {  // extension: final = (1-f) * radiance + f * fog colour, f from the primary hit
            const float f = T.w;
            r *= 1.0f - f; g *= 1.0f - f; b *= 1.0f - f;
            if ((__float_as_uint(nn.w) >> 16) == 0) { r = fmaf(P.fog_r, f, r); g = fmaf(P.fog_g, f, g); b = fmaf(P.fog_b, f, b); }
        }
        add_radiance(P, __ldg(V.rd + rec).y, r, g, b);
    }
    stats_flush<STATS>(P, st);
}

// ---------------------------------------------------------------------------------------------
// plan: path regeneration.  After scatter the next queue holds kCtlNext survivors; the rest of it is filled with the
// frame's next primary rays, so every launch of the following iteration works on (nearly) `cap` paths until the
// frame's samples run out.
// ---------------------------------------------------------------------------------------------
__global__ void stream_plan_kernel(const StreamView V) {
    const unsigned int n_ext = V.ctl[kCtlNext];
    const unsigned long long done = *V.prim_cursor;
    const unsigned long long left = V.prim_total - done;
    const unsigned int n_new = (unsigned int)min((unsigned long long)(V.cap - n_ext), left);
    V.ctl[kCtlNew] = n_new;
    V.ctl[kCtlNextTotal] = n_ext + n_new;
    V.ctl[kCtlPrimStart] = (unsigned int)done;
    V.ctl[kCtlPrimStart + 1] = (unsigned int)(done >> 32);
    *V.prim_cursor = done + n_new;
}

// ---------------------------------------------------------------------------------------------
// sorted mode: key pass + radix sort of (key, entry) pairs.  Key = 23-bit Morton code of the entry's hit point on the
// BVH's quantisation grid (8 + 8 + 7 bits); a dead entry (miss / ended path) gets bit 23, so the live ones come first
// and scatter stops at their count.  The paths' arithmetic, Philox counters and accumulator adds do not depend on the
// order they are processed in: the frame is the same.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread_bits3(uint32_t v) {  // 8 bits -> every third bit
    v = (v | (v << 8)) & 0x0000F00Fu;
    v = (v | (v << 4)) & 0x000C30C3u;
    v = (v | (v << 2)) & 0x00249249u;
    return v;
}

__global__ void __launch_bounds__(256) stream_sort_keys_kernel(const StreamView V, uint32_t n, float qox, float qoy, float qoz, float icx, float icy, float icz,
                                                               uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const float4* __restrict__ qa = V.qa[V.cur];
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {  // block-uniform trip count
        const uint32_t i = base + threadIdx.x;
        bool live = false;
        if (i < n) {
            const float4 A = qa[i];
            live = __float_as_uint(A.w) != kDeadPrim;
            uint32_t key = 1u << 23;
            if (live) {
                const uint32_t gx = __float2uint_rz(fminf(fmaxf((A.x - qox) * icx, 0.f), 65535.f)) >> 8;
                const uint32_t gy = __float2uint_rz(fminf(fmaxf((A.y - qoy) * icy, 0.f), 65535.f)) >> 8;
                const uint32_t gz = __float2uint_rz(fminf(fmaxf((A.z - qoz) * icz, 0.f), 65535.f)) >> 8;
                key = (spread_bits3(gx) | (spread_bits3(gy) << 1) | (spread_bits3(gz) << 2)) >> 1;
            }
            keys[i] = key;
            idx[i] = i;
        }
        const unsigned lm = __ballot_sync(FULL_MASK, live);
        if ((threadIdx.x & 31) == 0 && lm) atomicAdd(V.ctl + kCtlLive, (unsigned int)__popc(lm));
    }
}

static size_t sort_temp_bytes(uint32_t cap) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), x(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, x, (int)cap, 0, 24);
    return (bytes + 255) / 256 * 256;
}

size_t stream_sort_bytes(uint32_t cap) { return 4 * (((size_t)cap * sizeof(uint32_t) + 255) / 256 * 256) + sort_temp_bytes(cap); }

cudaError_t stream_launch_sort(const TraceParams& p, StreamView& v, uint32_t n_cur, int begin_bit, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    v.order = nullptr;
    if (n_cur == 0) return cudaSuccess;
    if (n_cur > v.cap || begin_bit < 0 || begin_bit > 23 || scratch_bytes < stream_sort_bytes(v.cap)) return cudaErrorInvalidValue;
    const size_t arr = ((size_t)v.cap * sizeof(uint32_t) + 255) / 256 * 256;
    uint8_t* q = (uint8_t*)scratch;
    uint32_t* k0 = (uint32_t*)q; uint32_t* k1 = (uint32_t*)(q + arr);
    uint32_t* x0 = (uint32_t*)(q + 2 * arr); uint32_t* x1 = (uint32_t*)(q + 3 * arr);
    void* tmp = q + 4 * arr;
    size_t tmp_bytes = scratch_bytes - 4 * arr;
    const SceneView& S = p.scene;
    const int grid = (int)std::min<uint32_t>((n_cur + 255u) / 256u, 148u * 32u);
    stream_sort_keys_kernel<<<grid, 256, 0, st>>>(v, n_cur, S.qox, S.qoy, S.qoz, 1.0f / S.qcx, 1.0f / S.qcy, 1.0f / S.qcz, k0, x0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    cub::DoubleBuffer<uint32_t> dk(k0, k1), dx(x0, x1);
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dx, (int)n_cur, begin_bit, 24, st);
    if (e != cudaSuccess) return e;
    v.order = dx.Current();
    return cudaSuccess;
}

cudaError_t stream_launch_plan(const StreamView& v, cudaStream_t st) {
    stream_plan_kernel<<<1, 1, 0, st>>>(v);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
bool stream_wants_wide_nodes() { return GORT_POOL_WIDE != 0; }

size_t stream_bytes_per_slot() {
    return 2 * (3 * sizeof(float4) + sizeof(uint2)) + 4 * sizeof(float4) + sizeof(uint2) + kLC * (1 + 4 + 4 + 4 + 4 + 32);
}

template <int SRC, bool STATS, int GEOM, bool TOP>
static cudaError_t launch_pool_top(const TraceParams& p, const StreamView& v, int sm_count, cudaStream_t st) {
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        // no shared memory in this kernel (TOP: 5.4 KB per CTA): the rest of the SM's carve-out is L1 for the BVH
        int carveout = TOP ? 40 : cudaSharedmemCarveoutMaxL1;
        if (TOP) {
            // "prefer L1" leaves room for ONE CTA's block (and the occupancy query below answers 1): ask for the carve-out
            // that holds the blocks of all the CTAs the register file admits
            cudaFuncAttributes fa;
            int dev = 0, smem_sm = 0;
            if (cudaFuncGetAttributes(&fa, pool_trace_kernel<SRC, STATS, GEOM, TOP>) == cudaSuccess && cudaGetDevice(&dev) == cudaSuccess &&
                cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess && smem_sm > 0) {
                const size_t need = (size_t)GORT_POOL_MINB * (fa.sharedSizeBytes + 1024);  // + the 1 KB the system reserves per CTA
                carveout = (int)std::min<size_t>(100, (need * 100 + (size_t)smem_sm - 1) / (size_t)smem_sm + 1);
            }
        }
        cudaFuncSetAttribute(pool_trace_kernel<SRC, STATS, GEOM, TOP>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pool_trace_kernel<SRC, STATS, GEOM, TOP>, 128, 0);
        if (e != cudaSuccess) return e;
        ctas_per_sm = n > 0 ? n : 1;
        if (getenv("GORT_STREAM_DEBUG")) fprintf(stderr, "pool_trace<%d,%d,%d,%d>: carve-out %d, %d CTAs/SM\n", SRC, (int)STATS, GEOM, (int)TOP, carveout, ctas_per_sm);
    }
    pool_trace_kernel<SRC, STATS, GEOM, TOP><<<sm_count * ctas_per_sm, 128, 0, st>>>(p, v);
    return cudaGetLastError();
}

template <int SRC, bool STATS, int GEOM>
static cudaError_t launch_pool(const TraceParams& p, const StreamView& v, int sm_count, cudaStream_t st) {
#if GORT_POOL_WIDE
    if (!STATS && v.top) return launch_pool_top<SRC, false, GEOM, true>(p, v, sm_count, st);  // (a stats frame reads every node from global memory)
#endif
    return launch_pool_top<SRC, STATS, GEOM, false>(p, v, sm_count, st);
}

template <int SRC>
static cudaError_t launch_pool_variant(const TraceParams& p, const StreamView& v, int geom, bool stats, int sm_count, cudaStream_t st) {
    if (stats) {
        switch (geom) {
            case 1: return launch_pool<SRC, true, 1>(p, v, sm_count, st);
            case 2: return launch_pool<SRC, true, 2>(p, v, sm_count, st);
            default: return launch_pool<SRC, true, 3>(p, v, sm_count, st);
        }
    }
    switch (geom) {
        case 1: return launch_pool<SRC, false, 1>(p, v, sm_count, st);
        case 2: return launch_pool<SRC, false, 2>(p, v, sm_count, st);
        default: return launch_pool<SRC, false, 3>(p, v, sm_count, st);
    }
}

template <bool STATS, int GEOM>
static cudaError_t launch_cone_variant(const TraceParams& p, const StreamView& v, int sm_count, cudaStream_t st) {
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        cudaFuncSetAttribute(pool_cone_kernel<STATS, GEOM>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pool_cone_kernel<STATS, GEOM>, 128, 0);
        if (e != cudaSuccess) return e;
        ctas_per_sm = n > 0 ? n : 1;
    }
    pool_cone_kernel<STATS, GEOM><<<sm_count * ctas_per_sm, 128, 0, st>>>(p, v);
    return cudaGetLastError();
}

static cudaError_t launch_cone(const TraceParams& p, const StreamView& v, int geom, bool stats, int sm_count, cudaStream_t st) {
    if (stats) {
        switch (geom) {
            case 1: return launch_cone_variant<true, 1>(p, v, sm_count, st);
            case 2: return launch_cone_variant<true, 2>(p, v, sm_count, st);
            default: return launch_cone_variant<true, 3>(p, v, sm_count, st);
        }
    }
    switch (geom) {
        case 1: return launch_cone_variant<false, 1>(p, v, sm_count, st);
        case 2: return launch_cone_variant<false, 2>(p, v, sm_count, st);
        default: return launch_cone_variant<false, 3>(p, v, sm_count, st);
    }
}

cudaError_t stream_launch_primary(const TraceParams& p, const StreamView& v, int geom, bool stats, int sm_count, cudaStream_t st) {
    return launch_pool_variant<SRC_PRIMARY>(p, v, geom, stats, sm_count, st);
}

cudaError_t stream_launch_trace_ext(const TraceParams& p, const StreamView& v, int geom, bool stats, int sm_count, cudaStream_t st) {
    return launch_pool_variant<SRC_EXT>(p, v, geom, stats, sm_count, st);
}

#define GORT_GEOM_DISPATCH(KERNEL, GRID)                                                \
    if (stats) {                                                                        \
        if (geom == 1) KERNEL<true, 1><<<GRID, 128, 0, st>>>(p, v);                     \
        else if (geom == 2) KERNEL<true, 2><<<GRID, 128, 0, st>>>(p, v);                \
        else KERNEL<true, 3><<<GRID, 128, 0, st>>>(p, v);                               \
    } else {                                                                            \
        if (geom == 1) KERNEL<false, 1><<<GRID, 128, 0, st>>>(p, v);                    \
        else if (geom == 2) KERNEL<false, 2><<<GRID, 128, 0, st>>>(p, v);               \
        else KERNEL<false, 3><<<GRID, 128, 0, st>>>(p, v);                              \
    }

cudaError_t stream_launch_scatter(const TraceParams& p, const StreamView& v, int geom, bool stats, int sm_count, cudaStream_t st) {
    const int grid = sm_count * 16;
    GORT_GEOM_DISPATCH(stream_scatter_kernel, grid)
    return cudaGetLastError();
}

// one light chunk of calculateDirectLighting for the records of the current bounce
cudaError_t stream_launch_shade_chunk(const TraceParams& p, const StreamView& v, int geom, bool stats, int sm_count, cudaStream_t st) {
    const int grid = sm_count * 16;
    if (stats) stream_pair_setup_kernel<true><<<grid, 128, 0, st>>>(p, v);
    else stream_pair_setup_kernel<false><<<grid, 128, 0, st>>>(p, v);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = launch_pool_variant<SRC_HARD>(p, v, geom, stats, sm_count, st);
    if (e != cudaSuccess) return e;
    if (p.soft) {
        e = launch_cone(p, v, geom, stats, sm_count, st);
        if (e != cudaSuccess) return e;
        GORT_GEOM_DISPATCH(stream_soft_cand_kernel, grid)
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = launch_pool_variant<SRC_SOFT>(p, v, geom, stats, sm_count, st);
        if (e != cudaSuccess) return e;
    }
    if (stats) stream_shade_accum_kernel<true><<<grid, 128, 0, st>>>(p, v);
    else stream_shade_accum_kernel<false><<<grid, 128, 0, st>>>(p, v);
    return cudaGetLastError();
}

}  // namespace gort
