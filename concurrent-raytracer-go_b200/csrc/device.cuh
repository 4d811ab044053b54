// device.cuh — device-side building blocks shared by the sm_100a kernels of libgort (kernels.cu: the persistent
// per-warp-queue kernel; stream.cu: the global-queue wavefront pipeline for large scenes): fast math, Philox4x32-10,
// Sphere.Hit / Triangle.Hit, hitWorld over the flattened BVH, soft-shadow cone culling, fixed-point radiance adds.
#pragma once
#include <cfloat>
#include <cstdint>

#include "kernels.h"

namespace gort {

#define FULL_MASK 0xffffffffu

#ifdef GORT_DEBUG
constexpr bool kDbg = true;
#else
constexpr bool kDbg = false;
#endif

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
struct Stats {
    unsigned int v[kStatCount];
};

template <bool STATS>
__device__ __forceinline__ void stat_add(Stats& st, int idx, unsigned int n = 1) {
    if (STATS) st.v[idx] += n;
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// 256-bit read-only load (sm_100: LDG.E.ENL2.256): two consecutive float4 of a 32-byte-aligned record in ONE request.
// A BVH node is 64 bytes; walking lanes sit on different nodes, and the L1 serves a warp's load one distinct line per
// cycle — so the node fetch costs two passes through the L1 instead of four (the wavefront pipeline's walk was bound by
// exactly that: l1tex data-pipe wavefronts at 90 % of peak, profiles/r2_ncu_pool_c4_v1.txt).
__device__ __forceinline__ void ldg8(const float4* p, float4& a, float4& b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
}

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fmaf(az, bz, fmaf(ay, by, ax * bx));
}

// Single-MUFU approximations (<= 2 ulp).  IEEE-rounded 1/x and sqrt expand to ~10 instructions each
// and were ~18 % of all issued instructions in the first profile (profiles/r1_trace_v0_summary.md);
// fp32 against the float64 reference already differs by more than these 2 ulp.
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_fast(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Vec3.Normalize (vector.go:61-67): zero vector stays zero.
__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {
    const float l2 = dot3(x, y, z, x, y, z);
    const float inv = l2 > 0.f ? rsqrt_fast(l2) : 0.f;
    x *= inv;
    y *= inv;
    z *= inv;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 with host-precomputed round keys (the key schedule depends only on the seed).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox(const uint32_t* __restrict__ rk, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// BVH kernels: one out-of-line copy for the three stages that draw random numbers (the kernel is bound by instruction
// fetch: C2-view loses 7 % per 2.5 KB of code).  The round keys are rebuilt from the seed (they sit in the parameter bank,
// which a non-inlined function can only reach through generic loads).
static __device__ __noinline__ uint4 philox_out(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
template <bool SMALL>
__device__ __forceinline__ uint4 philox_at(const uint32_t* __restrict__ rk, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    if (SMALL) return philox(rk, c0, c1, c2, c3);
    return philox_out(rk[0], rk[1], c0, c1, c2, c3);
}

__device__ __forceinline__ float ex2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_fast(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// RandomVec3InUnitSphere (vector.go:132-139): a uniform point in the open unit ball.  The reference
// rejection-samples the cube; a data-dependent loop costs a warp its slowest lane (first profile: 13
// active lanes per instruction, RNG = 32 % of issued instructions), so the same distribution is drawn
// loop-free from one Philox block: z = 1-2u1, phi = 2 pi u2, r = cbrt(u3) (oracle.cpp Rng::in_unit_sphere).
__device__ __forceinline__ void ball_from_block(const uint4 r, float& bx, float& by, float& bz) {
    const float k = 1.0f / 16777216.0f;
    const float u1 = (float)(r.x >> 8) * k, u2 = (float)(r.y >> 8) * k, u3 = (float)(r.z >> 8) * k;
    const float z = fmaf(-2.0f, u1, 1.0f);
    const float sxy = sqrt_fast(fmaxf(0.f, fmaf(-z, z, 1.0f)));
    float sn, cs;
    __sincosf(6.2831853071795864769f * u2, &sn, &cs);
    const float rad = ex2_fast(lg2_fast(u3) * (1.0f / 3.0f));  // cbrt; u3 = 0 -> 0
    const float rs = rad * sxy;
    bx = rs * cs;
    by = rs * sn;
    bz = rad * z;
}

// ---------------------------------------------------------------------------------------------
// hitWorld (renderer.go:333-346) over the flattened BVH.  ANY = boolean query (shadow rays);
// otherwise closest hit with the linear scan's tie rule (equal t: later scan order wins).
// prim: sphere -> leaf-order index; triangle -> index | 0x80000000.
// ---------------------------------------------------------------------------------------------
// GEOM (BVH kernels are specialised by what the scene holds, see trace_kernel): 1 spheres only, 2 triangles only, 3 both
template <int GEOM = 3>
__device__ __forceinline__ int prim_order(const SceneView& S, int prim) {
    if (GEOM == 1 || (GEOM == 3 && prim >= 0)) return __ldg(&S.sphere_meta[prim]).y;
    return __float_as_int(ldg4(S.tris + 4 * (size_t)(prim & 0x7fffffff) + 1).w);
}

struct RayQuery {
    float ox, oy, oz, dx, dy, dz;
    float a, inv_a;      // |d|^2 and its reciprocal (ray.Direction.LengthSquared(), sphere.go:24)
    float tmin, tbest;   // tbest starts at tMax and shrinks (closestT, renderer.go:335-341)
    int best;
    bool found;
};

// ---- Sphere.Hit (geometry/sphere.go:22-59) for spheres [start, start+cnt) ----
template <bool STATS, int GEOM = 3>
__device__ __forceinline__ void test_spheres(const SceneView& S, const float4* __restrict__ spheres, RayQuery& q, uint32_t start, int cnt, Stats& st) {
    for (int i = 0; i < cnt; i++) {
        stat_add<STATS>(st, kStatSphereTests);
        const float4 s = ldg4(spheres + start + i);
        const float ocx = q.ox - s.x, ocy = q.oy - s.y, ocz = q.oz - s.z;
        const float hb = dot3(ocx, ocy, ocz, q.dx, q.dy, q.dz);
        // discriminant/a from the component of oc perpendicular to the ray: the same quantity as
        // halfB^2 - a*c (sphere.go:28) without fp32 cancellation.
        const float k = hb * q.inv_a;
        const float lx = fmaf(-k, q.dx, ocx), ly = fmaf(-k, q.dy, ocy), lz = fmaf(-k, q.dz, ocz);
        const float dn = fmaf(s.w, s.w, -dot3(lx, ly, lz, lx, ly, lz));
        if (dn < 0.f) continue;
        const float sq = sqrt_fast(dn * q.a);
        float root = (-hb - sq) * q.inv_a;
        if (root < q.tmin || q.tbest < root) {
            root = (-hb + sq) * q.inv_a;
            if (root < q.tmin || q.tbest < root) continue;
        }
        stat_add<STATS>(st, kStatSphereHits);
        const int pr = (int)(start + i);
        if (root == q.tbest && q.found) {
            if (prim_order<GEOM>(S, pr) < prim_order<GEOM>(S, q.best)) continue;
        }
        q.tbest = root;
        q.best = pr;
        q.found = true;
    }
}

// ---- Triangle.Hit (geometry/triangle.go:36-88), Moller-Trumbore, for triangles [start, start+cnt) ----
template <bool STATS, int GEOM = 3>
__device__ __forceinline__ void test_tris(const SceneView& S, const float4* __restrict__ tris, RayQuery& q, uint32_t start, int cnt, Stats& st) {
    for (int i = 0; i < cnt; i++) {
        stat_add<STATS>(st, kStatTriTests);
        const float4* tp = tris + 4 * (size_t)(start + i);
        const float4 v0 = ldg4(tp), e1 = ldg4(tp + 1), e2 = ldg4(tp + 2);
        const float hx = q.dy * e2.z - q.dz * e2.y, hy = q.dz * e2.x - q.dx * e2.z, hz = q.dx * e2.y - q.dy * e2.x;
        const float aa = dot3(e1.x, e1.y, e1.z, hx, hy, hz);
        if (aa > -1e-6f && aa < 1e-6f) { stat_add<STATS>(st, kStatTriRejA); continue; }
        const float f = rcp_fast(aa);
        const float sx = q.ox - v0.x, sy = q.oy - v0.y, sz = q.oz - v0.z;
        const float u = f * dot3(sx, sy, sz, hx, hy, hz);
        if (u < 0.0f || u > 1.0f) { stat_add<STATS>(st, kStatTriRejU); continue; }
        const float qx = sy * e1.z - sz * e1.y, qy = sz * e1.x - sx * e1.z, qz = sx * e1.y - sy * e1.x;
        const float vv = f * dot3(q.dx, q.dy, q.dz, qx, qy, qz);
        if (vv < 0.0f || u + vv > 1.0f) { stat_add<STATS>(st, kStatTriRejV); continue; }
        const float t = f * dot3(e2.x, e2.y, e2.z, qx, qy, qz);
        if (t < q.tmin || t > q.tbest) { stat_add<STATS>(st, kStatTriRejT); continue; }
        stat_add<STATS>(st, kStatTriHits);
        const int pr = (int)((start + i) | 0x80000000u);
        if (t == q.tbest && q.found) {
            if (prim_order<GEOM>(S, pr) < prim_order<GEOM>(S, q.best)) continue;
        }
        q.tbest = t;
        q.best = pr;
        q.found = true;
    }
}

// hitWorld over the BVH.  `any` (per lane, data not code: closest-hit and shadow rays share every
// instruction of a batch) ends the walk at the first accepted primitive — exactly how the renderer
// uses hitWorld for shadows (renderer.go:305,320: only `hit` is read).
// One copy per kernel (__noinline__): inlined at its five call sites the walk was 42 % of an 80 KB kernel,
// more than the 32 KB instruction cache holds; warps sit in different stages, so they thrashed it
// (profiles/r1_ncu_trace_c2view_v5.txt: stall_no_instruction 5.2 per issue).
template <bool STATS, int GEOM = 3>
__device__ __noinline__ bool traverse(const SceneView& S, float ox, float oy, float oz, float dx, float dy, float dz,
                                      float tmin, float tmax, bool any, float& t_out, int& prim_out, Stats& st) {
    stat_add<STATS>(st, any ? kStatShadow : kStatClosest);
    if (S.n_nodes == 0) return false;
    RayQuery q;
    q.ox = ox; q.oy = oy; q.oz = oz; q.dx = dx; q.dy = dy; q.dz = dz;
    q.a = dot3(dx, dy, dz, dx, dy, dz);
    q.inv_a = rcp_fast(q.a);
    q.tmin = tmin; q.tbest = tmax; q.best = 0; q.found = false;
    const float ooeps = 8.27180613e-25f;  // 2^-80
    const float idx = rcp_fast(fabsf(dx) > ooeps ? dx : copysignf(ooeps, dx));
    const float idy = rcp_fast(fabsf(dy) > ooeps ? dy : copysignf(ooeps, dy));
    const float idz = rcp_fast(fabsf(dz) > ooeps ? dz : copysignf(ooeps, dz));
    const float oodx = ox * idx, oody = oy * idy, oodz = oz * idz;

    // S sits in the kernel's parameter bank behind a reference (this function is not inlined): read the array
    // pointers once, not once per visit — a dependent load ahead of every node fetch otherwise
    const float4* __restrict__ nodes = S.nodes;
    const float4* __restrict__ spheres = GEOM != 2 ? S.spheres : nullptr;
    const float4* __restrict__ tris = GEOM != 1 ? S.tris : nullptr;
    int stack[64];
    int sp = 0;
    int node = 0;

    for (;;) {
        if (node >= 0) {
            stat_add<STATS>(st, kStatNodes);
            const float4* np = nodes + 4 * (size_t)node;
            const float4 n0 = ldg4(np), n1 = ldg4(np + 1), n2 = ldg4(np + 2), n3 = ldg4(np + 3);
            const float c0lox = fmaf(n0.x, idx, -oodx), c0hix = fmaf(n0.y, idx, -oodx);
            const float c0loy = fmaf(n0.z, idy, -oody), c0hiy = fmaf(n0.w, idy, -oody);
            const float c0loz = fmaf(n2.x, idz, -oodz), c0hiz = fmaf(n2.y, idz, -oodz);
            const float c1lox = fmaf(n1.x, idx, -oodx), c1hix = fmaf(n1.y, idx, -oodx);
            const float c1loy = fmaf(n1.z, idy, -oody), c1hiy = fmaf(n1.w, idy, -oody);
            const float c1loz = fmaf(n2.z, idz, -oodz), c1hiz = fmaf(n2.w, idz, -oodz);
            const float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tmin));
            const float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), q.tbest));
            const float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tmin));
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
                if (sp == 0) break;
                node = stack[--sp];
            }
        } else {
            const uint32_t v = ~(uint32_t)node;
            const uint32_t start = v & 0x3FFFFFFu;
            const int cnt = (int)((v >> 26) & 15u) + 1;
            if (GEOM == 1) test_spheres<STATS, GEOM>(S, spheres, q, start, cnt, st);
            else if (GEOM == 2) test_tris<STATS, GEOM>(S, tris, q, start, cnt, st);
            else if (((v >> 30) & 1u) == 0) test_spheres<STATS, GEOM>(S, spheres, q, start, cnt, st);
            else test_tris<STATS, GEOM>(S, tris, q, start, cnt, st);
            if ((q.found && any) || sp == 0) break;
            node = stack[--sp];
        }
    }
    t_out = q.tbest;
    prim_out = q.best;
    return q.found;
}

// ---------------------------------------------------------------------------------------------
// tiny sphere-only scenes (<= kSmallMax spheres, no triangles): the reference's own linear scan
// (hitWorld renderer.go:337-343), fully unrolled, with the spheres read straight from the kernel
// parameter bank (constant-bank operands: no load, no address arithmetic).  All tests of a ray are
// independent instruction streams; the closest-hit reduction keeps the scan order, which makes `<=`
// the reference's last-wins tie rule.  A shadow query is the same code: "some root lies in
// [tMin, tMax]" is exactly `found`.
// ---------------------------------------------------------------------------------------------
// MASKED: only the spheres whose bit is set in `mask` are tested (the shadow-cone candidates of a (hit, light) pair), in
// scan order and with the same arithmetic, so the answer equals the full scan's whenever the spheres left out cannot be hit.
template <bool STATS, bool MASKED>
__device__ __forceinline__ bool small_query(const TraceParams& P, uint32_t mask, float ox, float oy, float oz, float dx, float dy, float dz, float tmin,
                                            float tmax, bool any, float& t_out, int& prim_out, Stats& st) {
    stat_add<STATS>(st, any ? kStatShadow : kStatClosest);
    const float a = dot3(dx, dy, dz, dx, dy, dz);
    const float inv_a = rcp_fast(a);
    float tbest = tmax;
    int best = -1;
    // rolled on purpose: unrolled 12x at three call sites this loop was 30 % of the kernel's code and the
    // kernel did not fit the instruction cache (profiles/r1_ncu_trace_c1view_v5.txt)
#pragma unroll 1
    for (int k = 0; MASKED ? (mask != 0u) : (k < P.small_n); k++) {
        int i = k;
        if (MASKED) {
            i = __ffs(mask) - 1;
            mask &= mask - 1u;
        }
        {
            stat_add<STATS>(st, kStatSphereTests);
            const float4 s = P.small_sph[i];
            const float ocx = ox - s.x, ocy = oy - s.y, ocz = oz - s.z;
            const float hb = dot3(ocx, ocy, ocz, dx, dy, dz);
            const float k2 = hb * inv_a;
            const float lx = fmaf(-k2, dx, ocx), ly = fmaf(-k2, dy, ocy), lz = fmaf(-k2, dz, ocz);
            const float dn = fmaf(s.w, s.w, -dot3(lx, ly, lz, lx, ly, lz));  // discriminant / a (sphere.go:28)
            if (dn >= 0.f) {  // most tests miss: the roots are only worked out for the few that do not
                const float sq = sqrt_fast(dn * a);
                const float r0 = (-hb - sq) * inv_a, r1 = (-hb + sq) * inv_a;
                // sphere.go:35-40 with tMax = closestT: the near root if it is >= tMin, else the far root
                const float cand = (r0 < tmin) ? r1 : r0;
                const bool h = !(cand < tmin || tbest < cand);
                if (STATS && h) stat_add<STATS>(st, kStatSphereHits);
                tbest = h ? cand : tbest;
                best = h ? i : best;
            }
        }
    }
    t_out = tbest;
    prim_out = best;
    return best >= 0;
}

// STATS only: account the node visits a lane made since `before` to call site `site` (see kStatWalkLane0)
template <bool STATS>
__device__ __forceinline__ void walk_account(Stats& st, unsigned int before, int site) {
    if (STATS) {
        const unsigned int d = st.v[kStatNodes] - before;
        const unsigned int m = __activemask();
        const unsigned int mx = __reduce_max_sync(m, d);
        st.v[kStatWalkLane0 + site] += d;
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) st.v[kStatWalkWarp0 + site] += 32u * mx;
    }
}

template <bool STATS, bool SMALL, int GEOM = 3>
__device__ __forceinline__ bool query(const TraceParams& P, float ox, float oy, float oz, float dx, float dy, float dz, float tmin,
                                      float tmax, bool any, float& t_out, int& prim_out, Stats& st) {
    if (SMALL) return small_query<STATS, false>(P, 0u, ox, oy, oz, dx, dy, dz, tmin, tmax, any, t_out, prim_out, st);
    return traverse<STATS, GEOM>(P.scene, ox, oy, oz, dx, dy, dz, tmin, tmax, any, t_out, prim_out, st);
}

// material / light records: tiny scenes read them from the kernel parameter bank (no memory latency on
// the dependent chain of a deep path), BVH scenes from global memory through L1
template <bool SMALL>
__device__ __forceinline__ float4 mat4(const TraceParams& P, int mat, int k) {
    if (SMALL) return P.small_mats[mat][k];
    return ldg4(P.scene.mats + 4 * (size_t)mat + k);
}
template <bool SMALL>
__device__ __forceinline__ float4 light4(const TraceParams& P, int light, int k) {
    if (SMALL) return P.small_lights[light][k];
    return ldg4(P.scene.lights + 2 * (size_t)light + k);
}

// ---------------------------------------------------------------------------------------------
// Soft-shadow candidate culling.
//
// calculateSmartShadow casts 16 rays normalize(L + 0.1 * ball) from one hit point toward one light
// (renderer.go:311-328).  |ball| < 1, so every one of them lies inside the cone of half angle
// asin(0.1) about L, and only t in [0.001, lightDistance] counts.  A primitive that no ray of that
// cone can reach within that range cannot occlude any of the 16: it is dropped ONCE per (hit, light)
// pair and the 16 rays test only the survivors (most pairs keep 0-2 primitives).  The answer of every
// ray is unchanged; only tests that must fail are skipped.
// ---------------------------------------------------------------------------------------------
constexpr float kConeSin = 0.1f;
constexpr float kConeCos = 0.99498743710662f;  // sqrt(1 - 0.1^2)

// v = centre - apex, r >= 0, a = unit axis.  Conservative: false only if no cone ray can hit.
__device__ __forceinline__ bool cone_sphere_candidate(float vx, float vy, float vz, float r, float ax, float ay, float az, float tmax) {
    const float dv2 = dot3(vx, vy, vz, vx, vy, vz);
    const float inv_dv = rsqrt_fast(fmaxf(dv2, 1e-30f));
    const float dv = dv2 * inv_dv;
    const float ca = dot3(ax, ay, az, vx, vy, vz) * inv_dv;  // cos(angle between the axis and the centre)
    // exterior apex: the sphere subtends asin(r/dv); it meets the cone iff angle <= asin(0.1) + asin(r/dv)
    const float sp = fminf(1.0f, r * inv_dv);
    const float cp = sqrt_fast(fmaxf(0.f, fmaf(-sp, sp, 1.0f)));
    const bool ext = ca >= fmaf(kConeCos, cp, -kConeSin * sp) - 1e-4f;
    // apex on the surface up to rounding (the hit point's own sphere): a ray leaving the surface outward
    // at cos >= delta can only "hit" at t <= shell/delta < tMin = 0.001, which Sphere.Hit rejects (sphere.go:35).
    // -ca = axis . outward normal; the least outward cone ray has cos >= -ca - 0.105.
    const float shell = 2e-5f + 1e-6f * (r + dv);
    const bool sur = !(-ca - 0.105f > fmaxf(0.03f, 1000.0f * shell));
    bool cand = (dv < r - shell) ? true : ((dv <= r + shell) ? sur : ext);
    // entirely beyond the light
    if (dv - r > fmaf(tmax, 1.00001f, 1e-5f)) cand = false;
    return cand;
}

// Sphere.Hit (sphere.go:22-59) as a boolean for a UNIT direction: some root in [tmin, tmax]
__device__ __forceinline__ bool sphere_occludes_unit(const float4 s, float ox, float oy, float oz, float dx, float dy, float dz, float tmin,
                                                     float tmax) {
    const float ocx = ox - s.x, ocy = oy - s.y, ocz = oz - s.z;
    const float hb = dot3(ocx, ocy, ocz, dx, dy, dz);
    const float lx = fmaf(-hb, dx, ocx), ly = fmaf(-hb, dy, ocy), lz = fmaf(-hb, dz, ocz);
    const float dn = fmaf(s.w, s.w, -dot3(lx, ly, lz, lx, ly, lz));
    if (dn < 0.f) return false;
    const float sq = sqrt_fast(dn);
    const float r0 = -hb - sq, r1 = -hb + sq;
    return !(r0 < tmin || tmax < r0) || !(r1 < tmin || tmax < r1);
}

// Triangle.Hit (triangle.go:36-88) as a boolean
__device__ __forceinline__ bool tri_occludes(const float4* __restrict__ tp, float ox, float oy, float oz, float dx, float dy, float dz,
                                             float tmin, float tmax) {
    const float4 v0 = ldg4(tp), e1 = ldg4(tp + 1), e2 = ldg4(tp + 2);
    const float hx = dy * e2.z - dz * e2.y, hy = dz * e2.x - dx * e2.z, hz = dx * e2.y - dy * e2.x;
    const float aa = dot3(e1.x, e1.y, e1.z, hx, hy, hz);
    if (aa > -1e-6f && aa < 1e-6f) return false;
    const float f = rcp_fast(aa);
    const float sx = ox - v0.x, sy = oy - v0.y, sz = oz - v0.z;
    const float u = f * dot3(sx, sy, sz, hx, hy, hz);
    if (u < 0.0f || u > 1.0f) return false;
    const float qx = sy * e1.z - sz * e1.y, qy = sz * e1.x - sx * e1.z, qz = sx * e1.y - sy * e1.x;
    const float vv = f * dot3(dx, dy, dz, qx, qy, qz);
    if (vv < 0.0f || u + vv > 1.0f) return false;
    const float t = f * dot3(e2.x, e2.y, e2.z, qx, qy, qz);
    return !(t < tmin || t > tmax);
}

// Tangent-plane pruning of shadow-cone candidates.  With n = hit.Normal and a = the unit direction to the light, every
// ray of the pair's cone has d.n >= (a.n - 0.1) / 1.1 =: mu.  If a.n > 0.105 all of them leave the surface, and a point they
// reach at t >= tMin = 0.001 lies at height (p - o).n >= 0.001 * mu above the tangent plane: a primitive whose every point is
// lower cannot occlude the pair (the other faces of a convex object the hit point lies on, everything behind a wall).
// Returns the height below which a primitive is dropped (minus a rounding allowance for fp32 coordinates); -inf = keep all.
__device__ __forceinline__ float tangent_threshold(float a_dot_n, float ox, float oy, float oz) {
    if (!(a_dot_n > 0.105f)) return -__int_as_float(0x7f800000);
    return 0.001f * (a_dot_n - 0.1f) * (1.0f / 1.1f) - 2e-5f * (1.0f + fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz))));
}

constexpr int kMaxCand = 6;           // candidate primitives kept per (hit, light) pair on BVH scenes
constexpr uint32_t kCandOverflow = 0xFFu;  // more than kMaxCand: the pair's rays walk the BVH themselves

// One child box of a BVH node against the shadow cone of a (hit, light) pair (apex o, unit axis a, range tmax): the box is
// tested through its bounding sphere, and dropped when it lies wholly below the hit point's tangent plane (thr).
__device__ __forceinline__ bool cone_box_hit(float lox, float hix, float loy, float hiy, float loz, float hiz, float ox, float oy, float oz, float ax,
                                             float ay, float az, float tmax, float nx, float ny, float nz, float thr) {
    const float ex = 0.5f * (hix - lox), ey = 0.5f * (hiy - loy), ez = 0.5f * (hiz - loz);
    const float vx = fmaf(0.5f, hix + lox, -ox), vy = fmaf(0.5f, hiy + loy, -oy), vz = fmaf(0.5f, hiz + loz, -oz);
    const float rb2 = dot3(ex, ey, ez, ex, ey, ez);
    const float rb = rb2 * rsqrt_fast(fmaxf(rb2, 1e-30f)) * 1.00001f + 1e-6f;
    const float dv2 = dot3(vx, vy, vz, vx, vy, vz);
    const float inv_dv = rsqrt_fast(fmaxf(dv2, 1e-30f));
    const float dv = dv2 * inv_dv;
    const float ca = dot3(ax, ay, az, vx, vy, vz) * inv_dv;
    const float sphi = fminf(1.0f, rb * inv_dv);
    const float cphi = sqrt_fast(fmaxf(0.f, fmaf(-sphi, sphi, 1.0f)));
    bool hit = (dv <= rb) || (ca >= fmaf(kConeCos, cphi, -kConeSin * sphi) - 1e-4f);
    if (dv - rb > fmaf(tmax, 1.00001f, 1e-5f)) hit = false;
    if (!(ex >= 0.f)) hit = false;  // inverted box = empty child
    // highest point of the box above the hit point's tangent plane (see tangent_threshold)
    const float top = fmaxf(nx * (lox - ox), nx * (hix - ox)) + fmaxf(ny * (loy - oy), ny * (hiy - oy)) + fmaxf(nz * (loz - oz), nz * (hiz - oz));
    if (top < thr) hit = false;
    return hit;
}

// One leaf primitive against the cone: false only if no ray of the cone can hit it within [tMin, tmax].
template <int GEOM = 3>
__device__ __forceinline__ bool cone_prim_keep(const SceneView& S, bool is_tri, uint32_t idx, float ox, float oy, float oz, float ax, float ay, float az,
                                               float tmax, float nx, float ny, float nz, float thr) {
    bool keep = true;
    if (!is_tri) {
        const float4 s = ldg4(S.spheres + idx);
        keep = cone_sphere_candidate(s.x - ox, s.y - oy, s.z - oz, fabsf(s.w), ax, ay, az, tmax);
        if (dot3(nx, ny, nz, s.x - ox, s.y - oy, s.z - oz) + fabsf(s.w) < thr) keep = false;  // behind the tangent plane
    } else {
        // the triangle's plane: a cone whose every ray moves away from it (or crosses it below
        // tMin when the apex lies on it — the hit point's own face) cannot hit the triangle
        const float4* tp = S.tris + 4 * (size_t)idx;
        const float4 v0 = ldg4(tp), nn = ldg4(tp + 3);
        const float hgt = dot3(nn.x, nn.y, nn.z, ox - v0.x, oy - v0.y, oz - v0.z);
        const float x = dot3(nn.x, nn.y, nn.z, ax, ay, az);  // n . d ranges over [x - 0.105, x + 0.105]
        const float eps = 2e-5f + 1e-6f * (fabsf(ox) + fabsf(oy) + fabsf(oz) + fabsf(v0.x) + fabsf(v0.y) + fabsf(v0.z));
        if (hgt > eps) keep = !(x - 0.105f >= 0.f);
        else if (hgt < -eps) keep = !(x + 0.105f <= 0.f);
        else keep = !(fabsf(x) - 0.105f > fmaxf(0.03f, 1000.0f * eps));
        if (keep && thr > -3.0e38f) {
            // all three vertices below the tangent-plane threshold (see tangent_threshold)
            const float4 e1 = ldg4(tp + 1), e2 = ldg4(tp + 2);
            const float h0 = dot3(nx, ny, nz, v0.x - ox, v0.y - oy, v0.z - oz);
            const float h1 = h0 + dot3(nx, ny, nz, e1.x, e1.y, e1.z), h2 = h0 + dot3(nx, ny, nz, e2.x, e2.y, e2.z);
            if (fmaxf(h0, fmaxf(h1, h2)) < thr) keep = false;
        }
    }
    return keep;
}

// Walk the BVH with the cone (apex o, unit axis a, range tmax); boxes are tested through their bounding
// spheres.  Writes up to kMaxCand primitive references (sphere: index; triangle: index | 0x80000000) and
// returns their number, or kCandOverflow.
template <bool STATS, int GEOM = 3>
__device__ __forceinline__ uint32_t cone_candidates(const SceneView& S, float ox, float oy, float oz, float ax, float ay, float az, float tmax,
                                                    float nx, float ny, float nz, float thr, uint32_t* __restrict__ out, Stats& st) {
    if (S.n_nodes == 0) return 0;
    int stack[64];
    int sp = 0;
    int node = 0;
    uint32_t n = 0;
    for (;;) {
        if (node >= 0) {
            stat_add<STATS>(st, kStatConeTests, 2);
            const float4* np = S.nodes + 4 * (size_t)node;
            const float4 n0 = ldg4(np), n1 = ldg4(np + 1), n2 = ldg4(np + 2), n3 = ldg4(np + 3);
            const bool h0 = cone_box_hit(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, ox, oy, oz, ax, ay, az, tmax, nx, ny, nz, thr);
            const bool h1 = cone_box_hit(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, ox, oy, oz, ax, ay, az, tmax, nx, ny, nz, thr);
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
            const bool is_tri = GEOM == 2 || (GEOM == 3 && ((v >> 30) & 1u) != 0);
            for (int i = 0; i < cnt; i++) {
                stat_add<STATS>(st, kStatConeTests);
                if (cone_prim_keep<GEOM>(S, is_tri, start + i, ox, oy, oz, ax, ay, az, tmax, nx, ny, nz, thr)) {
                    if (n >= (uint32_t)kMaxCand) return kCandOverflow;
                    out[n++] = (start + i) | (is_tri ? 0x80000000u : 0u);
                }
            }
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
    return n;
}

// tracePixel's color.Add (renderer.go:159) into the pixel's fixed-point accumulators.  Integer adds
// commute, so the sum is independent of the schedule.  NaN contributions are dropped (a NaN sample
// makes the reference's pixel NaN -> undefined uint8).
__device__ __forceinline__ void add_radiance(const TraceParams& P, uint32_t pixl, float r, float g, float b) {
    unsigned long long* acc = P.accum + 3 * (size_t)pixl;
    const float scale = (float)(1u << kAccumFracBits);
    if (r != 0.f && r == r) atomicAdd(acc + 0, (unsigned long long)__float2ll_rn(fminf(fmaxf(r, -kSampleClamp), kSampleClamp) * scale));
    if (g != 0.f && g == g) atomicAdd(acc + 1, (unsigned long long)__float2ll_rn(fminf(fmaxf(g, -kSampleClamp), kSampleClamp) * scale));
    if (b != 0.f && b == b) atomicAdd(acc + 2, (unsigned long long)__float2ll_rn(fminf(fmaxf(b, -kSampleClamp), kSampleClamp) * scale));
}

// AtmosphereConfig.GetSkyColor (atmosphere/atmosphere.go:100-135; oracle.cpp sky_color), extension: the colour of a ray that
// leaves the scene.  Out of line: it runs once per finished path, and the trace kernels are bound by their code footprint.
static __device__ __noinline__ float3 sky_color(const float* __restrict__ q, float dx, float dy, float dz) {
    normalize3(dx, dy, dz);
    const float t = 0.5f * (dy + 1.0f);
    float r = fmaf(q[0], t, q[3] * (1.0f - t)), g = fmaf(q[1], t, q[4] * (1.0f - t)), b = fmaf(q[2], t, q[5] * (1.0f - t));
    const float atm = __expf(-fmaxf(0.f, dy) * q[20]);
    const float sr = fmaf(q[17], atm, q[14] * (1.0f - atm)), sg = fmaf(q[18], atm, q[15] * (1.0f - atm)), sb = fmaf(q[19], atm, q[16] * (1.0f - atm));
    r = fmaf(sr, 0.25f, r * 0.75f); g = fmaf(sg, 0.25f, g * 0.75f); b = fmaf(sb, 0.25f, b * 0.75f);
    const float sun_dot = dot3(dx, dy, dz, q[6], q[7], q[8]);
    if (sun_dot > 1.0f - q[13]) {
        const float x = (sun_dot - (1.0f - q[13])) / q[13];
        const float si = fminf(x * sqrt_fast(x), 1.0f) * q[12] * 0.9f;  // pow(x, 1.5)
        r = fmaf(q[9], si, r * (1.0f - si)); g = fmaf(q[10], si, g * (1.0f - si)); b = fmaf(q[11], si, b * (1.0f - si));
    }
    float tf = q[26];
    if (tf > 0.5f) tf = 1.0f - tf;
    const float dark = 1.0f - 2.0f * tf * 0.3f;
    r *= dark; g *= dark; b *= dark;
    if (q[21] > 0.f) {
        const float ff = __expf(-q[21]);
        r = fmaf(r, ff, q[22] * (1.0f - ff)); g = fmaf(g, ff, q[23] * (1.0f - ff)); b = fmaf(b, ff, q[24] * (1.0f - ff));
    }
    return make_float3(fminf(fmaxf(r, 0.1f), 0.98f), fminf(fmaxf(g, 0.1f), 0.98f), fminf(fmaxf(b, 0.1f), 0.98f));
}

__device__ __forceinline__ float pow5(float x) {  // math.Pow(x, 5): sign-preserving for negative x
    const float x2 = x * x;
    return x2 * x2 * x;
}

// Two uniform-ball points from ONE Philox block (the 16 soft-shadow samples of a (hit, light) pair take
// 8 blocks): sample A from (x,y), sample B from (z,w); u1 = 21 bits, u2 = 21 bits, u3 = 22 bits
// (oracle.cpp Rng::in_unit_sphere_half).
__device__ __forceinline__ void ball_from_bits(uint32_t a, uint32_t b, float& bx, float& by, float& bz) {
    const float u1 = (float)(a >> 11) * (1.0f / 2097152.0f), u2 = (float)(b >> 11) * (1.0f / 2097152.0f);
    const float u3 = (float)(((a & 0x7FFu) << 11) | (b & 0x7FFu)) * (1.0f / 4194304.0f);
    const float z = fmaf(-2.0f, u1, 1.0f);
    const float sxy = sqrt_fast(fmaxf(0.f, fmaf(-z, z, 1.0f)));
    float sn, cs;
    __sincosf(6.2831853071795864769f * u2, &sn, &cs);
    const float rad = ex2_fast(lg2_fast(u3) * (1.0f / 3.0f));  // cbrt; u3 = 0 -> 0
    const float rs = rad * sxy;
    bx = rs * cs;
    by = rs * sn;
    bz = rad * z;
}

}  // namespace gort
