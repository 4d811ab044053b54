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
#include "kernels.h"

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <vector>

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
__device__ __noinline__ uint4 philox_out(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
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
            bool h[2];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const float lox = c ? n1.x : n0.x, hix = c ? n1.y : n0.y, loy = c ? n1.z : n0.z, hiy = c ? n1.w : n0.w;
                const float loz = c ? n2.z : n2.x, hiz = c ? n2.w : n2.y;
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
                h[c] = hit;
            }
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h[0] && h[1]) {
                stack[sp++] = c1;
                node = c0;
            } else if (h[0]) {
                node = c0;
            } else if (h[1]) {
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
                bool keep = true;
                stat_add<STATS>(st, kStatConeTests);
                if (!is_tri) {
                    const float4 s = ldg4(S.spheres + start + i);
                    keep = cone_sphere_candidate(s.x - ox, s.y - oy, s.z - oz, fabsf(s.w), ax, ay, az, tmax);
                    if (dot3(nx, ny, nz, s.x - ox, s.y - oy, s.z - oz) + fabsf(s.w) < thr) keep = false;  // behind the tangent plane
                } else {
                    // the triangle's plane: a cone whose every ray moves away from it (or crosses it below
                    // tMin when the apex lies on it — the hit point's own face) cannot hit the triangle
                    const float4* tp = S.tris + 4 * (size_t)(start + i);
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
                if (keep) {
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
template <bool STATS, bool SMALL, int GEOM>
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
                    const float sinT = sqrt_fast(1.0f - cosT * cosT);
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
                        const float sin2 = eta * eta * (1.0f - cn * cn);
                        if (sin2 > 1.0f) {
                            const float d2 = dot3(ux, uy, uz, rnx, rny, rnz);
                            sx = fmaf(-2.0f * d2, rnx, ux); sy = fmaf(-2.0f * d2, rny, uy); sz = fmaf(-2.0f * d2, rnz, uz);
                        } else {
                            const float k = fmaf(eta, cn, sqrt_fast(1.0f - sin2));
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
                                                    unsigned int* __restrict__ active_count) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (uint32_t)P.n_local_tiles * 32u) return;
    const SceneView& S = P.scene;
    const uint32_t ltile = id >> 5, block = id & 31u;
    // The frame's accumulators are not cleared wholesale (12 MB for 800x600): a kept block clears its own 32 pixels
    // here, a culled block is marked and resolve_kernel writes black for it without reading them.
    P.block_active[id] = 0;
    if (S.n_nodes == 0 || P.max_depth <= 0) return;
    const uint32_t gtile = (uint32_t)P.shard_rank + ltile * (uint32_t)P.shard_count;
    const uint32_t tx = gtile % (uint32_t)P.tiles_x, ty = gtile / (uint32_t)P.tiles_x;
    const uint32_t x0 = tx * kTile + ((block & 3u) << 3), y0 = ty * kTile + ((block >> 2) << 2);
    if (x0 >= (uint32_t)P.width || y0 >= (uint32_t)P.height) return;
    const uint32_t x1 = min(x0 + 8u, (uint32_t)P.width), y1 = min(y0 + 4u, (uint32_t)P.height);
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
    bool active = degenerate, deep = false;
    int stack[64];
    int sp = 0;
    int node = 0;
    while (!deep) {
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

cudaError_t launch_cull(const TraceParams& p, uint32_t* active_list, unsigned int* active_count, cudaStream_t stream) {
    const unsigned int n = (unsigned int)p.n_local_tiles * 32u;
    if (n == 0) return cudaSuccess;
    cull_kernel<<<(n + 127) / 128, 128, 0, stream>>>(p, active_list, active_count);
    return cudaGetLastError();
}


template <bool STATS, bool SMALL, int GEOM>
static cudaError_t launch_trace_variant(const TraceParams& p, int sm_count, cudaStream_t stream) {
    // persistent grid: as many CTAs as fit on the chip at once (occupancy is register-bound)
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_kernel<STATS, SMALL, GEOM>, kWarpsPerCta * 32, 0);
        if (e != cudaSuccess) return e;
        ctas_per_sm = n > 0 ? n : 1;
    }
    if (kDbg && p.debug_times) {  // -DGORT_DEBUG build + GORT_DEBUG_TIMES=1: per-warp timeline of this launch on stderr
        const int nw = sm_count * ctas_per_sm * kWarpsPerCta;
        cudaMemsetAsync(p.debug_times, 0xff, 8, stream);
        cudaMemsetAsync(p.debug_times + 1, 0, (size_t)nw * 128, stream);
        trace_kernel<STATS, SMALL, GEOM><<<sm_count * ctas_per_sm, kWarpsPerCta * 32, 0, stream>>>(p);
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
    trace_kernel<STATS, SMALL, GEOM><<<sm_count * ctas_per_sm, kWarpsPerCta * 32, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_trace(const TraceParams& p, bool stats, int sm_count, cudaStream_t stream) {
    if (p.n_local_tiles == 0) return cudaSuccess;
    // kernel variant: tiny sphere scenes scan the parameter bank; BVH scenes run the kernel specialised for what they hold
    const int geom = p.small_n > 0 ? 0 : ((p.scene.n_spheres > 0 ? 1 : 0) | (p.scene.n_tris > 0 ? 2 : 0));
    if (stats) {
        switch (geom) {
            case 0: return launch_trace_variant<true, true, 1>(p, sm_count, stream);
            case 1: return launch_trace_variant<true, false, 1>(p, sm_count, stream);
            case 2: return launch_trace_variant<true, false, 2>(p, sm_count, stream);
            default: return launch_trace_variant<true, false, 3>(p, sm_count, stream);
        }
    }
    switch (geom) {
        case 0: return launch_trace_variant<false, true, 1>(p, sm_count, stream);
        case 1: return launch_trace_variant<false, false, 1>(p, sm_count, stream);
        case 2: return launch_trace_variant<false, false, 2>(p, sm_count, stream);
        default: return launch_trace_variant<false, false, 3>(p, sm_count, stream);
    }
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
}

cudaError_t launch_resolve(const ResolveParams& p, cudaStream_t stream, int max_blocks) {
    const int n = p.n_local_tiles * kTilePixels;
    if (n == 0) return cudaSuccess;
    if (max_blocks > 0) resolve_kernel<<<std::min(max_blocks, (n + 127) / 128), 128, 0, stream>>>(p);
    else resolve_kernel<<<(n + 255) / 256, 256, 0, stream>>>(p);
    return cudaGetLastError();
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
__global__ void link_signal_kernel(unsigned int* flag) {
    __threadfence_system();
    atomicAdd_system(flag, 1u);
}
__global__ void link_store_kernel(unsigned int* flag, unsigned int value) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
    __threadfence_system();
}
__global__ void link_wait_kernel(const unsigned int* flag, unsigned int target, unsigned int* timed_out) {
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
        if (t - t0 > 20000000000ull) {  // 20 s: a rank died; never hang the GPU — flag it and let the host report
            atomicExch_system(timed_out, 1u);
            break;
        }
    }
    __threadfence_system();
}
cudaError_t launch_link_signal(unsigned int* flag, cudaStream_t stream) {
    link_signal_kernel<<<1, 1, 0, stream>>>(flag);
    return cudaGetLastError();
}
cudaError_t launch_link_store(unsigned int* flag, unsigned int value, cudaStream_t stream) {
    link_store_kernel<<<1, 1, 0, stream>>>(flag, value);
    return cudaGetLastError();
}
cudaError_t launch_link_wait(const unsigned int* flag, unsigned int target, unsigned int* timed_out, cudaStream_t stream) {
    link_wait_kernel<<<1, 1, 0, stream>>>(flag, target, timed_out);
    return cudaGetLastError();
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
