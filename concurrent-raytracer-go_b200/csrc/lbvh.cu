// lbvh.cu — BVH build on the device for large scenes (LBVH: Morton order + Karras' parallel topology).
//
// The reference renderer has no acceleration structure at all (hitWorld is a linear scan, /root/reference
// internal/renderer/renderer.go:333-346), so the BVH is this implementation's own and its only contract is "same closest
// hit as the linear scan".  The host builder (bvh.cpp, binned SAH) needs ~0.45 s for a million primitives on the 16 host
// threads — a third of an end-to-end 4K frame on one GPU, and it does not shrink when the frame is sharded over 8 GPUs
// (every rank process builds the same tree on the same cores).  On the device the same million primitives take a few
// milliseconds: 63-bit Morton codes of the primitive centroids, one radix sort (CUB), the internal nodes of the binary radix
// tree in parallel (T. Karras, "Maximizing parallelism in the construction of BVHs, octrees and k-d trees", HPG 2012), boxes
// bottom-up with one atomic counter per node, then the same flat layout bvh.h describes (two child boxes per 64-byte node,
// one primitive per leaf, per-type primitive arrays in leaf order, the quantised 32-byte copy).
#include <cub/cub.cuh>

#include "lbvh.h"

#include <algorithm>

namespace gort {

namespace {

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {  // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

struct PrimBox {
    float lox, loy, loz, hix, hiy, hiz;
};

// primitive i of the build: spheres first (scene order), then triangles
__device__ __forceinline__ PrimBox prim_box(const LbvhIn& in, uint32_t i) {
    PrimBox b;
    if (i < in.n_spheres) {
        const float4 s = in.spheres[i];
        const float r = fabsf(s.w);
        b.lox = s.x - r; b.loy = s.y - r; b.loz = s.z - r;
        b.hix = s.x + r; b.hiy = s.y + r; b.hiz = s.z + r;
    } else {
        const float4* t = in.tris + 4 * (size_t)(i - in.n_spheres);
        const float4 v0 = t[0], e1 = t[1], e2 = t[2];
        b.lox = fminf(v0.x, fminf(v0.x + e1.x, v0.x + e2.x)); b.hix = fmaxf(v0.x, fmaxf(v0.x + e1.x, v0.x + e2.x));
        b.loy = fminf(v0.y, fminf(v0.y + e1.y, v0.y + e2.y)); b.hiy = fmaxf(v0.y, fmaxf(v0.y + e1.y, v0.y + e2.y));
        b.loz = fminf(v0.z, fminf(v0.z + e1.z, v0.z + e2.z)); b.hiz = fmaxf(v0.z, fmaxf(v0.z + e1.z, v0.z + e2.z));
    }
    return b;
}

__global__ void lbvh_codes_kernel(const LbvhIn in, unsigned long long* __restrict__ keys, uint32_t* __restrict__ ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = in.n_spheres + in.n_tris;
    if (i >= n) return;
    const PrimBox b = prim_box(in, i);
    const float sx = 2097151.0f / fmaxf(in.world_hi[0] - in.world_lo[0], 1e-30f), sy = 2097151.0f / fmaxf(in.world_hi[1] - in.world_lo[1], 1e-30f),
                sz = 2097151.0f / fmaxf(in.world_hi[2] - in.world_lo[2], 1e-30f);
    const float cx = (0.5f * (b.lox + b.hix) - in.world_lo[0]) * sx, cy = (0.5f * (b.loy + b.hiy) - in.world_lo[1]) * sy,
                cz = (0.5f * (b.loz + b.hiz) - in.world_lo[2]) * sz;
    const unsigned long long qx = (unsigned long long)fminf(fmaxf(cx, 0.f), 2097151.0f), qy = (unsigned long long)fminf(fmaxf(cy, 0.f), 2097151.0f),
                             qz = (unsigned long long)fminf(fmaxf(cz, 0.f), 2097151.0f);
    keys[i] = (expand21(qx) << 2) | (expand21(qy) << 1) | expand21(qz);
    ids[i] = i;
}

// common-prefix length of sorted keys i and j (ties broken by the position, so every key is unique); -1 outside [0, n)
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ k, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = k[i], b = k[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// Karras 2012, section 3: internal node i of the binary radix tree over the sorted keys.  child code = index | leaf << 31
__global__ void lbvh_topology_kernel(const unsigned long long* __restrict__ keys, int n, int2* __restrict__ children, int* __restrict__ parent_internal,
                                     int* __restrict__ parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    int2 c;
    if (lo == gamma) { c.x = gamma | (int)0x80000000; parent_leaf[gamma] = i; }
    else { c.x = gamma; parent_internal[gamma] = i; }
    if (hi == gamma + 1) { c.y = (gamma + 1) | (int)0x80000000; parent_leaf[gamma + 1] = i; }
    else { c.y = gamma + 1; parent_internal[gamma + 1] = i; }
    children[i] = c;
    if (i == 0) parent_internal[0] = -1;
}

// boxes bottom-up: the second thread to arrive at a node owns it (its sibling subtree is complete)
__global__ void lbvh_refit_kernel(const LbvhIn in, const uint32_t* __restrict__ ids, int n, const int2* __restrict__ children,
                                  const int* __restrict__ parent_internal, const int* __restrict__ parent_leaf, unsigned int* __restrict__ arrived,
                                  float* __restrict__ boxes /* 6 per internal node */) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int node = parent_leaf[leaf];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(arrived + node, 1u) == 0u) return;  // first arrival: the sibling is not done yet
        const int2 c = children[node];
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int cc = k ? c.y : c.x;
            if (cc < 0) {
                const PrimBox b = prim_box(in, ids[cc & 0x7fffffff]);
                lo[0] = fminf(lo[0], b.lox); lo[1] = fminf(lo[1], b.loy); lo[2] = fminf(lo[2], b.loz);
                hi[0] = fmaxf(hi[0], b.hix); hi[1] = fmaxf(hi[1], b.hiy); hi[2] = fmaxf(hi[2], b.hiz);
            } else {
                const volatile float* b = boxes + 6 * (size_t)cc;
                lo[0] = fminf(lo[0], b[0]); lo[1] = fminf(lo[1], b[1]); lo[2] = fminf(lo[2], b[2]);
                hi[0] = fmaxf(hi[0], b[3]); hi[1] = fmaxf(hi[1], b[4]); hi[2] = fmaxf(hi[2], b[5]);
            }
        }
        float* o = boxes + 6 * (size_t)node;
        o[0] = lo[0]; o[1] = lo[1]; o[2] = lo[2]; o[3] = hi[0]; o[4] = hi[1]; o[5] = hi[2];
        node = parent_internal[node];
    }
}

// depth of the tree: every leaf counts its way up
__global__ void lbvh_depth_kernel(int n, const int* __restrict__ parent_internal, const int* __restrict__ parent_leaf, int* __restrict__ max_depth) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    int d = 0;
    if (leaf < n) {
        d = 1;
        for (int node = parent_leaf[leaf]; node >= 0; node = parent_internal[node]) d++;
    }
    // one atomic per warp
    for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_depth, d);
}

__global__ void lbvh_flags_kernel(const uint32_t* __restrict__ ids, int n, uint32_t n_spheres, uint32_t* __restrict__ is_sphere) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_sphere[i] = ids[i] < n_spheres ? 1u : 0u;
}

// per-type primitive arrays in leaf (= Morton) order
__global__ void lbvh_gather_kernel(const LbvhIn in, const uint32_t* __restrict__ ids, int n, const uint32_t* __restrict__ sphere_rank, LbvhOut out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = ids[i], sr = sphere_rank[i];
    if (id < in.n_spheres) {
        out.spheres[sr] = in.spheres[id];
        out.sphere_meta[sr] = in.sphere_meta[id];
    } else {
        const uint32_t tr = (uint32_t)i - sr;
        const float4* s = in.tris + 4 * (size_t)(id - in.n_spheres);
        float4* d = out.tris + 4 * (size_t)tr;
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
    }
}

__device__ __forceinline__ float pad_down(float v, float pad) { return __fadd_rd(v, -pad); }
__device__ __forceinline__ float pad_up(float v, float pad) { return __fadd_ru(v, pad); }

// the flat nodes (bvh.h): both children's boxes in the parent, links, and the quantised copy behind them
__global__ void lbvh_emit_kernel(const LbvhIn in, const uint32_t* __restrict__ ids, int n, const int2* __restrict__ children, const float* __restrict__ boxes,
                                 const uint32_t* __restrict__ sphere_rank, LbvhOut out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 c = children[i];
    float lo[2][3], hi[2][3];
    int link[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int cc = k ? c.y : c.x;
        if (cc < 0) {
            const int leaf = cc & 0x7fffffff;
            const uint32_t id = ids[leaf];
            const PrimBox b = prim_box(in, id);
            lo[k][0] = b.lox; lo[k][1] = b.loy; lo[k][2] = b.loz; hi[k][0] = b.hix; hi[k][1] = b.hiy; hi[k][2] = b.hiz;
            const bool sph = id < in.n_spheres;
            const uint32_t start = sph ? sphere_rank[leaf] : (uint32_t)leaf - sphere_rank[leaf];
            link[k] = (int)~(start | (sph ? 0u : (1u << 30)));  // one primitive per leaf: count - 1 = 0
        } else {
            const float* b = boxes + 6 * (size_t)cc;
            lo[k][0] = b[0]; lo[k][1] = b[1]; lo[k][2] = b[2]; hi[k][0] = b[3]; hi[k][1] = b[4]; hi[k][2] = b[5];
            link[k] = cc;
        }
#pragma unroll
        for (int a = 0; a < 3; a++) {  // conservative padding for fp32 slab arithmetic, as the host builder's
            lo[k][a] = pad_down(lo[k][a], in.pad);
            hi[k][a] = pad_up(hi[k][a], in.pad);
        }
    }
    float4* o = out.nodes + 4 * (size_t)i;
    o[0] = make_float4(lo[0][0], hi[0][0], lo[0][1], hi[0][1]);
    o[1] = make_float4(lo[1][0], hi[1][0], lo[1][1], hi[1][1]);
    o[2] = make_float4(lo[0][2], hi[0][2], lo[1][2], hi[1][2]);
    o[3] = make_float4(__int_as_float(link[0]), __int_as_float(link[1]), 0.f, 0.f);
    uint32_t qw[2][4];
#pragma unroll
    for (int k = 0; k < 2; k++) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const double ql = floor(((double)lo[k][a] - (double)in.qorigin[a]) / (double)in.qcell[a]) - 2.0;
            const double qh = ceil(((double)hi[k][a] - (double)in.qorigin[a]) / (double)in.qcell[a]) + 2.0;
            qw[k][a] = (uint32_t)fmin(65535.0, fmax(0.0, ql)) | ((uint32_t)fmin(65535.0, fmax(0.0, qh)) << 16);
        }
        qw[k][3] = (uint32_t)link[k];
    }
    float4* q = out.nodes + 4 * (size_t)(n - 1) + 2 * (size_t)i;
    q[0] = make_float4(__uint_as_float(qw[0][0]), __uint_as_float(qw[0][1]), __uint_as_float(qw[0][2]), __uint_as_float(qw[0][3]));
    q[1] = make_float4(__uint_as_float(qw[1][0]), __uint_as_float(qw[1][1]), __uint_as_float(qw[1][2]), __uint_as_float(qw[1][3]));
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

// ---------------------------------------------------------------------------------------------
// 4-wide collapse (lbvh.h)
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void wide_parent_kernel(const float4* __restrict__ nodes, int n2, int* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const float4 l = nodes[4 * (size_t)i + 3];
    const int c0 = __float_as_int(l.x), c1 = __float_as_int(l.y);
    if (c0 >= 0) parent[c0] = i;
    if (c1 >= 0 && c1 != c0) parent[c1] = i;
    if (i == 0) parent[0] = -1;
}

__global__ void wide_mark_kernel(const int* __restrict__ parent, int n2, uint32_t* __restrict__ mark) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    int d = 0;
    for (int p = parent[i]; p >= 0; p = parent[p]) d++;
    mark[i] = (d & 1) ? 0u : 1u;
}

struct QGrid {
    float o[3], c[3];
};

__device__ __forceinline__ uint32_t quantise(float lo, float hi, float o, float c) {
    const double ql = floor(((double)lo - (double)o) / (double)c) - 2.0, qh = ceil(((double)hi - (double)o) / (double)c) + 2.0;
    return (uint32_t)fmin(65535.0, fmax(0.0, ql)) | ((uint32_t)fmin(65535.0, fmax(0.0, qh)) << 16);
}

__global__ void wide_emit_kernel(const float4* __restrict__ nodes, int n2, const uint32_t* __restrict__ mark, const uint32_t* __restrict__ widx,
                                 const QGrid g, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2 || !mark[i]) return;
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 12; k++) w[k] = 0u;
#pragma unroll
    for (int k = 0; k < 4; k++) w[12 + k] = (uint32_t)kWideNoChild;
    int n = 0;
    auto add = [&](float lox, float hix, float loy, float hiy, float loz, float hiz, int link) {
        w[3 * n + 0] = quantise(lox, hix, g.o[0], g.c[0]);
        w[3 * n + 1] = quantise(loy, hiy, g.o[1], g.c[1]);
        w[3 * n + 2] = quantise(loz, hiz, g.o[2], g.c[2]);
        w[12 + n] = link >= 0 ? widx[link] : (uint32_t)link;  // an inner grandchild sits at even depth: it is a wide node itself
        n++;
    };
    const float4* np = nodes + 4 * (size_t)i;
    const float4 n0 = np[0], n1 = np[1], n2v = np[2], n3 = np[3];
    const int c[2] = {__float_as_int(n3.x), __float_as_int(n3.y)};
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (k == 1 && c[1] == c[0]) break;  // (the single-primitive scene references its leaf from both slots)
        if (c[k] >= 0) {
            const float4* cp = nodes + 4 * (size_t)c[k];
            const float4 m0 = cp[0], m1 = cp[1], m2 = cp[2], m3 = cp[3];
            const int g0 = __float_as_int(m3.x), g1 = __float_as_int(m3.y);
            add(m0.x, m0.y, m0.z, m0.w, m2.x, m2.y, g0);
            if (g1 != g0) add(m1.x, m1.y, m1.z, m1.w, m2.z, m2.w, g1);
        } else if (k == 0) {
            add(n0.x, n0.y, n0.z, n0.w, n2v.x, n2v.y, c[0]);
        } else {
            add(n1.x, n1.y, n1.z, n1.w, n2v.z, n2v.w, c[1]);
        }
    }
    float4* o = out + 4 * (size_t)widx[i];
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = make_float4(__uint_as_float(w[4 * k]), __uint_as_float(w[4 * k + 1]), __uint_as_float(w[4 * k + 2]), __uint_as_float(w[4 * k + 3]));
}

}  // namespace

namespace {

// one thread: 85 nodes, once per uploaded scene
__global__ void wide_top_kernel(const float4* __restrict__ wide, float4* __restrict__ top) {
    int ids[kTopNodes];
    unsigned char depth[kTopNodes];
    int cnt = 1;
    ids[0] = 0;  // the root is wide node 0
    depth[0] = 0;
    for (int h = 0; h < cnt; h++) {
        const float4* np = wide + 4 * (size_t)ids[h];
        const float4 a = np[0], b = np[1], c = np[2];
        float4 d = np[3];
        if (depth[h] < kTopLevels - 1) {
            int l[4] = {__float_as_int(d.x), __float_as_int(d.y), __float_as_int(d.z), __float_as_int(d.w)};
            for (int k = 0; k < 4; k++) {
                if (l[k] >= 0 && l[k] != kWideNoChild && cnt < kTopNodes) {
                    ids[cnt] = l[k];
                    depth[cnt] = depth[h] + 1;
                    l[k] = kTopBase + cnt;
                    cnt++;
                }
            }
            d = make_float4(__int_as_float(l[0]), __int_as_float(l[1]), __int_as_float(l[2]), __int_as_float(l[3]));
        }
        top[4 * h + 0] = a; top[4 * h + 1] = b; top[4 * h + 2] = c; top[4 * h + 3] = d;
    }
    const float4 none = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 links = make_float4(__int_as_float(kWideNoChild), __int_as_float(kWideNoChild), __int_as_float(kWideNoChild), __int_as_float(kWideNoChild));
    for (int h = cnt; h < kTopNodes; h++) { top[4 * h + 0] = none; top[4 * h + 1] = none; top[4 * h + 2] = none; top[4 * h + 3] = links; }
}

}  // namespace

cudaError_t wide_top_block(const float4* wide, float4* top_out, cudaStream_t st) {
    wide_top_kernel<<<1, 1, 0, st>>>(wide, top_out);
    return cudaGetLastError();
}

size_t bvh_collapse_scratch_bytes(int n2) {
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, n2);
    return align256(scan_bytes) + 3 * align256((size_t)n2 * 4) + 1024;
}

cudaError_t bvh_collapse_wide(const float4* nodes2, int n2, float4* wide_out, const float* qorigin3, const float* qcell3, void* scratch, size_t scratch_bytes,
                              cudaStream_t st) {
    if (n2 <= 0) return cudaSuccess;
    if (scratch_bytes < bvh_collapse_scratch_bytes(n2)) return cudaErrorInvalidValue;
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, n2);
    uint8_t* p = (uint8_t*)scratch;
    auto take = [&](size_t bytes) { uint8_t* r = p; p += align256(bytes); return r; };
    void* cub_tmp = take(scan_bytes);
    int* parent = (int*)take((size_t)n2 * 4);
    uint32_t* mark = (uint32_t*)take((size_t)n2 * 4);
    uint32_t* widx = (uint32_t*)take((size_t)n2 * 4);
    const int T = 256, G = (n2 + T - 1) / T;
    wide_parent_kernel<<<G, T, 0, st>>>(nodes2, n2, parent);
    wide_mark_kernel<<<G, T, 0, st>>>(parent, n2, mark);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(cub_tmp, scan_bytes, mark, widx, n2, st);
    if (e != cudaSuccess) return e;
    QGrid g;
    for (int a = 0; a < 3; a++) { g.o[a] = qorigin3[a]; g.c[a] = qcell3[a]; }
    wide_emit_kernel<<<G, T, 0, st>>>(nodes2, n2, mark, widx, g, wide_out);
    return cudaGetLastError();
}

size_t lbvh_scratch_bytes(uint32_t n) {
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return align256(std::max(sort_bytes, scan_bytes)) + 2 * align256((size_t)n * 8) + 4 * align256((size_t)n * 4) + align256((size_t)n * 8) +
           2 * align256((size_t)n * 4) + align256((size_t)n * 4) + align256((size_t)n * 24) + 1024;
}

cudaError_t lbvh_build(const LbvhIn& in, const LbvhOut& out, void* scratch, size_t scratch_bytes, int* max_depth_out, cudaStream_t st) {
    const uint32_t n = in.n_spheres + in.n_tris;
    if (n < 2) return cudaErrorInvalidValue;
    if (scratch_bytes < lbvh_scratch_bytes(n)) return cudaErrorInvalidValue;
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    uint8_t* p = (uint8_t*)scratch;
    auto take = [&](size_t bytes) { uint8_t* r = p; p += align256(bytes); return r; };
    void* cub_tmp = take(std::max(sort_bytes, scan_bytes));
    size_t cub_bytes = std::max(sort_bytes, scan_bytes);
    unsigned long long* keys_in = (unsigned long long*)take((size_t)n * 8);
    unsigned long long* keys = (unsigned long long*)take((size_t)n * 8);
    uint32_t* ids_in = (uint32_t*)take((size_t)n * 4);
    uint32_t* ids = (uint32_t*)take((size_t)n * 4);
    uint32_t* is_sphere = (uint32_t*)take((size_t)n * 4);
    uint32_t* sphere_rank = (uint32_t*)take((size_t)n * 4);
    int2* children = (int2*)take((size_t)n * 8);
    int* parent_internal = (int*)take((size_t)n * 4);
    int* parent_leaf = (int*)take((size_t)n * 4);
    unsigned int* arrived = (unsigned int*)take((size_t)n * 4);
    float* boxes = (float*)take((size_t)n * 24);
    int* d_depth = (int*)take(256);

    const int T = 256, G = (int)((n + T - 1) / T);
    lbvh_codes_kernel<<<G, T, 0, st>>>(in, keys_in, ids_in);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys, ids_in, ids, (int)n, 0, 63, st);
    if (e != cudaSuccess) return e;
    lbvh_topology_kernel<<<G, T, 0, st>>>(keys, (int)n, children, parent_internal, parent_leaf);
    e = cudaMemsetAsync(arrived, 0, (size_t)n * 4, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_depth, 0, 4, st);
    if (e != cudaSuccess) return e;
    lbvh_refit_kernel<<<G, T, 0, st>>>(in, ids, (int)n, children, parent_internal, parent_leaf, arrived, boxes);
    lbvh_depth_kernel<<<G, T, 0, st>>>((int)n, parent_internal, parent_leaf, d_depth);
    lbvh_flags_kernel<<<G, T, 0, st>>>(ids, (int)n, in.n_spheres, is_sphere);
    e = cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, is_sphere, sphere_rank, (int)n, st);
    if (e != cudaSuccess) return e;
    lbvh_gather_kernel<<<G, T, 0, st>>>(in, ids, (int)n, sphere_rank, out);
    lbvh_emit_kernel<<<G, T, 0, st>>>(in, ids, (int)n, children, boxes, sphere_rank, out);
    e = cudaMemcpyAsync(max_depth_out, d_depth, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace gort
