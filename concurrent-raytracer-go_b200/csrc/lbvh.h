// lbvh.h — device-side BVH build for large scenes (lbvh.cu): launch interface for capi.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace gort {

struct LbvhIn {
    const float4* spheres;      // (cx, cy, cz, r) in scene order
    const int2* sphere_meta;    // (material, scan order)
    const float4* tris;         // 4 float4 per triangle, packed as bvh.h describes, in scene order
    uint32_t n_spheres, n_tris;
    float world_lo[3], world_hi[3];  // bounds of all primitives
    float pad;                       // conservative padding of every node box (fp32 slab arithmetic)
    float qorigin[3], qcell[3];      // grid of the quantised node copy (bvh.h)
};

struct LbvhOut {
    float4* nodes;        // 6 float4 per inner node: (n - 1) x 4 fp32 nodes, then (n - 1) x 2 quantised ones
    float4* spheres;      // leaf order
    int2* sphere_meta;
    float4* tris;
};

size_t lbvh_scratch_bytes(uint32_t n_prims);
// Enqueues the build on `st`; *max_depth_out (host, pinned or pageable) is valid after the stream has been synchronised.
cudaError_t lbvh_build(const LbvhIn& in, const LbvhOut& out, void* scratch, size_t scratch_bytes, int* max_depth_out, cudaStream_t st);

// 4-wide collapse of a flat BVH2 (bvh.h layout, either builder) for the wavefront pipeline's walk: every inner node at even depth
// becomes a 64-byte node holding its (up to four) grandchildren: word 3k + a = child k's box on axis a as lo | hi << 16 on the
// quantisation grid (rounded outward by two cells), word 12 + k = child k's link (>= 0 wide node, < 0 leaf as in bvh.h,
// kWideNoChild = empty slot).  One visit = two 256-bit loads and four slab tests: half the dependent node fetches of the binary walk.
constexpr int kWideNoChild = 0x7FFFFFFF;
// Top of the wide tree as one contiguous block (wide_top_block): the nodes of the first kTopLevels wide levels (= eight binary
// levels) in breadth-first order, kTopNodes x 64 bytes, links between them rewritten to kTopBase + position in the block; links
// that leave the block are the global ones.  A traversal kernel stages the block in shared memory with one bulk copy.
constexpr int kTopLevels = 4;
constexpr int kTopNodes = 1 + 4 + 16 + 64;
constexpr int kTopBase = 0x40000000;
cudaError_t wide_top_block(const float4* wide, float4* top_out, cudaStream_t st);
size_t bvh_collapse_scratch_bytes(int n_nodes2);
cudaError_t bvh_collapse_wide(const float4* nodes2, int n_nodes2, float4* wide_out, const float* qorigin3, const float* qcell3, void* scratch,
                              size_t scratch_bytes, cudaStream_t st);

}  // namespace gort
