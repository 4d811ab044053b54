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

}  // namespace gort
