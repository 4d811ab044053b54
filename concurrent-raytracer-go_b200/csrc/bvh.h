// bvh.h — host-side BVH build and the flat structure-of-arrays layout the sm_100a kernels read.
//
// The reference renderer has NO acceleration structure: hitWorld is a linear scan
// (/root/reference internal/renderer/renderer.go:333-346) and internal/optimization's BVH is
// unwired and does not compile.  This BVH is therefore this implementation's own
// design; its only contract is "same closest hit as the linear scan" (tests/test_gpu_bvh.py).
//
// Layout (all arrays 16-byte aligned, read with 128-bit loads):
//   nodes  : 4 x float4 per inner node, breadth-first order (top levels first, so a prefix of the
//            array can be staged in shared memory):
//              n[0] = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//              n[1] = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//              n[2] = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//              n[3] = (child0, child1, 0, 0) as int bits
//            child >= 0 : inner node index.  child < 0 : leaf, v = ~child:
//              bits 0..25 first primitive (in the per-type leaf-ordered array), bits 26..29 count-1,
//              bit 30 type (0 sphere, 1 triangle).  An empty child has an inverted box (never hit).
//   spheres: float4 (cx, cy, cz, r) in leaf order; sphere_meta int2 (material, scan order)
//   tris   : 4 x float4 per triangle in leaf order:
//              (v0.xyz, material bits) (e1.xyz, scan-order bits) (e2.xyz, 0) (unit face normal xyz, 0)
#pragma once
#include <cstdint>
#include <vector>

#include "host_scene.h"

namespace gort {

struct F4 {
    float x, y, z, w;
};
struct I2 {
    int32_t x, y;
};

constexpr int kLeafTypeBit = 30;
constexpr int kLeafCountShift = 26;
constexpr uint32_t kLeafStartMask = (1u << 26) - 1;
constexpr int kMaxLeafPrims = 4;
constexpr int kMaxBvhDepth = 56;  // traversal stack is 64 entries

struct FlatBvh {
    std::vector<F4> nodes;        // 4 per inner node, followed by the quantised copy (2 per inner node, see qorigin)
    std::vector<F4> spheres;      // 1 per sphere
    std::vector<I2> sphere_meta;  // (material, order)
    std::vector<F4> tris;         // 4 per triangle
    // Quantised copy of the nodes for the wavefront pipeline's walk (stream.cu): 32 bytes per node = ONE 256-bit load,
    //   word k (k = 0..2: x, y, z) of child c: lo | hi << 16 on the grid  coordinate = qorigin[k] + q * qcell[k];  word 3: the child link
    // boxes are rounded outward by two cells, so a quantised box always contains the fp32 box (same hits, a few more visits)
    float qorigin[3] = {0, 0, 0}, qcell[3] = {1, 1, 1};
    int32_t n_nodes = 0;
    int32_t max_depth = 0;
    double build_ms = 0;
};

// The device records of one primitive (the layout above), shared by the host builder's leaves and the device builder's input.
void pack_sphere(const HostSphere& s, F4& geom, I2& meta);
void pack_triangle(const HostTriangle& t, F4 out[4]);

// Binned-SAH build (16 bins, all 3 axes), leaves of <= kMaxLeafPrims primitives of one type.
void build_bvh(const HostScene& scene, FlatBvh& out);

}  // namespace gort
