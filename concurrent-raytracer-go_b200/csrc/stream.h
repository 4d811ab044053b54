// stream.h — the global-queue wavefront pipeline for large scenes (stream.cu), launch interface for capi.cu.
//
// The persistent per-warp-queue kernel (kernels.cu) keeps every ray of a warp in shared memory: right for scenes
// whose geometry fits the parameter bank / L1 (the reference's own scenes), wrong for 10^5..10^6 primitives — there
// the BVH walk is bound by node-fetch latency, and 190 KB of per-warp queues per SM leave ~60 KB of L1 for the nodes
// while rays of very different walk length share a warp (8 of 32 lanes active, profiles/r1_ncu_trace_c4small_v6.txt).
// On a B200 the queues cost nothing in HBM (180 GB, ~7 TB/s; a path is 56 bytes), so for large scenes the recursion of
// traceRay (/root/reference internal/renderer/renderer.go:165-227) runs as a chain of small kernels over queues in
// global memory, one bounce at a time:
//
//   per iteration (the path queue holds paths of every depth; it is topped up with new primary rays — path
//   regeneration — so that every launch works on a full queue until the frame's samples run out):
//     scatter             hit record + Material.Scatter: one shade record per hit, survivors -> next queue (:175-189)
//     plan                how many primary rays fit behind the survivors
//     pool_trace<EXT>     hitWorld of the scattered rays, in place in the next queue
//     pool_trace<PRIMARY> tracePixel/getRay + hitWorld of the new primary rays, behind them        (:150-173,377-390)
//     per chunk of <= 4 lights:
//       pair_setup        light direction, back-facing cull, hard-shadow ray list             (:248-257,303-309)
//       pool_trace<HARD>  calculateSmartShadow's hard ray (boolean hitWorld)
//       pool_cone         shadow-cone walk of every lit pair: its candidate primitives (device.cuh), lanes refilled
//       soft_cand         16 jittered rays against <= 6 candidates, a quarter warp per pair             (:311-328)
//       pool_trace<SOFT>  the 16 rays of the pairs whose cone holds more than that, through the BVH
//       shade_accum       calculateDirectLighting's arithmetic + traceRay's weighting, one fixed-point add per hit (:258-297,191-226)
//
// pool_trace is ONE traversal kernel (four ray sources): persistent warps fetch rays from the stage's global pool
// and REFILL finished lanes while the pool lasts, so a warp's 32 lanes stay busy whatever the spread of walk
// lengths; it has no shared memory, so the SM's whole 256 KB is L1 for the BVH.  Same arithmetic, same Philox
// counters and the same commutative fixed-point accumulators as the per-warp-queue kernel: both paths produce the
// same image (tests/test_gpu_stream.py).
#pragma once
#include "kernels.h"

namespace gort {

constexpr int kStreamLightChunk = 4;  // lights per shading pass
constexpr int kStreamMaxChunks = 7;   // scenes with more than 28 lights use the per-warp-queue kernel
constexpr int kCtlWords = 64;         // uint32 counters per iteration
constexpr bool kStreamTopDefault = true;    // shared-memory staging of the top of the wide tree unless GORT_TOP=0
constexpr int kStreamSortDefault = 0;     // sorted mode (stream_launch_sort) unless GORT_SORT says otherwise
// counter block of one iteration (a ring of two; plus the frame's primary cursor behind the ring)
enum StreamCtl {
    kCtlNext = 0,       // paths appended to the next queue by scatter == rays of pool_trace<EXT>
    kCtlRec = 1,        // shade records of this iteration
    kCtlFetchExt = 2,   // fetch cursor of pool_trace<EXT>
    kCtlFetchPrimary = 3,
    kCtlNew = 4,        // primary rays generated this iteration (path regeneration), written by the plan kernel
    kCtlNextTotal = 5,  // kCtlNext + kCtlNew: entries of the next queue
    kCtlPrimStart = 6,  // (two words, low first) index of the first new primary ray in the frame's primary sequence
    kCtlLive = 15,      // sorted mode: live entries of the current queue (the spare eighth word of light chunk 0's block)
    kCtlChunk0 = 8,     // per light chunk c at kCtlChunk0 + 8 c:
    kCtlHard = 0, kCtlFetchHard = 1, kCtlWalk = 2, kCtlFetchWalk = 3,
    kCtlLit = 4,        // lit pairs (hard shadow ray unoccluded) == cone walks of pool_cone
    kCtlFetchCone = 5,
    kCtlCand = 6        // pairs with 1..6 shadow-cone candidates (records of soft_cand)
};

struct StreamView {
    // path queues, two buffers (current iteration / next iteration), structure of arrays:
    //   qa (hit point xyz, primitive | dead marker)   qb (incoming direction xyz, fog factor; primary entries: t)
    //   qc (throughput rgb, sample | depth << 16)     qd (global pixel, local accumulator index)
    float4* qa[2];
    float4* qb[2];
    float4* qc[2];
    uint2* qd[2];
    // shade records of the current iteration: ra (point, material) rb (normal, sample | depth << 16) rc (throughput, fog) rd (pixels)
    float4* ra;
    float4* rb;
    float4* rc;
    uint2* rd;
    float4* racc;            // running direct-light sum of a record across light chunks (scenes with > 4 lights)
    uint8_t* lit;            // [cap * 4] hard shadow ray of (record, light of the chunk) unoccluded
    unsigned int* cnt;       // [cap * 4] unoccluded soft shadow rays (of 16)
    uint32_t* hard_list;     // [cap * 4] record | light-in-chunk << 30
    uint32_t* walk_list;     // [cap * 4] lit pairs whose 16 rays walk the BVH themselves
    uint32_t* lit_list;      // [cap * 4] lit pairs, in the order their hard shadow rays finished
    uint4* cand_recs;        // [cap * 4] x 2: (pair, candidate count, candidates 0-1) (candidates 2-5)
    // sorted mode (stream_launch_sort): the live entries of the current queue in Morton order of their hit points;
    // nullptr = scatter reads the queue in the order it was written
    const uint32_t* order;
    // top of the wide tree as one block (lbvh.h: wide_top_block); nullptr = every node is read from global memory
    const float4* top;
    int top_plain;           // A/B switch: stage it with plain loads instead of the bulk copy
    unsigned int* ctl;       // this iteration's counter block
    const unsigned int* ctl_prev;  // previous iteration's block (kCtlNextTotal = entries of the current queue)
    unsigned long long* prim_cursor;  // primary rays of the frame generated so far
    unsigned long long prim_total;    // active blocks x 32 x samples
    uint32_t cap;            // path slots
    int cur;                 // buffer of the current queue
    int l0, lc, chunk, last_chunk;
    uint32_t n_active, n_deep;  // pixel blocks kept by the cull pass (read back by the host)
};

size_t stream_bytes_per_slot();
bool stream_wants_wide_nodes();  // the traversal kernel of this build walks the 4-wide collapse (lbvh.h)
// geom: 1 spheres only, 2 triangles only, 3 both (template specialisation, like the per-warp-queue kernel)
// plan: how many primary rays join the next queue this iteration (fills it up to `cap`), 1 thread
cudaError_t stream_launch_plan(const StreamView& v, cudaStream_t st);
cudaError_t stream_launch_primary(const TraceParams& p, const StreamView& v, int geom, bool stats, int sm_count, cudaStream_t st);
// Sorted mode: order the current queue's live entries by the Morton code of their hit points before scatter reads them, so
// that the shade records, the shadow-ray lists and the scattered rays that derive from 32 consecutive entries start close
// together (coherent walks).  n_cur = entries of the current queue (the host has read it back); sets v.order.
size_t stream_sort_bytes(uint32_t cap);
// begin_bit: 0 = the full 23-bit code; 8 = 32x coarser cells (one radix pass fewer); 23 = live entries first, order kept
cudaError_t stream_launch_sort(const TraceParams& p, StreamView& v, uint32_t n_cur, int begin_bit, void* scratch, size_t scratch_bytes, cudaStream_t st);
cudaError_t stream_launch_scatter(const TraceParams& p, const StreamView& v, int geom, bool stats, int sm_count, cudaStream_t st);
cudaError_t stream_launch_trace_ext(const TraceParams& p, const StreamView& v, int geom, bool stats, int sm_count, cudaStream_t st);
cudaError_t stream_launch_shade_chunk(const TraceParams& p, const StreamView& v, int geom, bool stats, int sm_count, cudaStream_t st);

}  // namespace gort
