// bvh.cpp — binned-SAH BVH2 build on the host and breadth-first flattening (see bvh.h).
#include "bvh.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <limits>
#include <atomic>
#include <queue>
#include <thread>

namespace gort {
namespace {

struct BPrim {
    double lo[3], hi[3], c[3];
    int32_t type;  // 0 sphere, 1 triangle
    int32_t idx;   // index into HostScene::spheres / ::tris
};

struct BNode {
    double lo[3], hi[3];
    int32_t left = -1, right = -1;  // tree-node indices; -1 => leaf
    int32_t first = 0, count = 0;   // leaf: range in the prim permutation
    int32_t type = 0;               // leaf: primitive type
    int32_t depth = 0;
};

struct Box {
    double lo[3], hi[3];
    void reset() {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::numeric_limits<double>::infinity();
            hi[a] = -std::numeric_limits<double>::infinity();
        }
    }
    void grow(const double* l, const double* h) {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], l[a]);
            hi[a] = std::max(hi[a], h[a]);
        }
    }
    void grow(const Box& b) { grow(b.lo, b.hi); }
    double area() const {
        double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

constexpr int kBins = 16;
static double kNodeCost = 1.0;  // GORT_BVH_NODE_COST (in primitive tests)

// A subtree whose construction was deferred to a worker thread: the slot of the parent that will point to it
struct Deferred {
    int parent, which;  // nodes[parent].left (0) / .right (1) receives the subtree's root
    int first, count, depth;
};

struct Bins3 {
    Box box[3][kBins];
    int cnt[3][kBins];
    void reset() {
        for (int a = 0; a < 3; a++)
            for (int k = 0; k < kBins; k++) {
                box[a][k].reset();
                cnt[a][k] = 0;
            }
    }
};

// threads the top of the tree may use for one node's passes (set by build_bvh; worker builders stay sequential)
static unsigned g_node_threads = 1;
constexpr int kParallelNode = 1 << 16;  // ranges of at least this many primitives are scanned in parallel chunks

template <class F>
static void for_chunks(int first, int count, unsigned threads, F&& f) {
    const unsigned nt = std::max(1u, std::min<unsigned>(threads, (unsigned)(count / (kParallelNode / 4)) + 1u));
    if (nt <= 1) {
        f(0u, first, first + count);
        return;
    }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; t++) {
        const int lo = first + (int)((long long)count * t / nt), hi = first + (int)((long long)count * (t + 1) / nt);
        if (t + 1 < nt) pool.emplace_back([&f, t, lo, hi]() { f(t, lo, hi); });
        else f(t, lo, hi);
    }
    for (auto& th : pool) th.join();
}

static void bin_range(const std::vector<BPrim>& prims, int lo, int hi, const Box& cb, const double scale3[3], const bool axis_ok[3], Bins3& b) {
    b.reset();
    for (int i = lo; i < hi; i++) {
        const BPrim& p = prims[i];
        for (int axis = 0; axis < 3; axis++) {
            if (!axis_ok[axis]) continue;
            const int k = std::min(kBins - 1, std::max(0, (int)((p.c[axis] - cb.lo[axis]) * scale3[axis])));
            b.cnt[axis][k]++;
            b.box[axis][k].grow(p.lo, p.hi);
        }
    }
}

static void fill_bins(const std::vector<BPrim>& prims, int first, int count, const Box& cb, const double scale3[3], const bool axis_ok[3],
                      Bins3& out) {
    const unsigned threads = count >= kParallelNode ? g_node_threads : 1u;
    if (threads <= 1) {
        bin_range(prims, first, first + count, cb, scale3, axis_ok, out);
        return;
    }
    std::vector<Bins3> part(threads);
    std::vector<char> used(threads, 0);
    for_chunks(first, count, threads, [&](unsigned t, int lo, int hi) {
        bin_range(prims, lo, hi, cb, scale3, axis_ok, part[t]);
        used[t] = 1;
    });
    out.reset();
    for (unsigned t = 0; t < threads; t++) {
        if (!used[t]) continue;
        for (int a = 0; a < 3; a++)
            for (int k = 0; k < kBins; k++) {
                out.cnt[a][k] += part[t].cnt[a][k];
                out.box[a][k].grow(part[t].box[a][k]);
            }
    }
}

// bounds of the primitives and of their centroids over [first, first + count), and whether both types occur
static void range_bounds(const std::vector<BPrim>& prims, int first, int count, Box& b, Box& cb, bool& mixed) {
    const unsigned threads = count >= kParallelNode ? g_node_threads : 1u;
    const int32_t type0 = prims[first].type;
    auto scan = [&prims, type0](int lo, int hi, Box& bb_out, Box& cc_out, bool& mx_out) {
        // accumulated in locals and stored once: the threads' result slots share cache lines
        Box bb, cc;
        bb.reset();
        cc.reset();
        bool mx = false;
        for (int i = lo; i < hi; i++) {
            bb.grow(prims[i].lo, prims[i].hi);
            cc.grow(prims[i].c, prims[i].c);
            if (prims[i].type != type0) mx = true;
        }
        bb_out = bb;
        cc_out = cc;
        mx_out = mx;
    };
    if (threads <= 1) {
        scan(first, first + count, b, cb, mixed);
        return;
    }
    std::vector<Box> pb(threads), pc(threads);
    std::vector<char> pm(threads, 0), used(threads, 0);
    for_chunks(first, count, threads, [&](unsigned t, int lo, int hi) {
        bool mx;
        scan(lo, hi, pb[t], pc[t], mx);
        pm[t] = mx ? 1 : 0;
        used[t] = 1;
    });
    b.reset();
    cb.reset();
    mixed = false;
    for (unsigned t = 0; t < threads; t++) {
        if (!used[t]) continue;
        b.grow(pb[t]);
        cb.grow(pc[t]);
        mixed = mixed || pm[t];
    }
}

struct Builder {
    std::vector<BPrim>* prims_ptr = nullptr;  // shared permutation array; builders work on disjoint ranges
    std::vector<BNode> nodes;
    int max_depth = 0;
    // top-level builder only: subtrees with <= defer_below primitives are not built but recorded here
    int defer_below = 0;
    std::vector<Deferred> deferred;

    int make_leaf(int first, int count, int depth, const Box& b) {
        std::vector<BPrim>& prims = *prims_ptr;
        BNode n;
        memcpy(n.lo, b.lo, sizeof(n.lo));
        memcpy(n.hi, b.hi, sizeof(n.hi));
        n.first = first;
        n.count = count;
        n.type = prims[first].type;
        n.depth = depth;
        max_depth = std::max(max_depth, depth);
        nodes.push_back(n);
        return (int)nodes.size() - 1;
    }

    int build(int first, int count, int depth) {
        std::vector<BPrim>& prims = *prims_ptr;
        Box b, cb;
        bool mixed = false;
        range_bounds(prims, first, count, b, cb, mixed);
        int split = -1;  // prims [first, split) go left

        if (count <= kMaxLeafPrims && mixed) {
            // homogeneous leaves only: separate the types
            auto mid = std::stable_partition(prims.begin() + first, prims.begin() + first + count,
                                             [](const BPrim& p) { return p.type == 0; });
            split = (int)(mid - prims.begin());
        } else if (count == 1) {
            return make_leaf(first, count, depth, b);
        } else if (depth >= 32) {
            // depth guard: balanced median split on the longest centroid axis
            int axis = 0;
            for (int a = 1; a < 3; a++)
                if (cb.hi[a] - cb.lo[a] > cb.hi[axis] - cb.lo[axis]) axis = a;
            if (count <= kMaxLeafPrims) return make_leaf(first, count, depth, b);
            split = first + count / 2;
            std::nth_element(prims.begin() + first, prims.begin() + split, prims.begin() + first + count,
                             [axis](const BPrim& x, const BPrim& y) { return x.c[axis] < y.c[axis]; });
        } else {
            // binned SAH over the three axes
            double best_cost = std::numeric_limits<double>::infinity();
            int best_axis = -1, best_bin = -1;
            const double parent_area = b.area();
            // one pass over the primitives fills the bins of all three axes (large ranges: in parallel chunks, merged in a
            // fixed order — boxes and counts are order-independent, so the tree does not depend on the thread count)
            Bins3 bins;
            double scale3[3];
            bool axis_ok[3];
            for (int axis = 0; axis < 3; axis++) {
                axis_ok[axis] = cb.hi[axis] > cb.lo[axis];
                scale3[axis] = axis_ok[axis] ? kBins / (cb.hi[axis] - cb.lo[axis]) : 0.0;
            }
            fill_bins(prims, first, count, cb, scale3, axis_ok, bins);
            for (int axis = 0; axis < 3; axis++) {
                if (!axis_ok[axis]) continue;
                const Box* bin_box = bins.box[axis];
                const int* bin_cnt = bins.cnt[axis];
                double right_area[kBins];
                int right_cnt[kBins];
                Box acc;
                acc.reset();
                int cnt = 0;
                for (int k = kBins - 1; k > 0; k--) {
                    acc.grow(bin_box[k]);
                    cnt += bin_cnt[k];
                    right_area[k] = acc.area();
                    right_cnt[k] = cnt;
                }
                acc.reset();
                cnt = 0;
                for (int k = 0; k < kBins - 1; k++) {
                    acc.grow(bin_box[k]);
                    cnt += bin_cnt[k];
                    if (cnt == 0 || right_cnt[k + 1] == 0) continue;
                    double cost = acc.area() * cnt + right_area[k + 1] * right_cnt[k + 1];
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = axis;
                        best_bin = k;
                    }
                }
            }
            // depth 0 never becomes a leaf when it can be split: the root node stores its children's boxes
            if (best_axis >= 0 && count <= kMaxLeafPrims && depth > 0) {
                // leaf cost = count (one unit per primitive test); split cost = kNodeCost + SAH.  A node visit is a
                // dependent 64-byte fetch plus two slab tests: measured on the 100 k-sphere scene the walk is bound by
                // exactly those fetches, so a visit is priced at several primitive tests and leaves fill up to 4
                double split_cost = kNodeCost + (parent_area > 0 ? best_cost / parent_area : (double)count);
                if (split_cost >= (double)count) return make_leaf(first, count, depth, b);
            }
            if (best_axis < 0) {
                if (count <= kMaxLeafPrims && depth > 0) return make_leaf(first, count, depth, b);
                split = first + count / 2;  // identical centroids: split by index
            } else {
                const double cmin = cb.lo[best_axis], cmax = cb.hi[best_axis];
                const double scale = kBins / (cmax - cmin);
                auto mid = std::partition(prims.begin() + first, prims.begin() + first + count, [&](const BPrim& p) {
                    int k = std::min(kBins - 1, std::max(0, (int)((p.c[best_axis] - cmin) * scale)));
                    return k <= best_bin;
                });
                split = (int)(mid - prims.begin());
                if (split == first || split == first + count) split = first + count / 2;
            }
        }

        BNode n;
        memcpy(n.lo, b.lo, sizeof(n.lo));
        memcpy(n.hi, b.hi, sizeof(n.hi));
        n.depth = depth;
        n.first = first;
        n.count = count;
        nodes.push_back(n);
        int self = (int)nodes.size() - 1;
        const int lc = split - first, rc = first + count - split;
        int l = -2, r = -2;
        if (defer_below > 0 && lc <= defer_below && lc > kMaxLeafPrims) deferred.push_back(Deferred{self, 0, first, lc, depth + 1});
        else l = build(first, lc, depth + 1);
        if (defer_below > 0 && rc <= defer_below && rc > kMaxLeafPrims) deferred.push_back(Deferred{self, 1, split, rc, depth + 1});
        else r = build(split, rc, depth + 1);
        nodes[self].left = l;
        nodes[self].right = r;
        return self;
    }
};

inline float round_down(double v) {
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}
inline float round_up(double v) {
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}
inline float as_float(int32_t i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}

}  // namespace

void pack_sphere(const HostSphere& s, F4& geom, I2& meta) {
    geom = F4{(float)s.c[0], (float)s.c[1], (float)s.c[2], (float)s.r};
    meta = I2{s.mat, s.order};
}

void pack_triangle(const HostTriangle& t, F4 o[4]) {
    double e1[3], e2[3], nrm[3];
    for (int a = 0; a < 3; a++) {
        e1[a] = t.v[1][a] - t.v[0][a];
        e2[a] = t.v[2][a] - t.v[0][a];
    }
    // calculateNormal (triangle.go:30-34): normalize(e1 x e2), zero-safe (vector.go:61-67)
    nrm[0] = e1[1] * e2[2] - e1[2] * e2[1];
    nrm[1] = e1[2] * e2[0] - e1[0] * e2[2];
    nrm[2] = e1[0] * e2[1] - e1[1] * e2[0];
    double len = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    if (len == 0) nrm[0] = nrm[1] = nrm[2] = 0;
    else for (int a = 0; a < 3; a++) nrm[a] /= len;
    o[0] = F4{(float)t.v[0][0], (float)t.v[0][1], (float)t.v[0][2], as_float(t.mat)};
    o[1] = F4{(float)e1[0], (float)e1[1], (float)e1[2], as_float(t.order)};
    o[2] = F4{(float)e2[0], (float)e2[1], (float)e2[2], 0.f};
    o[3] = F4{(float)nrm[0], (float)nrm[1], (float)nrm[2], 0.f};
}

void build_bvh(const HostScene& scene, FlatBvh& out) {
    auto t0 = std::chrono::steady_clock::now();
    {
        static const char* const e = getenv("GORT_BVH_NODE_COST");
        kNodeCost = e ? atof(e) : 1.0;
    }
    out = FlatBvh();
    Builder B;
    const int nS = (int)scene.spheres.size(), nT = (int)scene.tris.size();
    std::vector<BPrim> prims_storage((size_t)nS + nT);
    // (asked once: the call reads /sys on every use, microseconds that a 7-sphere scene re-uploaded per frame would pay each time)
    static const unsigned hw_cores = std::thread::hardware_concurrency();
    unsigned hw = hw_cores;
    if (const char* e = getenv("GORT_BVH_THREADS")) hw = (unsigned)std::max(1, atoi(e));
    auto prim_boxes = [&](int lo, int hi) {
        for (int j = lo; j < hi; j++) {
            BPrim& p = prims_storage[j];
            if (j < nS) {
                const HostSphere& s = scene.spheres[j];
                const double r = std::fabs(s.r);
                for (int a = 0; a < 3; a++) {
                    p.lo[a] = s.c[a] - r;
                    p.hi[a] = s.c[a] + r;
                    p.c[a] = s.c[a];
                }
                p.type = 0;
                p.idx = j;
            } else {
                const HostTriangle& t = scene.tris[j - nS];
                for (int a = 0; a < 3; a++) {
                    p.lo[a] = std::min(t.v[0][a], std::min(t.v[1][a], t.v[2][a]));
                    p.hi[a] = std::max(t.v[0][a], std::max(t.v[1][a], t.v[2][a]));
                    p.c[a] = 0.5 * (p.lo[a] + p.hi[a]);
                }
                p.type = 1;
                p.idx = j - nS;
            }
        }
    };
    {
        const int np = nS + nT;
        const unsigned nt = (np >= 20000 && hw > 1) ? hw : 1u;
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; t++) {
            const int lo = (int)((long long)np * t / nt), hi = (int)((long long)np * (t + 1) / nt);
            if (t + 1 < nt) pool.emplace_back(prim_boxes, lo, hi);
            else prim_boxes(lo, hi);
        }
        for (auto& th : pool) th.join();
    }
    Box world;
    world.reset();
    for (const BPrim& p : prims_storage) world.grow(p.lo, p.hi);
    const int n = nS + nT;
    if (n == 0) return;

    // Large scenes: the top of the tree is built here, subtrees below ~n/64 primitives by worker threads (each on
    // its own disjoint range of the permutation array, into its own node vector), then spliced in.  The tree is
    // the same as the sequential one: every split depends only on the primitives of its own range.
    B.prims_ptr = &prims_storage;
    B.nodes.reserve((size_t)n);
    const bool parallel = n >= 20000 && hw > 1;
    if (parallel) B.defer_below = std::max(1024, n / 64);
    g_node_threads = parallel ? hw : 1u;  // the top of the tree scans its large ranges in parallel chunks
    static const bool times = getenv("GORT_BVH_TIMES") != nullptr;
    auto lap = [&](const char* what) {
        if (times) fprintf(stderr, "[gort bvh] %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    lap("primitive boxes");
    int root = B.build(0, n, 0);
    lap("top of the tree");
    if (parallel && !B.deferred.empty()) {
        const size_t nj = B.deferred.size();
        std::vector<Builder> subs(nj);
        std::vector<int> sub_root(nj, -1);
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                const size_t j = next.fetch_add(1);
                if (j >= nj) break;
                subs[j].prims_ptr = &prims_storage;
                subs[j].nodes.reserve((size_t)B.deferred[j].count);
                sub_root[j] = subs[j].build(B.deferred[j].first, B.deferred[j].count, B.deferred[j].depth);
            }
        };
        std::vector<std::thread> pool;
        const unsigned nt = (unsigned)std::min<size_t>(hw, nj);
        for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
        // splice: subtree j's nodes go to [off[j], off[j] + size_j) of the top builder's node array (copied in parallel)
        std::vector<int> off(nj);
        size_t total = B.nodes.size();
        for (size_t j = 0; j < nj; j++) {
            off[j] = (int)total;
            total += subs[j].nodes.size();
        }
        B.nodes.resize(total);
        std::atomic<size_t> next_copy{0};
        auto copy = [&]() {
            for (;;) {
                const size_t j = next_copy.fetch_add(1);
                if (j >= nj) break;
                BNode* dst = B.nodes.data() + off[j];
                for (size_t i = 0; i < subs[j].nodes.size(); i++) {
                    BNode nd = subs[j].nodes[i];
                    if (nd.left >= 0) nd.left += off[j];
                    if (nd.right >= 0) nd.right += off[j];
                    dst[i] = nd;
                }
            }
        };
        pool.clear();
        for (unsigned t = 1; t < nt; t++) pool.emplace_back(copy);
        copy();
        for (auto& th : pool) th.join();
        for (size_t j = 0; j < nj; j++) {
            const Deferred& d = B.deferred[j];
            (d.which ? B.nodes[d.parent].right : B.nodes[d.parent].left) = sub_root[j] + off[j];
            B.max_depth = std::max(B.max_depth, subs[j].max_depth);
        }
    }
    lap("subtrees + splice");
    // Nodes store the children's boxes in the parent, so a leaf root (single-primitive scene) needs
    // a wrapper; both slots reference the same leaf (the second test ties and changes nothing).
    if (B.nodes[root].left < 0) {
        BNode w = B.nodes[root];
        w.left = root;
        w.right = root;
        w.depth = 0;
        B.nodes.push_back(w);
        root = (int)B.nodes.size() - 1;
    }

    // conservative padding for fp32 slab arithmetic (see DESIGN.md "BVH")
    double extent = 0;
    for (int a = 0; a < 3; a++) extent = std::max(extent, std::max(std::fabs(world.lo[a]), std::fabs(world.hi[a])));
    const double pad = 4e-7 * std::max(1.0, extent);

    // Breadth-first numbering of the inner nodes (the order vector is its own queue).  The same traversal fixes where each
    // leaf's primitives go in the per-type arrays (leaves are laid out in the order their parents are numbered, child 0
    // before child 1), so the arrays can then be filled in parallel.
    std::vector<int32_t> flat_index(B.nodes.size(), -1), leaf_start(B.nodes.size(), -1);
    std::vector<int32_t> order;
    order.reserve(B.nodes.size());
    order.push_back(root);
    flat_index[root] = 0;
    uint32_t n_sph_out = 0, n_tri_out = 0;
    for (size_t head = 0; head < order.size(); head++) {
        const BNode& nd = B.nodes[order[head]];
        const int32_t kids[2] = {nd.left, nd.right};
        for (int c = 0; c < 2; c++) {
            const BNode& k = B.nodes[kids[c]];
            if (k.left != -1) {
                flat_index[kids[c]] = (int32_t)order.size();
                order.push_back(kids[c]);
            } else if (leaf_start[kids[c]] < 0) {  // (the single-primitive scene references its leaf from both root slots)
                uint32_t& cursor = k.type == 0 ? n_sph_out : n_tri_out;
                leaf_start[kids[c]] = (int32_t)cursor;
                cursor += (uint32_t)k.count;
            }
        }
    }
    out.n_nodes = (int32_t)order.size();
    out.nodes.resize((size_t)out.n_nodes * 6);
    // 16-bit grid over the (padded) world box, two spare cells on every side
    double qo[3], qc[3];
    for (int a = 0; a < 3; a++) {
        const double lo = world.lo[a] - pad, hi = world.hi[a] + pad;
        qc[a] = std::max((hi - lo) / 65520.0, 1e-30);
        qo[a] = lo - 8.0 * qc[a];
        out.qcell[a] = (float)qc[a];
        out.qorigin[a] = (float)qo[a];
        // the kernel decodes with the float values: quantise against exactly those
        qc[a] = (double)out.qcell[a];
        qo[a] = (double)out.qorigin[a];
    }
    F4* qbase = out.nodes.data() + (size_t)out.n_nodes * 4;
    out.spheres.resize(n_sph_out);
    out.sphere_meta.resize(n_sph_out);
    out.tris.resize((size_t)n_tri_out * 4);
    out.max_depth = B.max_depth + 1;

    auto emit_leaf = [&](const BNode& leaf, uint32_t start) -> int32_t {
        if (leaf.type == 0) {
            for (int i = 0; i < leaf.count; i++) pack_sphere(scene.spheres[prims_storage[leaf.first + i].idx], out.spheres[start + i], out.sphere_meta[start + i]);
        } else {
            for (int i = 0; i < leaf.count; i++) pack_triangle(scene.tris[prims_storage[leaf.first + i].idx], &out.tris[4 * (size_t)(start + i)]);
        }
        uint32_t v = (start & kLeafStartMask) | ((uint32_t)(leaf.count - 1) << kLeafCountShift) | ((uint32_t)leaf.type << kLeafTypeBit);
        return (int32_t)~v;
    };

    auto fill = [&](int32_t f0, int32_t f1) {
        for (int32_t fi = f0; fi < f1; fi++) {
            const BNode& nd = B.nodes[order[fi]];
            float lo[2][3], hi[2][3];
            int32_t child[2];
            const int32_t kids[2] = {nd.left, nd.right};
            for (int c = 0; c < 2; c++) {
                const BNode& k = B.nodes[kids[c]];
                for (int a = 0; a < 3; a++) {
                    lo[c][a] = round_down(k.lo[a] - pad);
                    hi[c][a] = round_up(k.hi[a] + pad);
                }
                child[c] = k.left == -1 ? emit_leaf(k, (uint32_t)leaf_start[kids[c]]) : flat_index[kids[c]];
            }
            out.nodes[(size_t)fi * 4 + 0] = F4{lo[0][0], hi[0][0], lo[0][1], hi[0][1]};
            out.nodes[(size_t)fi * 4 + 1] = F4{lo[1][0], hi[1][0], lo[1][1], hi[1][1]};
            out.nodes[(size_t)fi * 4 + 2] = F4{lo[0][2], hi[0][2], lo[1][2], hi[1][2]};
            out.nodes[(size_t)fi * 4 + 3] = F4{as_float(child[0]), as_float(child[1]), 0.f, 0.f};
            uint32_t qw[2][4];
            for (int c = 0; c < 2; c++) {
                for (int a = 0; a < 3; a++) {
                    const double ql = std::floor(((double)lo[c][a] - qo[a]) / qc[a]) - 2.0, qh = std::ceil(((double)hi[c][a] - qo[a]) / qc[a]) + 2.0;
                    const uint32_t l = (uint32_t)std::min(65535.0, std::max(0.0, ql)), h = (uint32_t)std::min(65535.0, std::max(0.0, qh));
                    qw[c][a] = l | (h << 16);
                }
                qw[c][3] = (uint32_t)child[c];
            }
            auto bits = [](uint32_t u) { float f; memcpy(&f, &u, 4); return f; };
            qbase[(size_t)fi * 2 + 0] = F4{bits(qw[0][0]), bits(qw[0][1]), bits(qw[0][2]), bits(qw[0][3])};
            qbase[(size_t)fi * 2 + 1] = F4{bits(qw[1][0]), bits(qw[1][1]), bits(qw[1][2]), bits(qw[1][3])};
        }
    };
    {
        const unsigned nt = (parallel && out.n_nodes >= 32768) ? hw : 1u;
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; t++) {
            const int32_t f0 = (int32_t)((long long)out.n_nodes * t / nt), f1 = (int32_t)((long long)out.n_nodes * (t + 1) / nt);
            if (t + 1 < nt) pool.emplace_back(fill, f0, f1);
            else fill(f0, f1);
        }
        for (auto& th : pool) th.join();
    }
    lap("flatten");
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace gort
