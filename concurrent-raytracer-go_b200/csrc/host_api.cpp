// host_api.cpp — host-only entry points of the C ABI (include/gort.h "host-only scene model"):
// the scene loader and the BVH builder can be exercised without a CUDA device.
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/gort.h"
#include "bvh.h"
#include "host_scene.h"

using namespace gort;

struct gort_host_scene {
    HostScene s;
};

static void put_err(char* buf, size_t len, const std::string& msg) {
    if (!buf || len == 0) return;
    size_t n = std::min(len - 1, msg.size());
    memcpy(buf, msg.data(), n);
    buf[n] = 0;
}

extern "C" {

int gort_host_scene_parse(const char* json_text, size_t json_len, uint32_t options, gort_host_scene** out, char* errbuf, size_t errbuf_len) {
    if (!out || !json_text) return GORT_ERR_INVALID;
    *out = nullptr;
    gort_host_scene* hs = new gort_host_scene();
    std::string err = scene_from_json(json_text, json_len, options, hs->s);
    if (!err.empty()) {
        put_err(errbuf, errbuf_len, err);
        delete hs;
        return GORT_ERR_PARSE;
    }
    *out = hs;
    return GORT_OK;
}

int gort_host_scene_from_desc(const gort_scene_desc* desc, gort_host_scene** scene_inout, char* errbuf, size_t errbuf_len) {
    if (!desc || !scene_inout) return GORT_ERR_INVALID;
    gort_host_scene* hs = *scene_inout ? *scene_inout : new gort_host_scene();
    std::string err = scene_from_desc(*desc, hs->s);
    if (!err.empty()) {
        put_err(errbuf, errbuf_len, err);
        if (!*scene_inout) delete hs;
        return GORT_ERR_INVALID;
    }
    *scene_inout = hs;
    return GORT_OK;
}

void gort_host_scene_free(gort_host_scene* scene) { delete scene; }

int gort_host_scene_counts(const gort_host_scene* scene, int32_t* c) {
    if (!scene || !c) return GORT_ERR_INVALID;
    c[0] = (int32_t)scene->s.spheres.size();
    c[1] = (int32_t)scene->s.tris.size();
    c[2] = (int32_t)scene->s.mats.size();
    c[3] = (int32_t)scene->s.lights.size();
    c[4] = scene->s.n_hittables;
    return GORT_OK;
}

int gort_host_scene_get_sphere(const gort_host_scene* scene, int32_t i, double* o, int32_t* material, int32_t* order) {
    if (!scene || !o || i < 0 || i >= (int32_t)scene->s.spheres.size()) return GORT_ERR_INVALID;
    const HostSphere& s = scene->s.spheres[i];
    o[0] = s.c[0]; o[1] = s.c[1]; o[2] = s.c[2]; o[3] = s.r;
    if (material) *material = s.mat;
    if (order) *order = s.order;
    return GORT_OK;
}

int gort_host_scene_get_triangle(const gort_host_scene* scene, int32_t i, double* v9, int32_t* material, int32_t* order) {
    if (!scene || !v9 || i < 0 || i >= (int32_t)scene->s.tris.size()) return GORT_ERR_INVALID;
    const HostTriangle& t = scene->s.tris[i];
    memcpy(v9, t.v, 9 * sizeof(double));
    if (material) *material = t.mat;
    if (order) *order = t.order;
    return GORT_OK;
}

int gort_host_scene_get_material(const gort_host_scene* scene, int32_t i, int32_t* type, double* o) {
    if (!scene || !o || i < 0 || i >= (int32_t)scene->s.mats.size()) return GORT_ERR_INVALID;
    const HostMaterial& m = scene->s.mats[i];
    if (type) *type = m.type;
    o[0] = m.color[0]; o[1] = m.color[1]; o[2] = m.color[2];
    o[3] = m.roughness; o[4] = m.metallic; o[5] = m.specular; o[6] = m.ior;
    return GORT_OK;
}

int gort_host_scene_get_light(const gort_host_scene* scene, int32_t i, double* o) {
    if (!scene || !o || i < 0 || i >= (int32_t)scene->s.lights.size()) return GORT_ERR_INVALID;
    const HostLight& l = scene->s.lights[i];
    memcpy(o, l.pos, 3 * sizeof(double));
    memcpy(o + 3, l.color, 3 * sizeof(double));
    o[6] = l.intensity;
    return GORT_OK;
}

int gort_host_scene_get_camera(const gort_host_scene* scene, double* o) {
    if (!scene || !o) return GORT_ERR_INVALID;
    memcpy(o, scene->s.cam_pos, 3 * sizeof(double));
    memcpy(o + 3, scene->s.cam_look_at, 3 * sizeof(double));
    memcpy(o + 6, scene->s.cam_up, 3 * sizeof(double));
    o[9] = scene->s.cam_fov;
    o[10] = scene->s.cam_aspect;
    return GORT_OK;
}

int gort_host_scene_bvh_validate(const gort_host_scene* scene, int64_t* info4, char* errbuf, size_t errbuf_len) {
    if (!scene) return GORT_ERR_INVALID;
    const HostScene& hs = scene->s;
    FlatBvh b;
    build_bvh(hs, b);
    const size_t nS = hs.spheres.size(), nT = hs.tris.size();
    auto bad = [&](const std::string& m) {
        put_err(errbuf, errbuf_len, m);
        return (int)GORT_ERR_INVALID;
    };
    if (b.spheres.size() != nS || b.sphere_meta.size() != nS || b.tris.size() != 4 * nT) return bad("leaf arrays do not cover every primitive");
    if (nS + nT == 0) {
        if (b.n_nodes != 0) return bad("empty scene must have no nodes");
        if (info4) info4[0] = info4[1] = info4[2] = info4[3] = 0;
        return GORT_OK;
    }
    if (b.n_nodes <= 0) return bad("no nodes");
    std::vector<char> seen_order(nS + nT, 0);
    std::vector<char> seen_node((size_t)b.n_nodes, 0);
    int64_t leaves = 0;
    int max_depth = 0;
    std::string err;
    auto as_int = [](float f) { int32_t i; memcpy(&i, &f, 4); return i; };
    // returns the exact bounds of the subtree under `child` and checks them against the stored box
    std::function<bool(int32_t, int, float*, float*)> walk = [&](int32_t child, int depth, float* lo, float* hi) -> bool {
        max_depth = std::max(max_depth, depth);
        if (depth > kMaxBvhDepth + 4) { err = "tree deeper than the traversal stack allows"; return false; }
        for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; }
        if (child < 0) {
            const uint32_t v = ~(uint32_t)child;
            const uint32_t start = v & kLeafStartMask;
            const int cnt = (int)((v >> kLeafCountShift) & 15u) + 1;
            const int type = (int)((v >> kLeafTypeBit) & 1u);
            if (cnt > kMaxLeafPrims) { err = "leaf larger than kMaxLeafPrims"; return false; }
            leaves++;
            for (int i = 0; i < cnt; i++) {
                int32_t order;
                if (type == 0) {
                    if (start + i >= nS) { err = "sphere leaf out of range"; return false; }
                    const F4& s = b.spheres[start + i];
                    order = b.sphere_meta[start + i].y;
                    const float r = std::fabs(s.w);
                    const float c[3] = {s.x, s.y, s.z};
                    for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], c[a] - r); hi[a] = std::max(hi[a], c[a] + r); }
                } else {
                    if (start + i >= nT) { err = "triangle leaf out of range"; return false; }
                    const F4* t = &b.tris[4 * (size_t)(start + i)];
                    order = as_int(t[1].w);
                    const float v0[3] = {t[0].x, t[0].y, t[0].z};
                    const float v1[3] = {t[0].x + t[1].x, t[0].y + t[1].y, t[0].z + t[1].z};
                    const float v2[3] = {t[0].x + t[2].x, t[0].y + t[2].y, t[0].z + t[2].z};
                    for (int a = 0; a < 3; a++) {
                        lo[a] = std::min(lo[a], std::min(v0[a], std::min(v1[a], v2[a])));
                        hi[a] = std::max(hi[a], std::max(v0[a], std::max(v1[a], v2[a])));
                    }
                }
                if (order < 0 || (size_t)order >= nS + nT) { err = "scan order out of range"; return false; }
                if (seen_order[order]) {
                    // the single-primitive scene duplicates its leaf in both root slots
                    if (!(nS + nT == 1)) { err = "primitive referenced twice"; return false; }
                }
                seen_order[order] = 1;
            }
            return true;
        }
        if (child >= b.n_nodes) { err = "child index out of range"; return false; }
        if (seen_node[child]) { err = "inner node referenced twice"; return false; }
        seen_node[child] = 1;
        const F4* n = &b.nodes[4 * (size_t)child];
        const int32_t c0 = as_int(n[3].x), c1 = as_int(n[3].y);
        const float blo[2][3] = {{n[0].x, n[0].z, n[2].x}, {n[1].x, n[1].z, n[2].z}};
        const float bhi[2][3] = {{n[0].y, n[0].w, n[2].y}, {n[1].y, n[1].w, n[2].w}};
        const int32_t kids[2] = {c0, c1};
        for (int k = 0; k < 2; k++) {
            float clo[3], chi[3];
            if (!walk(kids[k], depth + 1, clo, chi)) return false;
            const float slack = 1e-4f;  // triangle vertices are re-derived from fp32 edges
            for (int a = 0; a < 3; a++) {
                if (!(blo[k][a] <= clo[a] + slack * (1 + std::fabs(clo[a]))) || !(bhi[k][a] >= chi[a] - slack * (1 + std::fabs(chi[a])))) {
                    err = "child box does not enclose its subtree";
                    return false;
                }
                lo[a] = std::min(lo[a], clo[a]);
                hi[a] = std::max(hi[a], chi[a]);
            }
        }
        return true;
    };
    float lo[3], hi[3];
    if (!walk(0, 0, lo, hi)) return bad(err);
    for (size_t i = 0; i < seen_order.size(); i++)
        if (!seen_order[i]) return bad("primitive missing from the tree");
    for (int i = 0; i < b.n_nodes; i++)
        if (!seen_node[i]) return bad("unreachable inner node");
    if (info4) {
        info4[0] = b.n_nodes;
        info4[1] = max_depth;
        info4[2] = leaves;
        info4[3] = (int64_t)(b.nodes.size() * sizeof(F4) + b.spheres.size() * sizeof(F4) + b.sphere_meta.size() * sizeof(I2) + b.tris.size() * sizeof(F4));
    }
    return GORT_OK;
}

}  // extern "C"
