// host_scene.cpp — host mirror of the reference's scene package for the render path:
// scene.LoadFromFile / GetHittables / createMaterial / createCube
// (/root/reference internal/scene/scene.go:45-90,104-190), producing the flat HostScene.
#include "host_scene.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "json.h"

namespace gort {

static std::string lower(std::string s) {
    for (auto& c : s) c = (char)tolower((unsigned char)c);
    return s;
}

// Case-insensitive field lookup like encoding/json's struct decoding (exact match preferred).
static const json::Value* field(const json::Value* o, const char* name) {
    if (!o || !o->is_object()) return nullptr;
    if (const json::Value* v = o->get(name)) return v;
    std::string want = lower(name);
    const json::Value* found = nullptr;
    for (auto& kv : o->obj)
        if (lower(kv.first) == want) found = kv.second.get();
    return found;
}

// Vec3.UnmarshalJSON (internal/math/vector.go:176-193): [x,y,z] or {"X":..,"Y":..,"Z":..}
static bool vec3_field(const json::Value* v, double out[3], std::string& err, const char* what) {
    out[0] = out[1] = out[2] = 0;
    if (!v || v->kind == json::Value::Null) return true;  // absent -> zero value
    if (v->is_array()) {
        if (v->arr.size() != 3) {
            err = std::string(what) + ": expected 3 elements for Vec3, got " + std::to_string(v->arr.size());
            return false;
        }
        for (int i = 0; i < 3; i++) {
            if (!v->arr[i]->is_number()) {
                err = std::string(what) + ": Vec3 element is not a number";
                return false;
            }
            out[i] = v->arr[i]->num;
        }
        return true;
    }
    if (v->is_object()) {
        const char* names[3] = {"X", "Y", "Z"};
        for (int i = 0; i < 3; i++) {
            const json::Value* c = field(v, names[i]);
            if (c && c->is_number()) out[i] = c->num;
        }
        return true;
    }
    err = std::string(what) + ": cannot unmarshal into Vec3";
    return false;
}

static double num_field(const json::Value* o, const char* name, double dflt) {
    const json::Value* v = field(o, name);
    return (v && v->is_number()) ? v->num : dflt;
}

HostMaterial make_material(const std::string& type, bool has_color, const double color[3], bool has_rough, double rough,
                           bool has_metal, double metal, bool has_spec, double spec, bool has_ior, double ior) {
    HostMaterial m = kDefaultMaterial;
    // Declared extension: a material without "color" gets (1,1,1); the reference panics
    // on the unchecked assertion at scene.go:109/113/120/127/132/141/145.
    double c[3] = {1, 1, 1};
    if (has_color) memcpy(c, color, sizeof(c));
    auto set_color = [&]() { memcpy(m.color, c, sizeof(c)); };
    if (type == "metal") {  // scene.go:112-117 + NewMetal material.go:65-73
        m.type = GORT_MAT_METAL;
        set_color();
        m.roughness = std::fmin(has_rough ? rough : 0.0, 1.0);
        m.metallic = std::fmin(has_metal ? metal : 1.0, 1.0);
        m.specular = std::fmin(has_spec ? spec : 1.0, 1.0);
        m.ior = 1.5;
    } else if (type == "shiny") {  // scene.go:119-124 + NewShinyMaterial material.go:159-167
        m.type = GORT_MAT_SHINY;
        set_color();
        m.roughness = std::fmin(has_rough ? rough : 0.0, 1.0);
        m.metallic = std::fmin(has_metal ? metal : 0.0, 1.0);
        m.specular = std::fmin(has_spec ? spec : 1.0, 1.0);
        m.ior = 1.5;
    } else if (type == "perfectmirror") {  // scene.go:126-129 + NewPerfectMirror advanced_materials.go:117-123
        m.type = GORT_MAT_PERFECTMIRROR;
        set_color();
        m.roughness = std::fmin(has_rough ? rough : 0.0, 1.0);
        m.metallic = 1.0;  // GetMetallic advanced_materials.go:165
        m.specular = 1.0;
        m.ior = 2.0;
    } else if (type == "glass") {  // scene.go:131-134
        m.type = GORT_MAT_GLASS;
        set_color();
        m.ior = has_ior ? ior : 1.5;
        m.specular = 1.0;
    } else if (type == "dielectric") {  // scene.go:136-138
        m.type = GORT_MAT_DIELECTRIC;
        m.color[0] = m.color[1] = m.color[2] = 1.0;  // GetAlbedo material.go:266
        m.ior = has_ior ? ior : 1.5;
        m.specular = 1.0;
    } else if (type == "diffuselight") {  // scene.go:140-142
        m.type = GORT_MAT_DIFFUSELIGHT;
        set_color();  // Emit
        m.roughness = 1.0;
    } else {  // "lambertian" and the default branch scene.go:108-110,144-146
        m.type = GORT_MAT_LAMBERTIAN;
        set_color();
        m.roughness = 1.0;
    }
    return m;
}

static bool material_from_json(const json::Value* md, HostMaterial& out, std::string& err) {
    const json::Value* t = md ? md->get("type") : nullptr;
    if (!t || !t->is_string()) {
        err = "material.type missing or not a string (reference panics at scene.go:105)";
        return false;
    }
    double color[3] = {0, 0, 0};
    bool has_color = false;
    if (const json::Value* c = md->get("color")) {
        if (!c->is_array() || c->arr.size() < 3 || !c->arr[0]->is_number() || !c->arr[1]->is_number() || !c->arr[2]->is_number()) {
            err = "material.color is not an array of >= 3 numbers (reference panics in parseVec3, scene.go:211-217)";
            return false;
        }
        for (int i = 0; i < 3; i++) color[i] = c->arr[i]->num;
        has_color = true;
    }
    auto opt = [&](const char* key, bool& has, double& val) -> bool {
        has = false;
        val = 0;
        if (const json::Value* v = md->get(key)) {
            if (!v->is_number()) {
                err = std::string("material.") + key + " is not a number (reference panics in getFloat, scene.go:219-224)";
                return false;
            }
            has = true;
            val = v->num;
        }
        return true;
    };
    bool hr, hm, hs, hi;
    double r, m, s, i;
    if (!opt("roughness", hr, r) || !opt("metallic", hm, m) || !opt("specular", hs, s) || !opt("refractionIndex", hi, i)) return false;
    out = make_material(t->str, has_color, color, hr, r, hm, m, hs, s, hi, i);
    return true;
}

void add_sphere(HostScene& s, const double pos[3], double radius, int32_t mat) {
    HostSphere sp;
    memcpy(sp.c, pos, sizeof(sp.c));
    sp.r = radius;
    sp.mat = mat;
    sp.order = s.next_order();
    s.spheres.push_back(sp);
    s.n_hittables++;
}

static void push_tri(HostScene& s, const double* a, const double* b, const double* c, int32_t mat) {
    HostTriangle t;
    memcpy(t.v[0], a, 3 * sizeof(double));
    memcpy(t.v[1], b, 3 * sizeof(double));
    memcpy(t.v[2], c, 3 * sizeof(double));
    t.mat = mat;
    t.order = s.next_order();
    s.tris.push_back(t);
}

void add_cube(HostScene& s, const double pos[3], const double size[3], int32_t mat) {
    const double h[3] = {size[0] / 2.0, size[1] / 2.0, size[2] / 2.0};  // size.DivScalar(2.0) scene.go:151
    static const int sgn[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
    double v[8][3];
    for (int i = 0; i < 8; i++)
        for (int a = 0; a < 3; a++) v[i][a] = pos[a] + (sgn[i][a] < 0 ? -h[a] : h[a]);  // scene.go:153-162
    static const int faces[6][4] = {{0, 1, 2, 3}, {1, 5, 6, 2}, {5, 4, 7, 6}, {4, 0, 3, 7}, {3, 2, 6, 7}, {4, 5, 1, 0}};  // scene.go:164-171
    for (auto& f : faces) {  // scene.go:175-185
        push_tri(s, v[f[0]], v[f[1]], v[f[2]], mat);
        push_tri(s, v[f[0]], v[f[2]], v[f[3]], mat);
    }
    s.n_hittables++;
}

void add_prism(HostScene& s, const double verts[6][3], int32_t mat) {
    push_tri(s, verts[0], verts[1], verts[2], mat);
    push_tri(s, verts[3], verts[5], verts[4], mat);
    static const int quads[3][4] = {{0, 1, 4, 3}, {1, 2, 5, 4}, {2, 0, 3, 5}};
    for (auto& q : quads) {
        push_tri(s, verts[q[0]], verts[q[1]], verts[q[2]], mat);
        push_tri(s, verts[q[0]], verts[q[2]], verts[q[3]], mat);
    }
    s.n_hittables++;
}

std::string scene_from_json(const char* text, size_t len, uint32_t options, HostScene& out) {
    std::string err;
    json::ValuePtr root = json::parse(text, len, err);
    if (!root) return "error parsing JSON: " + err;  // scene.go:52-54
    if (!root->is_object()) return "error parsing JSON: top-level value is not an object";
    out = HostScene();

    const json::Value* cam = field(root.get(), "camera");
    if (!vec3_field(field(cam, "position"), out.cam_pos, err, "camera.position")) return err;
    if (!vec3_field(field(cam, "lookAt"), out.cam_look_at, err, "camera.lookAt")) return err;
    if (!vec3_field(field(cam, "up"), out.cam_up, err, "camera.up")) return err;
    out.cam_fov = num_field(cam, "fov", 0.0);
    out.cam_aspect = num_field(cam, "aspectRatio", 0.0);

    const json::Value* objs = field(root.get(), "objects");
    if (objs && objs->is_array()) {
        for (size_t i = 0; i < objs->arr.size(); i++) {
            const json::Value* o = objs->arr[i].get();
            const json::Value* t = field(o, "type");
            std::string type = (t && t->is_string()) ? t->str : "";
            bool prism = (type == "triangularPrism") && (options & 1u);
            if (type != "sphere" && type != "cube" && !prism) continue;  // "Unknown object type" scene.go:80-82
            HostMaterial m = kDefaultMaterial;
            if (!material_from_json(field(o, "material"), m, err)) return "object " + std::to_string(i + 1) + ": " + err;
            int32_t mi = (int32_t)out.mats.size();
            out.mats.push_back(m);
            double pos[3];
            if (!vec3_field(field(o, "position"), pos, err, "object.position")) return err;
            if (type == "sphere") {
                add_sphere(out, pos, num_field(o, "radius", 0.0), mi);
            } else if (type == "cube") {
                double size[3];
                if (!vec3_field(field(o, "size"), size, err, "object.size")) return err;
                add_cube(out, pos, size, mi);
            } else {
                const json::Value* vs = field(o, "vertices");
                if (!vs || !vs->is_array() || vs->arr.size() != 6) return "triangularPrism needs 6 vertices";
                double v[6][3];
                for (int k = 0; k < 6; k++)
                    if (!vec3_field(vs->arr[k].get(), v[k], err, "triangularPrism.vertices")) return err;
                add_prism(out, v, mi);
            }
        }
    }

    const json::Value* lights = field(root.get(), "lights");
    if (lights && lights->is_array()) {
        for (auto& lp : lights->arr) {
            HostLight l;
            if (!vec3_field(field(lp.get(), "position"), l.pos, err, "light.position")) return err;
            if (!vec3_field(field(lp.get(), "color"), l.color, err, "light.color")) return err;
            l.intensity = num_field(lp.get(), "intensity", 0.0);
            out.lights.push_back(l);
        }
    }

    {   // always parsed, never applied by the library: hosts that want the scene's own settings ask for them
        const json::Value* rb = field(root.get(), "renderer");
        if (rb) {
            const json::Value* v;
            if ((v = field(rb, "samples")) && v->kind == json::Value::Number) out.render_hints[0] = (int32_t)v->num;
            if ((v = field(rb, "maxDepth")) && v->kind == json::Value::Number) out.render_hints[1] = (int32_t)v->num;
            const char* flags[3] = {"antiAliasing", "recursiveReflections", "softShadows"};
            for (int k = 0; k < 3; k++)
                if ((v = field(rb, flags[k])) && v->kind == json::Value::Bool) out.render_hints[2 + k] = v->b ? 1 : 0;
        }
    }
    if (options & 2u) {
        const json::Value* fog = field(root.get(), "fog");
        const json::Value* en = field(fog, "enabled");
        if (en && en->kind == json::Value::Bool && en->b) {
            out.fog_enabled = 1;
            out.fog_density = num_field(fog, "density", 0.0);
            if (!vec3_field(field(fog, "color"), out.fog_color, err, "fog.color")) return err;
        }
    }
    if (options & 4u) {
        // "sky" block (extension): {"enabled": true, "preset": "default"|"white"|"sunset"|"night", per-field overrides} ->
        // AtmosphereConfig.  Presets: NewDefault/White/Sunset/NightAtmosphere (atmosphere/atmosphere.go:28-98).
        const json::Value* sky = field(root.get(), "sky");
        const json::Value* en = field(sky, "enabled");
        if (en && en->kind == json::Value::Bool && en->b) {
            static const double presets[4][27] = {
                {0.6, 0.8, 1.0, 0.9, 0.95, 1.0, 0.0, 0.8, -0.6, 1.0, 0.98, 0.95, 1.2, 0.015, 0.6, 0.8, 1.0, 1.0, 0.98, 0.95, 0.3, 0.0, 0.9, 0.92, 0.95, 0.05, 0.6},
                {0.98, 0.98, 1.0, 0.92, 0.92, 0.95, 0.0, 0.8, -0.6, 1.0, 0.99, 0.97, 0.8, 0.012, 0.9, 0.9, 0.95, 0.95, 0.95, 0.98, 0.2, 0.0, 0.95, 0.95, 0.98, 0.02, 0.6},
                {1.0, 0.4, 0.2, 1.0, 0.8, 0.6, 0.0, 0.3, -0.9, 1.0, 0.6, 0.3, 1.2, 0.03, 1.0, 0.4, 0.2, 1.0, 0.8, 0.6, 0.8, 0.1, 1.0, 0.8, 0.6, 0.3, 0.8},
                {0.1, 0.1, 0.3, 0.2, 0.2, 0.4, 0.0, -0.7, -0.7, 0.8, 0.8, 1.0, 0.3, 0.005, 0.1, 0.1, 0.3, 0.8, 0.8, 1.0, 0.2, 0.0, 0.1, 0.1, 0.2, 0.0, 0.0}};
            int which = 0;
            const json::Value* pv = field(sky, "preset");
            if (pv && pv->kind == json::Value::String) {
                const std::string name = lower(pv->str);
                if (name == "default") which = 0;
                else if (name == "white") which = 1;
                else if (name == "sunset") which = 2;
                else if (name == "night") which = 3;
                else return "sky.preset: unknown preset " + pv->str;
            }
            memcpy(out.sky_params, presets[which], sizeof(out.sky_params));
            struct F { const char* name; int off, n; };
            static const F fields[] = {{"skyColorTop", 0, 3}, {"skyColorBottom", 3, 3}, {"sunDirection", 6, 3}, {"sunColor", 9, 3},
                                       {"sunIntensity", 12, 1}, {"sunSize", 13, 1}, {"rayleighScattering", 14, 3}, {"mieScattering", 17, 3},
                                       {"atmosphericDepth", 20, 1}, {"fogDensity", 21, 1}, {"fogColor", 22, 3}, {"hazeIntensity", 25, 1},
                                       {"timeOfDay", 26, 1}};
            for (const F& f : fields) {
                const json::Value* v = field(sky, f.name);
                if (!v) continue;
                if (f.n == 3) {
                    double t[3];
                    if (!vec3_field(v, t, err, f.name)) return err;
                    memcpy(out.sky_params + f.off, t, sizeof(t));
                } else if (v->is_number()) {
                    out.sky_params[f.off] = v->num;
                }
            }
            out.sky_enabled = 1;
        }
    }
    return "";
}

std::string scene_from_desc(const gort_scene_desc& d, HostScene& out) {
    if (d.abi_version != GORT_ABI_VERSION) return "gort_scene_desc.abi_version mismatch";
    if (d.n_materials < 0 || d.n_spheres < 0 || d.n_triangles < 0 || d.n_lights < 0) return "negative count in gort_scene_desc";
    if (d.n_materials > 0 && (!d.mat_type || !d.mat_color || !d.mat_roughness || !d.mat_metallic || !d.mat_specular || !d.mat_ior))
        return "material arrays missing";
    if (d.n_spheres > 0 && (!d.sphere_center || !d.sphere_radius || !d.sphere_material || !d.sphere_order)) return "sphere arrays missing";
    if (d.n_triangles > 0 && (!d.tri_vertices || !d.tri_material || !d.tri_order)) return "triangle arrays missing";
    if (d.n_lights > 0 && (!d.light_position || !d.light_color || !d.light_intensity)) return "light arrays missing";
    {
        // `out` may be a scene built earlier: its arrays are reused (a re-upload of a scene of the same size then allocates
        // and faults in nothing), every other field starts from the defaults
        HostVec<HostMaterial> m = std::move(out.mats);
        HostVec<HostSphere> s = std::move(out.spheres);
        HostVec<HostTriangle> t = std::move(out.tris);
        out = HostScene();
        out.mats = std::move(m); out.spheres = std::move(s); out.tris = std::move(t);
    }
    memcpy(out.cam_pos, d.cam_position, sizeof(out.cam_pos));
    memcpy(out.cam_look_at, d.cam_look_at, sizeof(out.cam_look_at));
    memcpy(out.cam_up, d.cam_up, sizeof(out.cam_up));
    out.cam_fov = d.cam_fov;
    out.cam_aspect = d.cam_aspect;
    // large scenes (10^5..10^6 primitives, one material per object): the copies below run on all host threads
    const int64_t n_prims = (int64_t)d.n_spheres + d.n_triangles;
    const unsigned hw = (n_prims + d.n_materials) >= 100000 ? std::max(1u, std::min(16u, std::thread::hardware_concurrency())) : 1u;
    std::vector<const char*> errs(hw, nullptr);
    auto parallel = [&](int64_t count, auto&& body) {
        std::vector<std::thread> pool;
        auto run = [&](unsigned t) {
            for (int64_t i = count * t / hw; i < count * (t + 1) / hw && !errs[t]; i++)
                if (const char* e = body(i)) errs[t] = e;
        };
        for (unsigned t = 1; t < hw; t++) pool.emplace_back(run, t);
        run(0);
        for (auto& th : pool) th.join();
    };
    out.mats.resize(d.n_materials);
    parallel(d.n_materials, [&](int64_t i) -> const char* {
        HostMaterial& m = out.mats[i];
        m.type = d.mat_type[i];
        if (m.type < GORT_MAT_LAMBERTIAN || m.type > GORT_MAT_DIFFUSELIGHT) return "unknown material type";
        memcpy(m.color, d.mat_color + 3 * i, sizeof(m.color));
        m.roughness = d.mat_roughness[i];
        m.metallic = d.mat_metallic[i];
        m.specular = d.mat_specular[i];
        m.ior = d.mat_ior[i];
        if (m.type == GORT_MAT_PERFECTMIRROR) m.metallic = 1.0;  // GetMetallic advanced_materials.go:165
        if (m.type == GORT_MAT_DIELECTRIC) m.color[0] = m.color[1] = m.color[2] = 1.0;  // GetAlbedo material.go:266
        return nullptr;
    });
    std::vector<unsigned char> seen((size_t)n_prims, 0);
    auto check_order = [&](int32_t o) -> bool {  // every scan position taken exactly once (atomic: the loops run on several threads)
        if (o < 0 || o >= n_prims) return false;
        return __atomic_exchange_n(&seen[o], (unsigned char)1, __ATOMIC_RELAXED) == 0;
    };
    out.spheres.resize(d.n_spheres);
    parallel(d.n_spheres, [&](int64_t i) -> const char* {
        HostSphere& s = out.spheres[i];
        memcpy(s.c, d.sphere_center + 3 * i, sizeof(s.c));
        s.r = d.sphere_radius[i];
        s.mat = d.sphere_material[i];
        s.order = d.sphere_order[i];
        if (s.mat < 0 || s.mat >= d.n_materials) return "sphere material index out of range";
        if (!check_order(s.order)) return "sphere_order is not a permutation of the scan order";
        return nullptr;
    });
    out.tris.resize(d.n_triangles);
    parallel(d.n_triangles, [&](int64_t i) -> const char* {
        HostTriangle& t = out.tris[i];
        memcpy(t.v, d.tri_vertices + 9 * i, sizeof(t.v));
        t.mat = d.tri_material[i];
        t.order = d.tri_order[i];
        if (t.mat < 0 || t.mat >= d.n_materials) return "triangle material index out of range";
        if (!check_order(t.order)) return "tri_order is not a permutation of the scan order";
        return nullptr;
    });
    for (const char* e : errs)
        if (e) return e;
    out.lights.resize(d.n_lights);
    for (int i = 0; i < d.n_lights; i++) {
        HostLight& l = out.lights[i];
        memcpy(l.pos, d.light_position + 3 * i, sizeof(l.pos));
        memcpy(l.color, d.light_color + 3 * i, sizeof(l.color));
        l.intensity = d.light_intensity[i];
    }
    out.n_hittables = d.n_spheres + (d.n_triangles + 11) / 12;
    out.fog_enabled = d.fog_enabled;
    out.fog_density = d.fog_density;
    memcpy(out.fog_color, d.fog_color, sizeof(out.fog_color));
    out.sky_enabled = d.sky_enabled;
    memcpy(out.sky_params, d.sky_params, sizeof(out.sky_params));
    return "";
}

}  // namespace gort
