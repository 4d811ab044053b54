// host_scene.h — host-side scene model of libgort (float64, reference scan order) and the flat
// device layout produced from it.  Mirrors what scene.GetHittables()/GetLights() hand to Render
// (/root/reference internal/scene/scene.go:59-98) after createMaterial (:104-148) and createCube
// (:150-190).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gort.h"

namespace gort {

// The element types below are plain aggregates (no default member initialisers) kept in vectors whose resize() does not
// touch the new elements: scene_from_desc fills 10^6 of them from all host threads, and both the value-initialisation and
// the first touch of their pages would otherwise be one thread's work (60 of the 90 ms the call took for the 1 M-primitive
// scene on an 8-core host).
template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = default_init_allocator<U>;
    };
    default_init_allocator() = default;
    template <class U>
    default_init_allocator(const default_init_allocator<U>&) noexcept {}
    template <class U>
    void construct(U* p) {
        ::new (static_cast<void*>(p)) U;  // default-initialisation: nothing for a plain aggregate
    }
    template <class U, class... A>
    void construct(U* p, A&&... a) {
        ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
    }
};
template <class T>
using HostVec = std::vector<T, default_init_allocator<T>>;

struct HostMaterial {
    int32_t type;
    double color[3];
    double roughness, metallic, specular, ior;
};
// what a default-constructed material used to be (Lambertian, black, ior 1.5)
constexpr HostMaterial kDefaultMaterial = {GORT_MAT_LAMBERTIAN, {0, 0, 0}, 0, 0, 0, 1.5};

struct HostSphere {
    double c[3];
    double r;
    int32_t mat;
    int32_t order;  // position in the reference's flattened linear scan
};

struct HostTriangle {
    double v[3][3];
    int32_t mat;
    int32_t order;
};

struct HostLight {
    double pos[3], color[3], intensity;
};

struct HostScene {
    double cam_pos[3] = {0, 0, 0}, cam_look_at[3] = {0, 0, 0}, cam_up[3] = {0, 1, 0};
    double cam_fov = 60, cam_aspect = 1;
    HostVec<HostMaterial> mats;
    HostVec<HostSphere> spheres;
    HostVec<HostTriangle> tris;
    std::vector<HostLight> lights;
    int32_t n_hittables = 0;  // len(scene.GetHittables()): spheres + meshes
    int32_t fog_enabled = 0;
    double fog_density = 0, fog_color[3] = {0, 0, 0};
    // sky extension: AtmosphereConfig in declaration order (atmosphere/atmosphere.go:8-26), see gort_scene_desc::sky_params
    int32_t sky_enabled = 0;
    double sky_params[27] = {0};
    // the scene's own "renderer" block (README.md:285-291; ignored by the reference loader):
    // {samples, maxDepth, antiAliasing, recursiveReflections, softShadows}, -1 = key absent
    int32_t render_hints[5] = {-1, -1, -1, -1, -1};

    int32_t next_order() const { return (int32_t)(spheres.size() + tris.size()); }
};

// gort_scene_desc -> HostScene (validates indices and sizes). Returns "" or an error message.
std::string scene_from_desc(const gort_scene_desc& d, HostScene& out);

// Host mirror of scene.LoadFromFile + GetHittables for the reference's JSON schema
// (scene.go:12-39): returns "" or an error message.  options: bit0 prisms, bit1 fog, bit2 sky.
std::string scene_from_json(const char* text, size_t len, uint32_t options, HostScene& out);

// createMaterial's defaults and constructor clamps (scene.go:104-148; material.go:65-73,159-167;
// advanced_materials.go:14-19,117-123).  `has_*` = key present in the JSON object.
HostMaterial make_material(const std::string& type, bool has_color, const double color[3], bool has_rough, double rough,
                           bool has_metal, double metal, bool has_spec, double spec, bool has_ior, double ior);

// createCube (scene.go:150-190): 12 triangles appended in the reference's order.
void add_cube(HostScene& s, const double pos[3], const double size[3], int32_t mat);
// triangularPrism extension (README.md:227-240): 8 triangles.
void add_prism(HostScene& s, const double verts[6][3], int32_t mat);
void add_sphere(HostScene& s, const double pos[3], double radius, int32_t mat);

}  // namespace gort
