// Host-only timing of gort_scene_upload's first stage on a scene of the 1 M-primitive configuration's shape (no GPU):
// scene_from_desc into the spare scene + swap, as capi.cu does it.
//   g++ -O2 -std=c++17 -I concurrent-raytracer-go_b200/csrc -I include tools/desc_bench.cpp \
//       concurrent-raytracer-go_b200/build/host_scene.o -lpthread -o /tmp/desc_bench && /tmp/desc_bench 750000 249996
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "host_scene.h"
using namespace gort;
int main(int argc, char** argv) {
    const int nS = argc > 1 ? atoi(argv[1]) : 700000, nT = argc > 2 ? atoi(argv[2]) : 300000, reps = 8;
    const int nM = nS + nT / 12;
    std::vector<int32_t> mt(nM, 0), sm(nS), so(nS), tm(nT), to(nT);
    std::vector<double> mc(3 * (size_t)nM, 0.5), mr(nM, 0.1), mm(nM, 0.2), ms(nM, 0.3), mi(nM, 1.5), sc(3 * (size_t)nS, 1.0), sr(nS, 0.5), tv(9 * (size_t)nT, 2.0);
    for (int i = 0; i < nS; i++) { sm[i] = i; so[i] = i; }
    for (int i = 0; i < nT; i++) { tm[i] = nS + i / 12; to[i] = nS + i; }
    gort_scene_desc d{};
    d.abi_version = GORT_ABI_VERSION;
    d.n_materials = nM; d.mat_type = mt.data(); d.mat_color = mc.data(); d.mat_roughness = mr.data(); d.mat_metallic = mm.data(); d.mat_specular = ms.data(); d.mat_ior = mi.data();
    d.n_spheres = nS; d.sphere_center = sc.data(); d.sphere_radius = sr.data(); d.sphere_material = sm.data(); d.sphere_order = so.data();
    d.n_triangles = nT; d.tri_vertices = tv.data(); d.tri_material = tm.data(); d.tri_order = to.data();
    HostScene cur, spare;
    for (int r = 0; r < reps; r++) {
        auto t0 = std::chrono::steady_clock::now();
        std::string e = scene_from_desc(d, spare);
        std::swap(cur, spare);
        auto t1 = std::chrono::steady_clock::now();
        printf("scene_from_desc into the spare + swap %.2f ms  %s (%zu %zu %zu)\n", std::chrono::duration<double, std::milli>(t1 - t0).count(), e.c_str(), cur.mats.size(), cur.spheres.size(), cur.tris.size());
    }
}
