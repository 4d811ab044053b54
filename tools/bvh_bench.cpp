// Host-only timing of build_bvh on synthetic sphere/cube soups (no GPU): prints build time and an FNV hash of the flat
// arrays, so that a builder change can be checked for "same tree" and for speed.
//   g++ -O2 -std=c++17 -I concurrent-raytracer-go_b200/csrc tools/bvh_bench.cpp concurrent-raytracer-go_b200/build/bvh.o \
//       concurrent-raytracer-go_b200/build/host_scene.o -lpthread -o /tmp/bvh_bench && /tmp/bvh_bench 100000 0
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bvh.h"

using namespace gort;

static uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double uni(uint64_t& s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char** argv) {
    const int n_sph = argc > 1 ? atoi(argv[1]) : 100000, n_cubes = argc > 2 ? atoi(argv[2]) : 0, reps = argc > 3 ? atoi(argv[3]) : 3;
    const double ext = n_sph + 12 * n_cubes > 300000 ? 200.0 : 50.0;
    HostScene hs;
    hs.mats.push_back(kDefaultMaterial);
    uint64_t s = 20240601;
    for (int i = 0; i < n_sph; i++) {
        double p[3] = {(uni(s) * 2 - 1) * ext, (uni(s) * 2 - 1) * ext, (uni(s) * 2 - 1) * ext};
        add_sphere(hs, p, 0.1 + 0.4 * uni(s), 0);
    }
    for (int i = 0; i < n_cubes; i++) {
        double p[3] = {(uni(s) * 2 - 1) * ext, (uni(s) * 2 - 1) * ext, (uni(s) * 2 - 1) * ext};
        double sz[3] = {0.2 + uni(s), 0.2 + uni(s), 0.2 + uni(s)};
        add_cube(hs, p, sz, 0);
    }
    for (int r = 0; r < reps; r++) {
        FlatBvh b;
        auto t0 = std::chrono::steady_clock::now();
        build_bvh(hs, b);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](const void* p, size_t n) {
            const unsigned char* c = (const unsigned char*)p;
            for (size_t i = 0; i < n; i++) h = (h ^ c[i]) * 1099511628211ull;
        };
        mix(b.nodes.data(), b.nodes.size() * sizeof(F4));
        mix(b.spheres.data(), b.spheres.size() * sizeof(F4));
        mix(b.sphere_meta.data(), b.sphere_meta.size() * sizeof(I2));
        mix(b.tris.data(), b.tris.size() * sizeof(F4));
        printf("prims %d nodes %d depth %d build %.1f ms hash %016llx\n", n_sph + 12 * n_cubes, b.n_nodes, b.max_depth, ms, (unsigned long long)h);
    }
    return 0;
}
