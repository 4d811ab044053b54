// oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A float64 CPU restatement of the per-pixel render hot path of
// JoshElkind/concurrent-raytracer-go, written as a literal transcription (same
// operation order, same float64 arithmetic) of the Go sources cited next to every
// function (paths relative to /root/reference).  It is the parity checker for the
// CUDA path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.  The product (libgort.so) never links it.
//
// PARITY PIN STATUS: the reference ships no golden images and no test for
// intersection / scatter / shading / camera / tone-map; its RNG is
// wall-clock seeded and there is no Go toolchain in this image, so the reference
// itself cannot be run here.  What IS pinned: the nine known-answer vectors of
// internal/math/vector_test.go:8-105 and the value checks of
// internal/math/math_benchmarks_test.go:126-165 (tests/test_oracle_kat.py).  For
// every other function on the path: "parity unpinned" — the pin is this
// transcription, reviewable against the cited lines.
//
// Declared extensions (not in the reference; each OFF in reference mode):
//   * jitter=0  -> sub-pixel offset (0.5,0.5)          (reference always jitters, renderer.go:155-156)
//   * metal/shiny/... material without "color" -> (1,1,1) (reference panics, scene.go:113)
//   * rng_mode=PHILOX: counter-based Philox4x32-10 stream keyed on
//     (pixel, sample, bounce, purpose), identical to the CUDA path, so CPU and GPU can
//     be compared sample-for-sample; in this mode the unit-ball sampler is a loop-free
//     mapping with the same distribution as the reference's rejection loop (see Rng).  rng_mode=MT is a sequential mt19937_64 per worker
//     thread (the reference's global math/rand stream is not reproducible either).
//   * NaN after tone-map -> 0 (Go leaves uint8(NaN) implementation-defined).
//   * camera_mode=LOOKAT, triangularPrism objects, exponential fog, sky gradient: SURVEY §8(f).
//
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <vector>
#include <thread>
#include <atomic>
#include <random>
#include <limits>
#include <algorithm>
#include <memory>

namespace orc {

// ---------------------------------------------------------------------------
// Vec3 — internal/math/vector.go:9-122
// ---------------------------------------------------------------------------
struct V3 {
    double x = 0, y = 0, z = 0;
};
static inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }          // vector.go:17
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }          // vector.go:21
static inline V3 mul(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }          // vector.go:25
static inline V3 muls(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }           // vector.go:29
static inline V3 divs(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }           // vector.go:33
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      // vector.go:37
static inline V3 cross(V3 a, V3 b) {                                                     // vector.go:41
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline double length(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }  // vector.go:49
static inline double length_squared(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }     // vector.go:57
static inline V3 normalize(V3 a) {                                                       // vector.go:61-67
    double l = length(a);
    if (l == 0) return V3{};
    return divs(a, l);
}
static inline V3 reflect(V3 v, V3 n) { return sub(v, muls(n, 2 * dot(v, n))); }         // vector.go:77
static inline V3 refract(V3 v, V3 normal, double eta) {                                  // vector.go:81-96
    double cosTheta = dot(v, normal);
    if (cosTheta > 0) {
        normal = muls(normal, -1);
        eta = 1 / eta;
        cosTheta = -cosTheta;
    }
    double sinTheta2 = eta * eta * (1 - cosTheta * cosTheta);
    if (sinTheta2 > 1) return reflect(v, normal);
    double cosTheta2 = std::sqrt(1 - sinTheta2);
    return sub(muls(v, eta), muls(normal, eta * cosTheta + cosTheta2));
}
// Go's math.Max / math.Min propagate NaN (C fmax/fmin do not) — SURVEY §8c.
static inline double go_max(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return std::numeric_limits<double>::quiet_NaN();
    return a > b ? a : b;
}
static inline double go_min(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return std::numeric_limits<double>::quiet_NaN();
    return a < b ? a : b;
}
static inline V3 clamp(V3 v, double lo, double hi) {                                     // vector.go:98-104
    return {go_max(lo, go_min(hi, v.x)), go_max(lo, go_min(hi, v.y)), go_max(lo, go_min(hi, v.z))};
}
static inline uint8_t to_u8(double c) {                                                  // vector.go:108 uint8(c*255) truncating
    if (std::isnan(c)) return 0;  // declared: NaN -> 0
    return (uint8_t)(c * 255);
}
static inline void to_rgb(V3 v, uint8_t* rgb) {                                          // vector.go:106-109
    V3 c = clamp(v, 0, 1);
    rgb[0] = to_u8(c.x);
    rgb[1] = to_u8(c.y);
    rgb[2] = to_u8(c.z);
}
static inline bool near_zero(V3 v) {                                                     // vector.go:111-114
    const double s = 1e-8;
    return std::fabs(v.x) < s && std::fabs(v.y) < s && std::fabs(v.z) < s;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw — SC'11), the published algorithm.
// Same stream definition as the CUDA path (DESIGN.md "RNG streams").
// ---------------------------------------------------------------------------
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum RngMode { RNG_MT = 0, RNG_PHILOX = 1 };
enum Stream { STREAM_JITTER = 0, STREAM_SCATTER = 1, STREAM_SHADOW = 2 };

struct Counters {
    uint64_t samples = 0, hit_world = 0, sphere_tests = 0, tri_tests = 0, shadow_rays = 0, scatters = 0;
    void add(const Counters& o) {
        samples += o.samples; hit_world += o.hit_world; sphere_tests += o.sphere_tests;
        tri_tests += o.tri_tests; shadow_rays += o.shadow_rays; scatters += o.scatters;
    }
};

// Per-worker RNG context.  MT: sequential draws (≙ math/rand global stream, random.go:12-14).
// PHILOX: every draw site names (stream, seq); (pixel, sample, bounce) are set by the tracer.
struct Rng {
    int mode = RNG_MT;
    std::mt19937_64 mt;
    uint32_t key[2] = {0, 0};
    uint32_t pixel = 0, sample = 0, bounce = 0;

    double mt_float() {  // rand.Float64(): uniform [0,1) with 53 bits
        return (double)(mt() >> 11) * (1.0 / 9007199254740992.0);
    }
    void block(uint32_t stream, uint32_t seq, uint32_t out[4]) const {
        uint32_t ctr[4] = {pixel, sample, (bounce << 8) | stream, seq};
        philox4x32_10(ctr, key, out);
    }
    // RandomFloat() — random.go:12-14
    double uniform(uint32_t stream, uint32_t seq, int lane) {
        if (mode == RNG_MT) return mt_float();
        uint32_t r[4];
        block(stream, seq, r);
        return (double)(r[lane] >> 8) * (1.0 / 16777216.0);
    }
    // RandomVec3InUnitSphere() — vector.go:132-139: uniform point in the open unit ball.
    // MT mode: the reference's own rejection loop from [-1,1)^3, draw for draw.
    // PHILOX mode (the CUDA path's stream): the SAME distribution without a data-dependent loop — one
    // block gives three 24-bit uniforms (u1,u2,u3) and
    //     z = 1 - 2 u1,  phi = 2 pi u2,  r = cbrt(u3),  p = r (sqrt(1-z^2) cos phi, sqrt(1-z^2) sin phi, z)
    // (uniform direction by Archimedes' hat-box theorem, radius by inverting P(R<r) = r^3).  A rejection
    // loop makes a 32-wide warp wait for its slowest lane (measured: 13 active lanes per instruction);
    // tests/test_oracle_kat.py checks both samplers against the uniform-ball law.
    V3 in_unit_sphere(uint32_t stream, uint32_t seq) {
        if (mode == RNG_MT) {
            for (;;) {
                V3 r{mt_float(), mt_float(), mt_float()};                  // RandomVec3() vector.go:124-130
                V3 p = sub(muls(r, 2), V3{1, 1, 1});                        // vector.go:134
                if (length_squared(p) < 1) return p;
            }
        }
        uint32_t r[4];
        block(stream, seq, r);
        const double k = 1.0 / 16777216.0;
        const double u1 = (double)(r[0] >> 8) * k, u2 = (double)(r[1] >> 8) * k, u3 = (double)(r[2] >> 8) * k;
        const double z = 1.0 - 2.0 * u1;
        const double sxy = std::sqrt(std::fmax(0.0, 1.0 - z * z));
        const double phi = 6.283185307179586476925286766559 * u2;
        const double rad = std::cbrt(u3);
        return V3{rad * sxy * std::cos(phi), rad * sxy * std::sin(phi), rad * z};
    }
    // The soft-shadow samples (renderer.go:313-316) draw TWO ball points from one Philox block in PHILOX
    // mode (the CUDA path gives a lane two shadow rays per block): half 0 from words (0,1), half 1 from
    // words (2,3); u1 = 21 bits, u2 = 21 bits, u3 = 22 bits, then the same loop-free map as above.
    // MT mode: the reference's rejection loop, one call per sample.
    V3 in_unit_sphere_half(uint32_t stream, uint32_t seq, int half) {
        if (mode == RNG_MT) return in_unit_sphere(stream, seq);
        uint32_t r[4];
        block(stream, seq, r);
        const uint32_t a = r[2 * half], b = r[2 * half + 1];
        const double u1 = (double)(a >> 11) * (1.0 / 2097152.0), u2 = (double)(b >> 11) * (1.0 / 2097152.0);
        const double u3 = (double)(((a & 0x7FFu) << 11) | (b & 0x7FFu)) * (1.0 / 4194304.0);
        const double z = 1.0 - 2.0 * u1;
        const double sxy = std::sqrt(std::fmax(0.0, 1.0 - z * z));
        const double phi = 6.283185307179586476925286766559 * u2;
        const double rad = std::cbrt(u3);
        return V3{rad * sxy * std::cos(phi), rad * sxy * std::sin(phi), rad * z};
    }
};

// ---------------------------------------------------------------------------
// Ray / HitRecord — internal/geometry/ray.go:7-38
// ---------------------------------------------------------------------------
struct Ray {
    V3 o, d;  // NewRay does NOT normalise (ray.go:29-34)
};
static inline V3 ray_at(const Ray& r, double t) { return add(r.o, muls(r.d, t)); }  // ray.go:36-38

struct HitRecord {
    double t = 0;
    V3 point, normal;
    bool front_face = false;
    int material = -1;
    int prim = -1;  // flattened primitive index (diagnostic; not in the reference)
};

// ---------------------------------------------------------------------------
// Materials — internal/material/material.go, advanced_materials.go; factory scene.go:104-148
// ---------------------------------------------------------------------------
enum MatType { LAMBERTIAN = 0, METAL = 1, SHINY = 2, PERFECTMIRROR = 3, GLASS = 4, DIELECTRIC = 5, DIFFUSELIGHT = 6 };

struct Material {
    int type = LAMBERTIAN;
    V3 color;            // Albedo / Color / Emit
    double roughness = 0, metallic = 0, specular = 0, ior = 1.5;
};

// createMaterial — scene.go:104-148 (+ constructors material.go:65-73,159-167; advanced_materials.go:14-19,117-123)
// has_* flags say whether the JSON carried the key; absent colour -> (1,1,1) (declared extension).
static Material create_material(int type, int has_color, V3 color, int has_rough, double rough,
                                int has_metallic, double metallic, int has_spec, double spec,
                                int has_ior, double ior) {
    Material m;
    V3 c = has_color ? color : V3{1, 1, 1};
    switch (type) {
        case METAL:
            m.type = METAL; m.color = c;
            m.roughness = std::fmin(has_rough ? rough : 0.0, 1.0);
            m.metallic = std::fmin(has_metallic ? metallic : 1.0, 1.0);
            m.specular = std::fmin(has_spec ? spec : 1.0, 1.0);
            m.ior = 1.5;
            break;
        case SHINY:
            m.type = SHINY; m.color = c;
            m.roughness = std::fmin(has_rough ? rough : 0.0, 1.0);
            m.metallic = std::fmin(has_metallic ? metallic : 0.0, 1.0);
            m.specular = std::fmin(has_spec ? spec : 1.0, 1.0);
            m.ior = 1.5;
            break;
        case PERFECTMIRROR:
            m.type = PERFECTMIRROR; m.color = c;
            m.roughness = std::fmin(has_rough ? rough : 0.0, 1.0);
            m.ior = 2.0;
            break;
        case GLASS:
            m.type = GLASS; m.color = c; m.ior = has_ior ? ior : 1.5;
            break;
        case DIELECTRIC:
            m.type = DIELECTRIC; m.ior = has_ior ? ior : 1.5;
            break;
        case DIFFUSELIGHT:
            m.type = DIFFUSELIGHT; m.color = c;
            break;
        default:
            m.type = LAMBERTIAN; m.color = c;
            break;
    }
    return m;
}

static inline V3 mat_emitted(const Material& m) {            // material.go:37,131,207,262,300; adv:48,153
    return m.type == DIFFUSELIGHT ? m.color : V3{};
}
static inline V3 mat_albedo(const Material& m) {             // material.go:41,135,211,266,304; adv:52,157
    if (m.type == DIELECTRIC) return V3{1, 1, 1};
    if (m.type == DIFFUSELIGHT) return V3{};
    return m.color;
}
static inline double mat_metallic(const Material& m) {       // material.go:49,143,219,274,312; adv:60,165
    if (m.type == METAL || m.type == SHINY) return m.metallic;
    if (m.type == PERFECTMIRROR) return 1.0;
    return 0.0;
}

static inline double reflectance(double cosine, double refIdx) {  // material.go:282-286
    double r0 = (1 - refIdx) / (1 + refIdx);
    r0 = r0 * r0;
    return r0 + (1 - r0) * std::pow(1 - cosine, 5);
}
static inline double schlick(double ior, double cosTheta) {       // material.go:115-129,191-205; adv:146-151
    double f0 = std::pow((ior - 1.0) / (ior + 1.0), 2.0);
    return f0 + (1.0 - f0) * std::pow(1.0 - cosTheta, 5);
}

// Material.Scatter — returns false when the material does not scatter.
static bool scatter(const Material& m, const Ray& ray, const HitRecord& hit, Rng& rng, Ray& scattered, V3& atten) {
    switch (m.type) {
        case LAMBERTIAN: {                                           // material.go:26-35
            V3 dir = add(hit.normal, rng.in_unit_sphere(STREAM_SCATTER, 0));
            if (near_zero(dir)) dir = hit.normal;
            dir = normalize(dir);
            scattered = Ray{hit.point, dir};
            atten = m.color;
            return true;
        }
        case METAL: {                                                // material.go:75-113
            V3 reflected = reflect(ray.d, hit.normal);
            if (m.roughness > 0.001) {
                V3 pert = muls(rng.in_unit_sphere(STREAM_SCATTER, 0), m.roughness);
                reflected = normalize(add(reflected, pert));
            }
            V3 albedo = m.color;
            double cosTheta = std::fabs(dot(ray.d, hit.normal));
            double f = schlick(m.ior, cosTheta);
            V3 fresnel{f, f, f};
            double fs = 0.6 + m.metallic * 0.4;
            V3 e{albedo.x * (1.0 - fs) + fresnel.x * fs, albedo.y * (1.0 - fs) + fresnel.y * fs,
                 albedo.z * (1.0 - fs) + fresnel.z * fs};
            e = V3{go_max(0.0, go_min(1.0, e.x)), go_max(0.0, go_min(1.0, e.y)), go_max(0.0, go_min(1.0, e.z))};
            if (m.metallic > 0.8) {
                double mf = 0.4 + m.metallic * 0.5;
                e = V3{e.x * (1.0 - mf) + fresnel.x * mf, e.y * (1.0 - mf) + fresnel.y * mf,
                       e.z * (1.0 - mf) + fresnel.z * mf};
            }
            scattered = Ray{hit.point, reflected};
            atten = e;
            return true;
        }
        case SHINY: {                                                // material.go:169-189
            V3 reflected = reflect(ray.d, hit.normal);
            if (m.roughness > 0) {
                reflected = add(reflected, muls(rng.in_unit_sphere(STREAM_SCATTER, 0), m.roughness));
                reflected = normalize(reflected);
            }
            double cosTheta = std::fabs(dot(ray.d, hit.normal));
            double f = schlick(m.ior, cosTheta);
            double fs = 0.4 + m.specular * 0.4;
            V3 e{go_min(1.0, m.color.x * (1.0 - fs) + f * fs), go_min(1.0, m.color.y * (1.0 - fs) + f * fs),
                 go_min(1.0, m.color.z * (1.0 - fs) + f * fs)};
            scattered = Ray{hit.point, reflected};
            atten = e;
            return true;
        }
        case PERFECTMIRROR: {                                        // advanced_materials.go:125-144
            V3 reflected = reflect(ray.d, hit.normal);
            if (m.roughness > 0.001) {
                V3 pert = muls(rng.in_unit_sphere(STREAM_SCATTER, 0), m.roughness);
                reflected = normalize(add(reflected, pert));
            }
            double cosTheta = std::fabs(dot(ray.d, hit.normal));
            double f = schlick(m.ior, cosTheta);
            V3 e{m.color.x * (1.0 - 0.9) + f * 0.9, m.color.y * (1.0 - 0.9) + f * 0.9, m.color.z * (1.0 - 0.9) + f * 0.9};
            scattered = Ray{hit.point, reflected};
            atten = e;
            return true;
        }
        case GLASS:                                                  // advanced_materials.go:21-46
        case DIELECTRIC: {                                           // material.go:235-260
            atten = (m.type == GLASS) ? m.color : V3{1.0, 1.0, 1.0};
            double ratio = hit.front_face ? 1.0 / m.ior : m.ior;
            V3 unit = normalize(ray.d);
            double cosTheta = go_min(dot(muls(unit, -1), hit.normal), 1.0);
            double sinTheta = std::sqrt(1.0 - cosTheta * cosTheta);
            bool cannotRefract = ratio * sinTheta > 1.0;
            V3 dir;
            // Go evaluates `cannotRefract || reflectance(...) > RandomFloat()` left to right with
            // short-circuit: RandomFloat is NOT drawn when cannotRefract is true.
            if (cannotRefract || reflectance(cosTheta, ratio) > rng.uniform(STREAM_SCATTER, 0, 0)) {
                dir = reflect(unit, hit.normal);
            } else {
                dir = refract(unit, hit.normal, ratio);
            }
            scattered = Ray{hit.point, dir};
            return true;
        }
        case DIFFUSELIGHT:                                           // material.go:296-298
        default:
            return false;
    }
}

// ---------------------------------------------------------------------------
// Geometry — internal/geometry/sphere.go:22-59, triangle.go:13-88
// ---------------------------------------------------------------------------
struct Sphere {
    V3 c;
    double r;
    int mat;
};
struct Triangle {
    V3 v[3];
    V3 n[3];
    int mat;
};
enum PrimKind { PRIM_SPHERE = 0, PRIM_MESH = 1 };
struct Hittable {   // one entry of scene.GetHittables() (scene.go:59-90)
    int kind;
    int first, count;  // sphere: index into spheres; mesh: [first, first+count) into triangles
};

static bool sphere_hit(const Sphere& s, const Ray& ray, double tMin, double tMax, HitRecord& rec) {
    V3 oc = sub(ray.o, s.c);
    double a = length_squared(ray.d);
    double halfB = dot(oc, ray.d);
    double c = length_squared(oc) - s.r * s.r;
    double disc = halfB * halfB - a * c;
    if (disc < 0) return false;
    double sqrtd = std::sqrt(disc);
    double root = (-halfB - sqrtd) / a;
    if (root < tMin || tMax < root) {
        root = (-halfB + sqrtd) / a;
        if (root < tMin || tMax < root) return false;
    }
    double t = root;
    V3 point = ray_at(ray, t);
    V3 outward = divs(sub(point, s.c), s.r);
    bool front = dot(ray.d, outward) < 0;
    V3 normal = outward;
    if (!front) normal = muls(outward, -1);
    rec.t = t; rec.point = point; rec.normal = normal; rec.front_face = front; rec.material = s.mat;
    return true;
}

static V3 calculate_normal(V3 v0, V3 v1, V3 v2) {            // triangle.go:30-34
    return normalize(cross(sub(v1, v0), sub(v2, v0)));
}

static bool triangle_hit(const Triangle& tr, const Ray& ray, double tMin, double tMax, HitRecord& rec) {
    V3 edge1 = sub(tr.v[1], tr.v[0]);
    V3 edge2 = sub(tr.v[2], tr.v[0]);
    V3 h = cross(ray.d, edge2);
    double a = dot(edge1, h);
    if (a > -1e-6 && a < 1e-6) return false;
    double f = 1.0 / a;
    V3 s = sub(ray.o, tr.v[0]);
    double u = f * dot(s, h);
    if (u < 0.0 || u > 1.0) return false;
    V3 q = cross(s, edge1);
    double v = f * dot(ray.d, q);
    if (v < 0.0 || u + v > 1.0) return false;
    double t = f * dot(edge2, q);
    if (t < tMin || t > tMax) return false;
    V3 point = ray_at(ray, t);
    double w = 1.0 - u - v;                                                          // triangle.go:84-88
    V3 normal = normalize(add(add(muls(tr.n[0], w), muls(tr.n[1], u)), muls(tr.n[2], v)));
    bool front = dot(ray.d, normal) < 0;
    if (!front) normal = muls(normal, -1);
    rec.t = t; rec.point = point; rec.normal = normal; rec.front_face = front; rec.material = tr.mat;
    return true;
}

// ---------------------------------------------------------------------------
// Scene — internal/scene/scene.go
// ---------------------------------------------------------------------------
struct Light {          // scene.go:34-39 (Type is never read by the renderer — SURVEY F10)
    V3 pos, color;
    double intensity;
};
struct Camera {         // scene.go:18-24
    V3 pos, look_at, up{0, 1, 0};
    double fov = 60, aspect = 1.0;
};
struct Scene {
    Camera cam;
    std::vector<Material> mats;
    std::vector<Sphere> spheres;
    std::vector<Triangle> tris;
    std::vector<Hittable> hittables;
    std::vector<Light> lights;
    // fog extension: exponential, applied to the primary-hit distance
    int fog_enabled = 0;
    double fog_density = 0;
    V3 fog_color;
    // sky extension: AtmosphereConfig (atmosphere/atmosphere.go:8-26); a ray that leaves the scene returns
    // GetSkyColor(direction) instead of black
    // oracle-side BVH (use_accel), built on first use and kept while the primitive counts do not change
    mutable std::shared_ptr<void> accel_cache;
    mutable size_t accel_cache_spheres = 0, accel_cache_tris = 0;
    int sky_enabled = 0;
    V3 sky_top, sky_bottom, sun_dir, sun_color, rayleigh, mie, sky_fog_color;
    double sun_intensity = 0, sun_size = 0, atm_depth = 0, sky_fog_density = 0, haze = 0, time_of_day = 0;
};

// FastVec3Lerp is called by atmosphere.go but defined nowhere in the reference: the plain a(1-t) + bt
static inline V3 lerp3(V3 a, V3 b, double t) { return add(muls(a, 1.0 - t), muls(b, t)); }

// AtmosphereConfig.GetSkyColor — atmosphere/atmosphere.go:100-135
static V3 sky_color(const Scene& sc, V3 dir) {
    V3 u = normalize(dir);
    double t = 0.5 * (u.y + 1.0);
    V3 sky = lerp3(sc.sky_bottom, sc.sky_top, t);
    double depth = go_max(0.0, u.y);
    double atmospheric = std::exp(-depth * sc.atm_depth);
    V3 scat = lerp3(sc.rayleigh, sc.mie, atmospheric);
    sky = lerp3(sky, scat, 0.25);
    double sunDot = dot(u, sc.sun_dir);
    if (sunDot > (1.0 - sc.sun_size)) {
        double si = std::pow((sunDot - (1.0 - sc.sun_size)) / sc.sun_size, 1.5);
        si = go_min(si, 1.0);
        sky = lerp3(sky, sc.sun_color, si * sc.sun_intensity * 0.9);
    }
    double tf = sc.time_of_day;
    if (tf > 0.5) tf = 1.0 - tf;
    tf *= 2.0;
    double darkness = 1.0 - tf * 0.3;
    sky = muls(sky, darkness);
    if (sc.sky_fog_density > 0.0) {
        double ff = std::exp(-sc.sky_fog_density);
        sky = lerp3(sc.sky_fog_color, sky, ff);
    }
    return clamp(sky, 0.1, 0.98);
}

static void add_mesh_triangle(Scene& s, V3 v0, V3 v1, V3 v2, int mat) {   // NewTriangle triangle.go:13-20
    Triangle t;
    V3 n = calculate_normal(v0, v1, v2);
    t.v[0] = v0; t.v[1] = v1; t.v[2] = v2;
    t.n[0] = n; t.n[1] = n; t.n[2] = n;
    t.mat = mat;
    s.tris.push_back(t);
}

// createCube — scene.go:150-190: 8 corners, 6 quads, 2 triangles each, in this order.
static void add_cube(Scene& s, V3 position, V3 size, int mat) {
    V3 h = divs(size, 2.0);
    V3 vtx[8] = {
        add(position, V3{-h.x, -h.y, -h.z}), add(position, V3{h.x, -h.y, -h.z}),
        add(position, V3{h.x, h.y, -h.z}),   add(position, V3{-h.x, h.y, -h.z}),
        add(position, V3{-h.x, -h.y, h.z}),  add(position, V3{h.x, -h.y, h.z}),
        add(position, V3{h.x, h.y, h.z}),    add(position, V3{-h.x, h.y, h.z}),
    };
    static const int faces[6][4] = {{0, 1, 2, 3}, {1, 5, 6, 2}, {5, 4, 7, 6}, {4, 0, 3, 7}, {3, 2, 6, 7}, {4, 5, 1, 0}};
    Hittable hb{PRIM_MESH, (int)s.tris.size(), 12};
    for (auto& f : faces) {
        add_mesh_triangle(s, vtx[f[0]], vtx[f[1]], vtx[f[2]], mat);
        add_mesh_triangle(s, vtx[f[0]], vtx[f[2]], vtx[f[3]], mat);
    }
    s.hittables.push_back(hb);
}

// triangularPrism extension (README.md:227-240): 6 vertices, 8 triangles:
// caps (0,1,2),(3,5,4); sides (0,1,4,3),(1,2,5,4),(2,0,3,5) each as (a,b,c),(a,c,d).
static void add_prism(Scene& s, const V3 v[6], int mat) {
    Hittable hb{PRIM_MESH, (int)s.tris.size(), 8};
    add_mesh_triangle(s, v[0], v[1], v[2], mat);
    add_mesh_triangle(s, v[3], v[5], v[4], mat);
    static const int quads[3][4] = {{0, 1, 4, 3}, {1, 2, 5, 4}, {2, 0, 3, 5}};
    for (auto& q : quads) {
        add_mesh_triangle(s, v[q[0]], v[q[1]], v[q[2]], mat);
        add_mesh_triangle(s, v[q[0]], v[q[2]], v[q[3]], mat);
    }
    s.hittables.push_back(hb);
}

// ---------------------------------------------------------------------------
// Renderer — internal/renderer/renderer.go
// ---------------------------------------------------------------------------
enum CameraMode { CAMERA_REFERENCE = 0, CAMERA_LOOKAT = 1 };

struct Params {
    int width = 0, height = 0;
    int samples = 100, max_depth = 50;             // renderer.go:57-58
    int jitter = 1;                                // antiAliasing: unused by the reference (F7); 0 => (0.5,0.5)
    int recursive_reflections = 1, soft_shadows = 1;  // renderer.go:60-61
    int camera_mode = CAMERA_REFERENCE;
    int rng_mode = RNG_MT;
    uint64_t seed = 0;
    int threads = 1;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;            // crop (x1/y1 == 0 -> full frame); pixels outside are left untouched
    int use_accel = 0;                             // 0: linear-scan hitWorld (the reference); 1: oracle-side BVH (test speed-up only)
};

struct Accel;  // optional oracle-side BVH (defined below)

struct Tracer {
    const Scene& sc;
    const Params& p;
    const Accel* accel;
    Rng rng;
    Counters cnt;

    Tracer(const Scene& s, const Params& pp, const Accel* a) : sc(s), p(pp), accel(a) {}

    bool hittable_hit(const Hittable& h, const Ray& ray, double tMin, double tMax, HitRecord& rec) {
        if (h.kind == PRIM_SPHERE) {
            cnt.sphere_tests++;
            if (sphere_hit(sc.spheres[h.first], ray, tMin, tMax, rec)) { rec.prim = h.first; return true; }
            return false;
        }
        // Mesh.Hit — scene.go:196-209: nested linear scan, shrinking closestT, last-wins on equal t.
        bool any = false;
        double closest = tMax;
        HitRecord tmp;
        for (int i = 0; i < h.count; i++) {
            cnt.tri_tests++;
            if (triangle_hit(sc.tris[h.first + i], ray, tMin, closest, tmp)) {
                closest = tmp.t;
                rec = tmp;
                rec.prim = (int)sc.spheres.size() + h.first + i;
                any = true;
            }
        }
        return any;
    }

    bool hit_world_accel(const Ray& ray, double tMin, double tMax, HitRecord& rec);

    // hitWorld — renderer.go:333-346
    bool hit_world(const Ray& ray, double tMin, double tMax, HitRecord& rec) {
        cnt.hit_world++;
        if (accel) return hit_world_accel(ray, tMin, tMax, rec);
        bool any = false;
        double closest = tMax;
        HitRecord tmp;
        for (const Hittable& h : sc.hittables) {
            if (hittable_hit(h, ray, tMin, closest, tmp)) {
                closest = tmp.t;
                rec = tmp;
                any = true;
            }
        }
        return any;
    }

    // calculateSmartShadow — renderer.go:299-331
    double smart_shadow(const HitRecord& hit, const Light& light, uint32_t light_index) {
        V3 lightDir = normalize(sub(light.pos, hit.point));
        double lightDistance = length(sub(light.pos, hit.point));
        Ray shadowRay{hit.point, lightDir};
        HitRecord tmp;
        cnt.shadow_rays++;
        if (hit_world(shadowRay, 0.001, lightDistance, tmp)) return 0.0;
        if (p.soft_shadows) {
            const int shadowSamples = 16;
            double sum = 0.0;
            for (int i = 0; i < shadowSamples; i++) {
                V3 off = muls(rng.in_unit_sphere_half(STREAM_SHADOW, (light_index << 12) | ((uint32_t)(i >> 1) << 8), i & 1), 0.1);
                V3 softDir = normalize(add(lightDir, off));
                Ray softRay{hit.point, softDir};
                cnt.shadow_rays++;
                if (!hit_world(softRay, 0.001, lightDistance, tmp)) sum += 1.0;
            }
            return sum / (double)shadowSamples;
        }
        return 1.0;
    }

    // calculateDirectLighting — renderer.go:229-297
    V3 direct_lighting(const HitRecord& hit) {
        V3 total{};
        const Material& m = sc.mats[hit.material];
        V3 albedo = mat_albedo(m);
        double metallic = mat_metallic(m);
        double ambient = 0.1;
        if (metallic > 0.9) ambient = 0.05;
        else if (metallic > 0.7) ambient = 0.07;
        else if (metallic > 0.5) ambient = 0.08;
        total = add(total, V3{ambient, ambient, ambient});
        for (size_t li = 0; li < sc.lights.size(); li++) {
            const Light& light = sc.lights[li];
            V3 lightDir = normalize(sub(light.pos, hit.point));
            double lightDistance = length(sub(light.pos, hit.point));
            if (lightDistance < 0.001) continue;
            double shadow = smart_shadow(hit, light, (uint32_t)li);
            if (shadow > 0.0) {
                double cosTheta = go_max(0, dot(hit.normal, lightDir));
                double intensity = cosTheta * light.intensity / (lightDistance * lightDistance);
                double kd = 0.25;
                if (metallic > 0.95) kd = 0.05;
                else if (metallic > 0.9) kd = 0.08;
                else if (metallic > 0.8) kd = 0.12;
                else if (metallic > 0.7) kd = 0.15;
                else if (metallic > 0.5) kd = 0.2;
                V3 diffuse = muls(albedo, kd * intensity * shadow);
                total = add(total, diffuse);
                if (metallic > 0.5) {
                    V3 viewDir = normalize(muls(hit.point, -1));      // toward the WORLD ORIGIN (F10)
                    V3 halfDir = normalize(add(lightDir, viewDir));
                    double power = 32.0;
                    if (metallic > 0.9) power = 64.0;
                    else if (metallic > 0.8) power = 48.0;
                    double si = std::pow(go_max(0, dot(hit.normal, halfDir)), power);
                    V3 spec = muls(light.color, si * intensity * shadow * metallic * 3.0);
                    total = add(total, spec);
                }
            }
        }
        return total;
    }

    // traceRay — renderer.go:165-227
    V3 trace_ray(const Ray& ray, int depth, double* primary_t) {
        if (depth >= p.max_depth) return V3{};
        HitRecord hit;
        if (!hit_world(ray, 0.001, std::numeric_limits<double>::infinity(), hit))
            return sc.sky_enabled ? sky_color(sc, ray.d) : V3{0.0, 0.0, 0.0};   // renderer.go:171-173 (sky: extension)
        if (primary_t) *primary_t = hit.t * length(ray.d);   // fog extension: world-space distance
        rng.bounce = (uint32_t)depth;
        const Material& m = sc.mats[hit.material];
        V3 emitted = mat_emitted(m);
        V3 direct = direct_lighting(hit);
        Ray scattered;
        V3 atten;
        cnt.scatters++;
        if (!scatter(m, ray, hit, rng, scattered, atten)) return add(emitted, direct);
        V3 reflected{};
        if (p.recursive_reflections) {
            reflected = trace_ray(scattered, depth + 1, nullptr);
            rng.bounce = (uint32_t)depth;
        }
        double metallic = mat_metallic(m);
        double wr, wd;
        if (metallic > 0.95) { wr = 0.85; wd = 0.15; }
        else if (metallic > 0.9) { wr = 0.8; wd = 0.2; }
        else if (metallic > 0.8) { wr = 0.75; wd = 0.25; }
        else if (metallic > 0.7) { wr = 0.7; wd = 0.3; }
        else if (metallic > 0.5) { wr = 0.6; wd = 0.4; }
        else if (metallic > 0.2) { wr = 0.4; wd = 0.6; }
        else return add(add(emitted, direct), mul(atten, reflected));                   // renderer.go:225
        return add(add(emitted, muls(direct, wd)), muls(mul(atten, reflected), wr));    // renderer.go:196 etc.
    }

    // getRay — renderer.go:377-390 (reference camera: ignores lookAt/up/fov, looks down -Z, unnormalised)
    Ray get_ray(double u, double v) {
        const Camera& cam = sc.cam;
        if (p.camera_mode == CAMERA_LOOKAT) {
            // extension: classic look-at pinhole; v is un-flipped so images are upright.
            const double kPi = 3.14159265358979323846;
            double theta = cam.fov * kPi / 180.0;
            double hh = std::tan(theta / 2.0);
            double vh = 2.0 * hh, vw = vh * cam.aspect;
            V3 w = normalize(sub(cam.pos, cam.look_at));
            V3 uu = normalize(cross(cam.up, w));
            V3 vv = cross(w, uu);
            V3 horizontal = muls(uu, vw), vertical = muls(vv, vh);
            V3 ll = sub(sub(sub(cam.pos, divs(horizontal, 2)), divs(vertical, 2)), w);
            V3 dir = sub(add(add(ll, muls(horizontal, u)), muls(vertical, 1.0 - v)), cam.pos);
            return Ray{cam.pos, dir};
        }
        double viewportHeight = 2.0;
        double viewportWidth = viewportHeight * cam.aspect;
        double focalLength = 1.0;
        V3 origin = cam.pos;
        V3 horizontal{viewportWidth, 0, 0};
        V3 vertical{0, viewportHeight, 0};
        V3 ll = sub(sub(sub(origin, divs(horizontal, 2)), divs(vertical, 2)), V3{0, 0, focalLength});
        V3 dir = sub(add(add(ll, muls(horizontal, u)), muls(vertical, v)), origin);
        return Ray{origin, dir};
    }

    // tracePixel — renderer.go:150-163
    V3 trace_pixel(int x, int y) {
        V3 color{};
        int samples = p.samples;
        rng.pixel = (uint32_t)(y * p.width + x);
        for (int s = 0; s < samples; s++) {
            rng.sample = (uint32_t)s;
            rng.bounce = 0;
            double ju = 0.5, jv = 0.5;
            if (p.jitter) {
                if (rng.mode == RNG_MT) { ju = rng.mt_float(); jv = rng.mt_float(); }
                else {
                    // one Philox block serves two consecutive samples: lanes (0,1) the even one, (2,3) the odd one
                    rng.sample = (uint32_t)s >> 1;
                    ju = rng.uniform(STREAM_JITTER, 0, (s & 1) * 2);
                    jv = rng.uniform(STREAM_JITTER, 0, (s & 1) * 2 + 1);
                    rng.sample = (uint32_t)s;
                }
            }
            double u = ((double)x + ju) / (double)p.width;
            double v = ((double)y + jv) / (double)p.height;
            Ray ray = get_ray(u, v);
            cnt.samples++;
            double pt = -1.0;
            V3 c = trace_ray(ray, 0, sc.fog_enabled ? &pt : nullptr);
            if (sc.fog_enabled && pt >= 0.0) {
                // fog extension: f = 1 - exp(-density*d) (effects/atmospheric_effects.go:156-176, exponential),
                // colour lerp toward fog colour; misses stay black.
                double f = 1.0 - std::exp(-sc.fog_density * pt);
                c = add(muls(c, 1.0 - f), muls(sc.fog_color, f));
            }
            color = add(color, c);
        }
        return divs(color, (double)samples);
    }
};

// toneMap — renderer.go:348-367
static V3 tone_map(V3 c) {
    double exposure = 1.0, gamma = 2.2;
    c = muls(c, exposure);
    c.x = 1.0 - std::exp(-c.x); c.y = 1.0 - std::exp(-c.y); c.z = 1.0 - std::exp(-c.z);
    c.x = std::pow(c.x, 1.0 / gamma); c.y = std::pow(c.y, 1.0 / gamma); c.z = std::pow(c.z, 1.0 / gamma);
    c.x = go_max(0.0, go_min(1.0, c.x)); c.y = go_max(0.0, go_min(1.0, c.y)); c.z = go_max(0.0, go_min(1.0, c.z));
    return c;
}

// ---------------------------------------------------------------------------
// Optional oracle-side BVH: a median-split tree over flattened primitives used ONLY to make
// large-scene oracle runs finish (Params.use_accel=1).  It must return exactly what the
// linear scan returns (same t, last-wins on equal t) — tests/test_oracle_accel.py checks that.
// ---------------------------------------------------------------------------
struct Accel {
    struct Node { double lo[3], hi[3]; int left, right, first, count; };
    std::vector<Node> nodes;
    std::vector<int> prims;  // flattened primitive ids: [0,S) spheres, [S, S+T) triangles — reference scan order
    const Scene* sc = nullptr;

    void prim_bounds(int id, double lo[3], double hi[3]) const {
        int S = (int)sc->spheres.size();
        if (id < S) {
            const Sphere& s = sc->spheres[id];
            double r = std::fabs(s.r);
            lo[0] = s.c.x - r; lo[1] = s.c.y - r; lo[2] = s.c.z - r;
            hi[0] = s.c.x + r; hi[1] = s.c.y + r; hi[2] = s.c.z + r;
        } else {
            const Triangle& t = sc->tris[id - S];
            for (int a = 0; a < 3; a++) { lo[a] = 1e300; hi[a] = -1e300; }
            for (int k = 0; k < 3; k++) {
                double c[3] = {t.v[k].x, t.v[k].y, t.v[k].z};
                for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], c[a]); hi[a] = std::max(hi[a], c[a]); }
            }
        }
    }
    int build_rec(int first, int count) {
        Node n;
        for (int a = 0; a < 3; a++) { n.lo[a] = 1e300; n.hi[a] = -1e300; }
        for (int i = first; i < first + count; i++) {
            double lo[3], hi[3];
            prim_bounds(prims[i], lo, hi);
            for (int a = 0; a < 3; a++) { n.lo[a] = std::min(n.lo[a], lo[a]); n.hi[a] = std::max(n.hi[a], hi[a]); }
        }
        n.left = n.right = -1; n.first = first; n.count = count;
        int idx = (int)nodes.size();
        nodes.push_back(n);
        if (count > 4) {
            int axis = 0;
            double ext = -1;
            for (int a = 0; a < 3; a++) if (n.hi[a] - n.lo[a] > ext) { ext = n.hi[a] - n.lo[a]; axis = a; }
            int mid = first + count / 2;
            std::nth_element(prims.begin() + first, prims.begin() + mid, prims.begin() + first + count, [&](int A, int B) {
                double la[3], ha[3], lb[3], hb[3];
                prim_bounds(A, la, ha); prim_bounds(B, lb, hb);
                return la[axis] + ha[axis] < lb[axis] + hb[axis];
            });
            int l = build_rec(first, mid - first);
            int r = build_rec(mid, first + count - mid);
            nodes[idx].left = l; nodes[idx].right = r;
        }
        return idx;
    }
    void build(const Scene& s) {
        sc = &s;
        int n = (int)(s.spheres.size() + s.tris.size());
        prims.resize(n);
        // flattened reference order: hittables in order; a mesh contributes its triangles in order.
        // ids: spheres by their index, triangles offset by S; "order rank" is computed below.
        for (int i = 0; i < n; i++) prims[i] = i;
        nodes.clear();
        if (n > 0) build_rec(0, n);
    }
};

static inline bool slab(const Accel::Node& n, const Ray& r, double tMin, double tMax) {
    double o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    double t0 = tMin, t1 = tMax;
    for (int a = 0; a < 3; a++) {
        // conservative (slightly widened) slab so the tree never drops a hit the scan would find
        double lo = n.lo[a] - 1e-9 * (1 + std::fabs(n.lo[a])), hi = n.hi[a] + 1e-9 * (1 + std::fabs(n.hi[a]));
        if (d[a] == 0) { if (o[a] < lo || o[a] > hi) return false; continue; }
        double inv = 1.0 / d[a];
        double ta = (lo - o[a]) * inv, tb = (hi - o[a]) * inv;
        if (ta > tb) std::swap(ta, tb);
        ta -= 1e-9 * (1 + std::fabs(ta)); tb += 1e-9 * (1 + std::fabs(tb));
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
        if (t0 > t1) return false;
    }
    return true;
}

// Order rank of a flattened primitive in the reference's scan order (hittables order, mesh
// triangles nested): needed for the last-wins tie rule.
static std::vector<int> build_order_rank(const Scene& s) {
    int S = (int)s.spheres.size();
    std::vector<int> rank(S + s.tris.size(), 0);
    int r = 0;
    for (const Hittable& h : s.hittables) {
        if (h.kind == PRIM_SPHERE) rank[h.first] = r++;
        else for (int i = 0; i < h.count; i++) rank[S + h.first + i] = r++;
    }
    return rank;
}

struct AccelFull {
    Accel a;
    std::vector<int> rank;
};

bool Tracer::hit_world_accel(const Ray& ray, double tMin, double tMax, HitRecord& rec) {
    const AccelFull* af = reinterpret_cast<const AccelFull*>(accel);
    const Accel& A = af->a;
    if (A.nodes.empty()) return false;
    int S = (int)sc.spheres.size();
    bool any = false;
    double closest = tMax;
    int best_rank = -1;
    int stack[128];
    int sp = 0;
    stack[sp++] = 0;
    HitRecord tmp;
    while (sp > 0) {
        const Accel::Node& n = A.nodes[stack[--sp]];
        if (!slab(n, ray, tMin, closest)) continue;
        if (n.left < 0) {
            for (int i = n.first; i < n.first + n.count; i++) {
                int id = A.prims[i];
                bool h;
                if (id < S) { cnt.sphere_tests++; h = sphere_hit(sc.spheres[id], ray, tMin, closest, tmp); }
                else { cnt.tri_tests++; h = triangle_hit(sc.tris[id - S], ray, tMin, closest, tmp); }
                if (h) {
                    // linear scan semantics: a later primitive at EQUAL t replaces the earlier one
                    if (tmp.t < closest || !any || af->rank[id] > best_rank) {
                        closest = tmp.t; rec = tmp; rec.prim = id; best_rank = af->rank[id]; any = true;
                    }
                }
            }
        } else {
            stack[sp++] = n.left;
            stack[sp++] = n.right;
        }
    }
    return any;
}

// Render — renderer.go:67-126 with createRenderTasks :398-436 (32x32 tiles, row-major) and
// worker/renderTile :128-148; the collector's toneMap+ToRGB+img.Set :92-97.
static void render(const Scene& sc, const Params& p, uint8_t* rgba, double* radiance, Counters* counters_out) {
    const int W = p.width, H = p.height;
    int X0 = p.x0, Y0 = p.y0, X1 = p.x1 > 0 ? p.x1 : W, Y1 = p.y1 > 0 ? p.y1 : H;
    const int tileSize = 32;
    const int ntx = (W + tileSize - 1) / tileSize, nty = (H + tileSize - 1) / tileSize;
    std::atomic<int> next{0};
    int nthreads = std::max(1, p.threads);
    std::vector<Counters> cnts(nthreads);
    std::shared_ptr<AccelFull> af;
    if (p.use_accel) {
        if (!sc.accel_cache || sc.accel_cache_spheres != sc.spheres.size() || sc.accel_cache_tris != sc.tris.size()) {
            std::shared_ptr<AccelFull> fresh(new AccelFull());
            fresh->a.build(sc);
            fresh->rank = build_order_rank(sc);
            sc.accel_cache = fresh;
            sc.accel_cache_spheres = sc.spheres.size();
            sc.accel_cache_tris = sc.tris.size();
        }
        af = std::static_pointer_cast<AccelFull>(sc.accel_cache);
    }
    auto worker = [&](int tid) {
        Tracer tr(sc, p, af ? reinterpret_cast<const Accel*>(af.get()) : nullptr);
        tr.rng.mode = p.rng_mode;
        tr.rng.key[0] = (uint32_t)p.seed;
        tr.rng.key[1] = (uint32_t)(p.seed >> 32);
        tr.rng.mt.seed(p.seed * 0x9E3779B97F4A7C15ull + (uint64_t)tid + 1);
        for (;;) {
            int t = next.fetch_add(1);
            if (t >= ntx * nty) break;
            int tx = t % ntx, ty = t / ntx;
            int sx = tx * tileSize, sy = ty * tileSize;
            int ex = std::min(sx + tileSize, W), ey = std::min(sy + tileSize, H);
            if (ex <= X0 || sx >= X1 || ey <= Y0 || sy >= Y1) continue;
            for (int y = std::max(sy, Y0); y < std::min(ey, Y1); y++) {
                for (int x = std::max(sx, X0); x < std::min(ex, X1); x++) {
                    V3 c = tr.trace_pixel(x, y);
                    if (radiance) {
                        double* o = radiance + 3 * ((size_t)y * W + x);
                        o[0] = c.x; o[1] = c.y; o[2] = c.z;
                    }
                    if (rgba) {
                        V3 m = tone_map(c);
                        uint8_t* o = rgba + 4 * ((size_t)y * W + x);   // image.RGBA: Pix[y*Stride + 4x], Stride = 4W
                        to_rgb(m, o);
                        o[3] = 255;
                    }
                }
            }
        }
        cnts[tid] = tr.cnt;
    };
    if (nthreads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; i++) th.emplace_back(worker, i);
        for (auto& t : th) t.join();
    }
    if (counters_out) {
        Counters total;
        for (auto& c : cnts) total.add(c);
        *counters_out = total;
    }
}

}  // namespace orc

// ===========================================================================
// C ABI (ctypes) — test harness surface only.
// ===========================================================================
using namespace orc;

extern "C" {

struct orc_material {
    int type;
    int has_color; double color[3];
    int has_roughness; double roughness;
    int has_metallic; double metallic;
    int has_specular; double specular;
    int has_ior; double ior;
};

struct orc_params {
    int width, height, samples, max_depth;
    int jitter, recursive_reflections, soft_shadows;
    int camera_mode, rng_mode;
    unsigned long long seed;
    int threads;
    int x0, y0, x1, y1;
    int use_accel;
};

struct orc_counters {
    unsigned long long samples, hit_world, sphere_tests, tri_tests, shadow_rays, scatters;
};

void* orc_scene_new() { return new Scene(); }
void orc_scene_free(void* s) { delete (Scene*)s; }

void orc_scene_set_camera(void* sp, const double* pos, const double* look_at, const double* up, double fov, double aspect) {
    Scene* s = (Scene*)sp;
    s->cam.pos = V3{pos[0], pos[1], pos[2]};
    s->cam.look_at = V3{look_at[0], look_at[1], look_at[2]};
    s->cam.up = V3{up[0], up[1], up[2]};
    s->cam.fov = fov;
    s->cam.aspect = aspect;
}

static int add_material(Scene* s, const orc_material* m) {
    Material mm = create_material(m->type, m->has_color, V3{m->color[0], m->color[1], m->color[2]}, m->has_roughness,
                                  m->roughness, m->has_metallic, m->metallic, m->has_specular, m->specular, m->has_ior, m->ior);
    s->mats.push_back(mm);
    return (int)s->mats.size() - 1;
}

void orc_scene_add_sphere(void* sp, const double* pos, double radius, const orc_material* m) {
    Scene* s = (Scene*)sp;
    int mi = add_material(s, m);
    s->hittables.push_back(Hittable{PRIM_SPHERE, (int)s->spheres.size(), 1});
    s->spheres.push_back(Sphere{V3{pos[0], pos[1], pos[2]}, radius, mi});   // NewSphere sphere.go:14-20
}
void orc_scene_add_cube(void* sp, const double* pos, const double* size, const orc_material* m) {
    Scene* s = (Scene*)sp;
    int mi = add_material(s, m);
    add_cube(*s, V3{pos[0], pos[1], pos[2]}, V3{size[0], size[1], size[2]}, mi);
}
void orc_scene_add_prism(void* sp, const double* verts18, const orc_material* m) {
    Scene* s = (Scene*)sp;
    int mi = add_material(s, m);
    V3 v[6];
    for (int i = 0; i < 6; i++) v[i] = V3{verts18[3 * i], verts18[3 * i + 1], verts18[3 * i + 2]};
    add_prism(*s, v, mi);
}
// A Mesh hittable from explicit triangles (scene.Mesh, scene.go:192-209): used to mirror flat
// gort_scene_desc inputs (synthetic scenes) where cubes arrive already expanded.
void orc_scene_add_mesh(void* sp, const double* verts9n, int n_tris, const orc_material* m) {
    Scene* s = (Scene*)sp;
    int mi = add_material(s, m);
    Hittable hb{PRIM_MESH, (int)s->tris.size(), n_tris};
    for (int i = 0; i < n_tris; i++) {
        const double* v = verts9n + 9 * i;
        add_mesh_triangle(*s, V3{v[0], v[1], v[2]}, V3{v[3], v[4], v[5]}, V3{v[6], v[7], v[8]}, mi);
    }
    s->hittables.push_back(hb);
}
void orc_scene_add_light(void* sp, const double* pos, const double* color, double intensity) {
    Scene* s = (Scene*)sp;
    s->lights.push_back(Light{V3{pos[0], pos[1], pos[2]}, V3{color[0], color[1], color[2]}, intensity});
}
void orc_scene_set_fog(void* sp, int enabled, double density, const double* color) {
    Scene* s = (Scene*)sp;
    s->fog_enabled = enabled;
    s->fog_density = density;
    s->fog_color = V3{color[0], color[1], color[2]};
}
// params27: the AtmosphereConfig fields in declaration order (atmosphere.go:8-26)
void orc_scene_set_sky(void* sp, int enabled, const double* q) {
    Scene* s = (Scene*)sp;
    s->sky_enabled = enabled;
    s->sky_top = V3{q[0], q[1], q[2]}; s->sky_bottom = V3{q[3], q[4], q[5]};
    s->sun_dir = V3{q[6], q[7], q[8]}; s->sun_color = V3{q[9], q[10], q[11]};
    s->sun_intensity = q[12]; s->sun_size = q[13];
    s->rayleigh = V3{q[14], q[15], q[16]}; s->mie = V3{q[17], q[18], q[19]};
    s->atm_depth = q[20]; s->sky_fog_density = q[21]; s->sky_fog_color = V3{q[22], q[23], q[24]};
    s->haze = q[25]; s->time_of_day = q[26];
}
void orc_sky_color(void* sp, const double* dir, double* out) {
    Scene* s = (Scene*)sp;
    V3 c = sky_color(*s, V3{dir[0], dir[1], dir[2]});
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
int orc_scene_counts(void* sp, int* n_spheres, int* n_tris, int* n_hittables, int* n_lights) {
    Scene* s = (Scene*)sp;
    *n_spheres = (int)s->spheres.size();
    *n_tris = (int)s->tris.size();
    *n_hittables = (int)s->hittables.size();
    *n_lights = (int)s->lights.size();
    return 0;
}
// Flattened triangle i: 9 vertex doubles + 3 normal doubles + material index (for loader cross-checks).
void orc_scene_get_triangle(void* sp, int i, double* out12, int* mat) {
    Scene* s = (Scene*)sp;
    const Triangle& t = s->tris[i];
    for (int k = 0; k < 3; k++) { out12[3 * k] = t.v[k].x; out12[3 * k + 1] = t.v[k].y; out12[3 * k + 2] = t.v[k].z; }
    out12[9] = t.n[0].x; out12[10] = t.n[0].y; out12[11] = t.n[0].z;
    *mat = t.mat;
}
void orc_scene_get_material(void* sp, int i, int* type, double* out7) {
    Scene* s = (Scene*)sp;
    const Material& m = s->mats[i];
    *type = m.type;
    out7[0] = m.color.x; out7[1] = m.color.y; out7[2] = m.color.z;
    out7[3] = m.roughness; out7[4] = m.metallic; out7[5] = m.specular; out7[6] = m.ior;
}

void orc_render(void* sp, const orc_params* pp, unsigned char* rgba, double* radiance, orc_counters* cnt) {
    Scene* s = (Scene*)sp;
    Params p;
    p.width = pp->width; p.height = pp->height; p.samples = pp->samples; p.max_depth = pp->max_depth;
    p.jitter = pp->jitter; p.recursive_reflections = pp->recursive_reflections; p.soft_shadows = pp->soft_shadows;
    p.camera_mode = pp->camera_mode; p.rng_mode = pp->rng_mode; p.seed = pp->seed; p.threads = pp->threads;
    p.x0 = pp->x0; p.y0 = pp->y0; p.x1 = pp->x1; p.y1 = pp->y1; p.use_accel = pp->use_accel;
    Counters c;
    render(*s, p, rgba, radiance, &c);
    if (cnt) {
        cnt->samples = c.samples; cnt->hit_world = c.hit_world; cnt->sphere_tests = c.sphere_tests;
        cnt->tri_tests = c.tri_tests; cnt->shadow_rays = c.shadow_rays; cnt->scatters = c.scatters;
    }
}

// ---- unit-level entry points for known-answer tests ------------------------
static V3 v3(const double* p) { return V3{p[0], p[1], p[2]}; }
static void put(double* o, V3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

void orc_vec_add(const double* a, const double* b, double* o) { put(o, add(v3(a), v3(b))); }
void orc_vec_sub(const double* a, const double* b, double* o) { put(o, sub(v3(a), v3(b))); }
void orc_vec_mul(const double* a, const double* b, double* o) { put(o, mul(v3(a), v3(b))); }
double orc_vec_dot(const double* a, const double* b) { return dot(v3(a), v3(b)); }
void orc_vec_cross(const double* a, const double* b, double* o) { put(o, cross(v3(a), v3(b))); }
double orc_vec_length(const double* a) { return length(v3(a)); }
void orc_vec_normalize(const double* a, double* o) { put(o, normalize(v3(a))); }
void orc_vec_reflect(const double* a, const double* n, double* o) { put(o, reflect(v3(a), v3(n))); }
void orc_vec_refract(const double* a, const double* n, double eta, double* o) { put(o, refract(v3(a), v3(n), eta)); }
void orc_vec_clamp(const double* a, double lo, double hi, double* o) { put(o, clamp(v3(a), lo, hi)); }
void orc_vec_to_rgb(const double* a, unsigned char* rgb) { to_rgb(v3(a), rgb); }
double orc_reflectance(double cosine, double ref_idx) { return reflectance(cosine, ref_idx); }
double orc_schlick(double ior, double cos_theta) { return schlick(ior, cos_theta); }
void orc_tone_map(const double* c, double* o) { put(o, tone_map(v3(c))); }
void orc_tone_map_rgb(const double* c, unsigned char* rgb) { to_rgb(tone_map(v3(c)), rgb); }

// out: t, point(3), normal(3), front_face
int orc_sphere_hit(const double* center, double radius, const double* ro, const double* rd, double tmin, double tmax, double* out8) {
    Sphere s{v3(center), radius, 0};
    HitRecord h;
    if (!sphere_hit(s, Ray{v3(ro), v3(rd)}, tmin, tmax, h)) return 0;
    out8[0] = h.t; put(out8 + 1, h.point); put(out8 + 4, h.normal); out8[7] = h.front_face ? 1 : 0;
    return 1;
}
int orc_triangle_hit(const double* v9, const double* ro, const double* rd, double tmin, double tmax, double* out8) {
    Triangle t;
    for (int k = 0; k < 3; k++) t.v[k] = v3(v9 + 3 * k);
    V3 n = calculate_normal(t.v[0], t.v[1], t.v[2]);
    t.n[0] = t.n[1] = t.n[2] = n;
    t.mat = 0;
    HitRecord h;
    if (!triangle_hit(t, Ray{v3(ro), v3(rd)}, tmin, tmax, h)) return 0;
    out8[0] = h.t; put(out8 + 1, h.point); put(out8 + 4, h.normal); out8[7] = h.front_face ? 1 : 0;
    return 1;
}
// Scatter with an explicit Philox context; out: scattered origin(3), dir(3), attenuation(3). Returns 0 if no scatter.
int orc_scatter(const orc_material* m, const double* ro, const double* rd, const double* point, const double* normal,
                int front_face, unsigned long long seed, unsigned pixel, unsigned sample, unsigned bounce, double* out9) {
    Material mm = create_material(m->type, m->has_color, V3{m->color[0], m->color[1], m->color[2]}, m->has_roughness,
                                  m->roughness, m->has_metallic, m->metallic, m->has_specular, m->specular, m->has_ior, m->ior);
    Rng rng;
    rng.mode = RNG_PHILOX;
    rng.key[0] = (uint32_t)seed; rng.key[1] = (uint32_t)(seed >> 32);
    rng.pixel = pixel; rng.sample = sample; rng.bounce = bounce;
    HitRecord h;
    h.point = v3(point); h.normal = v3(normal); h.front_face = front_face != 0; h.material = 0;
    Ray sc;
    V3 att;
    if (!scatter(mm, Ray{v3(ro), v3(rd)}, h, rng, sc, att)) return 0;
    put(out9, sc.o); put(out9 + 3, sc.d); put(out9 + 6, att);
    return 1;
}
void orc_philox4x32_10(const unsigned* ctr, const unsigned* key, unsigned* out) { philox4x32_10(ctr, key, out); }
void orc_in_unit_sphere(unsigned long long seed, unsigned pixel, unsigned sample, unsigned bounce, unsigned stream,
                        unsigned seq_base, double* out3) {
    Rng rng;
    rng.mode = RNG_PHILOX;
    rng.key[0] = (uint32_t)seed; rng.key[1] = (uint32_t)(seed >> 32);
    rng.pixel = pixel; rng.sample = sample; rng.bounce = bounce;
    put(out3, rng.in_unit_sphere(stream, seq_base));
}
void orc_in_unit_sphere_half(unsigned long long seed, unsigned pixel, unsigned sample, unsigned bounce, unsigned stream,
                             unsigned seq_base, int half, double* out3) {
    Rng rng;
    rng.mode = RNG_PHILOX;
    rng.key[0] = (uint32_t)seed; rng.key[1] = (uint32_t)(seed >> 32);
    rng.pixel = pixel; rng.sample = sample; rng.bounce = bounce;
    put(out3, rng.in_unit_sphere_half(stream, seq_base, half));
}
void orc_get_ray(void* sp, int camera_mode, double u, double v, double* out6) {
    Scene* s = (Scene*)sp;
    Params p;
    p.camera_mode = camera_mode;
    Tracer tr(*s, p, nullptr);
    Ray r = tr.get_ray(u, v);
    put(out6, r.o); put(out6 + 3, r.d);
}
// hitWorld on the scene (linear scan or accel): out: t, point(3), normal(3), front_face, material, prim
int orc_hit_world(void* sp, const double* ro, const double* rd, double tmin, double tmax, int use_accel, double* out10) {
    Scene* s = (Scene*)sp;
    Params p;
    std::unique_ptr<AccelFull> af;
    if (use_accel) {
        af.reset(new AccelFull());
        af->a.build(*s);
        af->rank = build_order_rank(*s);
    }
    Tracer tr(*s, p, af ? reinterpret_cast<const Accel*>(af.get()) : nullptr);
    HitRecord h;
    if (!tr.hit_world(Ray{v3(ro), v3(rd)}, tmin, tmax, h)) return 0;
    out10[0] = h.t; put(out10 + 1, h.point); put(out10 + 4, h.normal); out10[7] = h.front_face ? 1 : 0;
    out10[8] = h.material; out10[9] = h.prim;
    return 1;
}

}  // extern "C"
