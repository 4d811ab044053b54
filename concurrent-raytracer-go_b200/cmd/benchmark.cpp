// benchmark — working mirror of the reference's cmd/benchmark (/root/reference cmd/benchmark/main.go) over the C ABI
// of libgort.so.  The reference parses its flags (main.go:290-301), ignores the list flags (parseIntSlice returns a
// constant, :330-332; parseStringSlice does not compile, :334-336) and sleeps instead of rendering (:114-127).  This
// program keeps the flag surface and the report schema (BenchmarkResult tags :33-46, report keys :166-172) and
// renders for real: "workers" = number of GPUs, each combination renders frames for at least -duration.
//
//   benchmark -width 800 -height 600 -workers 1,2,4,8 -samples 10,50,100 -max-depth 10,25,50
//             -scenes default[,file.json...] -duration 5s -output benchmark_results.json
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <ctime>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gort.h"
#include "cli_common.h"

static const char* kDefaultScene =  // demo-assets/sphere_reflections_light.json with the camera in front of the spheres
    "{\"camera\":{\"position\":[0,0,8],\"lookAt\":[0,0,0],\"up\":[0,1,0],\"fov\":60,\"aspectRatio\":1.33},\"objects\":["
    "{\"type\":\"sphere\",\"position\":[0,0,0],\"radius\":1.0,\"material\":{\"type\":\"metal\",\"color\":[0.8,0.8,0.9],\"roughness\":0.1}},"
    "{\"type\":\"sphere\",\"position\":[2,0,0],\"radius\":0.5,\"material\":{\"type\":\"metal\",\"refractionIndex\":1.5}},"
    "{\"type\":\"sphere\",\"position\":[-2,0,0],\"radius\":0.7,\"material\":{\"type\":\"glass\",\"color\":[0.8,0.2,0.2]}},"
    "{\"type\":\"sphere\",\"position\":[0,2,0],\"radius\":0.3,\"material\":{\"type\":\"metal\",\"color\":[0.9,0.9,0.1],\"roughness\":0.3}},"
    "{\"type\":\"sphere\",\"position\":[0,-2,0],\"radius\":0.4,\"material\":{\"type\":\"glass\",\"color\":[0.2,0.8,0.2]}}],"
    "\"lights\":[{\"type\":\"point\",\"position\":[5,5,5],\"color\":[1,1,1],\"intensity\":1.0},"
    "{\"type\":\"point\",\"position\":[-3,3,3],\"color\":[0.8,0.8,1],\"intensity\":0.5}]}";

static std::vector<std::string> split(const std::string& s) {
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ','))
        if (!item.empty()) out.push_back(item);
    return out;
}
static std::vector<int> split_int(const std::string& s) {  // what parseIntSlice was meant to do (main.go:330-332)
    std::vector<int> out;
    for (const std::string& t : split(s)) out.push_back(atoi(t.c_str()));
    return out;
}
// Go's flag.Duration syntax: 5s, 250ms, 1m30s, 1.5s
static double parse_duration(const std::string& s) {
    double total = 0;
    size_t i = 0;
    while (i < s.size()) {
        size_t j = i;
        while (j < s.size() && (isdigit((unsigned char)s[j]) || s[j] == '.')) j++;
        const double v = atof(s.substr(i, j - i).c_str());
        size_t k = j;
        while (k < s.size() && !isdigit((unsigned char)s[k]) && s[k] != '.') k++;
        const std::string u = s.substr(j, k - j);
        if (u == "h") total += v * 3600;
        else if (u == "m") total += v * 60;
        else if (u == "s" || u.empty()) total += v;
        else if (u == "ms") total += v * 1e-3;
        else if (u == "us" || u == "\xC2\xB5s") total += v * 1e-6;
        else if (u == "ns") total += v * 1e-9;
        i = k;
    }
    return total;
}

struct Result {
    int workers, samples, max_depth;
    std::string scene;
    double seconds;  // mean render time of one frame
    int frames;
    double rays_per_second, pixels_per_second, speedup, efficiency;
    unsigned long long memory;
};

int main(int argc, char** argv) {
    int width = 800, height = 600;
    std::string workers_s = "1,2,4,8", samples_s = "10,50,100", depth_s = "10,25,50", scenes_s = "default", output = "benchmark_results.json";
    double duration = 5.0;
    bool profile = false, metrics = true;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        while (!a.empty() && a[0] == '-') a.erase(0, 1);
        std::string v;
        const size_t eq = a.find('=');
        bool has_v = false;
        if (eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); has_v = true; }
        auto need = [&]() -> std::string {
            if (has_v) return v;
            if (i + 1 >= argc) { fprintf(stderr, "flag needs an argument: -%s\n", a.c_str()); exit(2); }
            return argv[++i];
        };
        if (a == "width") width = atoi(need().c_str());
        else if (a == "height") height = atoi(need().c_str());
        else if (a == "workers") workers_s = need();
        else if (a == "samples") samples_s = need();
        else if (a == "max-depth") depth_s = need();
        else if (a == "scenes") scenes_s = need();
        else if (a == "duration") duration = parse_duration(need());
        else if (a == "profile") profile = has_v ? (v != "false") : true;
        else if (a == "metrics") metrics = has_v ? (v != "false") : true;
        else if (a == "output") output = need();
        else { fprintf(stderr, "flag provided but not defined: -%s\n", a.c_str()); return 2; }
    }
    const std::vector<int> workers = split_int(workers_s), samples = split_int(samples_s), depths = split_int(depth_s);
    const std::vector<std::string> scenes = split(scenes_s);

    printf("Starting comprehensive benchmark suite...\n");  // main.go:61-63
    printf("Configuration: %dx%d image, %zu worker configurations\n", width, height, workers.size());
    const int avail = gort_device_count();
    if (avail <= 0) {
        fprintf(stderr, "Benchmark failed: no CUDA device (libgort has no CPU fallback): %s\n", gort_last_error(nullptr));
        return 1;
    }
    std::vector<Result> results;
    std::vector<uint8_t> pix((size_t)width * height * 4);
    for (int w : workers) {
        if (w > avail) {
            printf("Skipping %d workers: only %d GPU(s) visible\n", w, avail);
            continue;
        }
        gort_ctx* ctx = nullptr;
        if (gort_create(nullptr, w, &ctx) != GORT_OK) {
            fprintf(stderr, "Benchmark failed: %s\n", gort_last_error(nullptr));
            return 1;
        }
        for (int spp : samples)
            for (int depth : depths)
                for (const std::string& scene : scenes) {
                    int rc = scene == "default" ? gort_scene_load_json(ctx, kDefaultScene, strlen(kDefaultScene), 0) : gort_scene_load_file(ctx, scene.c_str(), 0);
                    if (rc != GORT_OK) {
                        fprintf(stderr, "Benchmark failed: scene %s: %s\n", scene.c_str(), gort_last_error(ctx));
                        gort_destroy(ctx);
                        return 1;
                    }
                    gort_render_params p;
                    memset(&p, 0, sizeof(p));
                    p.abi_version = GORT_ABI_VERSION;
                    p.width = width; p.height = height; p.samples = spp; p.max_depth = depth;
                    p.anti_aliasing = 1; p.recursive_reflections = 1; p.soft_shadows = 1;
                    p.camera_mode = GORT_CAMERA_REFERENCE; p.shard_count = 1; p.seed = 20240601;
                    gort_stats st;
                    if (gort_render(ctx, &p, pix.data(), pix.size(), &st) != GORT_OK) {  // warm-up frame
                        fprintf(stderr, "Benchmark failed: %s\n", gort_last_error(ctx));
                        gort_destroy(ctx);
                        return 1;
                    }
                    const auto t0 = std::chrono::steady_clock::now();
                    int frames = 0;
                    double el = 0;
                    do {
                        p.seed++;
                        if (gort_render(ctx, &p, pix.data(), pix.size(), &st) != GORT_OK) {
                            fprintf(stderr, "Benchmark failed: %s\n", gort_last_error(ctx));
                            gort_destroy(ctx);
                            return 1;
                        }
                        frames++;
                        el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    } while (el < duration);
                    Result r;
                    r.workers = w; r.samples = spp; r.max_depth = depth; r.scene = scene;
                    r.seconds = el / frames; r.frames = frames;
                    r.pixels_per_second = (double)width * height / r.seconds;  // main.go:125-127
                    r.rays_per_second = r.pixels_per_second * spp;
                    r.memory = st.bvh_bytes + (unsigned long long)width * height * 28;
                    r.speedup = 1.0; r.efficiency = 100.0;
                    for (const Result& b : results)  // measured baseline: the 1-worker run of the same combination (the reference assumes linear, :129-132)
                        if (b.workers == 1 && b.samples == spp && b.max_depth == depth && b.scene == scene) {
                            r.speedup = b.seconds / r.seconds;
                            r.efficiency = r.speedup / w * 100.0;
                        }
                    results.push_back(r);
                    printf("Completed: %d workers, %d samples, %d depth, %s\n", w, spp, depth, scene.c_str());
                }
        gort_destroy(ctx);
    }

    double best_speedup = 0, best_eff = 0, fastest = results.empty() ? 0 : results[0].seconds, avg = 0;
    for (const Result& r : results) {
        best_speedup = std::max(best_speedup, r.speedup);
        best_eff = std::max(best_eff, r.efficiency);
        fastest = std::min(fastest, r.seconds);
        avg += r.seconds;
    }
    if (!results.empty()) avg /= results.size();

    auto list_int = [](const std::vector<int>& v) {
        std::string s = "[";
        for (size_t i = 0; i < v.size(); i++) s += (i ? ", " : "") + std::to_string(v[i]);
        return s + "]";
    };
    std::string cfg = "{\"width\": " + std::to_string(width) + ", \"height\": " + std::to_string(height) + ", \"workers\": " + list_int(workers) +
                      ", \"samples\": " + list_int(samples) + ", \"max_depth\": " + list_int(depths) + ", \"scenes\": [";
    for (size_t i = 0; i < scenes.size(); i++) cfg += (i ? ", " : "") + cli::json_str(scenes[i]);
    cfg += "], \"duration\": " + std::to_string((long long)(duration * 1e9)) + ", \"enable_profiling\": " + (profile ? "true" : "false") +
           ", \"enable_metrics\": " + (metrics ? "true" : "false") + ", \"output_file\": " + cli::json_str(output) + "}";
    if (!output.empty()) {
        FILE* f = fopen(output.c_str(), "w");
        if (!f) {
            fprintf(stderr, "Benchmark failed: failed to create output file: %s\n", output.c_str());
            return 1;
        }
        char ts[64];
        time_t now = time(nullptr);
        strftime(ts, sizeof(ts), "%Y-%m-%dT%H:%M:%S%z", localtime(&now));
        fprintf(f, "{\n  \"config\": %s,\n  \"results\": [\n", cfg.c_str());
        for (size_t i = 0; i < results.size(); i++) {
            const Result& r = results[i];
            // time.Duration marshals as integer nanoseconds
            fprintf(f, "    {\"config\": %s, \"worker_count\": %d, \"samples\": %d, \"max_depth\": %d, \"scene\": %s, \"duration\": %lld, "
                       "\"rays_per_second\": %.6g, \"pixels_per_second\": %.6g, \"memory_usage\": %llu, \"cpu_usage\": 0, \"speedup\": %.6g, \"efficiency\": %.6g, "
                       "\"frames\": %d}%s\n",
                    cfg.c_str(), r.workers, r.samples, r.max_depth, cli::json_str(r.scene).c_str(), (long long)(r.seconds * 1e9), r.rays_per_second,
                    r.pixels_per_second, r.memory, r.speedup, r.efficiency, r.frames, i + 1 < results.size() ? "," : "");
        }
        fprintf(f, "  ],\n  \"summary\": {\"average_time\": %lld, \"best_efficiency\": %.6g, \"best_speedup\": %.6g, \"fastest_time\": %lld, \"total_benchmarks\": %zu},\n",
                (long long)(avg * 1e9), best_eff, best_speedup, (long long)(fastest * 1e9), results.size());
        fprintf(f, "  \"system_info\": {\"cpu_count\": %u, \"gpu_count\": %d, \"backend\": \"libgort sm_100a\", \"go_version\": \"n/a (C++ host mirror)\"},\n",
                std::thread::hardware_concurrency(), avail);
        fprintf(f, "  \"timestamp\": %s\n}\n", cli::json_str(ts).c_str());
        fclose(f);
        printf("Benchmark report written to: %s\n", output.c_str());
    }
    // printSummary, main.go:250-288
    printf("\n============================================================\nBENCHMARK SUMMARY\n============================================================\n");
    printf("Total benchmarks run: %zu\n", results.size());
    if (!results.empty()) {
        printf("Best speedup: %.2fx\n", best_speedup);
        printf("Best efficiency: %.1f%%\n", best_eff);
        printf("Average time: %s\n", cli::go_duration(avg).c_str());
    }
    printf("\nDetailed results:\n%-10s %-10s %-10s %-15s %-15s %-15s\n", "Workers", "Samples", "Depth", "Time", "Speedup", "Efficiency");
    printf("---------------------------------------------------------------------------\n");
    for (const Result& r : results)
        printf("%-10d %-10d %-10d %-15s %-15.2f %-15.1f%%\n", r.workers, r.samples, r.max_depth, cli::go_duration(r.seconds).c_str(), r.speedup, r.efficiency);
    return 0;
}
