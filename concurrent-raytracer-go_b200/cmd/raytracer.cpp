// raytracer — mirror of the reference's cmd/raytracer (/root/reference cmd/raytracer/main.go:14-69) over the
// C ABI of libgort.so: same positional arguments, same console messages, same outputs (PNG + benchmark_data.json
// next to it with the fields of renderer.go:31-42).  The reference's CLI is Go; there is no Go toolchain in this
// image, so the host program above the ABI is C++ (INTEGRATION.md has the Go/cgo version).
//
//   raytracer <scene_file> <output_file> <width> <height>
// Additive flags (before the positionals, like Go's flag package): -gpus N, -samples N, -max-depth N, -seed N,
// -camera-mode reference|lookat, -prisms, -fog, -scene-settings (use the scene's own "renderer" block),
// -readme-json FILE (README.md:50-71 schema).
#include <sys/stat.h>

#include <chrono>
#include <cstdlib>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/gort.h"
#include "cli_common.h"

static bool has_ext(const std::string& p) {
    size_t slash = p.find_last_of('/'), dot = p.find_last_of('.');
    return dot != std::string::npos && (slash == std::string::npos || dot > slash);
}

int main(int argc, char** argv) {
    int gpus = 0, samples = 100, max_depth = 50, camera_mode = GORT_CAMERA_REFERENCE;
    unsigned long long seed = (unsigned long long)std::chrono::system_clock::now().time_since_epoch().count();  // time-seeded like random.go:8-10
    uint32_t options = 0;
    bool scene_settings = false, samples_given = false, depth_given = false;
    int aa = 1, rr = 1, ss = 1;  // NewParallelRenderer defaults, renderer.go:54-65
    std::string readme_json;
    std::vector<std::string> args;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { fprintf(stderr, "flag needs an argument: %s\n", name); exit(2); }
            return argv[++i];
        };
        if (a == "-gpus" || a == "--gpus") gpus = atoi(val("-gpus"));
        else if (a == "-samples" || a == "--samples") { samples = atoi(val("-samples")); samples_given = true; }
        else if (a == "-max-depth" || a == "--max-depth") { max_depth = atoi(val("-max-depth")); depth_given = true; }
        else if (a == "-scene-settings" || a == "--scene-settings") scene_settings = true;
        else if (a == "-seed" || a == "--seed") seed = strtoull(val("-seed"), nullptr, 10);
        else if (a == "-camera-mode" || a == "--camera-mode") camera_mode = std::string(val("-camera-mode")) == "lookat" ? GORT_CAMERA_LOOKAT : GORT_CAMERA_REFERENCE;
        else if (a == "-prisms" || a == "--prisms") options |= 1u;
        else if (a == "-fog" || a == "--fog") options |= 2u;
        else if (a == "-sky" || a == "--sky") options |= 4u;
        else if (a == "-readme-json" || a == "--readme-json") readme_json = val("-readme-json");
        else args.push_back(a);
    }
    if (args.size() < 4) {
        printf("Usage: raytracer <scene_file> <output_file> <width> <height>\n");
        printf("Example: raytracer scene.json output.png 800 600\n");
        return 1;
    }
    const std::string scene_file = args[0];
    std::string output = args[1];
    char* end = nullptr;
    const long width = strtol(args[2].c_str(), &end, 10);
    if (*end || args[2].empty()) { printf("Invalid width: %s\n", args[2].c_str()); return 1; }
    const long height = strtol(args[3].c_str(), &end, 10);
    if (*end || args[3].empty()) { printf("Invalid height: %s\n", args[3].c_str()); return 1; }

    printf("Loading scene from: %s\n", scene_file.c_str());
    const int avail = gort_device_count();
    if (gpus <= 0) gpus = avail > 0 ? 1 : 0;
    gort_ctx* ctx = nullptr;
    if (gort_create(nullptr, gpus > 0 ? gpus : 1, &ctx) != GORT_OK) {
        printf("Error creating renderer: %s\n", gort_last_error(nullptr));
        return 1;
    }
    const auto t_setup0 = std::chrono::steady_clock::now();
    if (gort_scene_load_file(ctx, scene_file.c_str(), options) != GORT_OK) {
        printf("Error loading scene: %s\n", gort_last_error(ctx));
        gort_destroy(ctx);
        return 1;
    }
    const double setup_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_setup0).count();
    int32_t n_sph = 0, n_tri = 0, n_mat = 0, n_light = 0, n_hit = 0;
    gort_scene_counts(ctx, &n_sph, &n_tri, &n_mat, &n_light, &n_hit);
    printf("Created %d hittables total\n", n_hit);  // scene.go:88

    if (scene_settings) {  // extension: the scene's "renderer" block (README.md:285-291); explicit flags win
        int32_t h[5];
        if (gort_scene_render_hints(ctx, h) == GORT_OK) {
            if (h[0] > 0 && !samples_given) samples = h[0];
            if (h[1] >= 0 && !depth_given) max_depth = h[1];
            if (h[2] >= 0) aa = h[2];
            if (h[3] >= 0) rr = h[3];
            if (h[4] >= 0) ss = h[4];
        }
    }
    printf("Rendering at %ldx%ld resolution...\n", width, height);
    gort_render_params p;
    memset(&p, 0, sizeof(p));
    p.abi_version = GORT_ABI_VERSION;
    p.width = (int32_t)width; p.height = (int32_t)height;
    p.samples = samples; p.max_depth = max_depth;
    p.anti_aliasing = aa; p.recursive_reflections = rr; p.soft_shadows = ss;
    p.camera_mode = camera_mode; p.shard_rank = 0; p.shard_count = 1; p.seed = seed;
    std::vector<uint8_t> pix((size_t)width * height * 4);
    gort_stats st;
    const auto t0 = std::chrono::steady_clock::now();
    if (gort_render(ctx, &p, pix.data(), pix.size(), &st) != GORT_OK) {
        printf("Error rendering: %s\n", gort_last_error(ctx));
        gort_destroy(ctx);
        return 1;
    }
    const double render_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // renderer.go:119-123
    printf("Render completed in %s\n", cli::go_duration(render_s).c_str());
    printf("Resolution: %ldx%ld\n", width, height);
    printf("Samples per pixel: %d\n", samples);
    printf("Max depth: %d\n", max_depth);
    printf("Workers: %d\n", st.n_devices);

    if (!has_ext(output)) output += ".png";
    printf("Saving to: %s\n", output.c_str());
    const std::string dir = cli::dir_of(output);
    cli::mkdir_all(dir);  // os.MkdirAll (renderer.go:439-441)
    if (!cli::write_png(output, pix.data(), (int)width, (int)height)) {
        printf("Error saving image: cannot write %s\n", output.c_str());
        gort_destroy(ctx);
        return 1;
    }

    char ts[64];
    time_t now = time(nullptr);
    strftime(ts, sizeof(ts), "%Y-%m-%dT%H:%M:%S%z", localtime(&now));
    {   // BenchmarkData, renderer.go:31-42,103-117 (json.MarshalIndent, two spaces)
        const std::string bp = dir + "/benchmark_data.json";
        FILE* f = fopen(bp.c_str(), "w");
        if (!f) {
            printf("Error saving benchmark data: cannot write %s\n", bp.c_str());
        } else {
            fprintf(f, "{\n  \"scene_name\": \"demo_scene\",\n  \"resolution\": \"%ldx%ld\",\n  \"render_time_seconds\": %.9g,\n", width, height, render_s);
            fprintf(f, "  \"samples\": %d,\n  \"max_depth\": %d,\n  \"num_workers\": %d,\n  \"objects\": %d,\n  \"lights\": %d,\n", samples, max_depth, st.n_devices, n_hit, n_light);
            fprintf(f, "  \"timestamp\": %s,\n  \"features\": [\n", cli::json_str(ts).c_str());
            fprintf(f, "    \"Improved metallic reflections with Fresnel effect\",\n    \"Shiny materials with configurable roughness and specular\",\n");
            fprintf(f, "    \"Enhanced light source reflections\",\n    \"Better specular highlights for metallic surfaces\"\n  ]\n}");
            fclose(f);
            printf("Benchmark data saved\n");
        }
    }
    if (!readme_json.empty()) {  // README.md:50-71 (keys sorted as a Go map marshals)
        FILE* f = fopen(readme_json.c_str(), "w");
        if (f) {
            const double px = (double)width * height;
            // the switches as they were rendered (-scene-settings may have turned them off); atmosphere: what the loader enabled
            const char* atmosphere = (options & 2u) ? ((options & 4u) ? "fog+sky" : "fog") : ((options & 4u) ? "sky" : "none");
            fprintf(f, "{\n  \"anti_aliasing\": %s,\n  \"atmosphere\": \"%s\",\n  \"bvh_build_time\": %s,\n  \"cpu_usage\": 0,\n  \"depth_of_field\": false,\n",
                    aa ? "true" : "false", atmosphere, cli::json_str(cli::go_duration(st.bvh_build_ms * 1e-3)).c_str());
            fprintf(f, "  \"height\": %ld,\n  \"max_depth\": %d,\n  \"memory_usage\": %llu,\n  \"output_file\": %s,\n", height, max_depth,
                    (unsigned long long)(st.bvh_bytes + (unsigned long long)px * 28), cli::json_str(output).c_str());
            fprintf(f, "  \"pixels_per_second\": %.0f,\n  \"rays_per_second\": %.0f,\n  \"recursive_reflections\": %s,\n  \"render_time\": %s,\n", px / render_s,
                    px * samples / render_s, rr ? "true" : "false", cli::json_str(cli::go_duration(render_s)).c_str());
            fprintf(f, "  \"samples\": %d,\n  \"scene_file\": %s,\n  \"setup_time\": %s,\n  \"soft_shadows\": %s,\n  \"total_time\": %s,\n", samples,
                    cli::json_str(scene_file).c_str(), cli::json_str(cli::go_duration(setup_s)).c_str(), ss ? "true" : "false",
                    cli::json_str(cli::go_duration(setup_s + render_s)).c_str());
            fprintf(f, "  \"width\": %ld,\n  \"worker_count\": %d\n}\n", width, st.n_devices);
            fclose(f);
        }
    }
    gort_destroy(ctx);
    return 0;
}
