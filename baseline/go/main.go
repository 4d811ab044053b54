// gort_baseline — times the UNMODIFIED reference renderer (raytraceGo/internal/{scene,renderer}) on the configurations
// bench.py measures, so that anyone with a Go toolchain can produce the true Go number next to the C++ port's
// (SURVEY 8d; this image has no Go: the file is shipped uncompiled).
//
// The reference's internal packages can only be imported from inside its module, so this file is dropped INTO a checkout
// of JoshElkind/concurrent-raytracer-go (baseline/go/run.sh does it):
//
//	cp baseline/go/main.go <checkout>/cmd/gort_baseline/main.go
//	cd <checkout> && go run ./cmd/gort_baseline -config c1_view [-workers N] [-frames K]
//
// Nothing in the reference's packages is touched.  What differs from `cmd/raytracer` is applied to the scene JSON before
// scene.LoadFromFile reads it, exactly as bench.py / tests/common.py do for the GPU arm:
//   - camera.position.z is mirrored (c1_view: -8 -> +8, c2_view: -25 -> +25): with the committed camera the geometry is
//     behind the viewer and every pixel is black (SURVEY F4);
//   - a material without "color" gets [1,1,1]: createMaterial's unchecked type assertion panics on it (scene.go:113, SURVEY F5).
// c1_ref / c2_ref keep the committed camera.  The reference's RNG is wall-clock seeded: frames are not reproducible.
package main

import (
	"encoding/json"
	"flag"
	"fmt"
	"os"
	"path/filepath"
	"runtime"
	"time"

	"raytraceGo/internal/renderer"
	"raytraceGo/internal/scene"
)

type config struct {
	file                      string
	width, height, spp, depth int
	mirrorZ, soft             bool
}

var configs = map[string]config{
	"c1_view": {"demo-assets/sphere_reflections_light.json", 800, 600, 100, 50, true, true},
	"c1_ref":  {"demo-assets/sphere_reflections_light.json", 800, 600, 100, 50, false, true},
	"c2_view": {"demo-assets/final_silver_prism_purple_cube_.json", 1200, 900, 100, 50, true, true},
	"c2_ref":  {"demo-assets/final_silver_prism_purple_cube_.json", 1200, 900, 100, 50, false, true},
}

func patched(path string, mirrorZ bool) (string, error) {
	raw, err := os.ReadFile(path)
	if err != nil {
		return "", err
	}
	var doc map[string]interface{}
	if err := json.Unmarshal(raw, &doc); err != nil {
		return "", err
	}
	if cam, ok := doc["camera"].(map[string]interface{}); ok && mirrorZ {
		if pos, ok := cam["position"].([]interface{}); ok && len(pos) == 3 {
			if z, ok := pos[2].(float64); ok {
				pos[2] = -z
			}
		}
	}
	if objs, ok := doc["objects"].([]interface{}); ok {
		for _, o := range objs {
			if om, ok := o.(map[string]interface{}); ok {
				if mat, ok := om["material"].(map[string]interface{}); ok {
					if _, has := mat["color"]; !has {
						mat["color"] = []interface{}{1.0, 1.0, 1.0}
					}
				}
			}
		}
	}
	out, err := json.Marshal(doc)
	if err != nil {
		return "", err
	}
	tmp := filepath.Join(os.TempDir(), "gort_baseline_scene.json")
	return tmp, os.WriteFile(tmp, out, 0o644)
}

func main() {
	name := flag.String("config", "c1_view", "c1_view | c1_ref | c2_view | c2_ref")
	workers := flag.Int("workers", runtime.NumCPU(), "worker goroutines (cmd/raytracer uses runtime.NumCPU())")
	frames := flag.Int("frames", 3, "timed frames (after one warm-up frame)")
	flag.Parse()
	c, ok := configs[*name]
	if !ok {
		fmt.Fprintln(os.Stderr, "unknown config", *name)
		os.Exit(2)
	}
	path, err := patched(c.file, c.mirrorZ)
	if err != nil {
		fmt.Fprintln(os.Stderr, "scene:", err)
		os.Exit(1)
	}
	sc, err := scene.LoadFromFile(path)
	if err != nil {
		fmt.Fprintln(os.Stderr, "load:", err)
		os.Exit(1)
	}
	r := renderer.NewParallelRenderer(*workers) // defaults: 100 spp, depth 50, soft shadows, reflections (renderer.go:54-65)
	r.SetSamples(c.spp)
	r.SetMaxDepth(c.depth)
	r.SetSoftShadows(c.soft)
	r.Render(sc, c.width, c.height) // warm-up
	start := time.Now()
	for i := 0; i < *frames; i++ {
		r.Render(sc, c.width, c.height)
	}
	sec := time.Since(start).Seconds() / float64(*frames)
	rays := float64(c.width) * float64(c.height) * float64(c.spp)
	line := map[string]interface{}{
		"impl": "reference-go", "config": *name, "width": c.width, "height": c.height, "samples": c.spp, "max_depth": c.depth,
		"render_time_seconds": sec, "rays_per_second": rays / sec, "pixels_per_second": float64(c.width*c.height) / sec,
		"worker_count": *workers, "gomaxprocs": runtime.GOMAXPROCS(0), "num_cpu": runtime.NumCPU(), "go": runtime.Version(),
	}
	out, _ := json.Marshal(line)
	fmt.Println(string(out))
}
