// render_gpu.go — REPLACEMENT for the body of ParallelRenderer.Render (internal/renderer/renderer.go:67-126).
// Delete the goroutine tile pool there (createRenderTasks, worker, renderTile, tracePixel, traceRay, hitWorld,
// calculateDirectLighting, calculateSmartShadow, toneMap, getRay: renderer.go:128-436) and add this file; the
// struct, the setters (settings.go), SaveImage, SaveBenchmarkData and GetStats stay as they are.
// Add one field to ParallelRenderer:  gpu *gpurender.Context
//
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain in the build image); the Python mirror
// concurrent-raytracer-go_b200/__init__.py:ParallelRenderer.Render makes the same three calls over the same C ABI.
package renderer

import (
	"fmt"
	"image"
	"time"

	"raytraceGo/internal/gpurender"
	"raytraceGo/internal/scene"
)

func (r *ParallelRenderer) Render(sc *scene.Scene, width, height int) *image.RGBA {
	startTime := time.Now()
	if r.gpu == nil {
		n := r.numWorkers // "workers" are GPUs now; cmd/raytracer passes runtime.NumCPU(), so clamp
		if have := gpurender.DeviceCount(); n > have {
			n = have
		}
		ctx, err := gpurender.New(n)
		if err != nil {
			panic(err) // no CPU fallback, by design; the reference panics on bad input too (scene.go:105-146)
		}
		r.gpu = ctx
	}
	flat := sc.Flatten()
	if err := r.gpu.Upload(&flat); err != nil {
		panic(err)
	}
	img, st, err := r.gpu.Frame(gpurender.Params{
		Samples: r.samples, MaxDepth: r.maxDepth, AntiAliasing: r.antiAliasing,
		RecursiveReflections: r.recursiveReflections, SoftShadows: r.softShadows,
		Seed: uint64(time.Now().UnixNano()), // the reference seeds math/rand from the clock (math/random.go:8-10)
	}, width, height)
	if err != nil {
		panic(err)
	}
	renderTime := time.Since(startTime)

	r.benchmarkData = BenchmarkData{ // renderer.go:103-117, unchanged fields
		SceneName:         sc.GetSceneName(),
		Resolution:        fmt.Sprintf("%dx%d", width, height),
		RenderTimeSeconds: renderTime.Seconds(),
		Samples:           r.samples,
		MaxDepth:          r.maxDepth,
		NumWorkers:        st.Devices,
		Objects:           len(sc.Objects),
		Lights:            len(sc.Lights),
		Timestamp:         time.Now().Format(time.RFC3339),
		Features:          r.benchmarkData.Features,
	}
	fmt.Printf("Render completed in %v\n", renderTime) // renderer.go:119-123
	fmt.Printf("Resolution: %dx%d\n", width, height)
	fmt.Printf("Samples per pixel: %d\n", r.samples)
	fmt.Printf("Max depth: %d\n", r.maxDepth)
	fmt.Printf("Workers: %d\n", st.Devices)
	return img
}
