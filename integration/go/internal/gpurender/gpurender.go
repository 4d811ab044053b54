// Package gpurender is the cgo shim between the reference's Go host and libgort.so.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain (`go: command not
// found`).  The file is written against include/gort.h; tests/test_abi.py checks that every C
// symbol used here is exported with the documented signature.
//
// It keeps the reference's surface: renderer.ParallelRenderer.Render keeps its signature
// (internal/renderer/renderer.go:67) and calls gpurender.Frame instead of its goroutine tile pool.
package gpurender

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../../concurrent-raytracer-go_b200/lib -lgort -Wl,-rpath,${SRCDIR}/../../../../concurrent-raytracer-go_b200/lib
#include <stdlib.h>
#include "gort.h"
*/
import "C"

import (
	"fmt"
	"image"
	"sync"
	"unsafe"
)

// Material type tags of gort.h (createMaterial, internal/scene/scene.go:104-148).
const (
	MatLambertian = iota
	MatMetal
	MatShiny
	MatPerfectMirror
	MatGlass
	MatDielectric
	MatDiffuseLight
)

// Flat is scene.Scene after GetHittables()' object factory, as plain arrays (gort_scene_desc).
// internal/scene gains `func (s *Scene) Flatten() gpurender.Flat` that walks s.Objects exactly like
// GetHittables (scene.go:59-90): sphere -> one entry; cube -> the 12 triangles of createCube
// (scene.go:150-190) in the same order; *_Order = position in the reference's scan.
type Flat struct {
	CamPosition, CamLookAt, CamUp [3]float64
	CamFOV, CamAspect             float64
	MatType                       []int32
	MatColor                      []float64 // 3 per material
	MatRoughness, MatMetallic     []float64
	MatSpecular, MatIOR           []float64
	SphereCenter                  []float64 // 3 per sphere
	SphereRadius                  []float64
	SphereMaterial, SphereOrder   []int32
	TriVertices                   []float64 // 9 per triangle
	TriMaterial, TriOrder         []int32
	LightPosition, LightColor     []float64 // 3 per light
	LightIntensity                []float64
}

// Params mirrors the ParallelRenderer fields (renderer.go:20-29).
type Params struct {
	Samples, MaxDepth                                 int
	AntiAliasing, RecursiveReflections, SoftShadows   bool
	Seed                                              uint64
}

// Stats is the subset of gort_stats the Go side writes into its benchmark JSON.
type Stats struct {
	KernelMs, TotalMs, UploadMs, BvhBuildMs float64
	PrimaryRays                             uint64
	Devices                                 int
}

type Context struct{ ctx *C.gort_ctx }

func lastErr(ctx *C.gort_ctx) string { return C.GoString(C.gort_last_error(ctx)) }

// DeviceCount is the number of usable CUDA devices (0 when there is none).
func DeviceCount() int {
	n := int(C.gort_device_count())
	if n < 0 {
		return 0
	}
	return n
}

// New opens nGPUs devices (0..n-1).  There is no CPU fallback: without a CUDA device this fails.
func New(nGPUs int) (*Context, error) {
	var c *C.gort_ctx
	if rc := C.gort_create(nil, C.int(nGPUs), &c); rc != 0 {
		return nil, fmt.Errorf("gort_create: %d: %s", int(rc), lastErr(nil))
	}
	return &Context{ctx: c}, nil
}

func (c *Context) Close() { C.gort_destroy(c.ctx); c.ctx = nil }

func f64(p []float64) *C.double {
	if len(p) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&p[0]))
}
func i32(p []int32) *C.int32_t {
	if len(p) == 0 {
		return nil
	}
	return (*C.int32_t)(unsafe.Pointer(&p[0]))
}

// Upload copies the flat scene to every device and builds the BVH.  The descriptor lives in C
// memory so that it may hold Go pointers for the duration of the call (cgo pointer rules); the
// library keeps no pointer after returning.
func (c *Context) Upload(f *Flat) error {
	d := (*C.gort_scene_desc)(C.calloc(1, C.size_t(unsafe.Sizeof(C.gort_scene_desc{}))))
	defer C.free(unsafe.Pointer(d))
	d.abi_version = C.GORT_ABI_VERSION
	for i := 0; i < 3; i++ {
		d.cam_position[i] = C.double(f.CamPosition[i])
		d.cam_look_at[i] = C.double(f.CamLookAt[i])
		d.cam_up[i] = C.double(f.CamUp[i])
	}
	d.cam_fov, d.cam_aspect = C.double(f.CamFOV), C.double(f.CamAspect)
	d.n_materials = C.int32_t(len(f.MatType))
	d.mat_type, d.mat_color = i32(f.MatType), f64(f.MatColor)
	d.mat_roughness, d.mat_metallic = f64(f.MatRoughness), f64(f.MatMetallic)
	d.mat_specular, d.mat_ior = f64(f.MatSpecular), f64(f.MatIOR)
	d.n_spheres = C.int32_t(len(f.SphereRadius))
	d.sphere_center, d.sphere_radius = f64(f.SphereCenter), f64(f.SphereRadius)
	d.sphere_material, d.sphere_order = i32(f.SphereMaterial), i32(f.SphereOrder)
	d.n_triangles = C.int32_t(len(f.TriMaterial))
	d.tri_vertices, d.tri_material, d.tri_order = f64(f.TriVertices), i32(f.TriMaterial), i32(f.TriOrder)
	d.n_lights = C.int32_t(len(f.LightIntensity))
	d.light_position, d.light_color, d.light_intensity = f64(f.LightPosition), f64(f.LightColor), f64(f.LightIntensity)
	if rc := C.gort_scene_upload(c.ctx, d); rc != 0 {
		return fmt.Errorf("gort_scene_upload: %d: %s", int(rc), lastErr(c.ctx))
	}
	return nil
}

func b2i(b bool) C.int32_t {
	if b {
		return 1
	}
	return 0
}

// pinned remembers which frames live in page-locked C memory (gort_host_alloc) rather than on the Go heap.
var pinned sync.Map // *uint8 -> struct{}

// newFrame backs an *image.RGBA with page-locked memory: the GPU then stores the pixels straight into img.Pix (no
// staging copy; the black regions arrive while the frame is still being traced).  That memory is not garbage
// collected: release it with FreeFrame, or keep reusing the image.  Falls back to the Go heap if pinning fails.
func newFrame(width, height int) *image.RGBA {
	var p unsafe.Pointer
	n := width * height * 4
	if C.gort_host_alloc(C.size_t(n), &p) != 0 {
		return image.NewRGBA(image.Rect(0, 0, width, height))
	}
	pinned.Store((*uint8)(p), struct{}{})
	return &image.RGBA{Pix: unsafe.Slice((*uint8)(p), n), Stride: 4 * width, Rect: image.Rect(0, 0, width, height)}
}

// FreeFrame releases the pixel memory of an image returned by Frame (safe for Go-heap images too).
func FreeFrame(img *image.RGBA) {
	if img == nil || len(img.Pix) == 0 {
		return
	}
	if _, ok := pinned.LoadAndDelete(&img.Pix[0]); ok {
		C.gort_host_free(unsafe.Pointer(&img.Pix[0]))
		img.Pix = nil
	}
}

// Frame renders one image; img.Pix is written directly (Stride must be 4*width — renderer.go:70).
func (c *Context) Frame(p Params, width, height int) (*image.RGBA, Stats, error) {
	img := newFrame(width, height)
	var rp C.gort_render_params
	rp.abi_version = C.GORT_ABI_VERSION
	rp.width, rp.height = C.int32_t(width), C.int32_t(height)
	rp.samples, rp.max_depth = C.int32_t(p.Samples), C.int32_t(p.MaxDepth)
	// The reference stores antiAliasing but never reads it: every sample is jittered
	// (renderer.go:24,155-156).  The drop-in keeps that; anti_aliasing = 0 (fixed 0.5,0.5 offset) is
	// gort's declared extension for the deterministic check and is not reachable from this shim.
	rp.anti_aliasing = 1
	_ = p.AntiAliasing
	rp.recursive_reflections, rp.soft_shadows = b2i(p.RecursiveReflections), b2i(p.SoftShadows)
	rp.camera_mode = C.GORT_CAMERA_REFERENCE
	rp.shard_count = 1
	rp.seed = C.uint64_t(p.Seed)
	var st C.gort_stats
	rc := C.gort_render(c.ctx, &rp, (*C.uint8_t)(unsafe.Pointer(&img.Pix[0])), C.size_t(len(img.Pix)), &st)
	if rc != 0 {
		return nil, Stats{}, fmt.Errorf("gort_render: %d: %s", int(rc), lastErr(c.ctx))
	}
	return img, Stats{KernelMs: float64(st.kernel_ms), TotalMs: float64(st.total_ms), UploadMs: float64(st.upload_ms),
		BvhBuildMs: float64(st.bvh_build_ms), PrimaryRays: uint64(st.primary_rays), Devices: int(st.n_devices)}, nil
}
