// flatten.go — NEW file for the reference's package internal/scene: turns a loaded *Scene into the plain arrays
// libgort takes (gpurender.Flat = gort_scene_desc).  It walks s.Objects in the order GetHittables does
// (scene.go:59-90) and applies the same type switch and defaults as createMaterial (scene.go:104-148) and the same
// vertex/face tables as createCube (scene.go:150-190), but emits numbers instead of geometry.Hittable values.
//
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain in the build image).  The C++ mirror of exactly this logic,
// concurrent-raytracer-go_b200/csrc/host_scene.cpp, is what the tests exercise (tests/test_host_loader.py).
package scene

import (
	gomath "math"

	"raytraceGo/internal/gpurender"
)

func minf(a, b float64) float64 { return gomath.Min(a, b) }

func num(m map[string]interface{}, key string, def float64) float64 {
	if v, ok := m[key]; ok {
		return v.(float64) // panics on a non-number exactly like getFloat, scene.go:218-223
	}
	return def
}

// colour of a material block; a missing "color" yields (1,1,1) where the reference panics (scene.go:113) — the one
// declared deviation, shared with the C++ loader and the oracle (demo-assets/sphere_reflections_light.json:24-27).
func colour(m map[string]interface{}) [3]float64 {
	raw, ok := m["color"].([]interface{})
	if !ok || len(raw) < 3 {
		return [3]float64{1, 1, 1}
	}
	return [3]float64{raw[0].(float64), raw[1].(float64), raw[2].(float64)}
}

// addMaterial appends one material record and returns its index.  Values are the post-constructor ones:
// NewMetal / NewShinyMaterial clamp roughness, metallic and specular with min(x, 1) (material.go:63-73,157-167),
// NewPerfectMirror clamps roughness and reports metallic 1 (advanced_materials.go:117-123,165).
func addMaterial(f *gpurender.Flat, m map[string]interface{}) int32 {
	kind, _ := m["type"].(string)
	c := colour(m)
	typ, rough, metal, spec, ior := int32(gpurender.MatLambertian), 0.0, 0.0, 0.0, 1.5
	switch kind {
	case "metal":
		typ, rough, metal, spec = gpurender.MatMetal, minf(num(m, "roughness", 0), 1), minf(num(m, "metallic", 1), 1), minf(num(m, "specular", 1), 1)
	case "shiny":
		typ, rough, metal, spec = gpurender.MatShiny, minf(num(m, "roughness", 0), 1), minf(num(m, "metallic", 0), 1), minf(num(m, "specular", 1), 1)
	case "perfectmirror":
		typ, rough, metal, spec, ior = gpurender.MatPerfectMirror, minf(num(m, "roughness", 0), 1), 1, 1, 2.0
	case "glass":
		typ, ior = gpurender.MatGlass, num(m, "refractionIndex", 1.5)
	case "dielectric":
		typ, ior, c = gpurender.MatDielectric, num(m, "refractionIndex", 1.5), [3]float64{1, 1, 1}
	case "diffuselight":
		typ = gpurender.MatDiffuseLight
	default: // "lambertian" and anything unknown (scene.go:143-146)
	}
	f.MatType = append(f.MatType, typ)
	f.MatColor = append(f.MatColor, c[0], c[1], c[2])
	f.MatRoughness = append(f.MatRoughness, rough)
	f.MatMetallic = append(f.MatMetallic, metal)
	f.MatSpecular = append(f.MatSpecular, spec)
	f.MatIOR = append(f.MatIOR, ior)
	return int32(len(f.MatType) - 1)
}

// corner signs and face table of createCube (scene.go:153-171); each face (a,b,c,d) becomes triangles (a,b,c),(a,c,d)
var cubeSign = [8][3]float64{{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}}
var cubeFace = [6][4]int{{0, 1, 2, 3}, {1, 5, 6, 2}, {5, 4, 7, 6}, {4, 0, 3, 7}, {3, 2, 6, 7}, {4, 5, 1, 0}}

// Flatten is what Render hands to the GPU instead of GetHittables()/GetLights().
func (s *Scene) Flatten() gpurender.Flat {
	f := gpurender.Flat{
		CamPosition: [3]float64{s.Camera.Position.X, s.Camera.Position.Y, s.Camera.Position.Z},
		CamLookAt:   [3]float64{s.Camera.LookAt.X, s.Camera.LookAt.Y, s.Camera.LookAt.Z},
		CamUp:       [3]float64{s.Camera.Up.X, s.Camera.Up.Y, s.Camera.Up.Z},
		CamFOV:      s.Camera.FOV, CamAspect: s.Camera.AspectRatio,
	}
	order := int32(0) // position in the reference's linear scan: decides exact-tie winners (renderer.go:337-343)
	for _, obj := range s.Objects {
		switch obj.Type {
		case "sphere":
			m := addMaterial(&f, obj.Material)
			f.SphereCenter = append(f.SphereCenter, obj.Position.X, obj.Position.Y, obj.Position.Z)
			f.SphereRadius = append(f.SphereRadius, obj.Radius)
			f.SphereMaterial = append(f.SphereMaterial, m)
			f.SphereOrder = append(f.SphereOrder, order)
			order++
		case "cube":
			m := addMaterial(&f, obj.Material)
			h := [3]float64{obj.Size.X / 2, obj.Size.Y / 2, obj.Size.Z / 2}
			p := [3]float64{obj.Position.X, obj.Position.Y, obj.Position.Z}
			var v [8][3]float64
			for i := range v {
				for a := 0; a < 3; a++ {
					v[i][a] = p[a] + cubeSign[i][a]*h[a]
				}
			}
			for _, q := range cubeFace {
				for _, t := range [2][3]int{{q[0], q[1], q[2]}, {q[0], q[2], q[3]}} {
					for _, k := range t {
						f.TriVertices = append(f.TriVertices, v[k][0], v[k][1], v[k][2])
					}
					f.TriMaterial = append(f.TriMaterial, m)
					f.TriOrder = append(f.TriOrder, order)
					order++
				}
			}
		default: // "Unknown object type": skipped, scene.go:80-82
		}
	}
	for _, l := range s.Lights { // Type is never read by the renderer (renderer.go:248-294)
		f.LightPosition = append(f.LightPosition, l.Position.X, l.Position.Y, l.Position.Z)
		f.LightColor = append(f.LightColor, l.Color.X, l.Color.Y, l.Color.Z)
		f.LightIntensity = append(f.LightIntensity, l.Intensity)
	}
	return f
}
