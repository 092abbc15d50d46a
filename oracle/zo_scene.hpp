// ORACLE — test infrastructure, not product code (see zo_vec.hpp header).
//
// zo_scene.hpp: scalar CPU statement of the per-frame scene-update path that
// BASELINE.json.north_star names.  The reference has none of it (no Transform, Scene, Camera,
// Mesh, Aabb, Frustum or culler anywhere; its frame hooks are empty —
// Sandbox/src/SandboxApp.cpp:7-13, Renderer/src/D3D12Renderer.cpp:4-18), so per the north star
// "a scalar C++ oracle built on the reference's own math types stands in for it".  The types
// here follow the reference's idiom (Desc aggregates like AppDesc, Core/include/Application.hpp:9-14;
// noexcept / [[nodiscard]] maths, Core/include/math/matrix.hpp:11-55) and sit where the frame
// loop would call them: Scene update from OnUpdate (Core/src/Application.cpp:47), MVP / cull /
// mesh transform from OnRender between BeginFrame and EndFrame (Application.cpp:49-51).
// PARITY UNPINNED by the reference; pinned by tests/test_oracle_kat.py and zo_f64.cpp.
#pragma once
#include "zo_math.hpp"

#include <cstdint>
#include <cstring>

namespace zo {

inline constexpr uint32_t kNoParent = 0xFFFFFFFFu;

// Local TRS.  rotation = (pitch, yaw, roll) radians, as mat4::from_euler (matrix.hpp:24).
struct Transform {
	vec3 position{0.0f, 0.0f, 0.0f};
	vec3 rotation{0.0f, 0.0f, 0.0f};
	vec3 scale{1.0f, 1.0f, 1.0f};
};

// [chosen] local = T * R * S, written in closed form so no sign-of-zero noise enters:
//   col j (j<3) = R.col[j] * s_j,   col 3 = (t, 1).
// tests/test_oracle_kat.py checks it against from_translation * from_euler * from_scale.
[[nodiscard]] inline mat4 local_matrix(const Transform& t) noexcept {
	float r[9];
	euler_rotation(X(t.rotation), Y(t.rotation), Z(t.rotation), r);
	const float sx = X(t.scale), sy = Y(t.scale), sz = Z(t.scale);
	return mat4(vec4(r[0] * sx, r[3] * sx, r[6] * sx, 0.0f),
	            vec4(r[1] * sy, r[4] * sy, r[7] * sy, 0.0f),
	            vec4(r[2] * sz, r[5] * sz, r[8] * sz, 0.0f),
	            vec4(X(t.position), Y(t.position), Z(t.position), 1.0f));
}

// Camera in the Desc idiom.  Default viewport 1280x720 (Sandbox/include/SandboxApp.hpp:6).
struct CameraDesc {
	vec3 eye{0.0f, 0.0f, 0.0f};
	vec3 center{0.0f, 0.0f, -1.0f};
	vec3 up{0.0f, 1.0f, 0.0f};
	float fov_y_rad = 1.04719755f;
	float aspect = 1280.0f / 720.0f;
	float near_z = 0.1f;
	float far_z = 5000.0f;
};

struct Camera {
	mat4 view, proj, view_proj;
	explicit Camera(const CameraDesc& d) noexcept
		: view(look_at(d.eye, d.center, d.up)),
		  proj(perspective(d.fov_y_rad, d.aspect, d.near_z, d.far_z)),
		  view_proj(mul(proj, view)) {}
	Camera(const mat4& v, const mat4& p) noexcept : view(v), proj(p), view_proj(mul(p, v)) {}
};

// [chosen] Gribb-Hartmann planes from the rows of P*V, D3D depth (0..1), NOT normalised:
//   left r3+r0, right r3-r0, bottom r3+r1, top r3-r1, near r2, far r3-r2.
// A point p is inside when plane.dot(vec4(p,1)) >= 0.
struct Frustum {
	vec4 plane[6];
	explicit Frustum(const mat4& vp) noexcept {
		const float* d = raw(vp);
		const vec4 r0(d[0], d[4], d[8], d[12]);
		const vec4 r1(d[1], d[5], d[9], d[13]);
		const vec4 r2(d[2], d[6], d[10], d[14]);
		const vec4 r3(d[3], d[7], d[11], d[15]);
		plane[0] = r3 + r0;
		plane[1] = r3 - r0;
		plane[2] = r3 + r1;
		plane[3] = r3 - r1;
		plane[4] = r2;
		plane[5] = r3 - r2;
	}
};

struct Aabb {
	vec3 center{0.0f, 0.0f, 0.0f};
	vec3 half_extent{0.0f, 0.0f, 0.0f};
};

// [chosen] local AABB -> world AABB by centre / half-extent (Arvo):
//   c' = M * (c, 1),   e'_r = (|M[r,0]|*e.x + |M[r,1]|*e.y) + |M[r,2]|*e.z
[[nodiscard]] inline Aabb world_aabb(const mat4& m, const Aabb& b) noexcept {
	const float* d = raw(m);
	const float ex = X(b.half_extent), ey = Y(b.half_extent), ez = Z(b.half_extent);
	Aabb w;
	w.center = transform_point(m, b.center);
	w.half_extent = vec3((std::fabs(d[0]) * ex + std::fabs(d[4]) * ey) + std::fabs(d[8]) * ez,
	                     (std::fabs(d[1]) * ex + std::fabs(d[5]) * ey) + std::fabs(d[9]) * ez,
	                     (std::fabs(d[2]) * ex + std::fabs(d[6]) * ey) + std::fabs(d[10]) * ez);
	return w;
}

// Signed margin of a world AABB against one plane:  plane.(c,1) + |n|.e
// (vec4::dot and vec3::dot are the reference's DPPS sums, vector.cpp:88,125).
[[nodiscard]] inline float plane_margin(const vec4& plane, const Aabb& w) noexcept {
	const float dist = plane.dot(vec4(X(w.center), Y(w.center), Z(w.center), 1.0f));
	const vec3 an(std::fabs(plane.x()), std::fabs(plane.y()), std::fabs(plane.z()));
	const float radius = an.dot(w.half_extent);
	return dist + radius;
}

// [chosen] visible iff margin >= 0 on all six planes (inclusive).
[[nodiscard]] inline bool is_visible(const Frustum& f, const Aabb& w) noexcept {
	for (int p = 0; p < 6; ++p)
		if (!(plane_margin(f.plane[p], w) >= 0.0f)) return false;
	return true;
}

// ---------------------------------------------------------------------------------------------
// Scene storage as the oracle sees it: packed float[3] SoA + parent indices, level-ordered
// (every parent precedes its children; level l is [level_offsets[l], level_offsets[l+1])).
struct SceneView {
	uint32_t entity_count = 0;
	const float* position = nullptr;     // [E][3]
	const float* rotation = nullptr;     // [E][3] pitch,yaw,roll
	const float* scale = nullptr;        // [E][3]
	const float* aabb_center = nullptr;  // [E][3]
	const float* aabb_half = nullptr;    // [E][3]
	const uint32_t* parent = nullptr;    // [E], kNoParent for roots; may be null (all roots)
};

inline vec3 load3(const float* p, uint32_t i) noexcept { return vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
inline void store_mat(float* dst, const mat4& m) noexcept { std::memcpy(dst, raw(m), 64); }
inline mat4 load_mat(const float* src) noexcept {
	return mat4(vec4(src[0], src[1], src[2], src[3]), vec4(src[4], src[5], src[6], src[7]),
	            vec4(src[8], src[9], src[10], src[11]), vec4(src[12], src[13], src[14], src[15]));
}

// One entity of the frame: world = parent.world * local (roots: world = local),
// mvp = (P*V) * world, flag = AABB-vs-frustum.  `world` already holds every earlier level.
inline bool update_entity(const SceneView& s, uint32_t i, const mat4& view_proj, const Frustum& fr,
                          float* world /*[E][16]*/, float* mvp /*[E][16]*/) noexcept {
	Transform t;
	t.position = load3(s.position, i);
	t.rotation = load3(s.rotation, i);
	t.scale = load3(s.scale, i);
	const mat4 local = local_matrix(t);
	const uint32_t p = s.parent ? s.parent[i] : kNoParent;
	const mat4 w = (p == kNoParent) ? local : mul(load_mat(world + 16 * size_t(p)), local);
	store_mat(world + 16 * size_t(i), w);
	if (mvp) store_mat(mvp + 16 * size_t(i), mul(view_proj, w));
	Aabb b;
	b.center = load3(s.aabb_center, i);
	b.half_extent = load3(s.aabb_half, i);
	return is_visible(fr, world_aabb(w, b));
}

// ---------------------------------------------------------------------------------------------
// Mesh batch: V vertices, I instances; instance k owns vertices [first[k], first[k+1]).
//   pos' = model.transform_point(pos)
//   nrm' = safe_normalize( model.inverse_transpose().transform_normal(nrm) )
inline void mesh_transform_instance(const mat4& model, const float* pos_in, const float* nrm_in,
                                    float* pos_out, float* nrm_out, uint32_t v0, uint32_t v1) noexcept {
	const mat4 nm = inverse_transpose(model);
	for (uint32_t v = v0; v < v1; ++v) {
		const vec3 p = transform_point(model, load3(pos_in, v));
		pos_out[3 * v] = X(p); pos_out[3 * v + 1] = Y(p); pos_out[3 * v + 2] = Z(p);
		const vec3 n = safe_normalize(transform_normal(nm, load3(nrm_in, v)));
		nrm_out[3 * v] = X(n); nrm_out[3 * v + 1] = Y(n); nrm_out[3 * v + 2] = Z(n);
	}
}

} // namespace zo
