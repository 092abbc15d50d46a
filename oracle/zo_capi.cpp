// ORACLE — test infrastructure, not product code (see zo_vec.hpp header).
//
// zo_capi.cpp: extern "C" surface of the float32 oracle, loaded with ctypes by tests/,
// __graft_entry__.smoke() and bench.py (cpu_baseline leg and --impl reference arm) — nowhere
// else.  The product library (zenyth_b200/lib/libzenyth_b200.so) neither links nor loads it.
//
// Built twice from the same sources by oracle/Makefile:
//   oracle/libzo_oracle.so        restated vec3/vec4/mat4 (zo_vec.hpp default backend)
//   oracle/_ref/libzo_oracle_ref.so   -DZO_REFERENCE_TYPES + the reference's own
//                                 Core/src/math/{vector,matrix}.cpp compiled in place
#include "zo_scene.hpp"
#include "zo_synth.hpp"

#include <algorithm>
#include <thread>
#include <vector>

using namespace zo;

namespace {

mat4 in_mat(const float* p) { return load_mat(p); }
void out_mat(float* dst, const mat4& m) { store_mat(dst, m); }
vec3 in_v3(const float* p) { return vec3(p[0], p[1], p[2]); }
void out_v3(float* dst, const vec3& v) { dst[0] = X(v); dst[1] = Y(v); dst[2] = Z(v); }

template <class F>
void parallel_for(uint32_t begin, uint32_t end, unsigned threads, F&& body) {
	const uint32_t n = end - begin;
	if (threads <= 1 || n < 4096) { body(begin, end, 0u); return; }
	const unsigned t = std::min<unsigned>(threads, (n + 1023) / 1024);
	std::vector<std::thread> pool;
	pool.reserve(t);
	for (unsigned k = 0; k < t; ++k) {
		const uint32_t b = begin + uint32_t((uint64_t(n) * k) / t);
		const uint32_t e = begin + uint32_t((uint64_t(n) * (k + 1)) / t);
		pool.emplace_back([&, b, e, k] { body(b, e, k); });
	}
	for (auto& th : pool) th.join();
}

} // namespace

extern "C" {

// 0 = restated types, 1 = the reference's own zenyth::math types.
int zo_backend(void) {
#ifdef ZO_REFERENCE_TYPES
	return 1;
#else
	return 0;
#endif
}

unsigned zo_hardware_threads(void) { const unsigned n = std::thread::hardware_concurrency(); return n ? n : 1; }

// ---- vector ops (exercise the backend's vec3 / vec4 directly) --------------------------------
void zo_vec3_binary(int op, const float* a, const float* b, float* out) {
	const vec3 x = in_v3(a), y = in_v3(b);
	switch (op) {
	case 0: out_v3(out, x + y); break;
	case 1: out_v3(out, x - y); break;
	case 2: out_v3(out, x.cross(y)); break;
	case 3: out_v3(out, x * b[0]); break;
	case 4: out_v3(out, x / b[0]); break;
	default: out_v3(out, -x); break;
	}
}
float zo_vec3_dot(const float* a, const float* b) { return in_v3(a).dot(in_v3(b)); }
float zo_vec3_length(const float* a) { return in_v3(a).length(); }
void zo_vec4_binary(int op, const float* a, const float* b, float* out) {
	const vec4 x(a[0], a[1], a[2], a[3]), y(b[0], b[1], b[2], b[3]);
	vec4 r;
	switch (op) {
	case 0: r = x + y; break;
	case 1: r = x - y; break;
	case 3: r = x * b[0]; break;
	case 4: r = x / b[0]; break;
	default: r = -x; break;
	}
	out[0] = r.x(); out[1] = r.y(); out[2] = r.z(); out[3] = r.w();
}
float zo_vec4_dot(const float* a, const float* b) { return vec4(a[0], a[1], a[2], a[3]).dot(vec4(b[0], b[1], b[2], b[3])); }
float zo_vec4_length(const float* a) { return vec4(a[0], a[1], a[2], a[3]).length(); }

// ---- completed mat4 contract -----------------------------------------------------------------
void zo_sincos(float x, float* s, float* c) { sincos(x, *s, *c); }
void zo_sincos_array(uint32_t n, const float* x, float* s, float* c) { for (uint32_t i = 0; i < n; ++i) sincos(x[i], s[i], c[i]); }
void zo_identity(float* out) { out_mat(out, identity()); }
void zo_diagonal(float d, float* out) { out_mat(out, diagonal(d)); }
void zo_from_basis(const float* r, const float* u, const float* f, float* out) { out_mat(out, from_basis(in_v3(r), in_v3(u), in_v3(f))); }
void zo_from_axis_angle(const float* axis, float angle, float* out) { out_mat(out, from_axis_angle(in_v3(axis), angle)); }
void zo_from_euler(float pitch, float yaw, float roll, float* out) { out_mat(out, from_euler(pitch, yaw, roll)); }
void zo_from_scale(const float* s, float* out) { out_mat(out, from_scale(in_v3(s))); }
void zo_from_translation(const float* t, float* out) { out_mat(out, from_translation(in_v3(t))); }
void zo_mat_add(const float* a, const float* b, float* out) { out_mat(out, add(in_mat(a), in_mat(b))); }
void zo_mat_sub(const float* a, const float* b, float* out) { out_mat(out, sub(in_mat(a), in_mat(b))); }
void zo_mat_scale(const float* a, float s, float* out) { out_mat(out, mul(in_mat(a), s)); }
void zo_mat_div(const float* a, float s, float* out) { out_mat(out, div(in_mat(a), s)); }
void zo_mat_mul(const float* a, const float* b, float* out) { out_mat(out, mul(in_mat(a), in_mat(b))); }
void zo_mat_mul_vec4(const float* a, const float* v, float* out) {
	const vec4 r = mul(in_mat(a), vec4(v[0], v[1], v[2], v[3]));
	out[0] = r.x(); out[1] = r.y(); out[2] = r.z(); out[3] = r.w();
}
void zo_transpose(const float* a, float* out) { out_mat(out, transpose(in_mat(a))); }
float zo_determinant(const float* a) { return determinant(in_mat(a)); }
void zo_inverse(const float* a, float* out) { out_mat(out, inverse(in_mat(a))); }
void zo_inverse_transpose(const float* a, float* out) { out_mat(out, inverse_transpose(in_mat(a))); }
void zo_transform_point(const float* a, const float* p, float* out) { out_v3(out, transform_point(in_mat(a), in_v3(p))); }
void zo_transform_normal(const float* a, const float* n, float* out) { out_v3(out, transform_normal(in_mat(a), in_v3(n))); }
void zo_look_at(const float* eye, const float* center, const float* up, float* out) { out_mat(out, look_at(in_v3(eye), in_v3(center), in_v3(up))); }
void zo_perspective(float fov_y, float aspect, float near_z, float far_z, float* out) { out_mat(out, perspective(fov_y, aspect, near_z, far_z)); }
void zo_orthographic(float l, float r, float b, float t, float n, float f, float* out) { out_mat(out, orthographic(l, r, b, t, n, f)); }
void zo_perspective_reverse_z(float fov_y, float aspect, float near_z, float* out) { out_mat(out, perspective_reverse_z(fov_y, aspect, near_z)); }

// ---- scene pieces ----------------------------------------------------------------------------
void zo_local_matrix(const float* t, const float* r, const float* s, float* out) {
	Transform tr;
	tr.position = in_v3(t); tr.rotation = in_v3(r); tr.scale = in_v3(s);
	out_mat(out, local_matrix(tr));
}
void zo_frustum_planes(const float* view_proj, float* out24) {
	const Frustum f(in_mat(view_proj));
	for (int p = 0; p < 6; ++p) { out24[4 * p] = f.plane[p].x(); out24[4 * p + 1] = f.plane[p].y(); out24[4 * p + 2] = f.plane[p].z(); out24[4 * p + 3] = f.plane[p].w(); }
}
void zo_world_aabb(const float* m, const float* c, const float* e, float* c_out, float* e_out) {
	Aabb b; b.center = in_v3(c); b.half_extent = in_v3(e);
	const Aabb w = world_aabb(in_mat(m), b);
	out_v3(c_out, w.center); out_v3(e_out, w.half_extent);
}
int zo_aabb_visible(const float* view_proj, const float* c, const float* e) {
	Aabb b; b.center = in_v3(c); b.half_extent = in_v3(e);
	return is_visible(Frustum(in_mat(view_proj)), b) ? 1 : 0;
}

// One frame over a level-ordered scene.  Levels run in order; inside a level the entity range
// is split over `threads` host threads (the CPU baseline of SURVEY.md §8d).  Outputs:
// world/mvp [E][16] column-major (mvp may be null), flags [E] (may be null), visible [<=E]
// ascending entity indices + *visible_count (both may be null).
void zo_scene_update(uint32_t entity_count, const uint32_t* level_offsets, uint32_t level_count,
                     const float* position, const float* rotation, const float* scale,
                     const float* aabb_center, const float* aabb_half, const uint32_t* parent,
                     const float* view16, const float* proj16,
                     float* world, float* mvp, uint8_t* flags,
                     uint32_t* visible, uint32_t* visible_count, unsigned threads) {
	SceneView s;
	s.entity_count = entity_count;
	s.position = position; s.rotation = rotation; s.scale = scale;
	s.aabb_center = aabb_center; s.aabb_half = aabb_half; s.parent = parent;
	const Camera cam(in_mat(view16), in_mat(proj16));
	const Frustum fr(cam.view_proj);
	std::vector<uint8_t> own_flags;
	if (!flags) { own_flags.resize(entity_count); flags = own_flags.data(); }
	const uint32_t one_level[2] = {0u, entity_count};
	if (!level_offsets) { level_offsets = one_level; level_count = 1; }
	for (uint32_t l = 0; l < level_count; ++l) {
		parallel_for(level_offsets[l], level_offsets[l + 1], threads, [&](uint32_t b, uint32_t e, unsigned) {
			for (uint32_t i = b; i < e; ++i)
				flags[i] = update_entity(s, i, cam.view_proj, fr, world, mvp) ? 1 : 0;
		});
	}
	if (visible_count || visible) {
		// ordered compaction: per-chunk counts, exclusive scan, scatter
		const unsigned t = std::max(1u, threads);
		std::vector<uint32_t> counts(t + 1, 0);
		parallel_for(0, entity_count, threads, [&](uint32_t b, uint32_t e, unsigned k) {
			uint32_t c = 0;
			for (uint32_t i = b; i < e; ++i) c += flags[i];
			counts[k + 1] = c;
		});
		for (unsigned k = 0; k < t; ++k) counts[k + 1] += counts[k];
		if (visible)
			parallel_for(0, entity_count, threads, [&](uint32_t b, uint32_t e, unsigned k) {
				uint32_t o = counts[k];
				for (uint32_t i = b; i < e; ++i) if (flags[i]) visible[o++] = i;
			});
		if (visible_count) *visible_count = counts[t];
	}
}

void zo_mesh_transform(uint32_t instance_count, const uint32_t* first_vertex /*[I+1]*/,
                       const float* model /*[I][16]*/, const float* pos_in, const float* nrm_in,
                       float* pos_out, float* nrm_out, unsigned threads) {
	parallel_for(0, instance_count, threads, [&](uint32_t b, uint32_t e, unsigned) {
		for (uint32_t k = b; k < e; ++k)
			mesh_transform_instance(in_mat(model + 16 * size_t(k)), pos_in, nrm_in, pos_out, nrm_out,
			                        first_vertex[k], first_vertex[k + 1]);
	});
}

// ---- synthetic inputs ------------------------------------------------------------------------
uint32_t zo_synth_hash(uint32_t seed, uint32_t index, uint32_t field) { return synth_hash(seed, index, field); }

// Fills packed [E][3] arrays and parent[E] for a level-ordered scene.  fanout is the number of
// consecutive children per parent (children of level l hang off level l-1).
void zo_synth_scene(uint32_t seed, const uint32_t* level_offsets, uint32_t level_count, uint32_t fanout,
                    float center_range, float* position, float* rotation, float* scale,
                    float* aabb_center, float* aabb_half, uint32_t* parent, unsigned threads) {
	for (uint32_t l = 0; l < level_count; ++l) {
		const uint32_t off = level_offsets[l], end = level_offsets[l + 1];
		const uint32_t prev_off = l ? level_offsets[l - 1] : 0, prev_size = l ? off - prev_off : 0;
		parallel_for(off, end, threads, [&](uint32_t b, uint32_t e, unsigned) {
			for (uint32_t i = b; i < e; ++i) {
				const bool root = (l == 0);
				for (uint32_t k = 0; k < 3; ++k) {
					position[3 * size_t(i) + k] = synth_entity_field(seed, i, F_TX + k, root, center_range);
					rotation[3 * size_t(i) + k] = synth_entity_field(seed, i, F_RX + k, root, center_range);
					scale[3 * size_t(i) + k] = synth_entity_field(seed, i, F_SX + k, root, center_range);
					aabb_center[3 * size_t(i) + k] = synth_entity_field(seed, i, F_CX + k, root, center_range);
					aabb_half[3 * size_t(i) + k] = synth_entity_field(seed, i, F_HX + k, root, center_range);
				}
				if (parent) parent[i] = root ? kNoParent : synth_parent(prev_off, prev_size, i - off, fanout);
			}
		});
	}
}

void zo_synth_mesh(uint32_t seed, uint32_t vertex_count, float* pos, float* nrm, unsigned threads) {
	parallel_for(0, vertex_count, threads, [&](uint32_t b, uint32_t e, unsigned) {
		for (uint32_t v = b; v < e; ++v) synth_vertex(seed, v, pos + 3 * size_t(v), nrm + 3 * size_t(v));
	});
}

} // extern "C"
