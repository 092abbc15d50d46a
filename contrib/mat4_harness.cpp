// mat4_harness.cpp — extern "C" access to the COMPLETED zenyth::math::mat4 (contrib/mat4_complete.cpp
// linked with the reference's own vector.cpp) for tests/test_mat4_complete.py.  Test infrastructure.
#include "pch.hpp"
#include "math/matrix.hpp"

#include <cstring>

using namespace zenyth::math;

namespace {
mat4 in(const float* p) { return mat4(vec4(p[0], p[1], p[2], p[3]), vec4(p[4], p[5], p[6], p[7]), vec4(p[8], p[9], p[10], p[11]), vec4(p[12], p[13], p[14], p[15])); }
void out(float* d, const mat4& m) { std::memcpy(d, m.data(), 64); }
vec3 v3(const float* p) { return vec3(p[0], p[1], p[2]); }
void o3(float* d, const vec3& v) { vec3 t = v; d[0] = t.x(); d[1] = t.y(); d[2] = t.z(); }
}

extern "C" {
void mc_identity(float* o) { out(o, mat4::identity()); }
void mc_diagonal(float d, float* o) { out(o, mat4(d)); }
void mc_from_basis(const float* r, const float* u, const float* f, float* o) { out(o, mat4::from_basis(v3(r), v3(u), v3(f))); }
void mc_from_axis_angle(const float* a, float ang, float* o) { out(o, mat4::from_axis_angle(v3(a), ang)); }
void mc_from_euler(float p, float y, float r, float* o) { out(o, mat4::from_euler(p, y, r)); }
void mc_from_scale(const float* s, float* o) { out(o, mat4::from_scale(v3(s))); }
void mc_from_translation(const float* t, float* o) { out(o, mat4::from_translation(v3(t))); }
void mc_add(const float* a, const float* b, float* o) { out(o, in(a) + in(b)); }
void mc_sub(const float* a, const float* b, float* o) { out(o, in(a) - in(b)); }
void mc_scale(const float* a, float s, float* o) { out(o, in(a) * s); }
void mc_div(const float* a, float s, float* o) { out(o, in(a) / s); }
void mc_mul(const float* a, const float* b, float* o) { out(o, in(a) * in(b)); }
void mc_mul_assign(const float* a, const float* b, float* o) { mat4 m = in(a); m *= in(b); out(o, m); }
void mc_mul_vec4(const float* a, const float* v, float* o) { const vec4 r = in(a) * vec4(v[0], v[1], v[2], v[3]); o[0] = r.x(); o[1] = r.y(); o[2] = r.z(); o[3] = r.w(); }
void mc_transpose(const float* a, float* o) { out(o, in(a).transpose()); }
float mc_determinant(const float* a) { return in(a).determinant(); }
void mc_inverse(const float* a, float* o) { out(o, in(a).inverse()); }
void mc_inverse_transpose(const float* a, float* o) { out(o, in(a).inverse_transpose()); }
void mc_transform_point(const float* a, const float* p, float* o) { o3(o, in(a).transform_point(v3(p))); }
void mc_transform_normal(const float* a, const float* n, float* o) { o3(o, in(a).transform_normal(v3(n))); }
void mc_look_at(const float* e, const float* c, const float* u, float* o) { out(o, mat4::look_at(v3(e), v3(c), v3(u))); }
void mc_perspective(float f, float a, float n, float fz, float* o) { out(o, mat4::perspective(f, a, n, fz)); }
void mc_orthographic(float l, float r, float b, float t, float n, float f, float* o) { out(o, mat4::orthographic(l, r, b, t, n, f)); }
void mc_perspective_reverse_z(float f, float a, float n, float* o) { out(o, mat4::perspective_reverse_z(f, a, n)); }
}
