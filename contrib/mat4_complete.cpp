// mat4_complete.cpp — SURVEY.md §8(f) N2: the missing bodies of zenyth::math::mat4, written against
// the reference's own headers in the SSE style of its Core/src/math/vector.cpp.
//
// The reference declares the whole mat4 API (Core/include/math/matrix.hpp:11-55) but every body in
// Core/src/math/matrix.cpp:5-121 is a stub (`return {};`, `return *this;`, raw m_data slots).  This
// translation unit is a drop-in REPLACEMENT for that file: same class, same signatures, real
// arithmetic.  It follows the conventions the GPU path and its CPU oracle use (DESIGN.md §3:
// column-major, right-handed, column vectors, Ry*Rx*Rz Euler order, 0..1 clip depth, no fused
// multiply-add, left-to-right column sums), so a Zenyth build that adopts it computes bit-identical
// matrices to libzenyth_b200 — tests/test_mat4_complete.py checks exactly that.
//
// Build (only where the reference tree exists; nothing of it is copied here):
//   make -C oracle mat4      (g++ -std=c++23 -O2 -msse4.1 -mfma -ffp-contract=off -Ioracle/shim
//                             -I$REF/Core/include ... ; takes the place of $REF/Core/src/math/matrix.cpp)
#include "pch.hpp"
#include "math/matrix.hpp"

#include <cmath>

namespace zenyth::math {

namespace {

inline __m128 splat(float s) noexcept { return _mm_set1_ps(s); }

// vec3::y() const returns z in the reference (Core/src/math/vector.cpp:54); the non-const
// overload is correct (vector.cpp:58), so read components from a copy.
struct xyz { float x, y, z; };
inline xyz read(const vec3& v) noexcept { vec3 t = v; return {t.x(), t.y(), t.z()}; }

// One sincos for host and device: fma/mul/add/float->int only.
inline void sincos_det(float x, float& s_out, float& c_out) noexcept {
	const float kf = std::fmaf(x, 0.636619747f, 12582912.0f) - 12582912.0f;
	const int q = static_cast<int>(kf);
	float r = std::fmaf(kf, -1.57079637e+00f, x);
	r = std::fmaf(kf, 4.37113883e-08f, r);
	r = std::fmaf(kf, 1.71512451e-15f, r);
	const float z = r * r;
	float sp = std::fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
	sp = std::fmaf(sp, z, -1.6666654611e-1f);
	const float s = std::fmaf(sp * z, r, r);
	float cp = std::fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
	cp = std::fmaf(cp, z, 4.166664568298827e-2f);
	const float c = std::fmaf(z, std::fmaf(z, cp, -0.5f), 1.0f);
	float so = (q & 1) ? c : s;
	float co = (q & 1) ? s : c;
	if (q & 2) so = 0.0f - so;
	if ((q + 1) & 2) co = 0.0f - co;
	s_out = so;
	c_out = co;
}

inline vec3 safe_normalize(const vec3& v) noexcept {
	const float len = v.length();
	return len > 0.0f ? v / len : v;
}

} // namespace

mat4::mat4() noexcept {}

mat4::mat4(const float diagonal) noexcept
	: m_col{_mm_set_ps(0.0f, 0.0f, 0.0f, diagonal), _mm_set_ps(0.0f, 0.0f, diagonal, 0.0f),
	        _mm_set_ps(0.0f, diagonal, 0.0f, 0.0f), _mm_set_ps(diagonal, 0.0f, 0.0f, 0.0f)} {}

mat4::mat4(const vec4& col0, const vec4& col1, const vec4& col2, const vec4& col3) noexcept
	: m_col{col0.m_simd, col1.m_simd, col2.m_simd, col3.m_simd} {}

mat4::mat4(const __m128 c0, const __m128 c1, const __m128 c2, const __m128 c3) noexcept
	: m_col{c0, c1, c2, c3} {}

mat4 mat4::identity() noexcept { return mat4(1.0f); }
mat4 mat4::zero() noexcept { return {}; }

mat4 mat4::from_basis(const vec3& right, const vec3& up, const vec3& forward) noexcept {
	const xyz r = read(right), u = read(up), f = read(forward);
	return {_mm_set_ps(0.0f, r.z, r.y, r.x), _mm_set_ps(0.0f, u.z, u.y, u.x), _mm_set_ps(0.0f, f.z, f.y, f.x),
	        _mm_set_ps(1.0f, 0.0f, 0.0f, 0.0f)};
}

// Rodrigues about the normalised axis.
mat4 mat4::from_axis_angle(const vec3& axis, const float angle_rad) noexcept {
	const xyz a = read(safe_normalize(axis));
	float s, c;
	sincos_det(angle_rad, s, c);
	const float t = 1.0f - c;
	return {_mm_set_ps(0.0f, t * a.x * a.z - s * a.y, t * a.x * a.y + s * a.z, t * a.x * a.x + c),
	        _mm_set_ps(0.0f, t * a.y * a.z + s * a.x, t * a.y * a.y + c, t * a.x * a.y - s * a.z),
	        _mm_set_ps(0.0f, t * a.z * a.z + c, t * a.y * a.z - s * a.x, t * a.x * a.z + s * a.y),
	        _mm_set_ps(1.0f, 0.0f, 0.0f, 0.0f)};
}

// R = Ry(yaw) * Rx(pitch) * Rz(roll), radians.
mat4 mat4::from_euler(const float pitch, const float yaw, const float roll) noexcept {
	float sx, cx, sy, cy, sz, cz;
	sincos_det(pitch, sx, cx);
	sincos_det(yaw, sy, cy);
	sincos_det(roll, sz, cz);
	const float sysx = sy * sx;
	const float cysx = cy * sx;
	const float r00 = cy * cz + sysx * sz, r01 = sysx * cz - cy * sz, r02 = sy * cx;
	const float r10 = cx * sz, r11 = cx * cz, r12 = 0.0f - sx;
	const float r20 = cysx * sz - sy * cz, r21 = sy * sz + cysx * cz, r22 = cy * cx;
	return {_mm_set_ps(0.0f, r20, r10, r00), _mm_set_ps(0.0f, r21, r11, r01), _mm_set_ps(0.0f, r22, r12, r02),
	        _mm_set_ps(1.0f, 0.0f, 0.0f, 0.0f)};
}

mat4 mat4::from_scale(const vec3& scale) noexcept {
	const xyz s = read(scale);
	return {_mm_set_ps(0.0f, 0.0f, 0.0f, s.x), _mm_set_ps(0.0f, 0.0f, s.y, 0.0f), _mm_set_ps(0.0f, s.z, 0.0f, 0.0f),
	        _mm_set_ps(1.0f, 0.0f, 0.0f, 0.0f)};
}

mat4 mat4::from_translation(const vec3& t) noexcept {
	const xyz p = read(t);
	return {_mm_set_ps(0.0f, 0.0f, 0.0f, 1.0f), _mm_set_ps(0.0f, 0.0f, 1.0f, 0.0f), _mm_set_ps(0.0f, 1.0f, 0.0f, 0.0f),
	        _mm_set_ps(1.0f, p.z, p.y, p.x)};
}

mat4 mat4::operator+(const mat4& rhs) const noexcept {
	return {_mm_add_ps(m_col[0], rhs.m_col[0]), _mm_add_ps(m_col[1], rhs.m_col[1]), _mm_add_ps(m_col[2], rhs.m_col[2]), _mm_add_ps(m_col[3], rhs.m_col[3])};
}
mat4 mat4::operator-(const mat4& rhs) const noexcept {
	return {_mm_sub_ps(m_col[0], rhs.m_col[0]), _mm_sub_ps(m_col[1], rhs.m_col[1]), _mm_sub_ps(m_col[2], rhs.m_col[2]), _mm_sub_ps(m_col[3], rhs.m_col[3])};
}
mat4 mat4::operator*(const float scalar) const noexcept {
	const __m128 s = splat(scalar);
	return {_mm_mul_ps(m_col[0], s), _mm_mul_ps(m_col[1], s), _mm_mul_ps(m_col[2], s), _mm_mul_ps(m_col[3], s)};
}
mat4 mat4::operator/(const float scalar) const noexcept {
	const __m128 s = splat(scalar);
	return {_mm_div_ps(m_col[0], s), _mm_div_ps(m_col[1], s), _mm_div_ps(m_col[2], s), _mm_div_ps(m_col[3], s)};
}

// M*v = ((c0*x + c1*y) + c2*z) + c3*w, the columns summed left to right.
vec4 mat4::operator*(const vec4& v) const noexcept {
	const __m128 x = _mm_shuffle_ps(v.m_simd, v.m_simd, _MM_SHUFFLE(0, 0, 0, 0));
	const __m128 y = _mm_shuffle_ps(v.m_simd, v.m_simd, _MM_SHUFFLE(1, 1, 1, 1));
	const __m128 z = _mm_shuffle_ps(v.m_simd, v.m_simd, _MM_SHUFFLE(2, 2, 2, 2));
	const __m128 w = _mm_shuffle_ps(v.m_simd, v.m_simd, _MM_SHUFFLE(3, 3, 3, 3));
	return vec4{_mm_add_ps(_mm_add_ps(_mm_add_ps(_mm_mul_ps(m_col[0], x), _mm_mul_ps(m_col[1], y)), _mm_mul_ps(m_col[2], z)), _mm_mul_ps(m_col[3], w))};
}

mat4 mat4::operator*(const mat4& rhs) const noexcept {
	const vec4 c0 = *this * vec4{rhs.m_col[0]};
	const vec4 c1 = *this * vec4{rhs.m_col[1]};
	const vec4 c2 = *this * vec4{rhs.m_col[2]};
	const vec4 c3 = *this * vec4{rhs.m_col[3]};
	return {c0, c1, c2, c3};
}

mat4& mat4::operator+=(const mat4& rhs) noexcept { return *this = *this + rhs; }
mat4& mat4::operator-=(const mat4& rhs) noexcept { return *this = *this - rhs; }
mat4& mat4::operator*=(const float scalar) noexcept { return *this = *this * scalar; }
mat4& mat4::operator/=(const float scalar) noexcept { return *this = *this / scalar; }
mat4& mat4::operator*=(const mat4& rhs) noexcept { return *this = *this * rhs; }

mat4 mat4::transpose() const noexcept {
	__m128 r0 = m_col[0], r1 = m_col[1], r2 = m_col[2], r3 = m_col[3];
	_MM_TRANSPOSE4_PS(r0, r1, r2, r3);
	return {r0, r1, r2, r3};
}

namespace {
// 2x2 minors shared by determinant() and inverse(); a(r,c) = d[c*4+r].
struct minors_t { float s0, s1, s2, s3, s4, s5, c0, c1, c2, c3, c4, c5; };
inline minors_t minors(const float* d) noexcept {
	minors_t m;
	m.s0 = d[0] * d[5] - d[1] * d[4];
	m.s1 = d[0] * d[9] - d[1] * d[8];
	m.s2 = d[0] * d[13] - d[1] * d[12];
	m.s3 = d[4] * d[9] - d[5] * d[8];
	m.s4 = d[4] * d[13] - d[5] * d[12];
	m.s5 = d[8] * d[13] - d[9] * d[12];
	m.c5 = d[10] * d[15] - d[11] * d[14];
	m.c4 = d[6] * d[15] - d[7] * d[14];
	m.c3 = d[6] * d[11] - d[7] * d[10];
	m.c2 = d[2] * d[15] - d[3] * d[14];
	m.c1 = d[2] * d[11] - d[3] * d[10];
	m.c0 = d[2] * d[7] - d[3] * d[6];
	return m;
}
inline float det_of(const minors_t& m) noexcept {
	return ((((m.s0 * m.c5 - m.s1 * m.c4) + m.s2 * m.c3) + m.s3 * m.c2) - m.s4 * m.c1) + m.s5 * m.c0;
}
} // namespace

float mat4::determinant() const noexcept { return det_of(minors(m_data)); }

// Adjugate / determinant; a singular matrix maps to zero().
mat4 mat4::inverse() const noexcept {
	const float* d = m_data;
	const minors_t k = minors(d);
	const float det = det_of(k);
	if (det == 0.0f) return zero();
	const float inv = 1.0f / det;
	const float a00 = d[0], a10 = d[1], a20 = d[2], a30 = d[3];
	const float a01 = d[4], a11 = d[5], a21 = d[6], a31 = d[7];
	const float a02 = d[8], a12 = d[9], a22 = d[10], a32 = d[11];
	const float a03 = d[12], a13 = d[13], a23 = d[14], a33 = d[15];
	mat4 r;
	r.m_data[0] = ((a11 * k.c5 - a12 * k.c4) + a13 * k.c3) * inv;
	r.m_data[4] = ((a02 * k.c4 - a01 * k.c5) - a03 * k.c3) * inv;
	r.m_data[8] = ((a31 * k.s5 - a32 * k.s4) + a33 * k.s3) * inv;
	r.m_data[12] = ((a22 * k.s4 - a21 * k.s5) - a23 * k.s3) * inv;
	r.m_data[1] = ((a12 * k.c2 - a10 * k.c5) - a13 * k.c1) * inv;
	r.m_data[5] = ((a00 * k.c5 - a02 * k.c2) + a03 * k.c1) * inv;
	r.m_data[9] = ((a32 * k.s2 - a30 * k.s5) - a33 * k.s1) * inv;
	r.m_data[13] = ((a20 * k.s5 - a22 * k.s2) + a23 * k.s1) * inv;
	r.m_data[2] = ((a10 * k.c4 - a11 * k.c2) + a13 * k.c0) * inv;
	r.m_data[6] = ((a01 * k.c2 - a00 * k.c4) - a03 * k.c0) * inv;
	r.m_data[10] = ((a30 * k.s4 - a31 * k.s2) + a33 * k.s0) * inv;
	r.m_data[14] = ((a21 * k.s2 - a20 * k.s4) - a23 * k.s0) * inv;
	r.m_data[3] = ((a11 * k.c1 - a10 * k.c3) - a12 * k.c0) * inv;
	r.m_data[7] = ((a00 * k.c3 - a01 * k.c1) + a02 * k.c0) * inv;
	r.m_data[11] = ((a31 * k.s1 - a30 * k.s3) - a32 * k.s0) * inv;
	r.m_data[15] = ((a20 * k.s3 - a21 * k.s1) + a22 * k.s0) * inv;
	return r;
}

mat4 mat4::inverse_transpose() const noexcept { return inverse().transpose(); }

// "Transform a position; applies translation" (matrix.hpp:45): w = 1, no perspective divide.
vec3 mat4::transform_point(const vec3& p) const noexcept {
	const xyz q = read(p);
	const vec4 r = *this * vec4{q.x, q.y, q.z, 1.0f};
	return {r.x(), r.y(), r.z()};
}

// "Transform a direction: ignores translation" (matrix.hpp:47): w = 0.
vec3 mat4::transform_normal(const vec3& n) const noexcept {
	const xyz q = read(n);
	const vec4 r = *this * vec4{q.x, q.y, q.z, 0.0f};
	return {r.x(), r.y(), r.z()};
}

// Right-handed view matrix, the camera looks down -Z.
mat4 mat4::look_at(const vec3& eye, const vec3& center, const vec3& up) noexcept {
	const vec3 f = safe_normalize(center - eye);
	const vec3 s = safe_normalize(f.cross(up));
	const vec3 u = s.cross(f);
	const xyz fv = read(f), sv = read(s), uv = read(u);
	return {_mm_set_ps(0.0f, 0.0f - fv.x, uv.x, sv.x), _mm_set_ps(0.0f, 0.0f - fv.y, uv.y, sv.y), _mm_set_ps(0.0f, 0.0f - fv.z, uv.z, sv.z),
	        _mm_set_ps(1.0f, f.dot(eye), 0.0f - u.dot(eye), 0.0f - s.dot(eye))};
}

// Right-handed, clip depth 0..1 (D3D12).
mat4 mat4::perspective(const float fov_y, const float aspect, const float near_z, const float far_z) noexcept {
	const float t = std::tan(fov_y * 0.5f);
	mat4 m;
	m[0, 0] = 1.0f / (aspect * t);
	m[1, 1] = 1.0f / t;
	m[2, 2] = far_z / (near_z - far_z);
	m[3, 2] = -1.0f;
	m[2, 3] = (near_z * far_z) / (near_z - far_z);
	return m;
}

mat4 mat4::orthographic(const float left, const float right, const float bottom, const float top, const float near_z, const float far_z) noexcept {
	mat4 m;
	m[0, 0] = 2.0f / (right - left);
	m[1, 1] = 2.0f / (top - bottom);
	m[2, 2] = 1.0f / (near_z - far_z);
	m[0, 3] = (left + right) / (left - right);
	m[1, 3] = (bottom + top) / (bottom - top);
	m[2, 3] = near_z / (near_z - far_z);
	m[3, 3] = 1.0f;
	return m;
}

// Right-handed, infinite far plane, depth 1 at near and 0 at infinity.
mat4 mat4::perspective_reverse_z(const float fov_y, const float aspect, const float near_z) noexcept {
	const float t = std::tan(fov_y * 0.5f);
	mat4 m;
	m[0, 0] = 1.0f / (aspect * t);
	m[1, 1] = 1.0f / t;
	m[3, 2] = -1.0f;
	m[2, 3] = near_z;
	return m;
}

} // namespace zenyth::math
