// ORACLE — test infrastructure, not product code (see zo_vec.hpp header).
//
// zo_math.hpp: CPU restatement that COMPLETES the reference's mat4 contract.  Every method the
// reference declares in Core/include/math/matrix.hpp:19-55 has a stub body in
// Core/src/math/matrix.cpp:5-121 (returns {} / *this / raw m_data slots), so there is no
// reference arithmetic to follow for matrices: PARITY UNPINNED by the reference (it ships no
// tests, fixtures or golden vectors).  What the reference does pin, and
// what this file follows:
//   - storage: column-major, m_data[c*4+r], translation in m_data[12..14]   matrix.hpp:7-8,57-58
//   - right-handed                                                           matrix.hpp:7
//   - column-vector convention M*v                                           matrix.hpp:50
//   - transform_point applies translation, transform_normal ignores it       matrix.hpp:45-48
//   - radians                                                                matrix.hpp:23, utils.hpp:5-6
//   - vec3/vec4 arithmetic: SSE lane ops, no FMA, dot = (p0+p1)+(p2+p3)       vector.cpp:61-90,113-127
//   - D3D12 target + perspective_reverse_z  =>  clip depth in [0,1]          Core/CMakeLists.txt:14, matrix.hpp:55
// Everything else is [chosen] and written down in DESIGN.md "Conventions"; the pins are the
// closed-form known-answer tests in tests/test_oracle_kat.py, the algebraic property tests,
// and the independent float64 shadow implementation in zo_f64.cpp.
//
// All arithmetic goes through zo::vec3 / zo::vec4 (the reference's own classes when built with
// -DZO_REFERENCE_TYPES) or through plain float expressions evaluated left to right; the file is
// compiled with -ffp-contract=off so no multiply-add is ever fused, except inside zo::sincos
// where std::fmaf is written out explicitly.
#pragma once
#include "zo_vec.hpp"

namespace zo {

// ---------------------------------------------------------------------------------------------
// Deterministic sincos  [chosen]
// The reference has no trigonometry (from_euler / from_axis_angle are stubs, matrix.cpp:23-29)
// and would reach for <cmath>.  glibc sinf and CUDA sinf differ by 1-2 ulp, which would make
// visible lists differ near plane boundaries, so the conventions pin ONE algorithm built only
// from IEEE-754 fma / mul / add / float->int conversion, which rounds identically on x86 and on
// sm_100a.  Accuracy: <= 1.6 ulp vs float64 sin/cos for |x| <= 1e5 (tests/test_oracle_kat.py).
//   k  = rint(x * 2/pi)          via the 1.5*2^23 magic-number add/sub
//   r  = x - k*pi/2              three-term Cody-Waite with fma (hi, mid, lo = float32 split of pi/2)
//   sin(r), cos(r) on [-pi/4, pi/4]  degree-7 / degree-8 minimax polynomials (Cephes sinf/cosf)
//   quadrant fix-up from k & 3
inline void sincos(float x, float& s_out, float& c_out) noexcept {
	const float kf = std::fmaf(x, 0.636619747f /*0x3f22f983*/, 12582912.0f) - 12582912.0f;
	const int q = static_cast<int>(kf);
	float r = std::fmaf(kf, -1.57079637e+00f /*0x3fc90fdb*/, x);
	r = std::fmaf(kf, 4.37113883e-08f /* -(0xb33bbd2e) */, r);
	r = std::fmaf(kf, 1.71512451e-15f /* -(0xa6f72ced) */, r);
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

// ---------------------------------------------------------------------------------------------
// Factories — complete matrix.hpp:19-26 (stubs at matrix.cpp:11-37)

inline mat4 zero() noexcept { return mat4(); }                                     // matrix.hpp:20
inline mat4 diagonal(float d) noexcept {                                           // matrix.hpp:12 mat4(float)
	return mat4(vec4(d, 0, 0, 0), vec4(0, d, 0, 0), vec4(0, 0, d, 0), vec4(0, 0, 0, d));
}
inline mat4 identity() noexcept { return diagonal(1.0f); }                         // matrix.hpp:19

inline mat4 from_basis(const vec3& right, const vec3& up, const vec3& forward) noexcept { // matrix.hpp:22
	return mat4(vec4(X(right), Y(right), Z(right), 0.0f),
	            vec4(X(up), Y(up), Z(up), 0.0f),
	            vec4(X(forward), Y(forward), Z(forward), 0.0f),
	            vec4(0.0f, 0.0f, 0.0f, 1.0f));
}

inline mat4 from_translation(const vec3& t) noexcept {                             // matrix.hpp:26
	return mat4(vec4(1, 0, 0, 0), vec4(0, 1, 0, 0), vec4(0, 0, 1, 0), vec4(X(t), Y(t), Z(t), 1.0f));
}

inline mat4 from_scale(const vec3& s) noexcept {                                   // matrix.hpp:25
	return mat4(vec4(X(s), 0, 0, 0), vec4(0, Y(s), 0, 0), vec4(0, 0, Z(s), 0), vec4(0, 0, 0, 1));
}

// matrix.hpp:24.  [chosen] R = Ry(yaw) * Rx(pitch) * Rz(roll), radians, with the element
// expressions and their evaluation order fixed here (DESIGN.md "Conventions" C4):
//   r00 =  cy*cz + (sy*sx)*sz   r01 = (sy*sx)*cz - cy*sz   r02 = sy*cx
//   r10 =  cx*sz                r11 =  cx*cz               r12 = -sx
//   r20 = (cy*sx)*sz - sy*cz    r21 =  sy*sz + (cy*sx)*cz  r22 = cy*cx
inline void euler_rotation(float pitch, float yaw, float roll, float r[9] /*row-major r[row*3+col]*/) noexcept {
	float sx, cx, sy, cy, sz, cz;
	sincos(pitch, sx, cx);
	sincos(yaw, sy, cy);
	sincos(roll, sz, cz);
	const float sysx = sy * sx;
	const float cysx = cy * sx;
	r[0] = cy * cz + sysx * sz;
	r[1] = sysx * cz - cy * sz;
	r[2] = sy * cx;
	r[3] = cx * sz;
	r[4] = cx * cz;
	r[5] = 0.0f - sx;
	r[6] = cysx * sz - sy * cz;
	r[7] = sy * sz + cysx * cz;
	r[8] = cy * cx;
}

inline mat4 from_euler(float pitch, float yaw, float roll) noexcept {              // matrix.hpp:24
	float r[9];
	euler_rotation(pitch, yaw, roll, r);
	return mat4(vec4(r[0], r[3], r[6], 0.0f), vec4(r[1], r[4], r[7], 0.0f),
	            vec4(r[2], r[5], r[8], 0.0f), vec4(0.0f, 0.0f, 0.0f, 1.0f));
}

inline vec3 safe_normalize(const vec3& v) noexcept {
	const float len = v.length();                                                  // vector.cpp:90
	return len > 0.0f ? v / len : v;                                               // vector.cpp:64
}

// matrix.hpp:23.  Rodrigues about the normalised axis.
inline mat4 from_axis_angle(const vec3& axis, float angle_rad) noexcept {
	const vec3 a = safe_normalize(axis);
	const float x = X(a), y = Y(a), z = Z(a);
	float s, c;
	sincos(angle_rad, s, c);
	const float t = 1.0f - c;
	return mat4(vec4(t * x * x + c, t * x * y + s * z, t * x * z - s * y, 0.0f),
	            vec4(t * x * y - s * z, t * y * y + c, t * y * z + s * x, 0.0f),
	            vec4(t * x * z + s * y, t * y * z - s * x, t * z * z + c, 0.0f),
	            vec4(0.0f, 0.0f, 0.0f, 1.0f));
}

// ---------------------------------------------------------------------------------------------
// Arithmetic — complete matrix.hpp:28-50 (stubs at matrix.cpp:39-105)

inline mat4 add(const mat4& a, const mat4& b) noexcept {                           // matrix.hpp:28
	return mat4(column(a, 0) + column(b, 0), column(a, 1) + column(b, 1), column(a, 2) + column(b, 2), column(a, 3) + column(b, 3));
}
inline mat4 sub(const mat4& a, const mat4& b) noexcept {                           // matrix.hpp:29
	return mat4(column(a, 0) - column(b, 0), column(a, 1) - column(b, 1), column(a, 2) - column(b, 2), column(a, 3) - column(b, 3));
}
inline mat4 mul(const mat4& a, float s) noexcept {                                 // matrix.hpp:30
	return mat4(column(a, 0) * s, column(a, 1) * s, column(a, 2) * s, column(a, 3) * s);
}
inline mat4 div(const mat4& a, float s) noexcept {                                 // matrix.hpp:31
	return mat4(column(a, 0) / s, column(a, 1) / s, column(a, 2) / s, column(a, 3) / s);
}

// matrix.hpp:50 (stub at matrix.cpp:103-105).  Column form, summed left to right:
//   M*v = ((col0*v.x + col1*v.y) + col2*v.z) + col3*v.w
inline vec4 mul(const mat4& m, const vec4& v) noexcept {
	return ((column(m, 0) * v.x() + column(m, 1) * v.y()) + column(m, 2) * v.z()) + column(m, 3) * v.w();
}

// matrix.hpp:32 (stub at matrix.cpp:55-57).  out.col[j] = a * b.col[j].
inline mat4 mul(const mat4& a, const mat4& b) noexcept {
	return mat4(mul(a, column(b, 0)), mul(a, column(b, 1)), mul(a, column(b, 2)), mul(a, column(b, 3)));
}

inline mat4 transpose(const mat4& m) noexcept {                                    // matrix.hpp:40
	const float* d = raw(m);
	return mat4(vec4(d[0], d[4], d[8], d[12]), vec4(d[1], d[5], d[9], d[13]),
	            vec4(d[2], d[6], d[10], d[14]), vec4(d[3], d[7], d[11], d[15]));
}

// 2x2 minors shared by determinant() and inverse().  a(r,c) = d[c*4+r].
struct Minors {
	float s0, s1, s2, s3, s4, s5, c0, c1, c2, c3, c4, c5;
};
inline Minors minors(const float* d) noexcept {
	const float a00 = d[0], a10 = d[1], a20 = d[2], a30 = d[3];
	const float a01 = d[4], a11 = d[5], a21 = d[6], a31 = d[7];
	const float a02 = d[8], a12 = d[9], a22 = d[10], a32 = d[11];
	const float a03 = d[12], a13 = d[13], a23 = d[14], a33 = d[15];
	Minors m;
	m.s0 = a00 * a11 - a10 * a01;
	m.s1 = a00 * a12 - a10 * a02;
	m.s2 = a00 * a13 - a10 * a03;
	m.s3 = a01 * a12 - a11 * a02;
	m.s4 = a01 * a13 - a11 * a03;
	m.s5 = a02 * a13 - a12 * a03;
	m.c5 = a22 * a33 - a32 * a23;
	m.c4 = a21 * a33 - a31 * a23;
	m.c3 = a21 * a32 - a31 * a22;
	m.c2 = a20 * a33 - a30 * a23;
	m.c1 = a20 * a32 - a30 * a22;
	m.c0 = a20 * a31 - a30 * a21;
	return m;
}
inline float det_from(const Minors& m) noexcept {
	return ((((m.s0 * m.c5 - m.s1 * m.c4) + m.s2 * m.c3) + m.s3 * m.c2) - m.s4 * m.c1) + m.s5 * m.c0;
}

inline float determinant(const mat4& m) noexcept {                                 // matrix.hpp:41
	return det_from(minors(raw(m)));
}

// matrix.hpp:42.  Adjugate / determinant; [chosen] a singular matrix (det == 0) maps to zero().
inline mat4 inverse(const mat4& m) noexcept {
	const float* d = raw(m);
	const Minors k = minors(d);
	const float det = det_from(k);
	if (det == 0.0f) return zero();
	const float inv = 1.0f / det;
	const float a00 = d[0], a10 = d[1], a20 = d[2], a30 = d[3];
	const float a01 = d[4], a11 = d[5], a21 = d[6], a31 = d[7];
	const float a02 = d[8], a12 = d[9], a22 = d[10], a32 = d[11];
	const float a03 = d[12], a13 = d[13], a23 = d[14], a33 = d[15];
	const float b00 = ((a11 * k.c5 - a12 * k.c4) + a13 * k.c3) * inv;
	const float b01 = ((a02 * k.c4 - a01 * k.c5) - a03 * k.c3) * inv;
	const float b02 = ((a31 * k.s5 - a32 * k.s4) + a33 * k.s3) * inv;
	const float b03 = ((a22 * k.s4 - a21 * k.s5) - a23 * k.s3) * inv;
	const float b10 = ((a12 * k.c2 - a10 * k.c5) - a13 * k.c1) * inv;
	const float b11 = ((a00 * k.c5 - a02 * k.c2) + a03 * k.c1) * inv;
	const float b12 = ((a32 * k.s2 - a30 * k.s5) - a33 * k.s1) * inv;
	const float b13 = ((a20 * k.s5 - a22 * k.s2) + a23 * k.s1) * inv;
	const float b20 = ((a10 * k.c4 - a11 * k.c2) + a13 * k.c0) * inv;
	const float b21 = ((a01 * k.c2 - a00 * k.c4) - a03 * k.c0) * inv;
	const float b22 = ((a30 * k.s4 - a31 * k.s2) + a33 * k.s0) * inv;
	const float b23 = ((a21 * k.s2 - a20 * k.s4) - a23 * k.s0) * inv;
	const float b30 = ((a11 * k.c1 - a10 * k.c3) - a12 * k.c0) * inv;
	const float b31 = ((a00 * k.c3 - a01 * k.c1) + a02 * k.c0) * inv;
	const float b32 = ((a31 * k.s1 - a30 * k.s3) - a32 * k.s0) * inv;
	const float b33 = ((a20 * k.s3 - a21 * k.s1) + a22 * k.s0) * inv;
	return mat4(vec4(b00, b10, b20, b30), vec4(b01, b11, b21, b31), vec4(b02, b12, b22, b32), vec4(b03, b13, b23, b33));
}

inline mat4 inverse_transpose(const mat4& m) noexcept { return transpose(inverse(m)); } // matrix.hpp:43

// matrix.hpp:45-46 "Transform a position; applies translation" — w = 1, no perspective divide.
inline vec3 transform_point(const mat4& m, const vec3& p) noexcept {
	const vec4 r = mul(m, vec4(X(p), Y(p), Z(p), 1.0f));
	return vec3(r.x(), r.y(), r.z());
}
// matrix.hpp:47-48 "Transform a direction: ignores translation" — w = 0.
inline vec3 transform_normal(const mat4& m, const vec3& n) noexcept {
	const vec4 r = mul(m, vec4(X(n), Y(n), Z(n), 0.0f));
	return vec3(r.x(), r.y(), r.z());
}

// ---------------------------------------------------------------------------------------------
// Camera matrices — complete matrix.hpp:52-55 (stubs at matrix.cpp:107-121)

// matrix.hpp:52.  [chosen] RH, camera looks down -Z: f = norm(center-eye), s = norm(f x up),
// u = s x f; rows (s, -s.eye), (u, -u.eye), (-f, f.eye), (0,0,0,1).
inline mat4 look_at(const vec3& eye, const vec3& center, const vec3& up) noexcept {
	const vec3 f = safe_normalize(center - eye);
	const vec3 s = safe_normalize(f.cross(up));                                    // vector.cpp:73-86
	const vec3 u = s.cross(f);
	return mat4(vec4(X(s), X(u), 0.0f - X(f), 0.0f),
	            vec4(Y(s), Y(u), 0.0f - Y(f), 0.0f),
	            vec4(Z(s), Z(u), 0.0f - Z(f), 0.0f),
	            vec4(0.0f - s.dot(eye), 0.0f - u.dot(eye), f.dot(eye), 1.0f));      // vector.cpp:88
}

// matrix.hpp:53.  [chosen] RH, clip depth 0..1 (D3D12): m00 = 1/(aspect*tan(fov/2)),
// m11 = 1/tan(fov/2), m22 = far/(near-far), m32 = -1, m23 = near*far/(near-far).
inline mat4 perspective(float fov_y, float aspect, float near_z, float far_z) noexcept {
	const float t = std::tan(fov_y * 0.5f);
	mat4 m;
	m[0, 0] = 1.0f / (aspect * t);
	m[1, 1] = 1.0f / t;
	m[2, 2] = far_z / (near_z - far_z);
	m[3, 2] = -1.0f;
	m[2, 3] = (near_z * far_z) / (near_z - far_z);
	return m;
}

// matrix.hpp:54.  [chosen] RH, clip depth 0..1.
inline mat4 orthographic(float left, float right, float bottom, float top, float near_z, float far_z) noexcept {
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

// matrix.hpp:55.  [chosen] RH, infinite far plane, depth 1 at near and 0 at infinity.
inline mat4 perspective_reverse_z(float fov_y, float aspect, float near_z) noexcept {
	const float t = std::tan(fov_y * 0.5f);
	mat4 m;
	m[0, 0] = 1.0f / (aspect * t);
	m[1, 1] = 1.0f / t;
	m[3, 2] = -1.0f;
	m[2, 3] = near_z;
	return m;
}

} // namespace zo
