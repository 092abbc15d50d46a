// zn_hostmath.cpp — host-side camera maths of the C ABI: the part of mat4 that a Camera needs,
// completing the stubbed mat4::look_at / perspective / orthographic / perspective_reverse_z /
// operator* of the reference (Core/include/math/matrix.hpp:32,52-55; stub bodies at
// Core/src/math/matrix.cpp:55-57,107-121).  Compiled with -ffp-contract=off: plain IEEE float
// multiplies and adds in the order of DESIGN.md "Conventions", so the view-projection matrix and
// the frustum planes handed to the kernels are the same bits a CPU evaluation gets.
#include "../../include/zenyth_b200.h"

#include <cmath>

namespace {

struct V3 { float x, y, z; };

inline V3 sub3(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
// (p0+p1)+(p2+p3) with a zero fourth lane, the shape of the reference's vec3::dot
// (Core/src/math/vector.cpp:88).
inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + 0.0f); }
inline V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 normalize3(V3 v) {
	const float len = ::sqrtf(dot3(v, v));
	if (!(len > 0.0f)) return v;
	return {v.x / len, v.y / len, v.z / len};
}
inline void clear16(float* m) { for (int i = 0; i < 16; ++i) m[i] = 0.0f; }

} // namespace

extern "C" {

void zn_mat4_look_at(const float eye[3], const float center[3], const float up[3], float out[16]) {
	const V3 e{eye[0], eye[1], eye[2]};
	const V3 f = normalize3(sub3(V3{center[0], center[1], center[2]}, e));
	const V3 s = normalize3(cross3(f, V3{up[0], up[1], up[2]}));
	const V3 u = cross3(s, f);
	out[0] = s.x; out[1] = u.x; out[2] = 0.0f - f.x; out[3] = 0.0f;
	out[4] = s.y; out[5] = u.y; out[6] = 0.0f - f.y; out[7] = 0.0f;
	out[8] = s.z; out[9] = u.z; out[10] = 0.0f - f.z; out[11] = 0.0f;
	out[12] = 0.0f - dot3(s, e); out[13] = 0.0f - dot3(u, e); out[14] = dot3(f, e); out[15] = 1.0f;
}

void zn_mat4_perspective(float fov_y_rad, float aspect, float near_z, float far_z, float out[16]) {
	const float t = std::tan(fov_y_rad * 0.5f);
	clear16(out);
	out[0] = 1.0f / (aspect * t);
	out[5] = 1.0f / t;
	out[10] = far_z / (near_z - far_z);
	out[11] = -1.0f;
	out[14] = (near_z * far_z) / (near_z - far_z);
}

void zn_mat4_orthographic(float left, float right, float bottom, float top, float near_z, float far_z, float out[16]) {
	clear16(out);
	out[0] = 2.0f / (right - left);
	out[5] = 2.0f / (top - bottom);
	out[10] = 1.0f / (near_z - far_z);
	out[12] = (left + right) / (left - right);
	out[13] = (bottom + top) / (bottom - top);
	out[14] = near_z / (near_z - far_z);
	out[15] = 1.0f;
}

void zn_mat4_perspective_reverse_z(float fov_y_rad, float aspect, float near_z, float out[16]) {
	const float t = std::tan(fov_y_rad * 0.5f);
	clear16(out);
	out[0] = 1.0f / (aspect * t);
	out[5] = 1.0f / t;
	out[11] = -1.0f;
	out[14] = near_z;
}

// out.col[j] = ((a.col0*b0j + a.col1*b1j) + a.col2*b2j) + a.col3*b3j
void zn_mat4_mul(const float a[16], const float b[16], float out[16]) {
	float r[16];
	for (int j = 0; j < 4; ++j)
		for (int i = 0; i < 4; ++i)
			r[j * 4 + i] = ((a[0 + i] * b[j * 4 + 0] + a[4 + i] * b[j * 4 + 1]) + a[8 + i] * b[j * 4 + 2]) + a[12 + i] * b[j * 4 + 3];
	for (int i = 0; i < 16; ++i) out[i] = r[i];
}

void zn_frustum_planes(const float vp[16], float out[24]) {
	for (int k = 0; k < 4; ++k) {
		const float r0 = vp[k * 4 + 0], r1 = vp[k * 4 + 1], r2 = vp[k * 4 + 2], r3 = vp[k * 4 + 3];
		out[0 * 4 + k] = r3 + r0;
		out[1 * 4 + k] = r3 - r0;
		out[2 * 4 + k] = r3 + r1;
		out[3 * 4 + k] = r3 - r1;
		out[4 * 4 + k] = r2;
		out[5 * 4 + k] = r3 - r2;
	}
}

} // extern "C"
