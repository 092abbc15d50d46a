// zn_math.cuh — device arithmetic of the scene path.
//
// Every float operation is written with an explicit round-to-nearest intrinsic (__fmul_rn,
// __fadd_rn, __fsub_rn, __fmaf_rn, __fdiv_rn, __fsqrt_rn): nvcc never contracts those into FMA
// and never reorders them, so the operation sequence below IS the result.  The sequence follows
// DESIGN.md "Conventions" C1-C9 — column-major mat4, M*v summed left to right over columns
// (the shape the reference's SSE vec4 code gives: Core/src/math/vector.cpp:113-115), dot
// products summed (p0+p1)+(p2+p3) as DPPS does (vector.cpp:88,125) — which is what makes
// visible lists match a CPU evaluation bit for bit instead of "within tolerance".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace zn {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// Conventions C3: deterministic sincos built only from fma / mul / add / float->int, so it rounds
// the same on x86 and on sm_100a (<= 1.6 ulp vs float64 for |x| <= 1e5).  Not sincosf(): the CUDA
// and glibc libm results differ by 1-2 ulp.  Not __sincosf(): SFU accuracy breaks the 1e-5 gate.
__device__ __forceinline__ void sincos_det(float x, float& s_out, float& c_out) {
	const float kf = sub(fma_(x, 0.636619747f, 12582912.0f), 12582912.0f);
	const int q = __float2int_rz(kf);
	float r = fma_(kf, -1.57079637e+00f, x);
	r = fma_(kf, 4.37113883e-08f, r);
	r = fma_(kf, 1.71512451e-15f, r);
	const float z = mul(r, r);
	float sp = fma_(-1.9515295891e-4f, z, 8.3321608736e-3f);
	sp = fma_(sp, z, -1.6666654611e-1f);
	const float s = fma_(mul(sp, z), r, r);
	float cp = fma_(2.443315711809948e-5f, z, -1.388731625493765e-3f);
	cp = fma_(cp, z, 4.166664568298827e-2f);
	const float c = fma_(z, fma_(z, cp, -0.5f), 1.0f);
	float so = (q & 1) ? c : s;
	float co = (q & 1) ? s : c;
	if (q & 2) so = sub(0.0f, so);
	if ((q + 1) & 2) co = sub(0.0f, co);
	s_out = so;
	c_out = co;
}

// Affine 3x4 (the bottom row of every local / world matrix is exactly 0,0,0,1 for finite inputs,
// so it is carried implicitly).  m[r][c], r in 0..2, c in 0..3.
struct Affine {
	float m[3][4];
};

// Conventions C4/C5: local = T * Ry(yaw)*Rx(pitch)*Rz(roll) * S in closed form.
__device__ __forceinline__ Affine local_trs(float tx, float ty, float tz, float pitch, float yaw, float roll,
                                            float sx_, float sy_, float sz_) {
	float sx, cx, sy, cy, sz, cz;
	sincos_det(pitch, sx, cx);
	sincos_det(yaw, sy, cy);
	sincos_det(roll, sz, cz);
	const float sysx = mul(sy, sx);
	const float cysx = mul(cy, sx);
	const float r00 = add(mul(cy, cz), mul(sysx, sz));
	const float r01 = sub(mul(sysx, cz), mul(cy, sz));
	const float r02 = mul(sy, cx);
	const float r10 = mul(cx, sz);
	const float r11 = mul(cx, cz);
	const float r12 = sub(0.0f, sx);
	const float r20 = sub(mul(cysx, sz), mul(sy, cz));
	const float r21 = add(mul(sy, sz), mul(cysx, cz));
	const float r22 = mul(cy, cx);
	Affine a;
	a.m[0][0] = mul(r00, sx_); a.m[1][0] = mul(r10, sx_); a.m[2][0] = mul(r20, sx_);
	a.m[0][1] = mul(r01, sy_); a.m[1][1] = mul(r11, sy_); a.m[2][1] = mul(r21, sy_);
	a.m[0][2] = mul(r02, sz_); a.m[1][2] = mul(r12, sz_); a.m[2][2] = mul(r22, sz_);
	a.m[0][3] = tx; a.m[1][3] = ty; a.m[2][3] = tz;
	return a;
}

// world = parent * local for affine operands.  The generic column form is
//   out.col[j] = ((P.col0*l0j + P.col1*l1j) + P.col2*l2j) + P.col3*l3j
// with l3j = 0 for j<3 (adds an exact zero: dropped) and l33 = 1 (P.col3*1 = P.col3).
__device__ __forceinline__ Affine affine_mul(const Affine& p, const Affine& l) {
	Affine w;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
#pragma unroll
		for (int r = 0; r < 3; ++r) {
			float v = add(add(mul(p.m[r][0], l.m[0][j]), mul(p.m[r][1], l.m[1][j])), mul(p.m[r][2], l.m[2][j]));
			if (j == 3) v = add(v, p.m[r][3]);
			w.m[r][j] = v;
		}
	}
	return w;
}

// mvp = VP * world, VP a general 4x4 (column-major vp[c*4+r]), world affine.  out[c*4+r].
__device__ __forceinline__ void mvp_mul(const float* __restrict__ vp, const Affine& w, float out[16]) {
#pragma unroll
	for (int j = 0; j < 4; ++j) {
#pragma unroll
		for (int r = 0; r < 4; ++r) {
			float v = add(add(mul(vp[0 * 4 + r], w.m[0][j]), mul(vp[1 * 4 + r], w.m[1][j])), mul(vp[2 * 4 + r], w.m[2][j]));
			if (j == 3) v = add(v, vp[3 * 4 + r]);
			out[j * 4 + r] = v;
		}
	}
}

// Conventions C7/C8: world AABB by centre / half-extent, then six plane margins; visible iff all
// margins >= 0.  planes[p*4 + {0,1,2,3}] = (nx, ny, nz, d), not normalised.
__device__ __forceinline__ bool aabb_visible(const Affine& w, float cx, float cy, float cz, float hx, float hy, float hz,
                                             const float* __restrict__ planes) {
	float c[3], e[3];
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		c[r] = add(add(add(mul(w.m[r][0], cx), mul(w.m[r][1], cy)), mul(w.m[r][2], cz)), w.m[r][3]);
		e[r] = add(add(mul(fabsf(w.m[r][0]), hx), mul(fabsf(w.m[r][1]), hy)), mul(fabsf(w.m[r][2]), hz));
	}
	bool vis = true;
#pragma unroll
	for (int p = 0; p < 6; ++p) {
		const float nx = planes[p * 4 + 0], ny = planes[p * 4 + 1], nz = planes[p * 4 + 2], d = planes[p * 4 + 3];
		const float dist = add(add(mul(nx, c[0]), mul(ny, c[1])), add(mul(nz, c[2]), d));
		const float rad = add(add(mul(fabsf(nx), e[0]), mul(fabsf(ny), e[1])), mul(fabsf(nz), e[2]));
		vis = vis && (add(dist, rad) >= 0.0f);
	}
	return vis;
}

// ---- general 4x4 (mesh normal matrix) -----------------------------------------------------------
// Conventions C9: inverse by 2x2 minors (adjugate / determinant); det == 0 -> zero matrix.
// d and out are column-major [c*4+r].  Same operation order as the CPU statement.
__device__ __forceinline__ void mat4_inverse(const float* __restrict__ d, float* __restrict__ out) {
	const float a00 = d[0], a10 = d[1], a20 = d[2], a30 = d[3];
	const float a01 = d[4], a11 = d[5], a21 = d[6], a31 = d[7];
	const float a02 = d[8], a12 = d[9], a22 = d[10], a32 = d[11];
	const float a03 = d[12], a13 = d[13], a23 = d[14], a33 = d[15];
	const float s0 = sub(mul(a00, a11), mul(a10, a01));
	const float s1 = sub(mul(a00, a12), mul(a10, a02));
	const float s2 = sub(mul(a00, a13), mul(a10, a03));
	const float s3 = sub(mul(a01, a12), mul(a11, a02));
	const float s4 = sub(mul(a01, a13), mul(a11, a03));
	const float s5 = sub(mul(a02, a13), mul(a12, a03));
	const float c5 = sub(mul(a22, a33), mul(a32, a23));
	const float c4 = sub(mul(a21, a33), mul(a31, a23));
	const float c3 = sub(mul(a21, a32), mul(a31, a22));
	const float c2 = sub(mul(a20, a33), mul(a30, a23));
	const float c1 = sub(mul(a20, a32), mul(a30, a22));
	const float c0 = sub(mul(a20, a31), mul(a30, a21));
	const float det = add(sub(add(add(sub(mul(s0, c5), mul(s1, c4)), mul(s2, c3)), mul(s3, c2)), mul(s4, c1)), mul(s5, c0));
	if (det == 0.0f) {
#pragma unroll
		for (int i = 0; i < 16; ++i) out[i] = 0.0f;
		return;
	}
	const float inv = __fdiv_rn(1.0f, det);
	const float b00 = mul(add(sub(mul(a11, c5), mul(a12, c4)), mul(a13, c3)), inv);
	const float b01 = mul(sub(sub(mul(a02, c4), mul(a01, c5)), mul(a03, c3)), inv);
	const float b02 = mul(add(sub(mul(a31, s5), mul(a32, s4)), mul(a33, s3)), inv);
	const float b03 = mul(sub(sub(mul(a22, s4), mul(a21, s5)), mul(a23, s3)), inv);
	const float b10 = mul(sub(sub(mul(a12, c2), mul(a10, c5)), mul(a13, c1)), inv);
	const float b11 = mul(add(sub(mul(a00, c5), mul(a02, c2)), mul(a03, c1)), inv);
	const float b12 = mul(sub(sub(mul(a32, s2), mul(a30, s5)), mul(a33, s1)), inv);
	const float b13 = mul(add(sub(mul(a20, s5), mul(a22, s2)), mul(a23, s1)), inv);
	const float b20 = mul(add(sub(mul(a10, c4), mul(a11, c2)), mul(a13, c0)), inv);
	const float b21 = mul(sub(sub(mul(a01, c2), mul(a00, c4)), mul(a03, c0)), inv);
	const float b22 = mul(add(sub(mul(a30, s4), mul(a31, s2)), mul(a33, s0)), inv);
	const float b23 = mul(sub(sub(mul(a21, s2), mul(a20, s4)), mul(a23, s0)), inv);
	const float b30 = mul(sub(sub(mul(a11, c1), mul(a10, c3)), mul(a12, c0)), inv);
	const float b31 = mul(add(sub(mul(a00, c3), mul(a01, c1)), mul(a02, c0)), inv);
	const float b32 = mul(sub(sub(mul(a31, s1), mul(a30, s3)), mul(a32, s0)), inv);
	const float b33 = mul(add(sub(mul(a20, s3), mul(a21, s1)), mul(a22, s0)), inv);
	out[0] = b00; out[1] = b10; out[2] = b20; out[3] = b30;
	out[4] = b01; out[5] = b11; out[6] = b21; out[7] = b31;
	out[8] = b02; out[9] = b12; out[10] = b22; out[11] = b32;
	out[12] = b03; out[13] = b13; out[14] = b23; out[15] = b33;
}

// The upper-left 3x3 of mat4_inverse(d) for an AFFINE matrix — bottom row exactly (0, 0, 0, 1), every
// other entry finite — which is what a scene's world matrices are.  Not another algorithm: it is
// mat4_inverse with the operations that IEEE arithmetic makes exact under those conditions folded
// away (x*1 = x, x*0 = +-0, y + +-0 = y, y - +-0 = y), so every result equals the general routine's
// bit for bit (up to the sign of a zero).  c5 = a22, c4 = a21, c2 = a20, c3 = c1 = c0 = +-0; of the
// s-terms only s0, s1, s3 are needed; det = (s0*a22 - s1*a21) + s3*a20.  ~45 operations instead of
// ~170.  Returns false (and computes nothing) when the matrix does not qualify.
// inv3[r*3 + c] = inverse(r, c) for r, c < 3.
__device__ __forceinline__ bool mat4_inverse_affine3x3(const float* __restrict__ d, float* __restrict__ inv3) {
	const float a00 = d[0], a10 = d[1], a20 = d[2];
	const float a01 = d[4], a11 = d[5], a21 = d[6];
	const float a02 = d[8], a12 = d[9], a22 = d[10];
	const float a03 = d[12], a13 = d[13], a23 = d[14];
	if (!(d[3] == 0.0f && d[7] == 0.0f && d[11] == 0.0f && d[15] == 1.0f)) return false;
	const float mag = fabsf(a00) + fabsf(a10) + fabsf(a20) + fabsf(a01) + fabsf(a11) + fabsf(a21) + fabsf(a02) + fabsf(a12) + fabsf(a22) +
	                  fabsf(a03) + fabsf(a13) + fabsf(a23);
	// An infinity or a NaN somewhere, or entries so large that a product the general routine forms (and then multiplies
	// by an exact zero) could overflow: the general routine propagates NaNs there, so it has to run.
	if (!(mag < 1.0e18f)) return false;
	const float s0 = sub(mul(a00, a11), mul(a10, a01));
	const float s1 = sub(mul(a00, a12), mul(a10, a02));
	const float s3 = sub(mul(a01, a12), mul(a11, a02));
	const float det = add(sub(mul(s0, a22), mul(s1, a21)), mul(s3, a20));
	if (det == 0.0f) {
#pragma unroll
		for (int i = 0; i < 9; ++i) inv3[i] = 0.0f;
		return true;
	}
	const float inv = __fdiv_rn(1.0f, det);
	inv3[0] = mul(sub(mul(a11, a22), mul(a12, a21)), inv);   // b00
	inv3[1] = mul(sub(mul(a02, a21), mul(a01, a22)), inv);   // b01
	inv3[2] = mul(s3, inv);                                  // b02
	inv3[3] = mul(sub(mul(a12, a20), mul(a10, a22)), inv);   // b10
	inv3[4] = mul(sub(mul(a00, a22), mul(a02, a20)), inv);   // b11
	inv3[5] = mul(sub(0.0f, s1), inv);                       // b12
	inv3[6] = mul(sub(mul(a10, a21), mul(a11, a20)), inv);   // b20
	inv3[7] = mul(sub(mul(a01, a20), mul(a00, a21)), inv);   // b21
	inv3[8] = mul(s0, inv);                                  // b22
	return true;
}

} // namespace zn
