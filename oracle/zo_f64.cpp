// ORACLE — test infrastructure, not product code (see zo_vec.hpp header).
//
// zo_f64.cpp: an INDEPENDENT float64 shadow of the scene path.  It shares no code with
// zo_math.hpp / zo_scene.hpp: plain row/column loops, libm sin/cos/tan, textbook formulas.  It
// exists because the reference pins no numbers for this path (PARITY UNPINNED):
// agreement between the float32 oracle and this file bounds the float32 oracle's own error and
// catches a convention mistake made once in zo_math.hpp.  It also defines the documented
// plane-boundary set of SURVEY.md §7 hard part 2:
//   an entity is BOUNDARY when, for some frustum plane, |margin| <= eps * (|n|.|c'| + |n|.e' + |d|)
// where margin = n.c' + d + |n|.e' evaluated here in float64.  Visible lists are required to
// match exactly on every non-boundary entity.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct M4 { double a[4][4]; };   // a[row][col]

M4 m_identity() { M4 m{}; for (int i = 0; i < 4; ++i) m.a[i][i] = 1.0; return m; }
M4 m_mul(const M4& x, const M4& y) {
	M4 r{};
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j) {
			double s = 0.0;
			for (int k = 0; k < 4; ++k) s += x.a[i][k] * y.a[k][j];
			r.a[i][j] = s;
		}
	return r;
}
M4 from_colmajor_f32(const float* d) { M4 m; for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) m.a[r][c] = d[c * 4 + r]; return m; }
void to_colmajor(const M4& m, double* d) { for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) d[c * 4 + r] = m.a[r][c]; }

M4 rot_x(double t) { M4 m = m_identity(); m.a[1][1] = std::cos(t); m.a[1][2] = -std::sin(t); m.a[2][1] = std::sin(t); m.a[2][2] = std::cos(t); return m; }
M4 rot_y(double t) { M4 m = m_identity(); m.a[0][0] = std::cos(t); m.a[0][2] = std::sin(t); m.a[2][0] = -std::sin(t); m.a[2][2] = std::cos(t); return m; }
M4 rot_z(double t) { M4 m = m_identity(); m.a[0][0] = std::cos(t); m.a[0][1] = -std::sin(t); m.a[1][0] = std::sin(t); m.a[1][1] = std::cos(t); return m; }
M4 translate(double x, double y, double z) { M4 m = m_identity(); m.a[0][3] = x; m.a[1][3] = y; m.a[2][3] = z; return m; }
M4 scale(double x, double y, double z) { M4 m = m_identity(); m.a[0][0] = x; m.a[1][1] = y; m.a[2][2] = z; return m; }

// local = T * Ry(yaw) * Rx(pitch) * Rz(roll) * S, by generic 4x4 products.
M4 local_trs(const float* t, const float* r, const float* s) {
	M4 R = m_mul(m_mul(rot_y(r[1]), rot_x(r[0])), rot_z(r[2]));
	return m_mul(m_mul(translate(t[0], t[1], t[2]), R), scale(s[0], s[1], s[2]));
}

} // namespace

extern "C" {

// float64 sin/cos of a float32 argument (the pin for zo::sincos).
void zo64_sincos(float x, double* s, double* c) { *s = std::sin(double(x)); *c = std::cos(double(x)); }

void zo64_from_euler(float pitch, float yaw, float roll, double* out16) {
	to_colmajor(m_mul(m_mul(rot_y(yaw), rot_x(pitch)), rot_z(roll)), out16);
}

void zo64_local_matrix(const float* t, const float* r, const float* s, double* out16) {
	to_colmajor(local_trs(t, r, s), out16);
}

// Scene frame in float64.  Outputs (any may be null): world64/mvp64 [E][16] column-major,
// visible64 [E] (1 = inside or touching), boundary [E] (1 = within eps of some plane).
void zo64_scene_update(uint32_t entity_count,
                       const float* position, const float* rotation, const float* scale_,
                       const float* aabb_center, const float* aabb_half, const uint32_t* parent,
                       const float* view16, const float* proj16, double eps,
                       double* world64, double* mvp64, uint8_t* visible64, uint8_t* boundary) {
	const M4 vp = m_mul(from_colmajor_f32(proj16), from_colmajor_f32(view16));
	// planes from the rows of P*V (Gribb-Hartmann, depth 0..1)
	double pl[6][4];
	for (int k = 0; k < 4; ++k) {
		pl[0][k] = vp.a[3][k] + vp.a[0][k];
		pl[1][k] = vp.a[3][k] - vp.a[0][k];
		pl[2][k] = vp.a[3][k] + vp.a[1][k];
		pl[3][k] = vp.a[3][k] - vp.a[1][k];
		pl[4][k] = vp.a[2][k];
		pl[5][k] = vp.a[3][k] - vp.a[2][k];
	}
	std::vector<M4> own;
	M4* W = nullptr;
	own.resize(entity_count);
	W = own.data();
	for (uint32_t i = 0; i < entity_count; ++i) {
		const M4 local = local_trs(position + 3 * size_t(i), rotation + 3 * size_t(i), scale_ + 3 * size_t(i));
		const uint32_t p = parent ? parent[i] : 0xFFFFFFFFu;
		W[i] = (p == 0xFFFFFFFFu) ? local : m_mul(W[p], local);
		if (world64) to_colmajor(W[i], world64 + 16 * size_t(i));
		if (mvp64) to_colmajor(m_mul(vp, W[i]), mvp64 + 16 * size_t(i));
		if (!visible64 && !boundary) continue;
		const float* c = aabb_center + 3 * size_t(i);
		const float* e = aabb_half + 3 * size_t(i);
		double cw[3], ew[3];
		for (int r = 0; r < 3; ++r) {
			cw[r] = W[i].a[r][0] * c[0] + W[i].a[r][1] * c[1] + W[i].a[r][2] * c[2] + W[i].a[r][3];
			ew[r] = std::fabs(W[i].a[r][0]) * e[0] + std::fabs(W[i].a[r][1]) * e[1] + std::fabs(W[i].a[r][2]) * e[2];
		}
		bool vis = true, bnd = false;
		for (int k = 0; k < 6; ++k) {
			const double dist = pl[k][0] * cw[0] + pl[k][1] * cw[1] + pl[k][2] * cw[2] + pl[k][3];
			const double rad = std::fabs(pl[k][0]) * ew[0] + std::fabs(pl[k][1]) * ew[1] + std::fabs(pl[k][2]) * ew[2];
			const double mag = std::fabs(pl[k][0]) * std::fabs(cw[0]) + std::fabs(pl[k][1]) * std::fabs(cw[1]) +
			                   std::fabs(pl[k][2]) * std::fabs(cw[2]) + rad + std::fabs(pl[k][3]);
			const double margin = dist + rad;
			if (!(margin >= 0.0)) vis = false;
			if (std::fabs(margin) <= eps * mag) bnd = true;
		}
		if (visible64) visible64[i] = vis ? 1 : 0;
		if (boundary) boundary[i] = bnd ? 1 : 0;
	}
}

// Mesh batch in float64: pos' = M*(p,1); nrm' = normalise( (M^-1)^T * (n,0) ), the inverse taken
// of the upper 3x3 by cofactors (valid because the bottom row of a model matrix is 0,0,0,1).
void zo64_mesh_transform(uint32_t instance_count, const uint32_t* first_vertex /*[I+1]*/,
                         const float* model /*[I][16]*/, const float* pos_in, const float* nrm_in,
                         double* pos_out, double* nrm_out) {
	for (uint32_t k = 0; k < instance_count; ++k) {
		const M4 m = from_colmajor_f32(model + 16 * size_t(k));
		const double (*a)[4] = m.a;
		double cof[3][3];
		cof[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
		cof[0][1] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
		cof[0][2] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
		cof[1][0] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
		cof[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
		cof[1][2] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
		cof[2][0] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
		cof[2][1] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
		cof[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
		const double det = a[0][0] * cof[0][0] + a[0][1] * cof[0][1] + a[0][2] * cof[0][2];
		for (uint32_t v = first_vertex[k]; v < first_vertex[k + 1]; ++v) {
			const float* p = pos_in + 3 * size_t(v);
			const float* n = nrm_in + 3 * size_t(v);
			for (int r = 0; r < 3; ++r)
				pos_out[3 * size_t(v) + r] = a[r][0] * p[0] + a[r][1] * p[1] + a[r][2] * p[2] + a[r][3];
			double q[3];
			for (int r = 0; r < 3; ++r) q[r] = (cof[r][0] * n[0] + cof[r][1] * n[1] + cof[r][2] * n[2]) / det;
			const double len = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
			for (int r = 0; r < 3; ++r) nrm_out[3 * size_t(v) + r] = len > 0.0 ? q[r] / len : q[r];
		}
	}
}

} // extern "C"
