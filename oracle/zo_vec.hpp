// ORACLE — test infrastructure, not product code.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, link or call anything in oracle/.
//
// zo_vec.hpp: the oracle's vector/matrix storage types, in two interchangeable backends.
//
//   default            : a scalar RESTATEMENT of the reference's vec3 / vec4 / mat4 storage and
//                        arithmetic (this file), lane for lane, so the oracle can be built on the
//                        GPU box where /root/reference does not exist.
//   -DZO_REFERENCE_TYPES: the reference's OWN classes (zenyth::math::vec3/vec4/mat4), compiled
//                        from /root/reference/Core/src/math/{vector,matrix}.cpp behind
//                        oracle/shim/pch.hpp.  Built by oracle/Makefile into oracle/_ref/.
//
// tests/test_oracle_ref.py checks that both backends give bit-identical results, which pins
// the restatement to the reference's real SSE code (Core/src/math/vector.cpp:41-128).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

#ifdef ZO_REFERENCE_TYPES
#include "pch.hpp"            // oracle/shim/pch.hpp (must precede Core/include on -I)
#include "math/vector.hpp"    // /root/reference/Core/include/math/vector.hpp
#include "math/matrix.hpp"    // /root/reference/Core/include/math/matrix.hpp
namespace zo {
using vec3 = zenyth::math::vec3;
using vec4 = zenyth::math::vec4;
using mat4 = zenyth::math::mat4;
}
#else
namespace zo {

// Restates zenyth::math::vec3 (Core/include/math/vector.hpp:35-73): four float lanes, the
// fourth ("w") carried through every operation exactly as the __m128 does
// (vector.cpp:42,46,50 zero it in every public constructor).
class vec3 {
public:
	explicit vec3(float s) : m_x(s), m_y(s), m_z(s), m_w(0.0f) {}             // vector.cpp:45-47
	vec3(float x, float y, float z) : m_x(x), m_y(y), m_z(z), m_w(0.0f) {}    // vector.cpp:41-43
	vec3() : m_x(0), m_y(0), m_z(0), m_w(0) {}                                 // vector.cpp:49-51

	// NOTE: the reference's `vec3::y() const` returns m_z (vector.cpp:54) — a defect that is
	// recorded, not reproduced.  Oracle code never reads y through a const reference of the
	// reference type; see zo::Y() below.
	[[nodiscard]] float x() const noexcept { return m_x; }
	[[nodiscard]] float y() const noexcept { return m_y; }
	[[nodiscard]] float z() const noexcept { return m_z; }
	float& x() noexcept { return m_x; }
	float& y() noexcept { return m_y; }
	float& z() noexcept { return m_z; }

	// vector.cpp:61-64 — one _mm_{add,sub,mul,div}_ps over all four lanes.
	[[nodiscard]] vec3 operator+(const vec3& o) const noexcept { return lanes(m_x + o.m_x, m_y + o.m_y, m_z + o.m_z, m_w + o.m_w); }
	[[nodiscard]] vec3 operator-(const vec3& o) const noexcept { return lanes(m_x - o.m_x, m_y - o.m_y, m_z - o.m_z, m_w - o.m_w); }
	[[nodiscard]] vec3 operator*(float s) const noexcept { return lanes(m_x * s, m_y * s, m_z * s, m_w * s); }
	[[nodiscard]] vec3 operator/(float s) const noexcept { return lanes(m_x / s, m_y / s, m_z / s, m_w / s); }
	void operator+=(const vec3& o) noexcept { *this = *this + o; }
	void operator-=(const vec3& o) noexcept { *this = *this - o; }
	void operator*=(float s) noexcept { *this = *this * s; }
	void operator/=(float s) noexcept { *this = *this / s; }
	// vector.cpp:71 — 0 - v, so +0 lanes stay +0 (not -0).
	[[nodiscard]] vec3 operator-() const noexcept { return lanes(0.0f - m_x, 0.0f - m_y, 0.0f - m_z, 0.0f - m_w); }

	// vector.cpp:73-86 — left = a.yzx * b.zxy, right = a.zxy * b.yzx, left - right.
	[[nodiscard]] vec3 cross(const vec3& o) const noexcept {
		return lanes(m_y * o.m_z - m_z * o.m_y,
		             m_z * o.m_x - m_x * o.m_z,
		             m_x * o.m_y - m_y * o.m_x,
		             m_w * o.m_w - m_w * o.m_w);
	}
	// vector.cpp:88 — DPPS with mask 0xF1: all four products, summed (p0+p1)+(p2+p3).
	[[nodiscard]] float dot(const vec3& o) const noexcept {
		return (m_x * o.m_x + m_y * o.m_y) + (m_z * o.m_z + m_w * o.m_w);
	}
	[[nodiscard]] float length_sq() const noexcept { return dot(*this); }       // vector.cpp:89
	[[nodiscard]] float length() const noexcept { return ::sqrtf(length_sq()); } // vector.cpp:90

private:
	static vec3 lanes(float x, float y, float z, float w) { vec3 r; r.m_x = x; r.m_y = y; r.m_z = z; r.m_w = w; return r; }
	alignas(16) float m_x;
	float m_y, m_z, m_w;
};

// Restates zenyth::math::vec4 (Core/include/math/vector.hpp:75-114, vector.cpp:94-127).
class vec4 {
public:
	explicit vec4(float s) : m_x(s), m_y(s), m_z(s), m_w(s) {}
	vec4(float x, float y, float z, float w) : m_x(x), m_y(y), m_z(z), m_w(w) {}
	vec4() : m_x(0), m_y(0), m_z(0), m_w(0) {}

	[[nodiscard]] float x() const noexcept { return m_x; }
	[[nodiscard]] float y() const noexcept { return m_y; }
	[[nodiscard]] float z() const noexcept { return m_z; }
	[[nodiscard]] float w() const noexcept { return m_w; }
	float& x() noexcept { return m_x; }
	float& y() noexcept { return m_y; }
	float& z() noexcept { return m_z; }
	float& w() noexcept { return m_w; }

	[[nodiscard]] vec4 operator+(const vec4& o) const noexcept { return {m_x + o.m_x, m_y + o.m_y, m_z + o.m_z, m_w + o.m_w}; }
	[[nodiscard]] vec4 operator-(const vec4& o) const noexcept { return {m_x - o.m_x, m_y - o.m_y, m_z - o.m_z, m_w - o.m_w}; }
	[[nodiscard]] vec4 operator*(float s) const noexcept { return {m_x * s, m_y * s, m_z * s, m_w * s}; }
	[[nodiscard]] vec4 operator/(float s) const noexcept { return {m_x / s, m_y / s, m_z / s, m_w / s}; }
	void operator+=(const vec4& o) noexcept { *this = *this + o; }
	void operator-=(const vec4& o) noexcept { *this = *this - o; }
	void operator*=(float s) noexcept { *this = *this * s; }
	void operator/=(float s) noexcept { *this = *this / s; }
	[[nodiscard]] vec4 operator-() const noexcept { return {0.0f - m_x, 0.0f - m_y, 0.0f - m_z, 0.0f - m_w}; }

	// vector.cpp:125 — DPPS 0xF1: (p0+p1)+(p2+p3).
	[[nodiscard]] float dot(const vec4& o) const noexcept {
		return (m_x * o.m_x + m_y * o.m_y) + (m_z * o.m_z + m_w * o.m_w);
	}
	[[nodiscard]] float length_sq() const noexcept { return dot(*this); }
	[[nodiscard]] float length() const noexcept { return ::sqrtf(length_sq()); }

private:
	alignas(16) float m_x;
	float m_y, m_z, m_w;
};

// Restates the WORKING part of zenyth::math::mat4 (Core/include/math/matrix.hpp:7-8,13,57-61,
// 66-69; Core/src/math/matrix.cpp:8-9): column-major storage m_data[c*4+r], the four-column
// constructor, the (r,c) subscript and data().  Everything else in matrix.cpp is a stub
// (matrix.cpp:5-121) and is completed as free functions in zo_math.hpp.
class mat4 {
public:
	mat4() noexcept : m_data{} {}                                              // matrix.hpp:67 zero-init
	mat4(const vec4& c0, const vec4& c1, const vec4& c2, const vec4& c3) noexcept
		: m_data{c0.x(), c0.y(), c0.z(), c0.w(), c1.x(), c1.y(), c1.z(), c1.w(),
		         c2.x(), c2.y(), c2.z(), c2.w(), c3.x(), c3.y(), c3.z(), c3.w()} {}
	[[nodiscard]] float operator[](const std::size_t r, const std::size_t c) const noexcept { return m_data[c * 4 + r]; }
	[[nodiscard]] float& operator[](const std::size_t r, const std::size_t c) noexcept { return m_data[c * 4 + r]; }
	[[nodiscard]] const void* data() const noexcept { return m_data; }
	[[nodiscard]] void* data() noexcept { return m_data; }

private:
	alignas(16) float m_data[16];
};

} // namespace zo
#endif // ZO_REFERENCE_TYPES

namespace zo {

static_assert(sizeof(vec3) == 16 && sizeof(vec4) == 16 && sizeof(mat4) == 64, "reference layout (SURVEY §2)");

// Component reads that are safe on BOTH backends.  The reference's `vec3::y() const` returns z
// (Core/src/math/vector.cpp:54); its non-const overload is correct (vector.cpp:58), so copy
// first.
inline float X(const vec3& v) noexcept { return v.x(); }
inline float Y(const vec3& v) noexcept { vec3 t = v; return t.y(); }
inline float Z(const vec3& v) noexcept { return v.z(); }

inline const float* raw(const mat4& m) noexcept { return static_cast<const float*>(m.data()); }
inline float* raw(mat4& m) noexcept { return static_cast<float*>(m.data()); }
inline vec4 column(const mat4& m, int j) noexcept { const float* d = raw(m) + 4 * j; return vec4(d[0], d[1], d[2], d[3]); }

} // namespace zo
