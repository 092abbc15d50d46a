// ORACLE — test infrastructure, not product code (see zo_vec.hpp header).
//
// zo_synth.hpp: the synthetic-scene generator of SURVEY.md §8(d): a stateless integer hash of
// (seed, entity index, field id) mapped to float by (u32 >> 8) * 2^-24, so host and device
// produce identical bits without a transfer.  The device twin lives in
// zenyth_b200/csrc/zn_synth.cu; tests/test_synth.py checks they agree bit for bit.
#pragma once
#include <cmath>
#include <cstdint>

namespace zo {

inline constexpr uint32_t kSynthBaseSeed = 0x5A454E59u;   // "ZENY"

enum SynthField : uint32_t {
	F_TX = 0, F_TY, F_TZ, F_RX, F_RY, F_RZ, F_SX, F_SY, F_SZ,
	F_CX, F_CY, F_CZ, F_HX, F_HY, F_HZ,
};

inline uint32_t synth_hash(uint32_t seed, uint32_t index, uint32_t field) noexcept {
	uint32_t x = seed ^ (index * 0x9E3779B1u) ^ (field * 0x85EBCA77u);
	x ^= x >> 16; x *= 0x7FEB352Du;
	x ^= x >> 15; x *= 0x846CA68Bu;
	x ^= x >> 16;
	return x;
}
inline float synth_u01(uint32_t seed, uint32_t index, uint32_t field) noexcept {
	return static_cast<float>(synth_hash(seed, index, field) >> 8) * 5.9604644775390625e-08f;  // 2^-24
}
inline float synth_range(uint32_t seed, uint32_t index, uint32_t field, float lo, float span) noexcept {
	return lo + synth_u01(seed, index, field) * span;
}

// Ranges of SURVEY.md §8(d) configs 2/3/5.  Roots: T in [-1000,1000]^3, S in [0.5,2]^3.
// Children: T in [-10,10]^3, S in [0.9,1.1]^3.  Euler in [-pi,pi)^3.  AABB half-extent in
// [0.1,1]^3; AABB centre in [-center_range, center_range]^3 (0 for the bench configs).
inline float synth_entity_field(uint32_t seed, uint32_t i, uint32_t field, bool is_root, float center_range) noexcept {
	switch (field) {
	case F_TX: case F_TY: case F_TZ:
		return is_root ? synth_range(seed, i, field, -1000.0f, 2000.0f) : synth_range(seed, i, field, -10.0f, 20.0f);
	case F_RX: case F_RY: case F_RZ:
		return synth_range(seed, i, field, -3.14159274f, 6.28318548f);
	case F_SX: case F_SY: case F_SZ:
		return is_root ? synth_range(seed, i, field, 0.5f, 1.5f) : synth_range(seed, i, field, 0.9f, 0.2f);
	case F_CX: case F_CY: case F_CZ:
		return synth_range(seed, i, field, 0.0f - center_range, center_range * 2.0f);
	default:
		return synth_range(seed, i, field, 0.1f, 0.9f);
	}
}

// Parent of local index k in level l (l >= 1): children of a parent are contiguous and sorted.
//   fanout >= 1: parent = prev_off + min(k / fanout, prev_size - 1)
inline uint32_t synth_parent(uint32_t prev_off, uint32_t prev_size, uint32_t k, uint32_t fanout) noexcept {
	uint32_t p = k / fanout;
	if (p >= prev_size) p = prev_size - 1;
	return prev_off + p;
}

// Mesh vertices: position uniform [-1,1]^3 (fields 0..2), normal = normalised uniform
// [-1,1]^3 (fields 3..5); a zero-length draw maps to (0,0,1).
inline void synth_vertex(uint32_t seed, uint32_t v, float pos[3], float nrm[3]) noexcept {
	for (uint32_t k = 0; k < 3; ++k) pos[k] = synth_range(seed, v, k, -1.0f, 2.0f);
	const float x = synth_range(seed, v, 3, -1.0f, 2.0f);
	const float y = synth_range(seed, v, 4, -1.0f, 2.0f);
	const float z = synth_range(seed, v, 5, -1.0f, 2.0f);
	const float len = ::sqrtf((x * x + y * y) + z * z);
	if (len > 0.0f) { nrm[0] = x / len; nrm[1] = y / len; nrm[2] = z / len; }
	else { nrm[0] = 0.0f; nrm[1] = 0.0f; nrm[2] = 1.0f; }
}

} // namespace zo
