// zn_synth.cu — synthetic scene / mesh inputs generated in place on the device (bench and test
// support): a stateless integer hash of (seed, index, field) mapped to float by
// (u32 >> 8) * 2^-24.  Integer arithmetic plus exactly-rounded float mul/add/div/sqrt, so the bits
// equal the host generator's (tests/test_synth.py).
#include "zn_internal.hpp"
#include "zn_math.cuh"

namespace zn {

__device__ __forceinline__ uint32_t synth_hash(uint32_t seed, uint32_t index, uint32_t field) {
	uint32_t x = seed ^ (index * 0x9E3779B1u) ^ (field * 0x85EBCA77u);
	x ^= x >> 16; x *= 0x7FEB352Du;
	x ^= x >> 15; x *= 0x846CA68Bu;
	x ^= x >> 16;
	return x;
}
__device__ __forceinline__ float synth_range(uint32_t seed, uint32_t index, uint32_t field, float lo, float span) {
	const float u = mul(__uint2float_rn(synth_hash(seed, index, field) >> 8), 5.9604644775390625e-08f);
	return add(lo, mul(u, span));
}

__global__ void synth_level_kernel(uint32_t* __restrict__ blocks, uint32_t tile_base, uint32_t seed,
                                   uint32_t begin, uint32_t end, uint32_t prev_off, uint32_t prev_size, uint32_t fanout,
                                   int is_root, float center_range) {
	const uint32_t rel = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t i = begin + rel;
	if (i >= end) return;
	// tile-major input blocks: row k of this entity's tile, lane rel % kTile
	float* soa = reinterpret_cast<float*>(blocks + size_t(tile_base + rel / kTile) * kBlockWords) + (rel % kTile) - i;
	const size_t pitch = kTile;
	uint32_t* parent = reinterpret_cast<uint32_t*>(soa) + 15 * pitch;
	const float t_lo = is_root ? -1000.0f : -10.0f, t_span = is_root ? 2000.0f : 20.0f;
	const float s_lo = is_root ? 0.5f : 0.9f, s_span = is_root ? 1.5f : 0.2f;
#pragma unroll
	for (uint32_t k = 0; k < 3; ++k) {
		soa[(0 + k) * pitch + i] = synth_range(seed, i, 0 + k, t_lo, t_span);
		soa[(3 + k) * pitch + i] = synth_range(seed, i, 3 + k, -3.14159274f, 6.28318548f);
		soa[(6 + k) * pitch + i] = synth_range(seed, i, 6 + k, s_lo, s_span);
		soa[(9 + k) * pitch + i] = synth_range(seed, i, 9 + k, sub(0.0f, center_range), mul(center_range, 2.0f));
		soa[(12 + k) * pitch + i] = synth_range(seed, i, 12 + k, 0.1f, 0.9f);
	}
	uint32_t p = ZN_NO_PARENT;
	if (!is_root) {
		p = (i - begin) / fanout;
		if (p >= prev_size) p = prev_size - 1;
		p += prev_off;
	}
	parent[i] = p;
}

__global__ void synth_mesh_kernel(float* __restrict__ pos, float* __restrict__ nrm, uint32_t seed, uint32_t count) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= count) return;
#pragma unroll
	for (uint32_t k = 0; k < 3; ++k) pos[size_t(v) * 3 + k] = synth_range(seed, v, k, -1.0f, 2.0f);
	const float x = synth_range(seed, v, 3, -1.0f, 2.0f);
	const float y = synth_range(seed, v, 4, -1.0f, 2.0f);
	const float z = synth_range(seed, v, 5, -1.0f, 2.0f);
	const float len = __fsqrt_rn(add(add(mul(x, x), mul(y, y)), mul(z, z)));
	float nx = 0.0f, ny = 0.0f, nz = 1.0f;
	if (len > 0.0f) { nx = __fdiv_rn(x, len); ny = __fdiv_rn(y, len); nz = __fdiv_rn(z, len); }
	nrm[size_t(v) * 3 + 0] = nx; nrm[size_t(v) * 3 + 1] = ny; nrm[size_t(v) * 3 + 2] = nz;
}

} // namespace zn

using namespace zn;

extern "C" {

int zn_scene_fill_synthetic(zn_scene* s, uint32_t seed, uint32_t fanout, float aabb_center_range) {
	ZN_REQUIRE(s, "zn_scene_fill_synthetic: scene is NULL");
	ZN_REQUIRE(fanout >= 1, "zn_scene_fill_synthetic: fanout must be >= 1");
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	for (size_t l = 0; l < s->levels.size(); ++l) {
		const zn_level& lv = s->levels[l];
		const uint32_t prev_off = l ? s->levels[l - 1].begin : 0;
		const uint32_t prev_size = l ? s->levels[l - 1].end - s->levels[l - 1].begin : 0;
		const uint32_t n = lv.end - lv.begin;
		synth_level_kernel<<<(n + 255) / 256, 256, 0, s->ctx->stream>>>(s->blocks, lv.tile_base, seed, lv.begin, lv.end,
		                                                               prev_off, prev_size, fanout, l == 0 ? 1 : 0, aabb_center_range);
	}
	s->ranges_dirty = true;
	s->inputs_touched = true;
	ZN_CUDA(cudaGetLastError());
	ZN_CUDA(cudaStreamSynchronize(s->ctx->stream));
	return ZN_OK;
}

int zn_mesh_fill_synthetic(zn_mesh* m, uint32_t seed) {
	ZN_REQUIRE(m, "zn_mesh_fill_synthetic: mesh is NULL");
	ZN_CUDA(cudaSetDevice(m->ctx->device));
	synth_mesh_kernel<<<(m->vertex_count + 255) / 256, 256, 0, m->ctx->stream>>>(m->pos_in, m->nrm_in, seed, m->vertex_count);
	ZN_CUDA(cudaGetLastError());
	ZN_CUDA(cudaStreamSynchronize(m->ctx->stream));
	return ZN_OK;
}

} // extern "C"
