// zn_scene.cu — the per-frame scene update on sm_100a:
//   K1/K2  scene_level_kernel<HAS_PARENT>: TRS -> local -> world (= parent.world * local) -> MVP,
//          AABB-vs-frustum flag, per-tile visible count + bit mask.  One launch per hierarchy level.
//   K3a    scan_tiles_kernel: exclusive scan of the per-tile counts (single CTA).
//   K3b    emit_visible_kernel: ordered visible list from masks + offsets (warp per tile).
//
// Data layout in HBM (DESIGN.md "Layout"): inputs are 15 scalar SoA arrays (tx,ty,tz, pitch,yaw,
// roll, sx,sy,sz, aabb centre xyz, aabb half xyz) + parent[u32], so every warp load is one fully
// coalesced 128-byte line; outputs are AoS float[E][16] (what mat4::data() consumers expect,
// Core/include/math/matrix.hpp:60-61).  A thread owns one entity; its two 64-byte matrices go to
// shared memory in the TMA 64-byte-swizzle pattern (conflict-free 16-byte stores) and leave the SM
// as two cp.async.bulk.tensor stores per 256-entity tile, so HBM sees 16 KiB contiguous writes and
// no thread ever issues a strided store.  Parent matrices of a tile are one contiguous range
// (children are sorted by parent) fetched with a single cp.async.bulk into shared memory while
// the threads are busy with sincos.
#include "zn_internal.hpp"
#include "zn_math.cuh"
#include "zn_ptx.cuh"

#include <algorithm>
#include <cstring>

namespace zn {

struct SceneIn {
	const float* comp[15];
	const uint32_t* parent;
};
struct FrameConst {
	float vp[16];
	float planes[24];
};

constexpr int kWarpsPerTile = kTile / 32;
constexpr uint32_t kTileBytes = kTile * 64;   // one matrix tile

template <bool HAS_PARENT>
constexpr uint32_t level_smem_bytes() { return (HAS_PARENT ? 3u : 2u) * kTileBytes + 1024u; }

template <bool HAS_PARENT>
__global__ void __launch_bounds__(kTile, 4)
scene_level_kernel(const __grid_constant__ CUtensorMap tm_world, const __grid_constant__ CUtensorMap tm_mvp,
                   const __grid_constant__ SceneIn in, const __grid_constant__ FrameConst fc,
                   const float* __restrict__ world_rd, uint32_t begin, uint32_t end, uint32_t tile_base,
                   uint32_t* __restrict__ tile_vis_count, uint32_t* __restrict__ tile_mask) {
	extern __shared__ uint8_t smem_raw[];
	__shared__ uint64_t s_bar;
	__shared__ uint32_t s_cnt[kWarpsPerTile];
	__shared__ uint32_t s_pmin[kWarpsPerTile], s_pmax[kWarpsPerTile];

	// 1024-byte aligned tiles (the swizzle pattern is a function of the shared address bits).
	uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	float4* world_tile = reinterpret_cast<float4*>(smem);
	float4* mvp_tile = reinterpret_cast<float4*>(smem + kTileBytes);
	float4* parent_tile = reinterpret_cast<float4*>(smem + 2 * kTileBytes);

	const int t = threadIdx.x;
	const int lane = t & 31;
	const int warp = t >> 5;
	const uint32_t tile_first = begin + blockIdx.x * kTile;
	const uint32_t i = tile_first + t;
	const bool valid = i < end;
	const uint32_t ic = valid ? i : end - 1;   // clamp: tail threads redo the last entity, rows >= end are clipped by TMA

	if (t == 0) {
		prefetch_tensormap(&tm_world);
		prefetch_tensormap(&tm_mvp);
	}

	// ---- all global loads first (16 independent requests in flight per thread) ------------------
	float v[15];
#pragma unroll
	for (int k = 0; k < 15; ++k) v[k] = ld_stream(in.comp[k] + ic);

	// ---- parent range of this tile -> one bulk copy into shared memory -------------------------
	uint32_t p = ZN_NO_PARENT;
	uint32_t pmin = 0;
	bool staged = false;
	if (HAS_PARENT) {
		p = ld_stream(in.parent + ic);
		const uint32_t lo = __reduce_min_sync(0xffffffffu, p);                               // NO_PARENT is UINT_MAX
		const uint32_t hi = __reduce_max_sync(0xffffffffu, p == ZN_NO_PARENT ? 0u : p);
		if (lane == 0) { s_pmin[warp] = lo; s_pmax[warp] = hi; }
		if (t == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
		__syncthreads();
		uint32_t pmax = 0;
		pmin = ZN_NO_PARENT;
#pragma unroll
		for (int k = 0; k < kWarpsPerTile; ++k) { pmin = min(pmin, s_pmin[k]); pmax = max(pmax, s_pmax[k]); }
		staged = (pmin != ZN_NO_PARENT) && (pmax - pmin < uint32_t(kTile));
		if (staged && t == 0) {
			const uint32_t bytes = (pmax - pmin + 1u) * 64u;
			mbar_arrive_expect_tx(&s_bar, bytes);
			bulk_g2s(parent_tile, world_rd + size_t(pmin) * 16, bytes, &s_bar);
		}
	}

	// ---- local TRS ------------------------------------------------------------------------------
	const Affine local = local_trs(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);

	// ---- world = parent.world * local -------------------------------------------------------------
	Affine w = local;
	if (HAS_PARENT && p != ZN_NO_PARENT) {
		float4 c0, c1, c2, c3;
		if (staged) {
			mbar_wait(&s_bar, 0);
			const float4* src = parent_tile + size_t(p - pmin) * 4;
			c0 = src[0]; c1 = src[1]; c2 = src[2]; c3 = src[3];
		} else {
			const float4* src = reinterpret_cast<const float4*>(world_rd) + size_t(p) * 4;
			c0 = __ldg(src); c1 = __ldg(src + 1); c2 = __ldg(src + 2); c3 = __ldg(src + 3);
		}
		Affine P;
		P.m[0][0] = c0.x; P.m[1][0] = c0.y; P.m[2][0] = c0.z;
		P.m[0][1] = c1.x; P.m[1][1] = c1.y; P.m[2][1] = c1.z;
		P.m[0][2] = c2.x; P.m[1][2] = c2.y; P.m[2][2] = c2.z;
		P.m[0][3] = c3.x; P.m[1][3] = c3.y; P.m[2][3] = c3.z;
		w = affine_mul(P, local);
	}

	// ---- matrices -> swizzled shared tiles (chunk c of row t lives at chunk c ^ ((t>>1)&3)) -------
	{
		const int sw = (t >> 1) & 3;
		float4* wr = world_tile + t * 4;
		wr[0 ^ sw] = make_float4(w.m[0][0], w.m[1][0], w.m[2][0], 0.0f);
		wr[1 ^ sw] = make_float4(w.m[0][1], w.m[1][1], w.m[2][1], 0.0f);
		wr[2 ^ sw] = make_float4(w.m[0][2], w.m[1][2], w.m[2][2], 0.0f);
		wr[3 ^ sw] = make_float4(w.m[0][3], w.m[1][3], w.m[2][3], 1.0f);
		float m[16];
		mvp_mul(fc.vp, w, m);
		float4* mr = mvp_tile + t * 4;
		mr[0 ^ sw] = make_float4(m[0], m[1], m[2], m[3]);
		mr[1 ^ sw] = make_float4(m[4], m[5], m[6], m[7]);
		mr[2 ^ sw] = make_float4(m[8], m[9], m[10], m[11]);
		mr[3 ^ sw] = make_float4(m[12], m[13], m[14], m[15]);
	}
	fence_proxy_async_smem();

	// ---- cull -------------------------------------------------------------------------------------
	const bool vis = valid && aabb_visible(w, v[9], v[10], v[11], v[12], v[13], v[14], fc.planes);
	const uint32_t mask = __ballot_sync(0xffffffffu, vis);
	const uint32_t tile = tile_base + blockIdx.x;
	if (lane == 0) {
		s_cnt[warp] = __popc(mask);
		tile_mask[size_t(tile) * kWarpsPerTile + warp] = mask;
	}
	__syncthreads();

	if (t == 0) {
		tensor_store_2d(&tm_world, world_tile, 0, int32_t(tile_first));
		tensor_store_2d(&tm_mvp, mvp_tile, 0, int32_t(tile_first));
		bulk_commit();
		uint32_t c = 0;
#pragma unroll
		for (int k = 0; k < kWarpsPerTile; ++k) c += s_cnt[k];
		tile_vis_count[tile] = c;
		bulk_wait_read_all();
	}
}

// Exclusive scan of n per-tile counts; one CTA of 1024 threads, 4 items per thread per pass.
__global__ void __launch_bounds__(1024)
scan_tiles_kernel(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets, uint32_t n,
                  uint32_t* __restrict__ total) {
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_carry;
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	if (t == 0) s_carry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < n; base += 4096) {
		const uint32_t idx = base + t * 4;
		uint32_t c[4];
#pragma unroll
		for (int k = 0; k < 4; ++k) c[k] = (idx + k < n) ? counts[idx + k] : 0u;
		const uint32_t mine = c[0] + c[1] + c[2] + c[3];
		uint32_t incl = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d) incl += y;
		}
		if (lane == 31) s_warp[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint32_t ws = s_warp[lane];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t y = __shfl_up_sync(0xffffffffu, ws, d);
				if (lane >= d) ws += y;
			}
			s_warp[lane] = ws;   // inclusive over warps
		}
		__syncthreads();
		const uint32_t carry = s_carry;
		uint32_t excl = carry + (warp ? s_warp[warp - 1] : 0u) + incl - mine;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			if (idx + k < n) offsets[idx + k] = excl;
			excl += c[k];
		}
		__syncthreads();
		if (t == 1023) s_carry = carry + s_warp[31];
		__syncthreads();
	}
	if (t == 0) *total = s_carry;
}

// Ordered visible list: one warp per tile.
__global__ void __launch_bounds__(256)
emit_visible_kernel(const uint32_t* __restrict__ tile_mask, const uint32_t* __restrict__ tile_offset,
                    const uint32_t* __restrict__ tile_first_entity, uint32_t tiles, uint32_t index_base,
                    uint32_t* __restrict__ visible) {
	const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (tile >= tiles) return;
	const int lane = threadIdx.x & 31;
	const uint32_t word = (lane < kWarpsPerTile) ? tile_mask[size_t(tile) * kWarpsPerTile + lane] : 0u;
	const uint32_t cnt = __popc(word);
	uint32_t incl = cnt;
#pragma unroll
	for (int d = 1; d < kWarpsPerTile; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += y;
	}
	const uint32_t excl = incl - cnt;
	const uint32_t base = tile_offset[tile];
	const uint32_t first = tile_first_entity[tile] + index_base;
	const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
	for (int w = 0; w < kWarpsPerTile; ++w) {
		const uint32_t m = __shfl_sync(0xffffffffu, word, w);
		const uint32_t off = __shfl_sync(0xffffffffu, excl, w);
		if ((m >> lane) & 1u) visible[base + off + __popc(m & lt)] = first + w * 32 + lane;
	}
}

// Strided host records (3 floats every `stride_f` floats) -> three SoA component arrays.
__global__ void repack3_kernel(const float* __restrict__ src, uint32_t stride_f, uint32_t count,
                               float* __restrict__ dx, float* __restrict__ dy, float* __restrict__ dz) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const float* s = src + size_t(i) * stride_f;
	dx[i] = s[0]; dy[i] = s[1]; dz[i] = s[2];
}
__global__ void unpack3_kernel(const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sz,
                               uint32_t count, float* __restrict__ dst) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	dst[size_t(i) * 3 + 0] = sx[i]; dst[size_t(i) * 3 + 1] = sy[i]; dst[size_t(i) * 3 + 2] = sz[i];
}

static int ensure_staging(zn_ctx* ctx, size_t bytes) {
	if (ctx->staging_bytes >= bytes) return ZN_OK;
	ZN_CUDA(cudaStreamSynchronize(ctx->stream));
	if (ctx->staging) ZN_CUDA(cudaFree(ctx->staging));
	ctx->staging = nullptr;
	ctx->staging_bytes = 0;
	ZN_CUDA(cudaMalloc(&ctx->staging, bytes));
	ctx->staging_bytes = bytes;
	return ZN_OK;
}

// Upload `count` strided 3-float records into SoA components [k0, k0+3) at entity offset `first`.
static int upload3(zn_scene* s, int k0, uint32_t first, uint32_t count, const float* host, size_t stride_bytes) {
	if (!host || count == 0) return ZN_OK;
	zn_ctx* ctx = s->ctx;
	const uint32_t stride_f = uint32_t(stride_bytes / 4);
	const uint32_t chunk = 1u << 22;   // records per pass: <= 64 MiB of staging at stride 16
	int rc = ensure_staging(ctx, size_t(std::min(count, chunk)) * stride_bytes);
	if (rc != ZN_OK) return rc;
	for (uint32_t done = 0; done < count; done += chunk) {
		const uint32_t n = std::min(chunk, count - done);
		// the last record needs only 12 bytes, not a full stride
		const size_t bytes = size_t(n - 1) * stride_bytes + 12;
		ZN_CUDA(cudaMemcpyAsync(ctx->staging, reinterpret_cast<const uint8_t*>(host) + size_t(done) * stride_bytes, bytes,
		                        cudaMemcpyHostToDevice, ctx->stream));
		float* d0 = s->soa + size_t(k0) * s->soa_pitch + first + done;
		repack3_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(static_cast<const float*>(ctx->staging), stride_f, n, d0,
		                                                        d0 + s->soa_pitch, d0 + 2 * s->soa_pitch);
		ZN_CUDA(cudaGetLastError());
	}
	return ZN_OK;
}

static int download3(zn_scene* s, int k0, uint32_t first, uint32_t count, float* host) {
	if (!host || count == 0) return ZN_OK;
	zn_ctx* ctx = s->ctx;
	int rc = ensure_staging(ctx, size_t(count) * 12);
	if (rc != ZN_OK) return rc;
	const float* s0 = s->soa + size_t(k0) * s->soa_pitch + first;
	unpack3_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(s0, s0 + s->soa_pitch, s0 + 2 * s->soa_pitch, count,
	                                                            static_cast<float*>(ctx->staging));
	ZN_CUDA(cudaGetLastError());
	ZN_CUDA(cudaMemcpyAsync(host, ctx->staging, size_t(count) * 12, cudaMemcpyDeviceToHost, ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(ctx->stream));
	return ZN_OK;
}

} // namespace zn

using namespace zn;

extern "C" {

int zn_scene_create(zn_ctx* ctx, const zn_scene_desc* desc, zn_scene** out_scene) {
	ZN_REQUIRE(ctx && desc && out_scene, "zn_scene_create: NULL argument");
	*out_scene = nullptr;
	const uint32_t E = desc->entity_count;
	ZN_REQUIRE(E > 0 && E <= 0x7FFFFF00u, "zn_scene_create: entity_count %u out of range", E);
	std::vector<uint32_t> offs;
	if (desc->level_offsets) {
		ZN_REQUIRE(desc->level_count >= 1, "zn_scene_create: level_count must be >= 1");
		offs.assign(desc->level_offsets, desc->level_offsets + desc->level_count + 1);
		ZN_REQUIRE(offs.front() == 0 && offs.back() == E, "zn_scene_create: level_offsets must start at 0 and end at entity_count");
		for (size_t l = 0; l + 1 < offs.size(); ++l)
			ZN_REQUIRE(offs[l] < offs[l + 1], "zn_scene_create: level %zu is empty or level_offsets is not ascending", l);
	} else {
		offs = {0u, E};
	}
	ZN_CUDA(cudaSetDevice(ctx->device));
	zn_scene* s = new zn_scene();
	s->ctx = ctx;
	s->entity_count = E;
	s->soa_pitch = (size_t(E) + 63) & ~size_t(63);
	uint32_t tile_base = 0;
	std::vector<uint32_t> tile_first;
	for (size_t l = 0; l + 1 < offs.size(); ++l) {
		zn_level lv;
		lv.begin = offs[l];
		lv.end = offs[l + 1];
		lv.tiles = (lv.end - lv.begin + kTile - 1) / kTile;
		lv.tile_base = tile_base;
		for (uint32_t k = 0; k < lv.tiles; ++k) tile_first.push_back(lv.begin + k * kTile);
		tile_base += lv.tiles;
		s->levels.push_back(lv);
	}
	s->tile_count = tile_base;
	auto cleanup = [&](int rc) { zn_scene_destroy(s); return rc; };
#define ZN_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return cleanup(fail(e_ == cudaErrorMemoryAllocation ? ZN_ERR_NOMEM : ZN_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_))); } while (0)
	ZN_TRY(cudaMalloc(&s->soa, s->soa_pitch * 15 * sizeof(float)));
	ZN_TRY(cudaMalloc(&s->parent, size_t(E) * 4));
	ZN_TRY(cudaMalloc(&s->world, size_t(E) * 64));
	ZN_TRY(cudaMalloc(&s->mvp, size_t(E) * 64));
	ZN_TRY(cudaMalloc(&s->visible, size_t(E) * 4));
	ZN_TRY(cudaMalloc(&s->visible_count, 4));
	ZN_TRY(cudaMalloc(&s->tile_vis_count, size_t(s->tile_count) * 4));
	ZN_TRY(cudaMalloc(&s->tile_vis_offset, size_t(s->tile_count) * 4));
	ZN_TRY(cudaMalloc(&s->tile_mask, size_t(s->tile_count) * kWarpsPerTile * 4));
	ZN_TRY(cudaMalloc(&s->tile_first_entity, size_t(s->tile_count) * 4));
	ZN_TRY(cudaHostAlloc(&s->host_count, 4, cudaHostAllocDefault));
	ZN_TRY(cudaMemsetAsync(s->soa, 0, s->soa_pitch * 15 * sizeof(float), ctx->stream));
	ZN_TRY(cudaMemsetAsync(s->parent, 0xFF, size_t(E) * 4, ctx->stream));
	ZN_TRY(cudaMemsetAsync(s->visible_count, 0, 4, ctx->stream));
	ZN_TRY(cudaMemcpyAsync(s->tile_first_entity, tile_first.data(), tile_first.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	ZN_TRY(cudaStreamSynchronize(ctx->stream));
#undef ZN_TRY
	for (auto& lv : s->levels) {
		int rc = encode_matrix_tensor_map(&lv.tm_world, s->world, lv.end);
		if (rc == ZN_OK) rc = encode_matrix_tensor_map(&lv.tm_mvp, s->mvp, lv.end);
		if (rc != ZN_OK) return cleanup(rc);
	}
	static bool attr_set = false;
	if (!attr_set) {
		ZN_CUDA(cudaFuncSetAttribute(scene_level_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(level_smem_bytes<true>())));
		ZN_CUDA(cudaFuncSetAttribute(scene_level_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(level_smem_bytes<false>())));
		attr_set = true;
	}
	*out_scene = s;
	return ZN_OK;
}

int zn_scene_destroy(zn_scene* s) {
	if (!s) return ZN_OK;
	cudaSetDevice(s->ctx->device);
	cudaStreamSynchronize(s->ctx->stream);
	cudaFree(s->soa); cudaFree(s->parent); cudaFree(s->world); cudaFree(s->mvp); cudaFree(s->visible);
	cudaFree(s->visible_count); cudaFree(s->tile_vis_count); cudaFree(s->tile_vis_offset); cudaFree(s->tile_mask);
	cudaFree(s->tile_first_entity);
	if (s->host_count) cudaFreeHost(s->host_count);
	for (auto e : s->prof_events) cudaEventDestroy(e);
	delete s;
	return ZN_OK;
}

static int check_range(zn_scene* s, uint32_t first, uint32_t count, const char* who) {
	ZN_REQUIRE(s, "%s: scene is NULL", who);
	ZN_REQUIRE(uint64_t(first) + count <= s->entity_count, "%s: range [%u, %u+%u) exceeds entity_count %u", who, first, first, count, s->entity_count);
	return ZN_OK;
}
static int check_stride(size_t stride_bytes, const char* who) {
	ZN_REQUIRE(stride_bytes >= 12 && stride_bytes % 4 == 0, "%s: stride_bytes %zu must be >= 12 and a multiple of 4", who, stride_bytes);
	return ZN_OK;
}

int zn_scene_set_transforms(zn_scene* s, uint32_t first, uint32_t count, const float* position, const float* rotation,
                            const float* scale, size_t stride_bytes) {
	int rc = check_range(s, first, count, "zn_scene_set_transforms");
	if (rc == ZN_OK) rc = check_stride(stride_bytes, "zn_scene_set_transforms");
	if (rc != ZN_OK) return rc;
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	if ((rc = upload3(s, 0, first, count, position, stride_bytes)) != ZN_OK) return rc;
	if ((rc = upload3(s, 3, first, count, rotation, stride_bytes)) != ZN_OK) return rc;
	return upload3(s, 6, first, count, scale, stride_bytes);
}

int zn_scene_set_bounds(zn_scene* s, uint32_t first, uint32_t count, const float* center, const float* half_extent, size_t stride_bytes) {
	int rc = check_range(s, first, count, "zn_scene_set_bounds");
	if (rc == ZN_OK) rc = check_stride(stride_bytes, "zn_scene_set_bounds");
	if (rc != ZN_OK) return rc;
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	if ((rc = upload3(s, 9, first, count, center, stride_bytes)) != ZN_OK) return rc;
	return upload3(s, 12, first, count, half_extent, stride_bytes);
}

int zn_scene_set_parents(zn_scene* s, uint32_t first, uint32_t count, const uint32_t* parent) {
	int rc = check_range(s, first, count, "zn_scene_set_parents");
	if (rc != ZN_OK) return rc;
	ZN_REQUIRE(parent || count == 0, "zn_scene_set_parents: parent is NULL");
	// validate the level-order contract on the host: a parent must live in an earlier level
	size_t l = 0;
	for (uint32_t k = 0; k < count; ++k) {
		const uint32_t i = first + k;
		while (i >= s->levels[l].end) ++l;
		const uint32_t p = parent[k];
		ZN_REQUIRE(p == ZN_NO_PARENT || p < s->levels[l].begin,
		           "zn_scene_set_parents: entity %u (level %zu) has parent %u, which is not in an earlier level", i, l, p);
	}
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	ZN_CUDA(cudaMemcpyAsync(s->parent + first, parent, size_t(count) * 4, cudaMemcpyHostToDevice, s->ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(s->ctx->stream));
	return ZN_OK;
}

int zn_scene_set_camera(zn_scene* s, const float view[16], const float proj[16]) {
	ZN_REQUIRE(s && view && proj, "zn_scene_set_camera: NULL argument");
	zn_mat4_mul(proj, view, s->view_proj);
	zn_frustum_planes(s->view_proj, s->planes);
	s->has_camera = true;
	return ZN_OK;
}

int zn_scene_set_index_base(zn_scene* s, uint32_t index_base) {
	ZN_REQUIRE(s, "zn_scene_set_index_base: scene is NULL");
	s->index_base = index_base;
	return ZN_OK;
}

int zn_scene_update(zn_scene* s) {
	ZN_REQUIRE(s, "zn_scene_update: scene is NULL");
	if (!s->has_camera) return fail(ZN_ERR_STATE, "zn_scene_update: no camera set (call zn_scene_set_camera first)");
	zn_ctx* ctx = s->ctx;
	ZN_CUDA(cudaSetDevice(ctx->device));
	SceneIn in;
	for (int k = 0; k < 15; ++k) in.comp[k] = s->soa + size_t(k) * s->soa_pitch;
	in.parent = s->parent;
	FrameConst fc;
	std::memcpy(fc.vp, s->view_proj, sizeof(fc.vp));
	std::memcpy(fc.planes, s->planes, sizeof(fc.planes));
	const bool prof = s->profiling;
	for (size_t l = 0; l < s->levels.size(); ++l) {
		const zn_level& lv = s->levels[l];
		if (prof) ZN_CUDA(cudaEventRecord(s->prof_events[l], ctx->stream));
		if (l == 0)
			scene_level_kernel<false><<<lv.tiles, kTile, level_smem_bytes<false>(), ctx->stream>>>(
				lv.tm_world, lv.tm_mvp, in, fc, s->world, lv.begin, lv.end, lv.tile_base, s->tile_vis_count, s->tile_mask);
		else
			scene_level_kernel<true><<<lv.tiles, kTile, level_smem_bytes<true>(), ctx->stream>>>(
				lv.tm_world, lv.tm_mvp, in, fc, s->world, lv.begin, lv.end, lv.tile_base, s->tile_vis_count, s->tile_mask);
	}
	if (prof) ZN_CUDA(cudaEventRecord(s->prof_events[s->levels.size()], ctx->stream));
	scan_tiles_kernel<<<1, 1024, 0, ctx->stream>>>(s->tile_vis_count, s->tile_vis_offset, s->tile_count, s->visible_count);
	const uint32_t warps_per_cta = 8;
	emit_visible_kernel<<<(s->tile_count + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, ctx->stream>>>(
		s->tile_mask, s->tile_vis_offset, s->tile_first_entity, s->tile_count, s->index_base, s->visible);
	if (prof) ZN_CUDA(cudaEventRecord(s->prof_events[s->levels.size() + 1], ctx->stream));
	ZN_CUDA(cudaGetLastError());
	return ZN_OK;
}

int zn_scene_set_profiling(zn_scene* s, int enable) {
	ZN_REQUIRE(s, "zn_scene_set_profiling: scene is NULL");
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	if (enable && s->prof_events.empty()) {
		s->prof_events.resize(s->levels.size() + 2);
		for (auto& e : s->prof_events) ZN_CUDA(cudaEventCreate(&e));
	}
	s->profiling = enable != 0;
	return ZN_OK;
}

int zn_scene_level_times(zn_scene* s, float* out_ms, uint32_t capacity, float* out_tail_ms) {
	ZN_REQUIRE(s && out_ms, "zn_scene_level_times: NULL argument");
	ZN_REQUIRE(capacity >= s->levels.size(), "zn_scene_level_times: capacity %u < level_count %zu", capacity, s->levels.size());
	if (!s->profiling || s->prof_events.empty()) return fail(ZN_ERR_STATE, "zn_scene_level_times: profiling is not enabled");
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	ZN_CUDA(cudaEventSynchronize(s->prof_events.back()));
	for (size_t l = 0; l < s->levels.size(); ++l) ZN_CUDA(cudaEventElapsedTime(out_ms + l, s->prof_events[l], s->prof_events[l + 1]));
	if (out_tail_ms) ZN_CUDA(cudaEventElapsedTime(out_tail_ms, s->prof_events[s->levels.size()], s->prof_events[s->levels.size() + 1]));
	return ZN_OK;
}

int zn_scene_launches_per_update(zn_scene* s, uint32_t* out_launches) {
	ZN_REQUIRE(s && out_launches, "zn_scene_launches_per_update: NULL argument");
	*out_launches = uint32_t(s->levels.size()) + 2;
	return ZN_OK;
}

int zn_scene_algorithmic_bytes(zn_scene* s, uint32_t visible_count, uint64_t* out_bytes) {
	ZN_REQUIRE(s && out_bytes, "zn_scene_algorithmic_bytes: NULL argument");
	const uint64_t roots = s->levels[0].end - s->levels[0].begin;
	const uint64_t children = uint64_t(s->entity_count) - roots;
	*out_bytes = roots * 188 + children * 192 + uint64_t(visible_count) * 4;
	return ZN_OK;
}

int zn_scene_visible_count(zn_scene* s, uint32_t* out_count) {
	ZN_REQUIRE(s && out_count, "zn_scene_visible_count: NULL argument");
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	ZN_CUDA(cudaMemcpyAsync(s->host_count, s->visible_count, 4, cudaMemcpyDeviceToHost, s->ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(s->ctx->stream));
	*out_count = *s->host_count;
	return ZN_OK;
}

int zn_scene_read_visible(zn_scene* s, uint32_t* out_indices, uint32_t capacity, uint32_t* out_count) {
	ZN_REQUIRE(s && out_count, "zn_scene_read_visible: NULL argument");
	uint32_t n = 0;
	int rc = zn_scene_visible_count(s, &n);
	if (rc != ZN_OK) return rc;
	*out_count = n;
	if (!out_indices) return ZN_OK;
	ZN_REQUIRE(capacity >= n, "zn_scene_read_visible: capacity %u < visible count %u", capacity, n);
	if (n) {
		ZN_CUDA(cudaMemcpyAsync(out_indices, s->visible, size_t(n) * 4, cudaMemcpyDeviceToHost, s->ctx->stream));
		ZN_CUDA(cudaStreamSynchronize(s->ctx->stream));
	}
	return ZN_OK;
}

static int read_matrices(zn_scene* s, const float* dev, uint32_t first, uint32_t count, float* out, const char* who) {
	int rc = check_range(s, first, count, who);
	if (rc != ZN_OK) return rc;
	ZN_REQUIRE(out || count == 0, "%s: output is NULL", who);
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	if (count) {
		ZN_CUDA(cudaMemcpyAsync(out, dev + size_t(first) * 16, size_t(count) * 64, cudaMemcpyDeviceToHost, s->ctx->stream));
		ZN_CUDA(cudaStreamSynchronize(s->ctx->stream));
	}
	return ZN_OK;
}

int zn_scene_read_world(zn_scene* s, uint32_t first, uint32_t count, float* out) {
	return read_matrices(s, s ? s->world : nullptr, first, count, out, "zn_scene_read_world");
}
int zn_scene_read_mvp(zn_scene* s, uint32_t first, uint32_t count, float* out) {
	return read_matrices(s, s ? s->mvp : nullptr, first, count, out, "zn_scene_read_mvp");
}

int zn_scene_read_inputs(zn_scene* s, uint32_t first, uint32_t count, float* position, float* rotation, float* scale,
                         float* aabb_center, float* aabb_half, uint32_t* parent) {
	int rc = check_range(s, first, count, "zn_scene_read_inputs");
	if (rc != ZN_OK) return rc;
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	float* outs[5] = {position, rotation, scale, aabb_center, aabb_half};
	for (int k = 0; k < 5; ++k)
		if ((rc = download3(s, 3 * k, first, count, outs[k])) != ZN_OK) return rc;
	if (parent && count) {
		ZN_CUDA(cudaMemcpyAsync(parent, s->parent + first, size_t(count) * 4, cudaMemcpyDeviceToHost, s->ctx->stream));
		ZN_CUDA(cudaStreamSynchronize(s->ctx->stream));
	}
	return ZN_OK;
}

int zn_scene_device_buffers(zn_scene* s, zn_scene_buffers* out) {
	ZN_REQUIRE(s && out, "zn_scene_device_buffers: NULL argument");
	out->world = s->world;
	out->mvp = s->mvp;
	out->visible = s->visible;
	out->visible_count = s->visible_count;
	out->entity_count = s->entity_count;
	return ZN_OK;
}

int zn_scene_frame_host(zn_scene* s, zn_frame_host_io* io) {
	ZN_REQUIRE(s && io, "zn_scene_frame_host: NULL argument");
	zn_ctx* ctx = s->ctx;
	const uint32_t E = s->entity_count;
	int rc = zn_scene_set_transforms(s, 0, E, io->position, io->rotation, io->scale, 12);
	if (rc != ZN_OK) return rc;
	if ((rc = zn_scene_update(s)) != ZN_OK) return rc;
	ZN_CUDA(cudaMemcpyAsync(s->host_count, s->visible_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (io->world) ZN_CUDA(cudaMemcpyAsync(io->world, s->world, size_t(E) * 64, cudaMemcpyDeviceToHost, ctx->stream));
	if (io->mvp) ZN_CUDA(cudaMemcpyAsync(io->mvp, s->mvp, size_t(E) * 64, cudaMemcpyDeviceToHost, ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(ctx->stream));
	io->visible_count = *s->host_count;
	if (io->visible && io->visible_count) {
		ZN_CUDA(cudaMemcpyAsync(io->visible, s->visible, size_t(io->visible_count) * 4, cudaMemcpyDeviceToHost, ctx->stream));
		ZN_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	return ZN_OK;
}

} // extern "C"
