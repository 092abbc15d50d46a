// zn_scene.cu — host side of the scene path: scene objects, uploads, the per-frame launch
// sequence (one persistent level kernel per hierarchy level, PDL-chained, then the emit kernel),
// readbacks, the multi-GPU expansion entry point, draw-argument building and the host-buffer
// frame calls.  The kernels live in zn_scene_kernels.cuh.
#include "zn_scene_kernels.cuh"

#include <cstdlib>

namespace zn {

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

// Function attributes are per device: set them for every scene (a process may drive several GPUs).
static cudaError_t set_kernel_attributes() {
	cudaError_t e;
#define ZN_ATTR(kernel, bytes)                                                                                   \
	if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes))) != cudaSuccess) return e; \
	if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess) return e;
	ZN_ATTR((scene_level_kernel<true, true>), kLevelSmem)
	ZN_ATTR((scene_level_kernel<false, true>), kLevelSmem)
	ZN_ATTR((scene_level_kernel<true, false>), kWorldSmem)
	ZN_ATTR((scene_level_kernel<false, false>), kWorldSmem)
	ZN_ATTR(scene_cull_kernel, kCullSmem)
	ZN_ATTR(scene_top_kernel<true>, kTopSmemView)
	ZN_ATTR(scene_top_kernel<false>, kTopSmemWorld)
#undef ZN_ATTR
	return cudaSuccess;
}

// Calls fn(level, lo, hi) for every level that intersects entities [first, first+count).
template <class F>
static int for_each_level_range(zn_scene* s, uint32_t first, uint32_t count, F&& fn) {
	const uint32_t last = first + count;
	for (const zn_level& lv : s->levels) {
		const uint32_t lo = std::max(first, lv.begin), hi = std::min(last, lv.end);
		if (lo >= hi) continue;
		const int rc = fn(lv, lo, hi);
		if (rc != ZN_OK) return rc;
	}
	return ZN_OK;
}

// Upload `count` strided 3-float records into block rows [row0, row0+3) for entities from `first`.
static int upload3(zn_scene* s, int row0, uint32_t first, uint32_t count, const float* host, size_t stride_bytes) {
	if (!host || count == 0) return ZN_OK;
	zn_ctx* ctx = s->ctx;
	s->inputs_touched = true;
	const uint32_t stride_w = uint32_t(stride_bytes / 4);
	const uint32_t chunk = 1u << 22;   // records per pass: <= 64 MiB of staging at stride 16
	int rc = ensure_staging(ctx, size_t(std::min(count, chunk)) * stride_bytes);
	if (rc != ZN_OK) return rc;
	for (uint32_t done = 0; done < count; done += chunk) {
		const uint32_t n = std::min(chunk, count - done);
		const size_t bytes = size_t(n - 1) * stride_bytes + 12;   // the last record needs only 12 bytes
		ZN_CUDA(cudaMemcpyAsync(ctx->staging, reinterpret_cast<const uint8_t*>(host) + size_t(done) * stride_bytes, bytes,
		                        cudaMemcpyHostToDevice, ctx->stream));
		const uint32_t chunk_first = first + done;
		rc = for_each_level_range(s, chunk_first, n, [&](const zn_level& lv, uint32_t lo, uint32_t hi) {
			const uint32_t* src = static_cast<const uint32_t*>(ctx->staging) + size_t(lo - chunk_first) * stride_w;
			repack3_kernel<<<(hi - lo + 255) / 256, 256, 0, ctx->stream>>>(src, stride_w, hi - lo, s->blocks, lv.tile_base, lo - lv.begin, row0);
			return ZN_OK;
		});
		if (rc != ZN_OK) return rc;
		ZN_CUDA(cudaGetLastError());
	}
	return ZN_OK;
}

// A handful of entities (<= kPatchEntities): ONE launch carrying the values as kernel parameters — no staging copy, no
// repack passes.  arrays[a] (NULL = absent) holds 3-float records for block rows row0 + 3a .. row0 + 3a + 2.
static int patch_rows(zn_scene* s, uint32_t row0, uint32_t first, uint32_t count, const float* const* arrays, int n_arrays, size_t stride_bytes) {
	PatchRows patch = {};
	patch.count = count;
	patch.row0 = row0;
	for (int a = 0; a < n_arrays; ++a)
		if (arrays[a]) patch.rows |= 0x7u << (3 * a);
	if (count == 0 || patch.rows == 0u) return ZN_OK;
	size_t l = 0;
	for (uint32_t k = 0; k < count; ++k) {
		const uint32_t i = first + k;
		while (i >= s->levels[l].end) ++l;
		const uint32_t rel = i - s->levels[l].begin;
		patch.at[k] = (s->levels[l].tile_base + rel / kTile) * kBlockWords + rel % kTile;
		for (int a = 0; a < n_arrays; ++a)
			if (arrays[a]) std::memcpy(&patch.v[k][a * 3], reinterpret_cast<const uint8_t*>(arrays[a]) + size_t(k) * stride_bytes, 12);
	}
	s->inputs_touched = true;
	patch_rows_kernel<<<1, kPatchEntities * 9, 0, s->ctx->stream>>>(s->blocks, patch);
	ZN_CUDA(cudaGetLastError());
	return ZN_OK;
}

// Download block rows [row0, row0+nrows) of entities [first, first+count) as packed [count][nrows].
static int download_rows(zn_scene* s, int row0, int nrows, uint32_t first, uint32_t count, void* host) {
	if (!host || count == 0) return ZN_OK;
	zn_ctx* ctx = s->ctx;
	int rc = ensure_staging(ctx, size_t(count) * nrows * 4);
	if (rc != ZN_OK) return rc;
	rc = for_each_level_range(s, first, count, [&](const zn_level& lv, uint32_t lo, uint32_t hi) {
		unpack_kernel<<<(hi - lo + 255) / 256, 256, 0, ctx->stream>>>(s->blocks, lv.tile_base, lo - lv.begin, hi - lo, row0, nrows,
		                                                             static_cast<uint32_t*>(ctx->staging) + size_t(lo - first) * nrows);
		return ZN_OK;
	});
	if (rc != ZN_OK) return rc;
	ZN_CUDA(cudaGetLastError());
	ZN_CUDA(cudaMemcpyAsync(host, ctx->staging, size_t(count) * nrows * 4, cudaMemcpyDeviceToHost, ctx->stream));
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
	// function attributes are per device: set them for every scene (a process may drive several GPUs)
	ZN_CUDA(set_kernel_attributes());
	zn_scene* s = new zn_scene();
	s->ctx = ctx;
	s->entity_count = E;
	// Persistent grid: as many CTAs as fit at once.
	int per_sm_a = 0, per_sm_b = 0, per_sm_c = 0;
	ZN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, scene_level_kernel<true, true>, kTile, kLevelSmem));
	ZN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, scene_level_kernel<false, true>, kTile, kLevelSmem));
	ZN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, scene_cull_kernel, kTile, kCullSmem));
	s->max_ctas = std::min(per_sm_a, per_sm_b) * ctx->sm_count;
	{
		int wa = 0, wb = 0;
		ZN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wa, scene_level_kernel<true, false>, kTile, kWorldSmem));
		ZN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wb, scene_level_kernel<false, false>, kTile, kWorldSmem));
		s->max_ctas_world = std::max(1, std::min(wa, wb)) * ctx->sm_count;
	}
	s->max_ctas_cull = per_sm_c * ctx->sm_count;
	if (s->max_ctas < 1 || s->max_ctas_cull < 1) { delete s; return fail(ZN_ERR_CUDA, "zn_scene_create: a scene kernel does not fit on an SM"); }
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
	s->chunk_count = (s->tile_count + kChunkTiles - 1) / kChunkTiles;
	auto cleanup = [&](int rc) { zn_scene_destroy(s); return rc; };
#define ZN_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return cleanup(fail(e_ == cudaErrorMemoryAllocation ? ZN_ERR_NOMEM : ZN_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_))); } while (0)
	ZN_TRY(cudaMalloc(&s->blocks, size_t(s->tile_count) * kBlockWords * 4));
	ZN_TRY(cudaMalloc(&s->tile_prange, size_t(s->tile_count) * sizeof(uint2)));
	ZN_TRY(cudaMalloc(&s->world, size_t(E) * 64));
	ZN_TRY(cudaMalloc(&s->mvp, size_t(E) * 64));
	ZN_TRY(cudaMalloc(&s->visible_own, size_t(E) * 4));
	ZN_TRY(cudaMalloc(&s->visible_count_own, 4));
	s->visible = s->visible_own;
	s->visible_count = s->visible_count_own;
	ZN_TRY(cudaMalloc(&s->chunk_total, size_t(s->chunk_count) * 3 * 4));   // three sets, by frame number mod 3
	ZN_TRY(cudaMalloc(&s->level_ticket, (s->levels.size() + 1) * 4));
	ZN_TRY(cudaMemsetAsync(s->level_ticket, 0, (s->levels.size() + 1) * 4, ctx->stream));
	s->record_words = record_header_words(s->chunk_count) + s->tile_count * kWarpsPerTile;
	ZN_TRY(cudaMalloc(&s->record_own, size_t(s->record_words) * 2 * 4));
	ZN_TRY(cudaMemsetAsync(s->record_own, 0, size_t(s->record_words) * 2 * 4, ctx->stream));
	s->record = s->record_own;
	ZN_TRY(cudaMalloc(&s->tile_first_entity, size_t(s->tile_count) * 4));
	ZN_TRY(cudaHostAlloc(&s->host_count, 4, cudaHostAllocDefault));
	ZN_TRY(cudaMemsetAsync(s->blocks, 0, size_t(s->tile_count) * kBlockWords * 4, ctx->stream));
	init_parent_rows_kernel<<<(s->tile_count * kTile + 255) / 256, 256, 0, ctx->stream>>>(s->blocks, s->tile_count);
	ZN_TRY(cudaGetLastError());
	ZN_TRY(cudaMemsetAsync(s->chunk_total, 0, size_t(s->chunk_count) * 3 * 4, ctx->stream));
	ZN_TRY(cudaMemcpyAsync(s->tile_first_entity, tile_first.data(), tile_first.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	ZN_TRY(cudaMemsetAsync(s->visible_count, 0, 4, ctx->stream));
	ZN_TRY(cudaStreamSynchronize(ctx->stream));
#undef ZN_TRY
	for (auto& lv : s->levels) {
		int rc = encode_matrix_tensor_map(&lv.tm_world, s->world, lv.end);
		if (rc == ZN_OK) rc = encode_matrix_tensor_map(&lv.tm_mvp, s->mvp, lv.end);
		if (rc != ZN_OK) return cleanup(rc);
	}
	{
		int rc = encode_matrix_tensor_map(&s->tm_world_all, s->world, E);
		if (rc == ZN_OK) rc = encode_matrix_tensor_map(&s->tm_mvp_all, s->mvp, E);
		if (rc != ZN_OK) return cleanup(rc);
	}
	// the leading levels that scene_top_kernel composes in one launch: as many as fit kTopTiles tiles, at least two
	{
		uint32_t tiles = 0, n = 0;
		while (n < s->levels.size() && tiles + s->levels[n].tiles <= uint32_t(kTopTiles)) tiles += s->levels[n++].tiles;
		s->top_levels = n >= 2 ? n : 0u;
	}
	*out_scene = s;
	return ZN_OK;
}

static void pipe_destroy(zn_scene* s);

int zn_scene_destroy(zn_scene* s) {
	if (!s) return ZN_OK;
	cudaSetDevice(s->ctx->device);
	cudaStreamSynchronize(s->ctx->stream);
	cudaFree(s->blocks); cudaFree(s->tile_prange); cudaFree(s->world); cudaFree(s->mvp); cudaFree(s->visible_own);
	cudaFree(s->expand_entries); cudaFree(s->expand_count); cudaFree(s->visible_count_own); cudaFree(s->chunk_total); cudaFree(s->level_ticket); cudaFree(s->record_own); cudaFree(s->tile_first_entity);
	if (s->host_count) cudaFreeHost(s->host_count);
	for (auto e : s->prof_events) cudaEventDestroy(e);
	for (auto e : s->cull_events) if (e) cudaEventDestroy(e);
	pipe_destroy(s);
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
	if (count <= kPatchEntities && s->tile_count <= 0xFFFFFFFFu / kBlockWords) {
		const float* arrays[3] = {position, rotation, scale};
		return patch_rows(s, 0, first, count, arrays, 3, stride_bytes);
	}
	if ((rc = upload3(s, 0, first, count, position, stride_bytes)) != ZN_OK) return rc;
	if ((rc = upload3(s, 3, first, count, rotation, stride_bytes)) != ZN_OK) return rc;
	return upload3(s, 6, first, count, scale, stride_bytes);
}

int zn_scene_set_bounds(zn_scene* s, uint32_t first, uint32_t count, const float* center, const float* half_extent, size_t stride_bytes) {
	int rc = check_range(s, first, count, "zn_scene_set_bounds");
	if (rc == ZN_OK) rc = check_stride(stride_bytes, "zn_scene_set_bounds");
	if (rc != ZN_OK) return rc;
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	if (count <= kPatchEntities && s->tile_count <= 0xFFFFFFFFu / kBlockWords) {
		const float* arrays[3] = {center, half_extent, nullptr};
		return patch_rows(s, 9, first, count, arrays, 2, stride_bytes);
	}
	if ((rc = upload3(s, 9, first, count, center, stride_bytes)) != ZN_OK) return rc;
	return upload3(s, 12, first, count, half_extent, stride_bytes);
}

int zn_scene_set_parents(zn_scene* s, uint32_t first, uint32_t count, const uint32_t* parent) {
	int rc = check_range(s, first, count, "zn_scene_set_parents");
	if (rc != ZN_OK) return rc;
	ZN_REQUIRE(parent || count == 0, "zn_scene_set_parents: parent is NULL");
	if (count == 0) return ZN_OK;
	// validate the level-order contract on the host: a parent must live in an earlier level
	size_t l = 0;
	for (uint32_t k = 0; k < count; ++k) {
		const uint32_t i = first + k;
		while (i >= s->levels[l].end) ++l;
		const uint32_t p = parent[k];
		ZN_REQUIRE(p == ZN_NO_PARENT || p < s->levels[l].begin,
		           "zn_scene_set_parents: entity %u (level %zu) has parent %u, which is not in an earlier level", i, l, p);
	}
	zn_ctx* ctx = s->ctx;
	ZN_CUDA(cudaSetDevice(ctx->device));
	const uint32_t chunk = 1u << 24;
	if ((rc = ensure_staging(ctx, size_t(std::min(count, chunk)) * 4)) != ZN_OK) return rc;
	for (uint32_t done = 0; done < count; done += chunk) {
		const uint32_t n = std::min(chunk, count - done);
		ZN_CUDA(cudaMemcpyAsync(ctx->staging, parent + done, size_t(n) * 4, cudaMemcpyHostToDevice, ctx->stream));
		const uint32_t chunk_first = first + done;
		rc = for_each_level_range(s, chunk_first, n, [&](const zn_level& lv, uint32_t lo, uint32_t hi) {
			repack1_kernel<<<(hi - lo + 255) / 256, 256, 0, ctx->stream>>>(static_cast<const uint32_t*>(ctx->staging) + (lo - chunk_first),
			                                                              hi - lo, s->blocks, lv.tile_base, lo - lv.begin, 15);
			return ZN_OK;
		});
		if (rc != ZN_OK) return rc;
		ZN_CUDA(cudaGetLastError());
		ZN_CUDA(cudaStreamSynchronize(ctx->stream));   // the caller's array may be pageable and reused
	}
	s->ranges_dirty = true;
	s->inputs_touched = true;
	return ZN_OK;
}

int zn_scene_set_camera(zn_scene* s, const float view[16], const float proj[16]) {
	ZN_REQUIRE(s && view && proj, "zn_scene_set_camera: NULL argument");
	zn_mat4_mul(proj, view, s->view_proj);
	zn_frustum_planes(s->view_proj, s->planes);
	s->has_camera = true;
	return ZN_OK;
}

int zn_scene_bind_visible_output(zn_scene* s, void* visible_dev, uint32_t capacity, void* visible_count_dev) {
	ZN_REQUIRE(s, "zn_scene_bind_visible_output: scene is NULL");
	if (!visible_dev && !visible_count_dev) {
		s->visible = s->visible_own;
		s->visible_count = s->visible_count_own;
		return ZN_OK;
	}
	ZN_REQUIRE(visible_dev && visible_count_dev, "zn_scene_bind_visible_output: pass both pointers, or both NULL to restore the internal buffers");
	ZN_REQUIRE(capacity >= s->entity_count, "zn_scene_bind_visible_output: capacity %u < entity_count %u", capacity, s->entity_count);
	s->visible = static_cast<uint32_t*>(visible_dev);
	s->visible_count = static_cast<uint32_t*>(visible_count_dev);
	return ZN_OK;
}

int zn_scene_bind_visibility_record(zn_scene* s, void* record_dev) {
	ZN_REQUIRE(s, "zn_scene_bind_visibility_record: scene is NULL");
	s->record_bound = record_dev != nullptr;
	s->record = record_dev ? static_cast<uint32_t*>(record_dev) : s->record_own + size_t(s->record_slot) * s->record_words;
	return ZN_OK;
}

int zn_scene_bind_peer_records(zn_scene* s, void* const* peer_record_ptrs, uint32_t peer_count) {
	ZN_REQUIRE(s, "zn_scene_bind_peer_records: scene is NULL");
	ZN_REQUIRE(peer_count < uint32_t(kMaxRecordTargets), "zn_scene_bind_peer_records: at most %d peers", kMaxRecordTargets - 1);
	ZN_REQUIRE(peer_record_ptrs || peer_count == 0, "zn_scene_bind_peer_records: NULL pointer array");
	s->peer_records.clear();
	for (uint32_t k = 0; k < peer_count; ++k) {
		ZN_REQUIRE(peer_record_ptrs[k], "zn_scene_bind_peer_records: peer %u is NULL", k);
		s->peer_records.push_back(static_cast<uint32_t*>(peer_record_ptrs[k]));
	}
	return ZN_OK;
}

int zn_scene_bind_peer_signal(zn_scene* s, void* const* flag_ptrs, uint32_t flag_count) {
	ZN_REQUIRE(s, "zn_scene_bind_peer_signal: scene is NULL");
	ZN_REQUIRE(flag_count <= uint32_t(kMaxRecordTargets), "zn_scene_bind_peer_signal: at most %d flags", kMaxRecordTargets);
	ZN_REQUIRE(flag_ptrs || flag_count == 0, "zn_scene_bind_peer_signal: NULL pointer array");
	s->peer_flags.clear();
	for (uint32_t k = 0; k < flag_count; ++k) {
		ZN_REQUIRE(flag_ptrs[k] && reinterpret_cast<uintptr_t>(flag_ptrs[k]) % 4 == 0, "zn_scene_bind_peer_signal: flag %u is NULL or misaligned", k);
		s->peer_flags.push_back(static_cast<uint32_t*>(flag_ptrs[k]));
	}
	return ZN_OK;
}

int zn_scene_set_signal_value(zn_scene* s, uint32_t value) {
	ZN_REQUIRE(s, "zn_scene_set_signal_value: scene is NULL");
	s->signal_value = value;
	return ZN_OK;
}

} // extern "C"

// The expansion launch on a stream of the caller's choice (the C-ABI entry points use the context
// stream; zn_exchange_* runs it on its side stream so that it overlaps the next frame's update).
int zn::expand_visible_on(cudaStream_t stream, zn_scene* s, const char* who, const void* records_dev, uint32_t record_count,
                          uint32_t record_stride_words, int skip_index, uint32_t skip_count, const uint32_t* index_bases, void* lists_dev,
                          uint32_t list_stride, void* counts_dev, const void* flags_dev, uint32_t flag_stride_words, uint32_t records_per_flag,
                          uint32_t expected, uint32_t timeout_ms, void* status_dev, const void* abort_status, bool early_start) {
	ZN_REQUIRE(s && records_dev && lists_dev && counts_dev, "%s: NULL argument", who);
	ZN_REQUIRE(record_count >= 1 && record_count <= 64, "%s: record_count %u must be in [1, 64]", who, record_count);
	ZN_REQUIRE(record_stride_words >= s->record_words, "%s: record stride %u < record size %u words", who, record_stride_words, s->record_words);
	ZN_REQUIRE(list_stride >= s->entity_count, "%s: list stride %u < entity_count %u", who, list_stride, s->entity_count);
	if (skip_index < 0 || uint32_t(skip_index) >= record_count) { skip_index = int(record_count); skip_count = 0; }   // nothing left out
	skip_count = std::min(skip_count, record_count - uint32_t(skip_index));
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	ExpandBases bases;
	for (uint32_t r = 0; r < 64; ++r) bases.base[r] = (index_bases && r < record_count) ? index_bases[r] : 0u;
	SignalTargets sig = {};
	if (flags_dev) {   // the signalled form also raises this rank's own flags: everything before it on the stream is complete
		for (uint32_t* flag : s->peer_flags) sig.flag[sig.n++] = flag;
		sig.value = s->signal_value;
	}
	if (record_count == skip_count && !flags_dev) return ZN_OK;     // nothing to expand and nothing to signal
	// early_start (zn_exchange_expand on the context stream): the expansion of frame k-1 shares nothing with the update of
	// frame k before it on the stream — other record slot, other half of the lists — and reads a record only behind its
	// owner's flag, so it is launched programmatically: its CTAs start beside the tail of the emit kernel instead of behind a
	// drained GPU.  Only the ONE thread that raises this rank's flags waits for that update to complete (griddepcontrol.wait,
	// which also keeps the chain of completions intact for the kernels after the expansion).
	cudaLaunchAttribute early;
	early.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	early.val.programmaticStreamSerializationAllowed = 1;
	const unsigned early_attrs = (early_start && flags_dev && !abort_status) ? 1u : 0u;
	if (s->expand_mode == ZN_EXPAND_WORKLIST) {
		// Two launches: classify every (record, group) into a work list of the non-empty ones, then a persistent grid works it off.
		const uint32_t live = record_count - skip_count;
		const uint32_t groups = s->chunk_count * 8u;
		const uint32_t need = std::max<uint32_t>(1u, live) * groups;
		if (s->expand_capacity < need) {
			ZN_CUDA(cudaStreamSynchronize(stream));
			if (s->expand_entries) ZN_CUDA(cudaFree(s->expand_entries));
			s->expand_entries = nullptr; s->expand_capacity = 0;
			ZN_CUDA(cudaMalloc(&s->expand_entries, size_t(need) * sizeof(uint4)));
			s->expand_capacity = need;
		}
		if (!s->expand_count) {
			ZN_CUDA(cudaMalloc(&s->expand_count, 8));
			ZN_CUDA(cudaMemsetAsync(s->expand_count, 0, 8, stream));
			int per_sm = 0;
			ZN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, expand_work_kernel, kExpandWarps * 32, 0));
			s->max_ctas_expand = std::max(per_sm, 1) * s->ctx->sm_count;
		}
		ExpandWork work = {s->expand_entries, s->expand_count, s->expand_parity};
		s->expand_parity ^= 1u;
		const dim3 cgrid((groups + kClassifyThreads - 1) / kClassifyThreads, std::max<uint32_t>(1u, live));
		cudaLaunchConfig_t ccfg = {};
		ccfg.gridDim = cgrid;
		ccfg.blockDim = dim3(kClassifyThreads);
		ccfg.stream = stream;
		ccfg.attrs = &early;
		ccfg.numAttrs = early_attrs;
		ZN_CUDA(cudaLaunchKernelEx(&ccfg, expand_classify_kernel, static_cast<const uint32_t*>(records_dev), record_stride_words,
		                           (const uint32_t*)s->tile_first_entity, s->tile_count, s->chunk_count, bases, record_count, skip_index, int(skip_count),
		                           static_cast<uint32_t*>(counts_dev), static_cast<const uint32_t*>(flags_dev), flag_stride_words, expected,
		                           records_per_flag, uint64_t(timeout_ms) * 1000000ull, static_cast<uint32_t*>(status_dev),
		                           static_cast<const uint32_t*>(abort_status), sig, work));
		ZN_CUDA(cudaGetLastError());
		cudaLaunchAttribute pdl;
		pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
		pdl.val.programmaticStreamSerializationAllowed = 1;
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3(std::min<uint32_t>((need + kExpandWarps - 1) / kExpandWarps, uint32_t(s->max_ctas_expand)));
		cfg.blockDim = dim3(kExpandWarps * 32);
		cfg.stream = stream;
		cfg.attrs = &pdl;
		cfg.numAttrs = 1;
		ZN_CUDA(cudaLaunchKernelEx(&cfg, expand_work_kernel, static_cast<const uint32_t*>(records_dev), record_stride_words,
		                           (const uint32_t*)s->tile_first_entity, s->tile_count, s->chunk_count, bases, static_cast<uint32_t*>(lists_dev),
		                           list_stride, work));
		return ZN_OK;
	}
	// One CTA per (record, chunk), one warp per 4-tile group.
	const dim3 grid((s->chunk_count + kExpandChunks - 1) / kExpandChunks, std::max<uint32_t>(1u, record_count - skip_count));
	cudaLaunchConfig_t gcfg = {};
	gcfg.gridDim = grid;
	gcfg.blockDim = dim3(kExpandWarps * 32);
	gcfg.stream = stream;
	gcfg.attrs = &early;
	gcfg.numAttrs = early_attrs;
	ZN_CUDA(cudaLaunchKernelEx(&gcfg, expand_visible_kernel, static_cast<const uint32_t*>(records_dev), record_stride_words,
	                           (const uint32_t*)s->tile_first_entity, s->tile_count, s->chunk_count, bases, record_count, skip_index, int(skip_count),
	                           static_cast<uint32_t*>(lists_dev), list_stride, static_cast<uint32_t*>(counts_dev),
	                           static_cast<const uint32_t*>(flags_dev), flag_stride_words, expected, records_per_flag,
	                           uint64_t(timeout_ms) * 1000000ull, static_cast<uint32_t*>(status_dev), static_cast<const uint32_t*>(abort_status), sig));
	ZN_CUDA(cudaGetLastError());
	return ZN_OK;
}

extern "C" {

int zn_scene_expand_visible(zn_scene* s, const void* records_dev, uint32_t record_count, uint32_t record_stride_words, int skip_index,
                            const uint32_t* index_bases, void* lists_dev, uint32_t list_stride, void* counts_dev) {
	ZN_REQUIRE(s, "zn_scene_expand_visible: scene is NULL");
	const void* abort_status = s->abort_status;   // set by a zn_scene_wait_flags on this scene: a timed-out wait cancels this expansion
	s->abort_status = nullptr;
	return expand_visible_on(s->ctx->stream, s, "zn_scene_expand_visible", records_dev, record_count, record_stride_words, skip_index, 1, index_bases, lists_dev,
	                      list_stride, counts_dev, nullptr, 0, 1, 0, 0, nullptr, abort_status);
}

int zn_scene_expand_visible_signalled(zn_scene* s, const void* records_dev, uint32_t record_count, uint32_t record_stride_words, int skip_index,
                                      uint32_t skip_count, const uint32_t* index_bases, void* lists_dev, uint32_t list_stride, void* counts_dev,
                                      const void* flags_dev, uint32_t flag_stride_words, uint32_t records_per_flag, uint32_t expected,
                                      uint32_t timeout_ms, void* status_dev) {
	ZN_REQUIRE(flags_dev && status_dev, "zn_scene_expand_visible_signalled: flags_dev / status_dev is NULL");
	ZN_REQUIRE(records_per_flag >= 1 && timeout_ms >= 1, "zn_scene_expand_visible_signalled: records_per_flag and timeout_ms must be >= 1");
	ZN_REQUIRE(skip_count <= record_count, "zn_scene_expand_visible_signalled: skip_count %u > record_count %u", skip_count, record_count);
	return expand_visible_on(s->ctx->stream, s, "zn_scene_expand_visible_signalled", records_dev, record_count, record_stride_words, skip_index, skip_count, index_bases,
	                      lists_dev, list_stride, counts_dev, flags_dev, flag_stride_words, records_per_flag, expected, timeout_ms, status_dev, nullptr);
}

int zn_scene_wait_flags(zn_scene* s, const void* flags_dev, uint32_t flag_stride_words, uint32_t flag_count, uint32_t expected,
                        uint32_t timeout_ms, void* status_dev) {
	ZN_REQUIRE(s && flags_dev && status_dev, "zn_scene_wait_flags: NULL argument");
	ZN_REQUIRE(flag_count >= 1 && flag_count <= 64 && timeout_ms >= 1, "zn_scene_wait_flags: flag_count must be in [1, 64], timeout_ms >= 1");
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	SignalTargets sig = {};
	for (uint32_t* flag : s->peer_flags) sig.flag[sig.n++] = flag;
	sig.value = s->signal_value;
	wait_flags_kernel<<<1, 64, 0, s->ctx->stream>>>(static_cast<const uint32_t*>(flags_dev), flag_stride_words, flag_count, expected,
	                                                uint64_t(timeout_ms) * 1000000ull, static_cast<uint32_t*>(status_dev), sig);
	ZN_CUDA(cudaGetLastError());
	s->abort_status = status_dev;
	return ZN_OK;
}

int zn_scene_set_level_fusion(zn_scene* s, int enabled) {
	ZN_REQUIRE(s, "zn_scene_set_level_fusion: scene is NULL");
	s->level_fusion = enabled != 0;
	return ZN_OK;
}

int zn_scene_set_expand_mode(zn_scene* s, int mode) {
	ZN_REQUIRE(s, "zn_scene_set_expand_mode: scene is NULL");
	ZN_REQUIRE(mode == ZN_EXPAND_GRID || mode == ZN_EXPAND_WORKLIST, "zn_scene_set_expand_mode: unknown mode %d", mode);
	s->expand_mode = mode;
	return ZN_OK;
}

int zn_scene_set_index_base(zn_scene* s, uint32_t index_base) {
	ZN_REQUIRE(s, "zn_scene_set_index_base: scene is NULL");
	s->index_base = index_base;
	return ZN_OK;
}

namespace zn {

// The emit kernel of a frame (K3): ordered visible list + record header from the masks and chunk
// totals the level kernels (or the cull pass) left.  Launched with programmatic stream
// serialization behind the kernel that produced them.
static int launch_emit(zn_scene* s, const RecordTargets& rec, const uint32_t* totals, uint32_t* totals_next, uint32_t* visible,
                       uint32_t* visible_count, uint32_t remote_tiles_below, bool pdl_ok) {
	cudaLaunchAttribute pdl;
	pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	pdl.val.programmaticStreamSerializationAllowed = 1;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(s->chunk_count);
	cfg.blockDim = dim3(256);
	cfg.dynamicSmemBytes = 0;
	cfg.stream = s->ctx->stream;
	cfg.attrs = &pdl;
	cfg.numAttrs = pdl_ok ? 1 : 0;
	ZN_CUDA(cudaLaunchKernelEx(&cfg, emit_visible_kernel, rec, (const uint32_t*)s->tile_first_entity, s->tile_count, s->chunk_count,
	                           totals, totals_next, s->index_base, visible, visible_count, remote_tiles_below));
	s->tail_is_emit = true;
	return ZN_OK;
}

// Next record slot (own record only) and next set of chunk totals: frame k uses totals set k % 3, its emit clears set
// (k + 2) % 3 — the one frame k - 1 used, whose emit has completed by then (every kernel of a frame waits for the one
// before it, level 0 at the latest when it ends).
static void next_frame_buffers(zn_scene* s) {
	s->prev_frame_record = s->frame_record;
	s->frame_parity = (s->frame_parity + 1u) % 3u;
	if (!s->record_bound) {
		s->record_slot ^= 1u;
		s->record = s->record_own + size_t(s->record_slot) * s->record_words;
	}
	s->frame_record = s->record;
}

static RecordTargets record_targets(const zn_scene* s) {
	RecordTargets rec = {};
	rec.ptr[0] = s->record;
	rec.n = 1;
	for (uint32_t* peer : s->peer_records) rec.ptr[rec.n++] = peer;
	return rec;
}

// One frame on the context stream.  with_view: the fused form (world + MVP + masks, then emit);
// otherwise TRS -> world only.
static int run_update(zn_scene* s, bool with_view) {
	zn_ctx* ctx = s->ctx;
	ZN_CUDA(cudaSetDevice(ctx->device));
	if (s->ranges_dirty) {
		tile_prange_kernel<<<(s->tile_count + 7) / 8, 256, 0, ctx->stream>>>(s->blocks, s->tile_count, s->tile_prange);
		ZN_CUDA(cudaGetLastError());
		s->ranges_dirty = false;
	}
	FrameConst fc = {};
	if (with_view) {
		std::memcpy(fc.vp, s->view_proj, sizeof(fc.vp));
		std::memcpy(fc.planes, s->planes, sizeof(fc.planes));
	}
	const bool prof = s->profiling;
	const uint32_t header_words = record_header_words(s->chunk_count);
	// Peers get the masks of the leading run of small levels from the emit kernel, of every other
	// level from its level kernel (see emit_chunk).
	size_t small_levels = 0;
	const uint32_t n_targets = 1u + uint32_t(s->peer_records.size());
	uint32_t remote_tiles_below = 0;
	if (n_targets > 1)
		while (small_levels < s->levels.size() && s->levels[small_levels].tiles < kSmallLevelTiles) {
			remote_tiles_below = s->levels[small_levels].tile_base + s->levels[small_levels].tiles;
			++small_levels;
		}
	// A frame with a view takes the next record slot (own record only) and the next set of chunk totals, so that its level
	// kernels share nothing with the previous frame's emit kernel and may run beside it.
	if (with_view) next_frame_buffers(s);
	const RecordTargets rec = record_targets(s);
	RecordTargets rec_local = rec;
	rec_local.n = 1;
	uint32_t* totals = s->chunk_total + size_t(s->frame_parity) * s->chunk_count;
	uint32_t* totals_next = s->chunk_total + size_t((s->frame_parity + 2u) % 3u) * s->chunk_count;
	// Level 0 need not wait for the kernel before it (the previous frame's emit, or whatever reads the last results) as
	// long as it overwrites nothing that kernel uses: always true for a world-only update, true for a fused one when this
	// frame's record is not the previous frame's (the scene's own is double-buffered; an exchange binds another slot every
	// frame).  Not after an upload: then even the input blocks are still being written.
	// And only behind an EMIT kernel of this scene — by then every level kernel of the previous frame has completed (a
	// world-only update ends with a level kernel that may still be reading the parents level 0 is about to rewrite).
	// And never a launch that stores into PEER records (a level 0 too large to leave its masks to the emit kernel): when a
	// slot of a peer's buffer may be rewritten is settled by the expansion before this update on the stream (zn_exchange.cpp,
	// "Slot reuse"), so such a launch waits for it like every later level does.  (scene_top_kernel never writes to peers.)
	const bool first_launch_pushes = with_view && n_targets > 1 && small_levels == 0;
	const uint32_t defer_wait = (s->tail_is_emit && !s->inputs_touched && !s->profiling && !first_launch_pushes &&
	                             (!with_view || s->frame_record != s->prev_frame_record)) ? 1u : 0u;
	// Every kernel of a frame is launched with programmatic stream serialization: its prologue (and
	// the first input-block load) overlaps the tail of the kernel before it; it calls
	// griddepcontrol.wait before touching anything that kernel wrote.  That includes level 0, whose
	// start-up overlaps the previous frame's emit — unless an upload has written the input blocks in
	// between (then the block load itself must wait, and level 0 is a normal launch).
	cudaLaunchAttribute pdl;
	pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	pdl.val.programmaticStreamSerializationAllowed = 1;
	cudaLaunchConfig_t cfg = {};
	cfg.blockDim = dim3(kTile);
	cfg.dynamicSmemBytes = with_view ? kLevelSmem : kWorldSmem;
	cfg.stream = ctx->stream;
	cfg.attrs = &pdl;
	size_t l_first = 0;
	if (s->level_fusion && s->top_levels && !prof) {
		// the top of the tree in one launch of one CTA, in place of level 0 (same dependency rules)
		TopMaps maps;
		TopLevels top = {};
		top.levels = s->top_levels;
		for (uint32_t l = 0; l < s->top_levels; ++l) {
			const zn_level& lv = s->levels[l];
			maps.world[l] = lv.tm_world;
			maps.mvp[l] = lv.tm_mvp;
			top.begin[l] = lv.begin;
			top.end[l] = lv.end;
			top.tiles += lv.tiles;
		}
		cfg.gridDim = dim3(1);
		cfg.dynamicSmemBytes = with_view ? kTopSmemView : kTopSmemWorld;
		cfg.numAttrs = s->inputs_touched ? 0 : 1;
		if (with_view)
			ZN_CUDA(cudaLaunchKernelEx(&cfg, scene_top_kernel<true>, maps, fc, top, (const uint32_t*)s->blocks, rec.ptr[0], header_words, totals, defer_wait));
		else
			ZN_CUDA(cudaLaunchKernelEx(&cfg, scene_top_kernel<false>, maps, fc, top, (const uint32_t*)s->blocks, rec.ptr[0], header_words, totals, defer_wait));
		cfg.dynamicSmemBytes = with_view ? kLevelSmem : kWorldSmem;
		l_first = s->top_levels;
	}
	for (size_t l = l_first; l < s->levels.size(); ++l) {
		zn_level& lv = s->levels[l];
		const uint32_t grid = std::min<uint32_t>(lv.tiles, uint32_t(with_view ? s->max_ctas : s->max_ctas_world));
		cfg.gridDim = dim3(grid);
		cfg.numAttrs = ((l == 0 && s->inputs_touched) || prof) ? 0 : 1;
		if (prof) ZN_CUDA(cudaEventRecord(s->prof_events[l], ctx->stream));
		uint32_t* ticket = s->level_ticket + l;   // one counter per level: PDL lets consecutive levels overlap
		auto launch = [&](auto kernel) {
			return cudaLaunchKernelEx(&cfg, kernel, lv.tm_world, lv.tm_mvp, fc, (const uint32_t*)s->blocks, (const uint2*)s->tile_prange,
			                          (const float*)s->world, lv.begin, lv.end, lv.tile_base, lv.tiles, l < small_levels ? rec_local : rec,
			                          header_words, totals, ticket, lv.ticket_base, l == 0 ? defer_wait : 0u);
		};
		if (l == 0) ZN_CUDA(with_view ? launch(scene_level_kernel<false, true>) : launch(scene_level_kernel<false, false>));
		else ZN_CUDA(with_view ? launch(scene_level_kernel<true, true>) : launch(scene_level_kernel<true, false>));
		// every tile takes one ticket and every CTA one more (the draw that tells it to stop)
		lv.ticket_base += lv.tiles + grid;
	}
	if (prof) ZN_CUDA(cudaEventRecord(s->prof_events[s->levels.size()], ctx->stream));
	s->tail_is_emit = false;
	if (with_view) {
		const int rc = launch_emit(s, rec, totals, totals_next, s->visible, s->visible_count, remote_tiles_below, !prof);
		if (rc != ZN_OK) return rc;
	}
	if (prof) ZN_CUDA(cudaEventRecord(s->prof_events[s->levels.size() + 1], ctx->stream));
	s->world_valid = true;
	s->inputs_touched = false;
	return ZN_OK;
}

// The tensor map of a caller-owned float[E][16] MVP array (encoded once per pointer).
static int mvp_map_for(zn_scene* s, void* mvp_dev, const CUtensorMap** out) {
	for (auto& e : s->mvp_maps)
		if (e.first == mvp_dev) { *out = &e.second; return ZN_OK; }
	if (s->mvp_maps.size() >= 16) s->mvp_maps.erase(s->mvp_maps.begin());
	CUtensorMap m;
	const int rc = encode_matrix_tensor_map(&m, mvp_dev, s->entity_count);
	if (rc != ZN_OK) return rc;
	s->mvp_maps.emplace_back(mvp_dev, m);
	*out = &s->mvp_maps.back().second;
	return ZN_OK;
}

} // namespace zn

int zn_scene_update(zn_scene* s) {
	ZN_REQUIRE(s, "zn_scene_update: scene is NULL");
	return run_update(s, s->has_camera);
}

int zn_scene_update_world(zn_scene* s) {
	ZN_REQUIRE(s, "zn_scene_update_world: scene is NULL");
	return run_update(s, false);
}

int zn_scene_clear_camera(zn_scene* s) {
	ZN_REQUIRE(s, "zn_scene_clear_camera: scene is NULL");
	s->has_camera = false;
	return ZN_OK;
}

int zn_scene_cull(zn_scene* s, const float view[16], const float proj[16], const zn_cull_output* out) {
	ZN_REQUIRE(s && view && proj, "zn_scene_cull: NULL argument");
	if (!s->world_valid) return fail(ZN_ERR_STATE, "zn_scene_cull: no world matrices yet (call zn_scene_update or zn_scene_update_world first)");
	zn_ctx* ctx = s->ctx;
	ZN_CUDA(cudaSetDevice(ctx->device));
	uint32_t* visible = s->visible;
	uint32_t* visible_count = s->visible_count;
	const CUtensorMap* tm_mvp = &s->tm_mvp_all;
	if (out) {
		ZN_REQUIRE((out->visible_dev == nullptr) == (out->visible_count_dev == nullptr), "zn_scene_cull: pass visible_dev and visible_count_dev together");
		if (out->visible_dev) {
			ZN_REQUIRE(out->visible_capacity >= s->entity_count, "zn_scene_cull: visible_capacity %u < entity_count %u", out->visible_capacity, s->entity_count);
			visible = static_cast<uint32_t*>(out->visible_dev);
			visible_count = static_cast<uint32_t*>(out->visible_count_dev);
		}
		if (out->mvp_dev) {
			ZN_REQUIRE(reinterpret_cast<uintptr_t>(out->mvp_dev) % 16 == 0, "zn_scene_cull: mvp_dev must be 16-byte aligned");
			const int rc = mvp_map_for(s, out->mvp_dev, &tm_mvp);
			if (rc != ZN_OK) return rc;
		}
	}
	FrameConst fc;
	float vp[16];
	zn_mat4_mul(proj, view, vp);
	std::memcpy(fc.vp, vp, sizeof(fc.vp));
	zn_frustum_planes(vp, fc.planes);
	next_frame_buffers(s);
	// behind an emit kernel of this scene (a previous pass, or a fused update) the pass need not wait: see scene_cull_kernel
	// (not with peer records bound: the pass stores masks into the peers' slots, which must stay behind the expansion that
	// precedes it on the stream — zn_exchange.cpp, "Slot reuse")
	const uint32_t defer_wait = (s->tail_is_emit && !s->inputs_touched && !s->profiling && s->peer_records.empty() &&
	                             s->frame_record != s->prev_frame_record) ? 1u : 0u;
	s->tail_is_emit = false;
	const RecordTargets rec = record_targets(s);
	uint32_t* totals = s->chunk_total + size_t(s->frame_parity) * s->chunk_count;
	uint32_t* totals_next = s->chunk_total + size_t((s->frame_parity + 2u) % 3u) * s->chunk_count;
	const bool prof = s->profiling;
	if (prof) ZN_CUDA(cudaEventRecord(s->cull_events[0], ctx->stream));
	const uint32_t grid = std::min<uint32_t>(s->tile_count, uint32_t(s->max_ctas_cull));
	uint32_t* ticket = s->level_ticket + s->levels.size();
	{
		// programmatic launch behind the update (or the previous pass's emit): the prologue and the first AABB load overlap its tail
		cudaLaunchAttribute pdl;
		pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
		pdl.val.programmaticStreamSerializationAllowed = 1;
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3(grid);
		cfg.blockDim = dim3(kTile);
		cfg.dynamicSmemBytes = kCullSmem;
		cfg.stream = ctx->stream;
		cfg.attrs = &pdl;
		cfg.numAttrs = (prof || s->inputs_touched) ? 0 : 1;
		ZN_CUDA(cudaLaunchKernelEx(&cfg, scene_cull_kernel, s->tm_world_all, *tm_mvp, fc, (const uint32_t*)s->blocks,
		                           (const uint32_t*)s->tile_first_entity, s->tile_count, s->entity_count, rec,
		                           record_header_words(s->chunk_count), totals, ticket, s->cull_ticket_base, defer_wait));
	}
	s->cull_ticket_base += s->tile_count + grid;
	if (prof) ZN_CUDA(cudaEventRecord(s->cull_events[1], ctx->stream));
	const int rc = launch_emit(s, rec, totals, totals_next, visible, visible_count, 0u, !prof);
	if (rc != ZN_OK) return rc;
	if (prof) ZN_CUDA(cudaEventRecord(s->cull_events[2], ctx->stream));
	return ZN_OK;
}

int zn_scene_set_ticket_origin(zn_scene* s, uint32_t origin) {
	ZN_REQUIRE(s, "zn_scene_set_ticket_origin: scene is NULL");
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	ZN_CUDA(cudaStreamSynchronize(s->ctx->stream));
	std::vector<uint32_t> init(s->levels.size() + 1, origin);
	ZN_CUDA(cudaMemcpyAsync(s->level_ticket, init.data(), init.size() * 4, cudaMemcpyHostToDevice, s->ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(s->ctx->stream));
	for (zn_level& lv : s->levels) lv.ticket_base = origin;
	s->cull_ticket_base = origin;
	return ZN_OK;
}

int zn_scene_cull_times(zn_scene* s, float* out_cull_ms, float* out_emit_ms) {
	ZN_REQUIRE(s && out_cull_ms, "zn_scene_cull_times: NULL argument");
	if (!s->profiling || !s->cull_events[0]) return fail(ZN_ERR_STATE, "zn_scene_cull_times: profiling is not enabled");
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	ZN_CUDA(cudaEventSynchronize(s->cull_events[2]));
	ZN_CUDA(cudaEventElapsedTime(out_cull_ms, s->cull_events[0], s->cull_events[1]));
	if (out_emit_ms) ZN_CUDA(cudaEventElapsedTime(out_emit_ms, s->cull_events[1], s->cull_events[2]));
	return ZN_OK;
}

int zn_scene_set_profiling(zn_scene* s, int enable) {
	ZN_REQUIRE(s, "zn_scene_set_profiling: scene is NULL");
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	if (enable && s->prof_events.empty()) {
		s->prof_events.resize(s->levels.size() + 2);
		for (auto& e : s->prof_events) ZN_CUDA(cudaEventCreate(&e));
		for (auto& e : s->cull_events) ZN_CUDA(cudaEventCreate(&e));
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
	// one launch per level — the fused leading levels (scene_top_kernel) count as one — plus the emit kernel
	const uint32_t fused = (s->level_fusion && s->top_levels && !s->profiling) ? s->top_levels - 1u : 0u;
	*out_launches = uint32_t(s->levels.size()) - fused + 1;
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
		if ((rc = download_rows(s, 3 * k, 3, first, count, outs[k])) != ZN_OK) return rc;
	return download_rows(s, 15, 1, first, count, parent);
}

int zn_scene_build_draw(zn_scene* s, uint32_t index_count_per_instance, void* instance_mvp_dev, uint32_t instance_capacity,
                        void* draw_args_dev) {
	ZN_REQUIRE(s && instance_mvp_dev && draw_args_dev, "zn_scene_build_draw: NULL argument");
	ZN_REQUIRE(instance_capacity >= s->entity_count, "zn_scene_build_draw: instance capacity %u < entity_count %u", instance_capacity, s->entity_count);
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	const uint32_t grid = uint32_t(std::min<uint64_t>((uint64_t(s->entity_count) * 4 + 255) / 256, uint64_t(s->ctx->sm_count) * 16));
	build_draw_kernel<<<grid, 256, 0, s->ctx->stream>>>(reinterpret_cast<const float4*>(s->mvp), s->visible, s->visible_count, s->index_base,
	                                                   index_count_per_instance, static_cast<float4*>(instance_mvp_dev),
	                                                   static_cast<uint32_t*>(draw_args_dev));
	ZN_CUDA(cudaGetLastError());
	return ZN_OK;
}

int zn_scene_device_buffers(zn_scene* s, zn_scene_buffers* out) {
	ZN_REQUIRE(s && out, "zn_scene_device_buffers: NULL argument");
	out->world = s->world;
	out->mvp = s->mvp;
	out->visible = s->visible;
	out->visible_count = s->visible_count;
	out->visibility_record = s->record;
	out->visibility_record_words = s->record_words;
	out->entity_count = s->entity_count;
	return ZN_OK;
}

int zn_scene_frame_host(zn_scene* s, zn_frame_host_io* io) {
	ZN_REQUIRE(s && io, "zn_scene_frame_host: NULL argument");
	zn_ctx* ctx = s->ctx;
	const uint32_t E = s->entity_count;
	ZN_REQUIRE(io->first <= E && uint64_t(io->first) + io->count <= E, "zn_scene_frame_host: range [%u, %u+%u) exceeds entity_count %u", io->first, io->first, io->count, E);
	const uint32_t n_in = io->count ? io->count : E - io->first;
	int rc = zn_scene_set_transforms(s, io->first, n_in, io->position, io->rotation, io->scale, 12);
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

// ---- pipelined end-to-end frames --------------------------------------------------------------
// Two frames in flight over three streams: H2D of frame k+1 (upload stream) overlaps the update of
// frame k (context stream) and the D2H of frame k's visible list (download stream).  PCIe is full
// duplex, so in steady state a frame costs max(H2D, update, D2H) instead of their sum.  Each slot
// owns its staging (3 x E x 12 bytes) and its visible list / count buffers.
struct zn_pipe {
	cudaStream_t up = nullptr;
	cudaStream_t down[2] = {nullptr, nullptr};   // one per slot: frame k's list copy must not queue behind frame k+1's count copy
	uint8_t* staging[2] = {nullptr, nullptr};
	uint32_t* vis[2] = {nullptr, nullptr};
	uint32_t* cnt[2] = {nullptr, nullptr};
	uint32_t* host_cnt = nullptr;                 // pinned [2]
	cudaEvent_t h2d[2] = {nullptr, nullptr}, repacked[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, counted[2] = {nullptr, nullptr};
	uint64_t submitted = 0, collected = 0;
};

static int pipe_create(zn_scene* s) {
	if (s->pipe && s->pipe_ready) return ZN_OK;
	pipe_destroy(s);               // a half-built pipe from an earlier failed attempt
	zn_pipe* p = new zn_pipe();
	s->pipe = p;                   // owned by the scene from here on: a failure below cannot leak it
	const size_t E = s->entity_count;
	ZN_CUDA(cudaStreamCreateWithFlags(&p->up, cudaStreamNonBlocking));
	ZN_CUDA(cudaHostAlloc(&p->host_cnt, 8, cudaHostAllocDefault));
	for (int k = 0; k < 2; ++k) {
		ZN_CUDA(cudaStreamCreateWithFlags(&p->down[k], cudaStreamNonBlocking));
		ZN_CUDA(cudaMalloc(&p->staging[k], E * 36));
		ZN_CUDA(cudaMalloc(&p->vis[k], E * 4));
		ZN_CUDA(cudaMalloc(&p->cnt[k], 4));
		ZN_CUDA(cudaEventCreateWithFlags(&p->h2d[k], cudaEventDisableTiming));
		ZN_CUDA(cudaEventCreateWithFlags(&p->repacked[k], cudaEventDisableTiming));
		ZN_CUDA(cudaEventCreateWithFlags(&p->done[k], cudaEventDisableTiming));
		ZN_CUDA(cudaEventCreateWithFlags(&p->counted[k], cudaEventDisableTiming));
	}
	s->pipe_ready = true;
	return ZN_OK;
}

static void pipe_destroy(zn_scene* s) {
	zn_pipe* p = s->pipe;
	if (!p) return;
	if (p->up) cudaStreamSynchronize(p->up);
	for (int k = 0; k < 2; ++k) if (p->down[k]) cudaStreamSynchronize(p->down[k]);
	for (int k = 0; k < 2; ++k) {
		cudaFree(p->staging[k]); cudaFree(p->vis[k]); cudaFree(p->cnt[k]);
		for (cudaEvent_t e : {p->h2d[k], p->repacked[k], p->done[k], p->counted[k]})
			if (e) cudaEventDestroy(e);
	}
	if (p->host_cnt) cudaFreeHost(p->host_cnt);
	if (p->up) cudaStreamDestroy(p->up);
	for (int k = 0; k < 2; ++k) if (p->down[k]) cudaStreamDestroy(p->down[k]);
	delete p;
	s->pipe = nullptr;
	s->pipe_ready = false;
}

int zn_scene_frame_host_submit(zn_scene* s, const zn_frame_host_io* io) {
	ZN_REQUIRE(s && io, "zn_scene_frame_host_submit: NULL argument");
	if (!s->has_camera) return fail(ZN_ERR_STATE, "zn_scene_frame_host_submit: no camera set");
	ZN_REQUIRE(s->levels.size() <= 63, "zn_scene_frame_host_submit: at most 63 hierarchy levels");
	const uint32_t E32 = s->entity_count;
	ZN_REQUIRE(io->first <= E32 && uint64_t(io->first) + io->count <= E32, "zn_scene_frame_host_submit: range [%u, %u+%u) exceeds entity_count %u",
	           io->first, io->first, io->count, E32);
	zn_ctx* ctx = s->ctx;
	ZN_CUDA(cudaSetDevice(ctx->device));
	int rc = pipe_create(s);
	if (rc != ZN_OK) return rc;
	zn_pipe* p = s->pipe;
	if (p->submitted - p->collected >= 2) return fail(ZN_ERR_STATE, "zn_scene_frame_host_submit: two frames are already in flight; call zn_scene_frame_host_wait");
	const int k = int(p->submitted & 1);
	const uint32_t n_in = io->count ? io->count : E32 - io->first;
	const size_t E = s->entity_count, bytes = size_t(n_in) * 12;
	const float* src[3] = {io->position, io->rotation, io->scale};
	const bool any = (src[0] || src[1] || src[2]) && n_in > 0;
	if (any) {
		// upload stream: wait until the slot's previous contents have been repacked, then one copy per array that is present
		if (p->submitted >= 2) ZN_CUDA(cudaStreamWaitEvent(p->up, p->repacked[k], 0));
		for (int a = 0; a < 3; ++a)
			if (src[a]) ZN_CUDA(cudaMemcpyAsync(p->staging[k] + a * E * 12, src[a], bytes, cudaMemcpyHostToDevice, p->up));
		ZN_CUDA(cudaEventRecord(p->h2d[k], p->up));
		// context stream: ONE launch scatters the arrays into the tile-major blocks
		ZN_CUDA(cudaStreamWaitEvent(ctx->stream, p->h2d[k], 0));
		RepackLevels lv = {};
		lv.levels = uint32_t(s->levels.size());
		for (size_t l = 0; l < s->levels.size(); ++l) { lv.begin[l] = s->levels[l].begin; lv.tile_base[l] = s->levels[l].tile_base; }
		lv.begin[s->levels.size()] = E32;
		auto at = [&](int a) { return src[a] ? reinterpret_cast<const uint32_t*>(p->staging[k] + a * E * 12) : nullptr; };
		repack_frame_kernel<<<(n_in + 255) / 256, 256, 0, ctx->stream>>>(at(0), at(1), at(2), io->first, n_in, s->blocks, lv);
		ZN_CUDA(cudaGetLastError());
		s->inputs_touched = true;
	}
	ZN_CUDA(cudaEventRecord(p->repacked[k], ctx->stream));
	uint32_t* keep_vis = s->visible;
	uint32_t* keep_cnt = s->visible_count;
	s->visible = p->vis[k];
	s->visible_count = p->cnt[k];
	rc = zn_scene_update(s);
	s->visible = keep_vis;
	s->visible_count = keep_cnt;
	if (rc != ZN_OK) return rc;
	ZN_CUDA(cudaEventRecord(p->done[k], ctx->stream));
	// download stream: the count first (the list's size is not known on the host yet)
	ZN_CUDA(cudaStreamWaitEvent(p->down[k], p->done[k], 0));
	ZN_CUDA(cudaMemcpyAsync(p->host_cnt + k, p->cnt[k], 4, cudaMemcpyDeviceToHost, p->down[k]));
	ZN_CUDA(cudaEventRecord(p->counted[k], p->down[k]));
	p->submitted++;
	return ZN_OK;
}

int zn_scene_frame_host_wait(zn_scene* s, zn_frame_host_io* io) {
	ZN_REQUIRE(s && io, "zn_scene_frame_host_wait: NULL argument");
	zn_pipe* p = s->pipe;
	if (!p || p->collected == p->submitted) return fail(ZN_ERR_STATE, "zn_scene_frame_host_wait: no frame in flight");
	ZN_CUDA(cudaSetDevice(s->ctx->device));
	const int k = int(p->collected & 1);
	ZN_CUDA(cudaEventSynchronize(p->counted[k]));
	io->visible_count = p->host_cnt[k];
	if (io->visible && io->visible_count) {
		ZN_CUDA(cudaMemcpyAsync(io->visible, p->vis[k], size_t(io->visible_count) * 4, cudaMemcpyDeviceToHost, p->down[k]));
		ZN_CUDA(cudaStreamSynchronize(p->down[k]));
	}
	p->collected++;
	return ZN_OK;
}

} // extern "C"
