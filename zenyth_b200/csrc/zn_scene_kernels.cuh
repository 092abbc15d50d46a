// zn_scene_kernels.cuh — device code of the scene path (included by zn_scene.cu only).
//
//
//   scene_level_kernel<HAS_PARENT>  one launch per hierarchy level (K1/K2 of SURVEY.md §2):
//       TRS -> local -> world (= parent.world * local) -> MVP -> AABB-vs-frustum flag, plus the
//       tile's 256-bit visibility mask.
//   scene_top_kernel                the leading levels whose tiles fit one CTA's shared memory, in ONE launch in
//       place of level 0: parents are read from the shared tiles, not through L2.
//   scene_cull_kernel               one camera pass over existing world matrices (zn_scene_cull).
//   emit_visible_kernel             one launch per frame (K3): ordered visible list from the masks,
//       single pass — a chunk's base is the sum of the chunk totals before it, then an ascending
//       warp-cooperative scatter; also fills the header of the frame's visibility record.
//   expand_visible_kernel           multi-GPU: all-gathered visibility records -> every rank's list
//       (expand_classify_kernel + expand_work_kernel: the same through a work list, for sparse sets).
//   patch_rows_kernel / repack*     uploads: a handful of entities as kernel parameters, ranges through staging.
//
// scene_level_kernel is PERSISTENT (4 CTAs of 256 threads per SM, tiles claimed by atomic ticket)
// and software-pipelined with the TMA unit:
//   - a tile's inputs are ONE cp.async.bulk (SASS UBLKCP) of its 16 KiB tile-major block, plus one
//     bulk copy of its contiguous parent-matrix range; thread 0 issues the copies for tile n+1 as
//     soon as all threads have pulled tile n into registers, so HBM latency hides behind the
//     sincos / matrix maths of tile n;
//   - a thread owns one entity; its two 64-byte matrices go to shared memory in the TMA
//     64-byte-swizzle pattern (conflict-free 16-byte stores) and the tile leaves as two
//     cp.async.bulk.tensor stores (SASS UTMASTG) that drain while tile n+1 is computed.
// No thread issues a global load or store for entity data; HBM sees 16 KiB bursts both ways.
//
// Data layout in HBM (DESIGN.md "Layout"): inputs are tile-major blocks — for every 256-entity
// tile one contiguous 16 KiB record of 16 rows x 256 lanes (tx,ty,tz, pitch,yaw,roll, sx,sy,sz,
// aabb centre xyz, aabb half xyz as float, parent as u32).  Tiles never straddle a hierarchy
// level.  Outputs are AoS float[E][16] (what mat4::data() consumers expect,
// Core/include/math/matrix.hpp:60-61), clipped at the level end by a per-level tensor map.
#pragma once
#include "zn_internal.hpp"
#include "zn_math.cuh"
#include "zn_ptx.cuh"

#include <algorithm>
#include <cstring>


namespace zn {

struct FrameConst {
	float vp[16];
	float planes[24];
};

constexpr int kWarpsPerTile = kTile / 32;                  // 8
constexpr uint32_t kInBytes = kBlockWords * 4;             // 16 KiB input block
constexpr uint32_t kParSlots = 64;                         // parent matrices staged per tile
constexpr uint32_t kParBytes = kParSlots * 64;             // 4 KiB
constexpr uint32_t kOutBytes = 2 * kTile * 64;             // world tile + mvp tile
constexpr uint32_t kLevelSmem = kOutBytes + kInBytes + kParBytes + 1024;
constexpr uint32_t kWorldInBytes = 10u * kTile * 4u;       // WITH_VIEW = false stages rows 0..8 and the parent row only (10 KiB)
constexpr uint32_t kWorldSmem = kTile * 64 + kWorldInBytes + kParBytes + 1024;   // no MVP tile: 31 KiB, six CTAs per SM
constexpr uint32_t kGather = 0xFFFFFFFFu;                  // tile_prange.y: parents too scattered to stage
constexpr int kChunkTiles = 32;                            // tiles per emit CTA (8192 entities)
constexpr uint32_t kSmallLevelTiles = 4096;                // levels below 1Mi entities leave their peer stores to the emit kernel

// Where a frame's visibility record is written: ptr[0] is this rank's own copy; ptr[1..n) are the
// copies kept by peer GPUs (device pointers into their memory, mapped over NVLink through CUDA
// IPC).  The kernels store to all of them, which fuses the all-gather of the records into the
// compute kernels tile by tile.
constexpr int kMaxRecordTargets = 8;
struct RecordTargets {
	uint32_t* ptr[kMaxRecordTargets];
	int n;
};

// The completion signal of that fused all-gather, also without a collective: `value` — a frame
// sequence number, monotonically increasing — stored into one flag word in every rank's memory.
// It is raised by the first CTA of the kernel that FOLLOWS the update on the stream (the expansion
// of the previous frame): the kernel boundary has completed every store of the level and emit
// kernels, so one system-scope fence and one release store per peer are all a frame pays.  (Raising
// it from the emit kernel itself — every CTA fences, the last one to finish signals — was measured
// 13 us per frame slower: 2048 system-scope fences behind the scatter stores.)
struct SignalTargets {
	uint32_t* flag[kMaxRecordTargets];
	int n;
	uint32_t value;
};

// Raise every flag of `sig` to its value: ONE system-scope fence (everything this thread has
// observed — the completed update before this kernel on the stream — is ordered before the flags),
// then relaxed system-scope stores.  (st.release.sys per flag compiles to a MEMBAR.ALL.SYS each.)
__device__ __forceinline__ void raise_flags(const SignalTargets& sig) {
	__threadfence_system();
	for (int q = 0; q < sig.n; ++q) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(sig.flag[q]), "r"(sig.value) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
	uint64_t t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// WITH_VIEW = false is the camera-less update (zn_scene_update_world): TRS -> world only.  It pulls
// rows 0..8 and the parent row of a block (10 KiB instead of 16), stores the world tile and writes
// neither MVPs nor masks; a later scene_cull_kernel pass supplies those per camera.
template <bool HAS_PARENT, bool WITH_VIEW>
__global__ void __launch_bounds__(kTile, WITH_VIEW ? 4 : 6)
scene_level_kernel(const __grid_constant__ CUtensorMap tm_world, const __grid_constant__ CUtensorMap tm_mvp,
                   const __grid_constant__ FrameConst fc, const uint32_t* __restrict__ blocks,
                   const uint2* __restrict__ tile_prange, const float* __restrict__ world_rd,
                   uint32_t begin, uint32_t end, uint32_t tile_base, uint32_t tiles,
                   const __grid_constant__ RecordTargets rec, uint32_t header_words, uint32_t* __restrict__ chunk_total,
                   uint32_t* __restrict__ ticket, uint32_t ticket_base, uint32_t defer_wait) {
	extern __shared__ uint8_t smem_raw[];
	__shared__ uint64_t s_full;
	__shared__ uint2 s_meta;
	__shared__ uint32_t s_mask[kWarpsPerTile];
	__shared__ uint32_t s_tile[2];   // level-relative tile of iteration n (slot n&1) and n+1

	// 1024-byte aligned tiles (the swizzle pattern is a function of the shared address bits)
	uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	constexpr uint32_t out_bytes = WITH_VIEW ? kOutBytes : uint32_t(kTile) * 64u;
	float4* out_world = reinterpret_cast<float4*>(smem);
	float4* out_mvp = reinterpret_cast<float4*>(smem + kTile * 64);   // WITH_VIEW only
	const uint32_t* in_blk = reinterpret_cast<const uint32_t*>(smem + out_bytes);
	const float4* in_par = reinterpret_cast<const float4*>(smem + out_bytes + (WITH_VIEW ? kInBytes : kWorldInBytes));
	constexpr int kParentRow = WITH_VIEW ? 15 : 9;   // where the block's parent row is staged

	const int t = threadIdx.x;
	const int lane = t & 31;
	const int warp = t >> 5;

	// thread 0 is the TMA issuer: loads of tile `tl` into the (single) input stage.  `first_tile`:
	// the input block does not depend on the kernel before this one on the stream (the previous
	// level, or — for level 0 — the previous frame's emit) but the parent matrices, the masks and
	// the chunk totals do, so under programmatic dependent launch the first block load is issued
	// before griddepcontrol.wait and everything else after it.
	auto issue_loads = [&](uint32_t tl, bool first_tile) {
		const uint32_t gt = tile_base + tl;
		uint2 pr = make_uint2(0u, 0u);
		if (HAS_PARENT) pr = __ldg(tile_prange + gt);
		const bool staged = HAS_PARENT && pr.y != 0u && pr.y != kGather;
		s_meta = pr;
		if (WITH_VIEW) {
			mbar_arrive_expect_tx(&s_full, kInBytes + (staged ? pr.y * 64u : 0u));
			bulk_g2s(const_cast<uint32_t*>(in_blk), blocks + size_t(gt) * kBlockWords, kInBytes, &s_full);
		} else {
			mbar_arrive_expect_tx(&s_full, 10u * kTile * 4u + (staged ? pr.y * 64u : 0u));
			bulk_g2s(const_cast<uint32_t*>(in_blk), blocks + size_t(gt) * kBlockWords, 9u * kTile * 4u, &s_full);
			bulk_g2s(const_cast<uint32_t*>(in_blk) + kParentRow * kTile, blocks + size_t(gt) * kBlockWords + 15 * kTile, kTile * 4u, &s_full);
		}
		// everything from here on is ordered behind the kernel before this one on the stream — unless this launch touches
		// nothing that kernel reads or writes (defer_wait: level 0 of a frame running beside the previous frame's emit)
		if (first_tile && !defer_wait) grid_dep_wait();
		if (staged) bulk_g2s(const_cast<float4*>(in_par), world_rd + size_t(pr.x) * 16, pr.y * 64u, &s_full);
	};

	if (t == 0) {
		mbar_init(&s_full, 1);
		fence_mbar_init();
		prefetch_tensormap(&tm_world);
		prefetch_tensormap(&tm_mvp);
		grid_dep_launch();          // the next level's kernel may start its own prologue now
		// Tiles are claimed dynamically (one atomic ticket per tile): a CTA that starts late — e.g.
		// while a collective's kernel holds part of the SMs — simply takes fewer tiles.  The grid is
		// never larger than the tile count, so every CTA gets at least one tile.
		const uint32_t tl0 = atomicAdd(ticket, 1u) - ticket_base;
		s_tile[0] = tl0;
		if (tl0 < tiles) issue_loads(tl0, true);
	}
	__syncthreads();

	for (uint32_t n = 0;; ++n) {
		const uint32_t tl = s_tile[n & 1u];
		if (tl >= tiles) break;
		// ---- tile n: shared memory -> registers ------------------------------------------------
		mbar_wait(&s_full, n & 1u);
		float v[15];
#pragma unroll
		for (int k = 0; k < (WITH_VIEW ? 15 : 9); ++k) v[k] = __uint_as_float(in_blk[k * kTile + t]);
		const uint32_t p = HAS_PARENT ? in_blk[kParentRow * kTile + t] : ZN_NO_PARENT;
		Affine P;
		if (HAS_PARENT && p != ZN_NO_PARENT) {
			const uint2 pr = s_meta;
			float4 c0, c1, c2, c3;
			if (pr.y != kGather) {
				const float4* src = in_par + size_t(p - pr.x) * 4;
				c0 = src[0]; c1 = src[1]; c2 = src[2]; c3 = src[3];
			} else {
				const float4* src = reinterpret_cast<const float4*>(world_rd) + size_t(p) * 4;
				c0 = __ldg(src); c1 = __ldg(src + 1); c2 = __ldg(src + 2); c3 = __ldg(src + 3);
			}
			P.m[0][0] = c0.x; P.m[1][0] = c0.y; P.m[2][0] = c0.z;
			P.m[0][1] = c1.x; P.m[1][1] = c1.y; P.m[2][1] = c1.z;
			P.m[0][2] = c2.x; P.m[1][2] = c2.y; P.m[2][2] = c2.z;
			P.m[0][3] = c3.x; P.m[1][3] = c3.y; P.m[2][3] = c3.z;
		}
		if (t == 0) bulk_wait_read_all();   // tile n-1's tensor stores have finished reading the out tile
		__syncthreads();                    // input stage consumed by everyone; out tile free

		// ---- prefetch tile n+1 while tile n is computed ------------------------------------------
		if (t == 0) {
			const uint32_t next = atomicAdd(ticket, 1u) - ticket_base;
			s_tile[(n + 1u) & 1u] = next;   // read by everyone after the barrier below
			if (next < tiles) issue_loads(next, false);
		}

		// ---- tile n: maths in registers ----------------------------------------------------------
		const Affine local = local_trs(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
		Affine w = local;
		if (HAS_PARENT && p != ZN_NO_PARENT) w = affine_mul(P, local);
		const uint32_t first = begin + tl * kTile;
		if (WITH_VIEW) {
			const bool vis = (first + t < end) && aabb_visible(w, v[9], v[10], v[11], v[12], v[13], v[14], fc.planes);
			const uint32_t mask = __ballot_sync(0xffffffffu, vis);
			if (lane == 0) s_mask[warp] = mask;
		}

		// ---- results -> swizzled shared tile (chunk c of row r lives at chunk c ^ ((r>>1)&3)) ------
		{
			const int sw = (t >> 1) & 3;
			float4* wr = out_world + t * 4;
			wr[0 ^ sw] = make_float4(w.m[0][0], w.m[1][0], w.m[2][0], 0.0f);
			wr[1 ^ sw] = make_float4(w.m[0][1], w.m[1][1], w.m[2][1], 0.0f);
			wr[2 ^ sw] = make_float4(w.m[0][2], w.m[1][2], w.m[2][2], 0.0f);
			wr[3 ^ sw] = make_float4(w.m[0][3], w.m[1][3], w.m[2][3], 1.0f);
			if (WITH_VIEW) {
				float m[16];
				mvp_mul(fc.vp, w, m);
				float4* mr = out_mvp + t * 4;
				mr[0 ^ sw] = make_float4(m[0], m[1], m[2], m[3]);
				mr[1 ^ sw] = make_float4(m[4], m[5], m[6], m[7]);
				mr[2 ^ sw] = make_float4(m[8], m[9], m[10], m[11]);
				mr[3 ^ sw] = make_float4(m[12], m[13], m[14], m[15]);
			}
		}
		fence_proxy_async_smem();
		__syncthreads();
		if (t == 0) {
			tensor_store_2d(&tm_world, out_world, 0, int32_t(first));
			if (WITH_VIEW) tensor_store_2d(&tm_mvp, out_mvp, 0, int32_t(first));
			bulk_commit();
		}
		if (!WITH_VIEW) continue;
		if (warp == 0) {
			const uint32_t gt = tile_base + tl;
			const uint32_t word = (lane < kWarpsPerTile) ? s_mask[lane] : 0u;
			if (lane < kWarpsPerTile) rec.ptr[0][header_words + size_t(gt) * kWarpsPerTile + lane] = word;
			const uint32_t cnt = __reduce_add_sync(0xffffffffu, __popc(word));
			if (lane == 0) atomicAdd(chunk_total + gt / kChunkTiles, cnt);
		} else if (warp == kWarpsPerTile - 1 && rec.n > 1) {
			// ... and, over NVLink, straight into the copy of this rank's record that every peer keeps
			// (fused all-gather: plain stores to mapped peer memory).  Issued by the last warp so the
			// TMA-issuing warp never queues behind remote stores, and spread over its lanes — lane l
			// carries word l % 8 to peer 1 + l / 8 (+ 4 in a second round) — so seven peers cost two
			// store instructions per tile instead of seven on the CTA's critical path (7 us per frame
			// at N = 8, measured with local targets).
			const uint32_t word = s_mask[lane & 7];
			const size_t at = header_words + size_t(tile_base + tl) * kWarpsPerTile + (lane & 7);
			const int q0 = 1 + (lane >> 3);
			if (q0 < rec.n) rec.ptr[q0][at] = word;
			if (q0 + 4 < rec.n) rec.ptr[q0 + 4][at] = word;
		}
	}
	if (t == 0) {
		bulk_wait_read_all();
		// defer_wait: the dependency on the kernel before this one is only one of ORDER — the kernels after this one wait
		// for this grid to complete, and through this wait for everything before it (the previous frame's emit)
		if (defer_wait) grid_dep_wait();
	}
}

// scene_top_kernel — the TOP OF THE TREE in one launch of ONE CTA.  The leading levels of a hierarchy are
// tiny (7, 56, 448 entities in the bench scene) and each costs a level kernel its whole latency chain —
// wait for the level before, pull the parents back from L2, compute, store, complete — about 6 us per
// level however few entities it has.  Here the leading levels whose tiles (at most kTopTiles in total)
// fit one CTA's shared memory are composed back to back: every input block is requested up front (they
// depend on nothing), the world tiles stay in shared memory in the swizzled layout they are stored
// from, and a child reads its parent's matrix from there — no round trip through L2 between levels,
// only a block barrier.  Results leave exactly as the level kernel's do (tensor stores through the
// per-level maps, mask words, chunk total), bit for bit: same device functions, same operands.
// Launched in place of level 0 (same programmatic dependency, same defer_wait rule), so behind an emit
// kernel the whole top of the tree hides in that kernel's shadow.
constexpr int kTopTiles = 5;
constexpr uint32_t kTopSmemView = uint32_t(kTopTiles) * (kTile * 64u + kInBytes) + 2u * kTile * 64u + 1024u;   // 193 KiB
constexpr uint32_t kTopSmemWorld = uint32_t(kTopTiles) * (kTile * 64u + kInBytes) + 1024u;                     // 161 KiB
struct TopLevels {
	uint32_t levels;              // fused levels, >= 2 (levels <= tiles <= kTopTiles)
	uint32_t tiles;               // their tiles in total; tile i of the scene is tile i of the run (level 0 starts at tile 0)
	uint32_t begin[kTopTiles];    // entity range of each fused level
	uint32_t end[kTopTiles];
};
struct TopMaps {
	CUtensorMap world[kTopTiles];   // per level, clipped at the level end
	CUtensorMap mvp[kTopTiles];
};

template <bool WITH_VIEW>
__global__ void __launch_bounds__(kTile, 1)
scene_top_kernel(const __grid_constant__ TopMaps maps, const __grid_constant__ FrameConst fc, const __grid_constant__ TopLevels top,
                 const uint32_t* __restrict__ blocks, uint32_t* __restrict__ record, uint32_t header_words,
                 uint32_t* __restrict__ chunk_total, uint32_t defer_wait) {
	extern __shared__ uint8_t smem_raw[];
	__shared__ uint64_t s_full[kTopTiles];
	__shared__ uint32_t s_mask[2][kWarpsPerTile];

	uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	float4* world_tiles = reinterpret_cast<float4*>(smem);                                        // [kTopTiles][kTile * 4], swizzled
	float4* mvp_tiles = reinterpret_cast<float4*>(smem + kTopTiles * kTile * 64);                 // [2][kTile * 4], WITH_VIEW only
	const uint32_t* in_blocks = reinterpret_cast<const uint32_t*>(smem + kTopTiles * kTile * 64 + (WITH_VIEW ? 2 * kTile * 64 : 0));

	const int t = threadIdx.x;
	const int lane = t & 31;
	const int warp = t >> 5;

	if (t == 0) {
		for (uint32_t i = 0; i < top.tiles; ++i) mbar_init(&s_full[i], 1);
		fence_mbar_init();
		for (uint32_t l = 0; l < top.levels; ++l) {
			prefetch_tensormap(&maps.world[l]);
			if (WITH_VIEW) prefetch_tensormap(&maps.mvp[l]);
		}
		grid_dep_launch();
		// the input blocks depend on nothing before this kernel on the stream (an upload makes this a normal launch)
		for (uint32_t i = 0; i < top.tiles; ++i) {
			uint32_t* dst = const_cast<uint32_t*>(in_blocks) + size_t(i) * kBlockWords;
			const uint32_t* src = blocks + size_t(i) * kBlockWords;
			if (WITH_VIEW) {
				mbar_arrive_expect_tx(&s_full[i], kInBytes);
				bulk_g2s(dst, src, kInBytes, &s_full[i]);
			} else {
				mbar_arrive_expect_tx(&s_full[i], 10u * kTile * 4u);
				bulk_g2s(dst, src, 9u * kTile * 4u, &s_full[i]);
				bulk_g2s(dst + 15 * kTile, src + 15 * kTile, kTile * 4u, &s_full[i]);
			}
		}
		// everything this kernel WRITES is ordered behind the kernel before it — see scene_level_kernel for defer_wait
		if (!defer_wait) grid_dep_wait();
	}
	__syncthreads();

	uint32_t i = 0;          // tile of the run == global tile
	uint32_t total = 0;      // visible entities of the run (warp 0, lane 0); all its tiles belong to chunk 0
	for (uint32_t l = 0; l < top.levels; ++l) {
		const uint32_t begin = top.begin[l], end = top.end[l];
		for (uint32_t first = begin; first < end; first += kTile, ++i) {
			mbar_wait(&s_full[i], 0);
			const uint32_t* in_blk = in_blocks + size_t(i) * kBlockWords;
			float v[15];
#pragma unroll
			for (int k = 0; k < (WITH_VIEW ? 15 : 9); ++k) v[k] = __uint_as_float(in_blk[k * kTile + t]);
			const bool live = first + t < end;
			const uint32_t p = (l > 0 && live) ? in_blk[15 * kTile + t] : ZN_NO_PARENT;
			Affine P;
			if (p != ZN_NO_PARENT) {
				// the parent lives in an earlier level (upload contract), i.e. in one of the tiles already composed
				uint32_t pl = l - 1;
				while (p < top.begin[pl]) --pl;
				uint32_t pt = 0;
				for (uint32_t q = 0; q < pl; ++q) pt += (top.end[q] - top.begin[q] + kTile - 1) / kTile;
				const uint32_t row = p - top.begin[pl];
				const float4* src = world_tiles + (size_t(pt) + (row >> 8)) * (kTile * 4) + (row & 255u) * 4;
				const int sw = ((row & 255u) >> 1) & 3;
				const float4 c0 = src[0 ^ sw], c1 = src[1 ^ sw], c2 = src[2 ^ sw], c3 = src[3 ^ sw];
				P.m[0][0] = c0.x; P.m[1][0] = c0.y; P.m[2][0] = c0.z;
				P.m[0][1] = c1.x; P.m[1][1] = c1.y; P.m[2][1] = c1.z;
				P.m[0][2] = c2.x; P.m[1][2] = c2.y; P.m[2][2] = c2.z;
				P.m[0][3] = c3.x; P.m[1][3] = c3.y; P.m[2][3] = c3.z;
			}
			if (WITH_VIEW) {
				// the MVP tile is double-buffered: the stores of tile i-2 must have finished reading this one
				if (t == 0) bulk_wait_read_but<1>();
				__syncthreads();
			}
			const Affine local = local_trs(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
			Affine w = local;
			if (p != ZN_NO_PARENT) w = affine_mul(P, local);
			if (WITH_VIEW) {
				const bool vis = live && aabb_visible(w, v[9], v[10], v[11], v[12], v[13], v[14], fc.planes);
				const uint32_t mask = __ballot_sync(0xffffffffu, vis);
				if (lane == 0) s_mask[i & 1u][warp] = mask;
			}
			float4* out_world = world_tiles + size_t(i) * (kTile * 4);
			float4* out_mvp = mvp_tiles + size_t(i & 1u) * (kTile * 4);
			{
				const int sw = (t >> 1) & 3;
				float4* wr = out_world + t * 4;
				wr[0 ^ sw] = make_float4(w.m[0][0], w.m[1][0], w.m[2][0], 0.0f);
				wr[1 ^ sw] = make_float4(w.m[0][1], w.m[1][1], w.m[2][1], 0.0f);
				wr[2 ^ sw] = make_float4(w.m[0][2], w.m[1][2], w.m[2][2], 0.0f);
				wr[3 ^ sw] = make_float4(w.m[0][3], w.m[1][3], w.m[2][3], 1.0f);
				if (WITH_VIEW) {
					float m[16];
					mvp_mul(fc.vp, w, m);
					float4* mr = out_mvp + t * 4;
					mr[0 ^ sw] = make_float4(m[0], m[1], m[2], m[3]);
					mr[1 ^ sw] = make_float4(m[4], m[5], m[6], m[7]);
					mr[2 ^ sw] = make_float4(m[8], m[9], m[10], m[11]);
					mr[3 ^ sw] = make_float4(m[12], m[13], m[14], m[15]);
				}
			}
			fence_proxy_async_smem();
			__syncthreads();             // the tile is complete: its children may read it, the TMA unit may store it
			if (t == 0) {
				tensor_store_2d(&maps.world[l], out_world, 0, int32_t(first));
				if (WITH_VIEW) tensor_store_2d(&maps.mvp[l], out_mvp, 0, int32_t(first));
				bulk_commit();
			}
			if (WITH_VIEW && warp == 0) {
				const uint32_t word = (lane < kWarpsPerTile) ? s_mask[i & 1u][lane] : 0u;
				if (lane < kWarpsPerTile) record[header_words + size_t(i) * kWarpsPerTile + lane] = word;
				total += __reduce_add_sync(0xffffffffu, __popc(word));
			}
		}
	}
	if (WITH_VIEW && t == 0 && total) atomicAdd(chunk_total, total);
	if (t == 0) {
		bulk_wait_read_all();
		if (defer_wait) grid_dep_wait();
	}
}

// scene_cull_kernel — a camera pass over world matrices that already exist (zn_scene_cull): what
// Scene::Cull(camera) runs from OnRender (Core/src/Application.cpp:49-51) after OnUpdate (:47) has
// composed the hierarchy, and again for every further view (shadow pass, second camera) without
// redoing a sincos or rewriting a world matrix.  ONE persistent launch over all tiles of all levels
// (the world matrices are complete, so levels no longer order anything):
//   in   the tile's 256 world matrices — one 2-D tensor load in the 64-byte swizzle the level kernel
//        stored them with, so the per-thread 16-byte reads are conflict-free — and rows 9..14 of its
//        input block (AABB centre + half extent, 6 KiB, one bulk copy);
//   out  the MVP tile (tensor store), the tile's 256-bit mask, the chunk total.
// 152 algorithmic bytes per entity (64 + 24 in, 64 out) + the emit kernel's 4 per visible entity.
// The tensor maps span all E rows: the last tile of a level also covers the first entities of the
// next level, whose MVPs it computes from their real world matrices — bit for bit what the owning
// tile stores — so the overlap is harmless; only the mask is clipped to the tile's own entities.
constexpr uint32_t kAabbBytes = 6u * kTile * 4u;
constexpr uint32_t kCullSmem = 2u * kTile * 64u + kAabbBytes + 1024u;

__global__ void __launch_bounds__(kTile, 5)
scene_cull_kernel(const __grid_constant__ CUtensorMap tm_world, const __grid_constant__ CUtensorMap tm_mvp,
                  const __grid_constant__ FrameConst fc, const uint32_t* __restrict__ blocks,
                  const uint32_t* __restrict__ tile_first_entity, uint32_t tiles, uint32_t entity_count,
                  const __grid_constant__ RecordTargets rec, uint32_t header_words, uint32_t* __restrict__ chunk_total,
                  uint32_t* __restrict__ ticket, uint32_t ticket_base, uint32_t defer_wait) {
	extern __shared__ uint8_t smem_raw[];
	__shared__ uint64_t s_full;
	__shared__ uint2 s_meta[2];      // {first entity, entities} of the tile of iteration n (slot n&1) and n+1
	__shared__ uint32_t s_mask[kWarpsPerTile];
	__shared__ uint32_t s_tile[2];

	uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	const float4* in_world = reinterpret_cast<const float4*>(smem);
	float4* out_mvp = reinterpret_cast<float4*>(smem + kTile * 64);
	const float* in_aabb = reinterpret_cast<const float*>(smem + 2 * kTile * 64);

	const int t = threadIdx.x;
	const int lane = t & 31;
	const int warp = t >> 5;

	// `first_tile`: the AABB rows do not depend on the kernel before this one on the stream (the update
	// that composed the world matrices), the world matrices do: under programmatic dependent launch
	// only the second copy waits for it.
	auto issue_loads = [&](uint32_t gt, uint32_t slot, bool first_tile) {
		const uint32_t first = __ldg(tile_first_entity + gt);
		const uint32_t next = (gt + 1 < tiles) ? __ldg(tile_first_entity + gt + 1) : entity_count;
		s_meta[slot] = make_uint2(first, next - first);
		mbar_arrive_expect_tx(&s_full, kTile * 64u + kAabbBytes);
		bulk_g2s(const_cast<float*>(in_aabb), blocks + size_t(gt) * kBlockWords + 9 * kTile, kAabbBytes, &s_full);
		// defer_wait: the kernel before this one is an emit — the world matrices were complete before IT started, and this
		// pass shares no buffer with it (next record slot, next set of totals) — so the two run side by side
		if (first_tile && !defer_wait) grid_dep_wait();
		tensor_load_2d(const_cast<float4*>(in_world), &tm_world, 0, int32_t(first), &s_full);
	};

	if (t == 0) {
		mbar_init(&s_full, 1);
		fence_mbar_init();
		prefetch_tensormap(&tm_world);
		prefetch_tensormap(&tm_mvp);
		grid_dep_launch();
		const uint32_t g0 = atomicAdd(ticket, 1u) - ticket_base;
		s_tile[0] = g0;
		if (g0 < tiles) issue_loads(g0, 0, true);
	}
	__syncthreads();

	for (uint32_t n = 0;; ++n) {
		const uint32_t gt = s_tile[n & 1u];
		if (gt >= tiles) break;
		mbar_wait(&s_full, n & 1u);
		const uint2 meta = s_meta[n & 1u];
		const int sw = (t >> 1) & 3;
		const float4 c0 = in_world[t * 4 + (0 ^ sw)], c1 = in_world[t * 4 + (1 ^ sw)], c2 = in_world[t * 4 + (2 ^ sw)],
		             c3 = in_world[t * 4 + (3 ^ sw)];
		float a[6];
#pragma unroll
		for (int k = 0; k < 6; ++k) a[k] = in_aabb[k * kTile + t];
		if (t == 0) bulk_wait_read_all();   // tile n-1's tensor store has finished reading the out tile
		__syncthreads();                    // input stage consumed by everyone; out tile free
		if (t == 0) {
			const uint32_t next = atomicAdd(ticket, 1u) - ticket_base;
			s_tile[(n + 1u) & 1u] = next;
			if (next < tiles) issue_loads(next, (n + 1u) & 1u, false);
		}
		Affine w;
		w.m[0][0] = c0.x; w.m[1][0] = c0.y; w.m[2][0] = c0.z;
		w.m[0][1] = c1.x; w.m[1][1] = c1.y; w.m[2][1] = c1.z;
		w.m[0][2] = c2.x; w.m[1][2] = c2.y; w.m[2][2] = c2.z;
		w.m[0][3] = c3.x; w.m[1][3] = c3.y; w.m[2][3] = c3.z;
		const bool vis = (uint32_t(t) < meta.y) && aabb_visible(w, a[0], a[1], a[2], a[3], a[4], a[5], fc.planes);
		const uint32_t mask = __ballot_sync(0xffffffffu, vis);
		if (lane == 0) s_mask[warp] = mask;
		{
			float m[16];
			mvp_mul(fc.vp, w, m);
			float4* mr = out_mvp + t * 4;
			mr[0 ^ sw] = make_float4(m[0], m[1], m[2], m[3]);
			mr[1 ^ sw] = make_float4(m[4], m[5], m[6], m[7]);
			mr[2 ^ sw] = make_float4(m[8], m[9], m[10], m[11]);
			mr[3 ^ sw] = make_float4(m[12], m[13], m[14], m[15]);
		}
		fence_proxy_async_smem();
		__syncthreads();
		if (t == 0) {
			tensor_store_2d(&tm_mvp, out_mvp, 0, int32_t(meta.x));
			bulk_commit();
		}
		if (warp == 0) {
			// lane l carries mask word l % 8 to record target l / 8 (+ 4 in a second round): own copy and up to 7 peers
			const uint32_t word = s_mask[lane & 7];
			const size_t at = header_words + size_t(gt) * kWarpsPerTile + (lane & 7);
			const int q0 = lane >> 3;
			if (q0 < rec.n) rec.ptr[q0][at] = word;
			if (q0 + 4 < rec.n) rec.ptr[q0 + 4][at] = word;
			const uint32_t cnt = __reduce_add_sync(0xffffffffu, (lane < kWarpsPerTile) ? __popc(word) : 0u);
			if (lane == 0) atomicAdd(chunk_total + gt / kChunkTiles, cnt);
		}
	}
	if (t == 0) {
		bulk_wait_read_all();
		if (defer_wait) grid_dep_wait();   // keeps the chain of completions: whoever waits for this grid has waited for that emit
	}
}

// One chunk (32 tiles = 8192 entities) of an ordered visible list: `totals` holds every chunk's
// visible count, `masks` the 256-bit tile masks.  The CTA gets its base offset by summing the
// totals of the chunks before it — a few KiB of L2 reads, no serial dependency between CTAs —
// then scatters its set bits in ascending order.  Returns this chunk's own count (in every thread).
// Visibility record of a frame (what a multi-GPU exchange ships, 1 bit per entity):
//   [0, chunks)                  exclusive base of every chunk in the visible list
//   [chunks]                     total visible count
//   [chunks+1, chunks+1+8*chunks) exclusive offset of every 4-tile group within its chunk
//   (padding to a multiple of 4 words)
//   [header, header + 8*tiles)   the 256-bit tile masks
__host__ __device__ inline uint32_t record_header_words(uint32_t chunks) { return (chunks * 9u + 1u + 3u) & ~3u; }

// The scatter of one warp's 32 mask words (4 tiles), staged in shared memory as words w4[8] and
// list offsets o4[8] (four per LDS.128, broadcast) with the tiles' first indices in first[4].
// Lane l owns bit l of every word, so the stores of a word are one coalesced ascending run.
// STRAIGHT-LINE code: one predicated store per word and nothing else — no fast paths, no branches.
// A warp that scatters alone is bound by the latency of its instruction stream, not by issue slots:
// the earlier form with all-ones / all-zero fast paths compiled to a reconvergence region per word
// (BSSY ... BSYNC, ~30 dependent instructions each) and took ~3 us per group; 32 independent
// test-popc-add-store chains overlap freely.  (Full GROUPS never get here: the callers write them as
// runs of 1024 indices straight from the counts.)
__device__ __forceinline__ void scatter_warp_words(const uint4* __restrict__ w4, const uint4* __restrict__ o4,
                                                   const uint32_t* __restrict__ first, int lane, uint32_t lt,
                                                   uint32_t* __restrict__ visible) {
#pragma unroll
	for (int g = 0; g < 8; ++g) {   // words 4g..4g+3 belong to tile g/2 of this warp
		const uint4 m = w4[g];
		const uint4 o = o4[g];
		const uint32_t f = first[g >> 1] + (g & 1) * 128 + lane;
		st_global_if(visible + o.x + __popc(m.x & lt), f, (m.x >> lane) & 1u);
		st_global_if(visible + o.y + __popc(m.y & lt), f + 32, (m.y >> lane) & 1u);
		st_global_if(visible + o.z + __popc(m.z & lt), f + 64, (m.z >> lane) & 1u);
		st_global_if(visible + o.w + __popc(m.w & lt), f + 96, (m.w >> lane) & 1u);
	}
}

// 1024 consecutive indices v0, v0+1, ... into out[0..1024) by one warp: what a group of four fully
// visible tiles contributes.  Full tiles are contiguous in entity index (also across a level
// boundary: a level whose last tile is full ends where the next one begins), so neither the masks
// nor a scan are needed — the counts in the record header already say "all of them".
__device__ __forceinline__ void store_run_1024(uint32_t* __restrict__ out, uint32_t v0, int lane) {
	// `out` is only 4-byte aligned in general (a list position is a count of visible entities — at the
	// bench's headline camera every group starts 2 words off a 16-byte boundary): peel `a` leading
	// words so that the body leaves as 16-byte stores, 255 of them, and finish with the 4 - a trailing
	// words.  (Falling back to 32-bit stores for unaligned runs cost a quarter of the write bandwidth.)
	const uint32_t a = (4u - ((uint32_t(reinterpret_cast<uintptr_t>(out)) >> 2) & 3u)) & 3u;   // warp-uniform
	uint32_t* body = out + a;
	const uint32_t vb = v0 + a;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const uint32_t i = uint32_t(k) * 32u + lane;           // 16-byte slot of the body
		const uint32_t v = vb + i * 4u;
		if (a == 0u || i < 255u) *reinterpret_cast<uint4*>(body + i * 4u) = make_uint4(v, v + 1, v + 2, v + 3);
	}
	if (a != 0u) {
		if (uint32_t(lane) < a) out[lane] = v0 + lane;                                          // head
		else if (uint32_t(lane) < 4u) out[1020u + lane] = v0 + 1020u + lane;                    // tail: words 1020+a .. 1023
	}
}

__device__ __forceinline__ uint32_t emit_chunk(const uint32_t* __restrict__ totals, const uint32_t* __restrict__ masks,
                                               const uint32_t* __restrict__ tile_first_entity, uint32_t tiles, uint32_t chunks,
                                               uint32_t chunk, uint32_t index_base, uint32_t* __restrict__ visible,
                                               uint32_t* __restrict__ visible_count, const RecordTargets& record_out,
                                               uint32_t remote_tiles_below) {
	__shared__ uint32_t s_part[8];
	__shared__ uint32_t s_warp_excl[8];
	__shared__ uint32_t s_base, s_total;
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

	// thread t owns mask word t of the chunk: tile chunk*32 + t/8, word t%8.  Loaded up front, together with
	// the totals (one round of loads instead of two; a full or empty chunk simply does not use it).
	const uint32_t tile = chunk * kChunkTiles + (t >> 3);
	const uint32_t word = (tile < tiles) ? __ldcg(masks + size_t(tile) * kWarpsPerTile + (t & 7)) : 0u;
	const uint32_t ctot = __ldcg(totals + chunk);
	// first entity of this warp's 4-tile group: only a full chunk uses it, but fetching it here keeps a dependent
	// load off the path between the block barriers and the stores
	const uint32_t group_tile = chunk * kChunkTiles + uint32_t(warp) * 4u;
	const uint32_t group_first = (group_tile < tiles) ? __ldg(tile_first_entity + group_tile) : 0u;
	const uint32_t tile_first = ((t & 7) == 0 && tile < tiles) ? __ldg(tile_first_entity + tile) + index_base : 0u;   // mixed chunks
	uint32_t acc = 0;
	for (uint32_t j = t; j < chunk; j += 256) acc += __ldcg(totals + j);
	acc = __reduce_add_sync(0xffffffffu, acc);
	if (lane == 0) s_part[warp] = acc;

	// Culling is spatially coherent, so most chunks are entirely invisible or entirely visible, and
	// the chunk total (accumulated by the level kernels) says so before a single mask word is read:
	// the header is then known in closed form and the list segment is one run of 8192 indices.
	// (Chunks whose masks this kernel must push to peers take the general path: it loads them.)
	const bool pushes_masks = record_out.n > 1 && chunk * kChunkTiles < remote_tiles_below;
	if ((ctot == 0u || ctot == uint32_t(kChunkTiles * kTile)) && !pushes_masks) {
		__syncthreads();
		if (warp == 0) {
			const uint32_t base = __reduce_add_sync(0xffffffffu, (lane < 8) ? s_part[lane] : 0u);
			if (lane < 8) {
				const uint32_t group_off = ctot ? uint32_t(lane) * 4u * kTile : 0u;
				for (int q = 0; q < record_out.n; ++q) record_out.ptr[q][chunks + 1u + chunk * 8u + lane] = group_off;
			}
			if (lane == 0) {
				s_base = base;
				for (int q = 0; q < record_out.n; ++q) record_out.ptr[q][chunk] = base;
				if (chunk == chunks - 1) {
					*visible_count = base + ctot;
					for (int q = 0; q < record_out.n; ++q) record_out.ptr[q][chunks] = base + ctot;
				}
			}
		}
		if (ctot == 0u) return 0u;
		__syncthreads();
		store_run_1024(visible + s_base + uint32_t(warp) * 4u * kTile, group_first + index_base, lane);
		return ctot;
	}

	// Masks of the SMALL leading levels reach the peers from here (coalesced, one kernel boundary)
	// instead of from their level kernels, whose few microseconds would each end on an NVLink round
	// trip; the big levels push their own masks tile by tile while they run.
	if (tile < remote_tiles_below)
		for (int q = 1; q < record_out.n; ++q) record_out.ptr[q][record_header_words(chunks) + size_t(tile) * kWarpsPerTile + (t & 7)] = word;
	const uint32_t cnt = __popc(word);
	uint32_t incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += y;
	}
	if (lane == 31) s_warp_excl[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		const uint32_t wt = (lane < 8) ? s_warp_excl[lane] : 0u;
		uint32_t winc = wt;
#pragma unroll
		for (int d = 1; d < 8; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, winc, d);
			if (lane >= d) winc += y;
		}
		const uint32_t base = __reduce_add_sync(0xffffffffu, (lane < 8) ? s_part[lane] : 0u);
		const uint32_t total = __shfl_sync(0xffffffffu, winc, 7);
		if (lane < 8) {
			s_warp_excl[lane] = winc - wt;
			for (int q = 0; q < record_out.n; ++q) record_out.ptr[q][chunks + 1u + chunk * 8u + lane] = winc - wt;
		}
		if (lane == 0) {
			s_base = base;
			s_total = total;
			for (int q = 0; q < record_out.n; ++q) record_out.ptr[q][chunk] = base;
			if (chunk == chunks - 1) {
				*visible_count = base + total;
				for (int q = 0; q < record_out.n; ++q) record_out.ptr[q][chunks] = base + total;
			}
		}
	}
	__syncthreads();
	const uint32_t my_off = s_base + s_warp_excl[warp] + incl - cnt;   // exclusive offset of this thread's word
	// warp-cooperative scatter: a warp owns 32 words (4 tiles); every lane tests ITS bit of a word,
	// so the stores of a word are one coalesced run in ascending order.  The words and offsets are
	// staged in shared memory and fetched four at a time (one LDS.128 each, broadcast) — the loop
	// is issue-bound, not bandwidth-bound, so instructions per word are what count.
	__shared__ __align__(16) uint32_t s_word[256];
	__shared__ __align__(16) uint32_t s_off[256];
	__shared__ uint32_t s_first[kChunkTiles];
	s_word[t] = word;
	s_off[t] = my_off;
	if ((t & 7) == 0) s_first[t >> 3] = tile_first;
	__syncwarp();
	const uint32_t lt = (1u << lane) - 1u;
	const uint4* w4 = reinterpret_cast<const uint4*>(s_word + warp * 32);
	const uint4* o4 = reinterpret_cast<const uint4*>(s_off + warp * 32);
	scatter_warp_words(w4, o4, s_first + warp * 4, lane, lt, visible);
	return s_total;
}

// K3: the frame's own ordered visible list.  The level kernels accumulated every chunk's visible
// count (chunk_total, one atomicAdd per tile).  Each CTA also stores the offsets it derived (chunk
// base, group offsets) into the header of the frame's visibility record — the 1-bit-per-entity
// form of the result that a multi-GPU exchange ships instead of the 32-bit-per-visible-entity
// list.  The accumulators are double-buffered by frame parity: while reading this frame's set,
// every CTA clears its own entry of the OTHER set (last read by the previous frame's emit), so the
// next frame finds zeros without any cross-CTA handshake.
__global__ void __launch_bounds__(256)
emit_visible_kernel(const __grid_constant__ RecordTargets rec, const uint32_t* __restrict__ tile_first_entity,
                    uint32_t tiles, uint32_t chunks, const uint32_t* __restrict__ chunk_total, uint32_t* __restrict__ chunk_total_next,
                    uint32_t index_base, uint32_t* __restrict__ visible, uint32_t* __restrict__ visible_count, uint32_t remote_tiles_below) {
	const uint32_t chunk = blockIdx.x;
	grid_dep_wait();   // masks and totals of every level are complete and visible
	// The next frame's first kernel (level 0, or a cull pass) may start its prologue and first input
	// load now; it calls griddepcontrol.wait itself before it writes a mask or touches a chunk total.
	grid_dep_launch();
	// rec.ptr[0] is this rank's own record (the masks are read from it); the header goes to all targets
	emit_chunk(chunk_total, rec.ptr[0] + record_header_words(chunks), tile_first_entity, tiles, chunks, chunk, index_base, visible, visible_count, rec,
	           remote_tiles_below);
	if (threadIdx.x == 0) chunk_total_next[chunk] = 0u;
}

// Signal + wait as a kernel of their own (one thread per flag), for comparison with the form fused
// into expand_visible_kernel: the kernel boundary orders the expansion after it.  A flag that does
// not arrive within the timeout stores 1 + its index into *status; an expansion launched with that
// status word (zn_scene_expand_visible after zn_scene_wait_flags) then expands nothing.
__global__ void wait_flags_kernel(const uint32_t* __restrict__ flags, uint32_t flag_stride, uint32_t count, uint32_t expected,
                                  uint64_t timeout_ns, uint32_t* __restrict__ status, const __grid_constant__ SignalTargets sig) {
	const uint32_t r = threadIdx.x;
	if (r == 0 && sig.n > 0) raise_flags(sig);
	if (r >= count) return;
	const uint64_t t0 = global_timer_ns();
	while (int32_t(ld_acquire_sys(flags + size_t(r) * flag_stride) - expected) < 0) {
		if (global_timer_ns() - t0 > timeout_ns) {
			atomicCAS(status, 0u, 1u + r);
			break;
		}
		__nanosleep(200);
	}
}

// Expands visibility records (this rank's and its peers', as all-gathered) into their ordered
// index lists: lists[r*list_stride ...], counts[r].  Same arithmetic as emit_visible_kernel, so
// every rank derives bit-identical lists from the 27x smaller records.
struct ExpandBases {
	uint32_t base[64];
};
constexpr int kExpandChunks = 1;   // chunks per CTA of the expansion grid (4, i.e. 32-warp CTAs, measured slower: 40 / 70 us)
constexpr int kExpandWarps = 8 * kExpandChunks;   // one 4-tile group per warp

// A record already carries every offset its owner computed (chunk bases, group offsets), so a WARP
// is autonomous — no block barrier, no shared-memory reduction.  One CTA per (record, chunk), one
// 4-tile group (1024 entities) per warp.  A warp reads four header words (the chunk's base and end,
// the group's offset and the next one — L1-cached, shared by the CTA's warps and its neighbours); an
// empty group exits on them; a full group is written as one run of 1024 consecutive indices straight
// from the header (16-byte stores whatever the run's alignment); only a mixed group fetches its 32
// mask words, scans them and scatters with the straight-line scatter_warp_words.
//
// What was measured on the way here (7 records of the 16.7M-entity scene; camera inside, f = 0.21:
// 70% of the chunks empty, 14% of the groups mixed with 27 of their 32 words non-zero, a plain fill
// of the same bytes takes 16.5 us / headline camera, f = 0.86, nearly every group full or empty,
// fill 56 us).  Round 1's kernel — this structure, but header -> offsets -> masks as three dependent
// load rounds, 32-bit stores for unaligned runs and a scatter with all-ones / all-zero fast paths
// that compiled to a reconvergence region per word (a lone warp retires a dependent instruction
// every ~10 cycles: ~3 us per mixed group) —                                              37 / 64 us
//   persistent warps drawing chunks from one GLOBAL ticket counter (same-address atomics, ~4 ns each)  67 / 78
//   one warp per chunk, eight inlined copies of the scatter (27k instructions: instruction cache)     65 / 77
//   one warp per chunk, loop not unrolled, straight-line predicated scatter (~1.1 us per group)        36 / 78
//   eight chunks per CTA, scouted, non-empty groups shared through a shared-memory work list           37 / 80
//   persistent warps, static stride over the groups, 1 or 8 items in flight per warp (LDGSTS ring)     39-41 / 80-84
//   this structure, single speculative load round, L2-only (.cg) loads                                  39 / 68
//   this structure, L1-cached header loads, masks fetched by mixed groups only (what is below)         36 / 64
//   ... and the first-entity load (static table) ahead of the flag acquire and the header: 32 registers
//   instead of 40, eight CTAs per SM (what is below)                                                   33 / 62
// None of the coarser-grained or persistent forms beat a wide grid of small, short-lived warps: the
// kernel is bound by how many independent 4 KiB store runs and scatters are in flight, and at f = 0.86
// ncu shows the persistent form at 42% issue utilisation and 54% DRAM throughput with MIO / LSU
// throttle stalls (profiles/r02_expand_persistent_ncu.csv).  A single record takes 9 us (round 1: 11),
// three take 16.5 us (round 1: 20): the per-record cost above that is the stream of 16k mostly empty
// warps per record through the SMs.
//
// Signalled form (flags != nullptr): thread 0 of the CTA waits — ld.acquire.sys, bounded — for the
// flag of the record's owner.  If the flag does not arrive within the timeout the record is NOT
// expanded: 1 + r goes to *status (first failure wins), counts[r] becomes 0, so a half-written
// record never reaches the scatter.  `abort_status` (plain form after zn_scene_wait_flags): a non-zero word there means
// that wait failed; nothing is expanded and all counts are zeroed.
__global__ void __launch_bounds__(kExpandWarps * 32)
expand_visible_kernel(const uint32_t* __restrict__ records, uint32_t record_stride, const uint32_t* __restrict__ tile_first_entity,
                      uint32_t tiles, uint32_t chunks, const __grid_constant__ ExpandBases bases, uint32_t record_count, int skip, int skip_count,
                      uint32_t* __restrict__ lists, uint32_t list_stride, uint32_t* __restrict__ counts,
                      const uint32_t* __restrict__ flags, uint32_t flag_stride, uint32_t expected, uint32_t records_per_flag,
                      uint64_t timeout_ns, uint32_t* __restrict__ status, const uint32_t* __restrict__ abort_status,
                      const __grid_constant__ SignalTargets sig) {
	__shared__ __align__(16) uint32_t s_word[kExpandWarps][32];
	__shared__ __align__(16) uint32_t s_off[kExpandWarps][32];
	__shared__ uint32_t s_first[kExpandWarps][4];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (sig.n > 0 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
		// the update that precedes this kernel on the stream is complete, peer stores included: tell every rank.  (Launched
		// programmatically — zn_exchange_expand — this is the one thread that waits for that update; a no-op otherwise.)
		grid_dep_wait();
		raise_flags(sig);
	}
	uint32_t r = blockIdx.y;                                         // one grid row per expanded record
	if (int(r) >= skip) r += uint32_t(skip_count);
	if (r >= record_count) return;                                   // every record skipped: this launch only signals
	const uint32_t chunk = blockIdx.x * kExpandChunks + (warp >> 3);
	if (chunk >= chunks) return;
	const uint32_t group = chunk * 8u + (warp & 7);
	const bool first = group == 0u;                                  // the warp that also delivers the record's count
	if (abort_status && __ldcg(abort_status) != 0u) {
		if (first && lane == 0) counts[r] = 0u;
		return;
	}
	const uint32_t* rec = records + size_t(r) * record_stride;
	// What does not depend on the owner's flag goes first, so that its latency runs beside the flag's: the first entity of
	// this lane's tile comes from a static table.  (With it hoisted the kernel needs 32 registers instead of 40 and the
	// signalled form costs what the plain one does: 61.8 us for seven records, was 68 against 63.  Prefetching the header
	// lines towards L2 ahead of the acquire as well measured no different, at N = 1 and at N = 8.)
	const uint32_t tile = group * 4u + (lane >> 3);
	const uint32_t tfirst = (tile < tiles) ? __ldg(tile_first_entity + tile) + bases.base[r] : 0u;
	if (flags) {
		// ONE acquire per CTA (thread 0; bar.sync extends it to the CTA): ld.acquire.sys is LDG.STRONG.SYS +
		// CCTL.IVALL in SASS — it invalidates the SM's L1, which the header loads below want to hit
		__shared__ uint32_t s_ok;
		if (threadIdx.x == 0) {
			uint32_t ok = 1u;
			const uint32_t* f = flags + size_t(r / records_per_flag) * flag_stride;
			const uint64_t t0 = global_timer_ns();
			while (int32_t(ld_acquire_sys(f) - expected) < 0) {
				// someone else already gave up on this record, or the time is up
				if (*reinterpret_cast<volatile uint32_t*>(status) == 1u + r || global_timer_ns() - t0 > timeout_ns) {
					atomicCAS(status, 0u, 1u + r);
					ok = 0u;
					break;
				}
				__nanosleep(200);
			}
			s_ok = ok;
		}
		__syncthreads();
		if (!s_ok) {
			if (first && lane == 0) counts[r] = 0u;
			return;
		}
	}
	uint32_t* visible = lists + size_t(r) * list_stride;
	// ---- header ----------------------------------------------------------------------------------------------
	// lane 0 the chunk's base, lane 1 the next chunk's (entry `chunks` is the total), lane 2 this group's offset,
	// lane 3 the next group's (the header holds 8 offsets for every chunk), lane 4 the record's total (first warp only).
	// Plain (L1-cached) loads: the eight warps of a CTA and the next CTAs on this SM read the same few lines, so
	// most EMPTY groups — 70% of them at f = 0.21 — learn that they are empty at L1 latency and leave.  The
	// record is complete before these loads by the memory model itself: it was written by kernels that
	// finished before this one started (plain form), or is read after the acquire of its owner's flag above.
	uint32_t h = 0u;
	if (lane < 2) h = rec[chunk + lane];
	else if (lane < 4 && group + (lane - 2) < chunks * 8u) h = rec[chunks + 1u + group + (lane - 2)];
	else if (lane == 4 && first) h = rec[chunks];
	if (first && lane == 4) counts[r] = h;
	if (group * 4u >= tiles) return;
	const uint32_t chunk_base = __shfl_sync(0xffffffffu, h, 0);
	const uint32_t chunk_count = __shfl_sync(0xffffffffu, h, 1) - chunk_base;
	const uint32_t group_off = __shfl_sync(0xffffffffu, h, 2);
	const uint32_t group_end = ((warp & 7) < 7) ? __shfl_sync(0xffffffffu, h, 3) : chunk_count;
	const uint32_t group_count = group_end - group_off;
	if (chunk_count > uint32_t(kChunkTiles * kTile) || group_count == 0u || group_count > 4u * kTile) return;   // empty (or not a record)
	const uint32_t base = chunk_base + group_off;
	if (group_count == 4u * kTile) {
		store_run_1024(visible + base, __shfl_sync(0xffffffffu, tfirst, 0), lane);
		return;
	}
	const uint32_t word = (tile < tiles) ? rec[record_header_words(chunks) + size_t(tile) * kWarpsPerTile + (lane & 7)] : 0u;
	const uint32_t cnt = __popc(word);
	uint32_t incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += y;
	}
	s_word[warp][lane] = word;
	s_off[warp][lane] = base + incl - cnt;
	if ((lane & 7) == 0) s_first[warp][lane >> 3] = tfirst;
	__syncwarp();
	scatter_warp_words(reinterpret_cast<const uint4*>(s_word[warp]), reinterpret_cast<const uint4*>(s_off[warp]), s_first[warp], lane,
	                   (1u << lane) - 1u, visible);
}

// ---- the expansion as two launches: classify, then work only where there is work ---------------------
// expand_classify_kernel: one THREAD per (record, 4-tile group).  It reads the group's four header words
// (coalesced across the CTA), and every non-empty group is appended to a work list as
// {record << 26 | full << 25 | group, list position} — one global atomicAdd per CTA (a block scan
// places the entries).  115k groups of seven records are 448 CTAs instead of 115k warps.
// expand_work_kernel: a persistent grid of autonomous warps walks the work list with a static stride,
// the next entry in flight while the current one is handled: a full group is one run of 1024
// consecutive indices, a mixed group fetches its 32 mask words, scans and scatters.  No warp is ever
// spent on an empty group, and consecutive entries go to different warps, so a run of mixed chunks
// spreads over the whole GPU.  The two launches are chained with programmatic dependent launch; the
// list's counter is double-buffered by launch parity (the work kernel clears the other one).
constexpr int kClassifyThreads = 256;
struct ExpandWork {
	uint4* entries;        // [capacity]: {record << 26 | full << 25 | group, list position, first entity index (+ index base) of the group, -}
	uint32_t* count;       // [2], by launch parity
	uint32_t parity;
};

__global__ void __launch_bounds__(kClassifyThreads)
expand_classify_kernel(const uint32_t* __restrict__ records, uint32_t record_stride, const uint32_t* __restrict__ tile_first_entity,
                       uint32_t tiles, uint32_t chunks, const __grid_constant__ ExpandBases bases,
                       uint32_t record_count, int skip, int skip_count, uint32_t* __restrict__ counts,
                       const uint32_t* __restrict__ flags, uint32_t flag_stride, uint32_t expected, uint32_t records_per_flag,
                       uint64_t timeout_ns, uint32_t* __restrict__ status, const uint32_t* __restrict__ abort_status,
                       const __grid_constant__ SignalTargets sig, const __grid_constant__ ExpandWork work) {
	__shared__ uint32_t s_warp[kClassifyThreads / 32];
	__shared__ uint32_t s_slot, s_ok;
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	if (t == 0) grid_dep_launch();                               // the work kernel may start its prologue
	if (sig.n > 0 && blockIdx.x == 0 && blockIdx.y == 0 && t == 0) {
		// the update that precedes this kernel on the stream is complete, peer stores included: tell every rank.  (Launched
		// programmatically — zn_exchange_expand — this is the one thread that waits for that update; a no-op otherwise.)
		grid_dep_wait();
		raise_flags(sig);
	}
	uint32_t r = blockIdx.y;                                         // one grid row per expanded record
	if (int(r) >= skip) r += uint32_t(skip_count);
	if (r >= record_count) return;                                   // every record skipped: this launch only signals
	const bool first = blockIdx.x == 0 && t == 0;
	if (t == 0) {
		uint32_t ok = 1u;
		if (abort_status && __ldcg(abort_status) != 0u) ok = 0u;
		if (ok && flags) {
			const uint32_t* f = flags + size_t(r / records_per_flag) * flag_stride;
			const uint64_t t0 = global_timer_ns();
			while (int32_t(ld_acquire_sys(f) - expected) < 0) {
				if (*reinterpret_cast<volatile uint32_t*>(status) == 1u + r || global_timer_ns() - t0 > timeout_ns) {
					atomicCAS(status, 0u, 1u + r);
					ok = 0u;
					break;
				}
				__nanosleep(200);
			}
		}
		s_ok = ok;
	}
	__syncthreads();
	if (!s_ok) {
		if (first) counts[r] = 0u;
		return;
	}
	const uint32_t* rec = records + size_t(r) * record_stride;
	if (first) counts[r] = rec[chunks];
	const uint32_t group = blockIdx.x * kClassifyThreads + t;
	const uint32_t chunk = group >> 3;
	uint32_t gcount = 0u, position = 0u, first_entity = 0u;
	if (group < chunks * 8u && group * 4u < tiles) {
		first_entity = __ldg(tile_first_entity + group * 4u) + bases.base[r];
		const uint32_t chunk_base = rec[chunk];
		const uint32_t chunk_count = rec[chunk + 1u] - chunk_base;
		const uint32_t group_off = rec[chunks + 1u + group];
		const uint32_t group_end = ((group & 7u) < 7u) ? rec[chunks + 2u + group] : chunk_count;
		gcount = group_end - group_off;
		if (chunk_count > uint32_t(kChunkTiles * kTile) || gcount > 4u * kTile) gcount = 0u;   // not a record: expand nothing
		position = chunk_base + group_off;
	}
	// block scan of the non-empty flags -> slot of every entry, one global atomic per CTA
	const uint32_t mine = gcount != 0u ? 1u : 0u;
	const uint32_t ballot = __ballot_sync(0xffffffffu, mine);
	if (lane == 0) s_warp[warp] = __popc(ballot);
	__syncthreads();
	if (warp == 0) {
		uint32_t v = (lane < kClassifyThreads / 32) ? s_warp[lane] : 0u;
		uint32_t incl = v;
#pragma unroll
		for (int d = 1; d < kClassifyThreads / 32; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d) incl += y;
		}
		if (lane < kClassifyThreads / 32) s_warp[lane] = incl - v;
		const uint32_t total = __shfl_sync(0xffffffffu, incl, kClassifyThreads / 32 - 1);
		if (lane == 0) s_slot = total ? atomicAdd(work.count + work.parity, total) : 0u;
	}
	__syncthreads();
	if (mine) {
		const uint32_t slot = s_slot + s_warp[warp] + __popc(ballot & ((1u << lane) - 1u));
		work.entries[slot] = make_uint4((r << 26) | (gcount == 4u * kTile ? (1u << 25) : 0u) | group, position, first_entity, 0u);
	}
}

__global__ void __launch_bounds__(kExpandWarps * 32)
expand_work_kernel(const uint32_t* __restrict__ records, uint32_t record_stride, const uint32_t* __restrict__ tile_first_entity,
                   uint32_t tiles, uint32_t chunks, const __grid_constant__ ExpandBases bases, uint32_t* __restrict__ lists,
                   uint32_t list_stride, const __grid_constant__ ExpandWork work) {
	__shared__ __align__(16) uint32_t s_word[kExpandWarps][32];
	__shared__ __align__(16) uint32_t s_off[kExpandWarps][32];
	__shared__ uint32_t s_first[kExpandWarps][4];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t workers = gridDim.x * kExpandWarps;
	const uint32_t worker = blockIdx.x * kExpandWarps + warp;
	const uint32_t lt = (1u << lane) - 1u;
	grid_dep_wait();                                                 // the classification is complete and visible
	const uint32_t n = __ldcg(work.count + work.parity);
	if (blockIdx.x == 0 && threadIdx.x == 0) work.count[work.parity ^ 1u] = 0u;   // last read by the previous launch pair
	if (worker >= n) return;
	// two entries ahead of the one being handled: a full group needs nothing but its entry
	const uint4 none = make_uint4(0u, 0u, 0u, 0u);
	uint4 cur = __ldcg(work.entries + worker);
	uint4 nxt = (worker + workers < n) ? __ldcg(work.entries + worker + workers) : none;
	for (uint32_t i = worker; i < n; i += workers) {
		const uint32_t i2 = i + 2u * workers;
		const uint4 nxt2 = (i2 < n) ? __ldcg(work.entries + i2) : none;
		const uint32_t r = cur.x >> 26, group = cur.x & 0x1FFFFFFu;
		uint32_t* visible = lists + size_t(r) * list_stride;
		if (cur.x & (1u << 25)) {
			store_run_1024(visible + cur.y, cur.z, lane);
		} else {
			const uint32_t tile = group * 4u + (lane >> 3);
			const uint32_t* rec = records + size_t(r) * record_stride;
			const uint32_t word = (tile < tiles) ? __ldcg(rec + record_header_words(chunks) + size_t(tile) * kWarpsPerTile + (lane & 7)) : 0u;
			const uint32_t tfirst = (tile < tiles) ? __ldg(tile_first_entity + tile) + bases.base[r] : 0u;
			const uint32_t cnt = __popc(word);
			uint32_t incl = cnt;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
				if (lane >= d) incl += y;
			}
			__syncwarp();
			s_word[warp][lane] = word;
			s_off[warp][lane] = cur.y + incl - cnt;
			if ((lane & 7) == 0) s_first[warp][lane >> 3] = tfirst;
			__syncwarp();
			scatter_warp_words(reinterpret_cast<const uint4*>(s_word[warp]), reinterpret_cast<const uint4*>(s_off[warp]), s_first[warp], lane, lt, visible);
		}
		cur = nxt;
		nxt = nxt2;
	}
}

// Visible list -> what a draw submission consumes: the MVP matrices of the
// visible entities packed in list order, and one indexed-indirect argument record in the
// D3D12_DRAW_INDEXED_ARGUMENTS layout {IndexCountPerInstance, InstanceCount, StartIndexLocation,
// BaseVertexLocation, StartInstanceLocation}.  Four threads move one 64-byte matrix, so both the
// gather (ascending entity order) and the packed store are full 16-byte accesses; the count is
// read on the device, no host round trip.
__global__ void __launch_bounds__(256)
build_draw_kernel(const float4* __restrict__ mvp, const uint32_t* __restrict__ visible, const uint32_t* __restrict__ visible_count,
                  uint32_t index_base, uint32_t index_count_per_instance, float4* __restrict__ instance_mvp,
                  uint32_t* __restrict__ draw_args) {
	const uint32_t count = *visible_count;
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		draw_args[0] = index_count_per_instance;
		draw_args[1] = count;
		draw_args[2] = 0u;
		draw_args[3] = 0u;
		draw_args[4] = 0u;
	}
	const uint64_t chunks = uint64_t(count) * 4;
	for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < chunks; i += uint64_t(gridDim.x) * blockDim.x) {
		const uint32_t entity = visible[i >> 2] - index_base;
		instance_mvp[i] = __ldg(mvp + size_t(entity) * 4 + (i & 3));
	}
}

// {first parent, parent count} of every tile: one warp per tile over the parent row of its block.
__global__ void __launch_bounds__(256)
tile_prange_kernel(const uint32_t* __restrict__ blocks, uint32_t tiles, uint2* __restrict__ out) {
	const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (tile >= tiles) return;
	const int lane = threadIdx.x & 31;
	const uint32_t* row = blocks + size_t(tile) * kBlockWords + 15 * kTile;
	uint32_t lo = ZN_NO_PARENT, hi = 0;
#pragma unroll
	for (int k = 0; k < kTile / 32; ++k) {
		const uint32_t p = row[k * 32 + lane];
		lo = min(lo, p);
		hi = max(hi, p == ZN_NO_PARENT ? 0u : p);
	}
	lo = __reduce_min_sync(0xffffffffu, lo);
	hi = __reduce_max_sync(0xffffffffu, hi);
	if (lane == 0) {
		uint2 r = make_uint2(0u, 0u);
		if (lo != ZN_NO_PARENT) r = make_uint2(lo, (hi - lo < kParSlots) ? hi - lo + 1u : kGather);
		out[tile] = r;
	}
}

__global__ void init_parent_rows_kernel(uint32_t* __restrict__ blocks, uint32_t tiles) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < tiles * kTile) blocks[size_t(i / kTile) * kBlockWords + 15 * kTile + (i % kTile)] = ZN_NO_PARENT;
}

// Strided host records (3 words every `stride_w` words) of level-relative entities
// [rel_first, rel_first+count) -> rows row0..row0+2 of the tile-major blocks.
__global__ void repack3_kernel(const uint32_t* __restrict__ src, uint32_t stride_w, uint32_t count,
                               uint32_t* __restrict__ blocks, uint32_t tile_base, uint32_t rel_first, int row0) {
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= count) return;
	const uint32_t rel = rel_first + j;
	uint32_t* d = blocks + size_t(tile_base + rel / kTile) * kBlockWords + row0 * kTile + (rel % kTile);
	const uint32_t* s = src + size_t(j) * stride_w;
	d[0] = s[0]; d[kTile] = s[1]; d[2 * kTile] = s[2];
}
// A HANDFUL of entities (the few that a frame's game logic moved): the values travel in the kernel's
// parameters — no staging copy, one launch for position, rotation and scale together.
constexpr uint32_t kPatchEntities = 16;
struct PatchRows {
	uint32_t at[kPatchEntities];          // word offset of the entity's lane in row 0 of its block
	float v[kPatchEntities][9];           // up to nine consecutive block rows per entity
	uint32_t count;
	uint32_t row0;                        // v[e][r] belongs to block row row0 + r (0: position, rotation, scale; 9: AABB centre, half extent)
	uint32_t rows;                        // bit r: v[e][r] is present
};
__global__ void patch_rows_kernel(uint32_t* __restrict__ blocks, const __grid_constant__ PatchRows patch) {
	const uint32_t e = threadIdx.x / 9u, r = threadIdx.x % 9u;
	if (e < patch.count && ((patch.rows >> r) & 1u)) blocks[size_t(patch.at[e]) + (patch.row0 + r) * kTile] = __float_as_uint(patch.v[e][r]);
}
// One launch for a whole host frame (zn_scene_frame_host_submit): up to three packed [count][3]
// arrays (position / rotation / scale, NULL = absent) of entities [first, first+count) into rows
// 0..8 of the tile-major blocks, across every level the range touches.
struct RepackLevels {
	uint32_t begin[64];      // first entity of level l; begin[levels] = entity_count
	uint32_t tile_base[64];
	uint32_t levels;
};
__global__ void __launch_bounds__(256)
repack_frame_kernel(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ rot, const uint32_t* __restrict__ scl,
                    uint32_t first, uint32_t count, uint32_t* __restrict__ blocks, const __grid_constant__ RepackLevels lv) {
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= count) return;
	const uint32_t e = first + j;
	uint32_t l = 0;
	while (l + 1 < lv.levels && e >= lv.begin[l + 1]) ++l;
	const uint32_t rel = e - lv.begin[l];
	uint32_t* d = blocks + size_t(lv.tile_base[l] + rel / kTile) * kBlockWords + (rel % kTile);
	const size_t s = size_t(j) * 3;
	if (pos) { d[0 * kTile] = pos[s]; d[1 * kTile] = pos[s + 1]; d[2 * kTile] = pos[s + 2]; }
	if (rot) { d[3 * kTile] = rot[s]; d[4 * kTile] = rot[s + 1]; d[5 * kTile] = rot[s + 2]; }
	if (scl) { d[6 * kTile] = scl[s]; d[7 * kTile] = scl[s + 1]; d[8 * kTile] = scl[s + 2]; }
}
__global__ void repack1_kernel(const uint32_t* __restrict__ src, uint32_t count, uint32_t* __restrict__ blocks,
                               uint32_t tile_base, uint32_t rel_first, int row) {
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= count) return;
	const uint32_t rel = rel_first + j;
	blocks[size_t(tile_base + rel / kTile) * kBlockWords + row * kTile + (rel % kTile)] = src[j];
}
// rows row0..row0+nrows-1 of the blocks -> packed [count][nrows] words
__global__ void unpack_kernel(const uint32_t* __restrict__ blocks, uint32_t tile_base, uint32_t rel_first, uint32_t count,
                              int row0, int nrows, uint32_t* __restrict__ dst) {
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= count) return;
	const uint32_t rel = rel_first + j;
	const uint32_t* s = blocks + size_t(tile_base + rel / kTile) * kBlockWords + row0 * kTile + (rel % kTile);
	for (int r = 0; r < nrows; ++r) dst[size_t(j) * nrows + r] = s[r * kTile];
}


} // namespace zn
