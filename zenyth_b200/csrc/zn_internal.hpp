// zn_internal.hpp — host-side plumbing shared by the translation units of libzenyth_b200.so.
#pragma once
#include "../../include/zenyth_b200.h"

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

namespace zn {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define ZN_CUDA(expr)                                                                             \
	do {                                                                                          \
		cudaError_t zn_e_ = (expr);                                                               \
		if (zn_e_ != cudaSuccess)                                                                 \
			return ::zn::fail(ZN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(zn_e_), __FILE__, __LINE__); \
	} while (0)

#define ZN_REQUIRE(cond, ...)                                  \
	do {                                                       \
		if (!(cond)) return ::zn::fail(ZN_ERR_INVALID, __VA_ARGS__); \
	} while (0)

// Entities (or vertices /4) handled by one CTA.
constexpr int kTile = 256;
constexpr int kRows = 16;                         // rows of an input block
constexpr uint32_t kBlockWords = kRows * kTile;   // 4096 words = 16 KiB per tile
constexpr int kMeshTileVerts = 1024;

// Builds a 2-D tensor map over a float[rows][16] matrix array: box {16, kTile}, 64-byte swizzle.
int encode_matrix_tensor_map(CUtensorMap* out, void* base, uint64_t rows);

} // namespace zn

struct zn_scene;
namespace zn {
int expand_visible_on(cudaStream_t stream, zn_scene* s, const char* who, const void* records_dev, uint32_t record_count,
                      uint32_t record_stride_words, int skip_index, uint32_t skip_count, const uint32_t* index_bases, void* lists_dev,
                      uint32_t list_stride, void* counts_dev, const void* flags_dev, uint32_t flag_stride_words, uint32_t records_per_flag,
                      uint32_t expected, uint32_t timeout_ms, void* status_dev, const void* abort_status, bool early_start = false);
} // namespace zn

struct zn_ctx {
	int device = 0;
	int sm_count = 0;
	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr;
	void* staging = nullptr;       // device staging for strided uploads
	size_t staging_bytes = 0;
};

struct zn_level {
	uint32_t begin = 0, end = 0;   // entity range
	uint32_t tile_base = 0;        // first global tile id of this level
	uint32_t tiles = 0;
	uint32_t ticket_base = 0;      // tickets this level's counter has handed out in earlier frames
	CUtensorMap tm_world, tm_mvp;  // extent clipped to `end` rows
};

struct zn_scene {
	zn_ctx* ctx = nullptr;
	uint32_t entity_count = 0;
	std::vector<zn_level> levels;
	uint32_t tile_count = 0;
	uint32_t index_base = 0;
	int max_ctas = 0;              // co-resident CTAs of the persistent level kernel (4 per SM)
	int max_ctas_world = 0;        // ... of its world-only form (5 per SM: no MVP tile in shared memory)
	bool has_camera = false;
	bool world_valid = false;      // an update has run: the world matrices exist (zn_scene_cull needs them)
	bool ranges_dirty = true;      // tile_prange must be rebuilt before the next update
	bool inputs_touched = true;    // an upload / fill has written the input blocks since the last update: level 0 must not
	                               // start its first block load early (programmatic launch) behind the kernel that wrote them
	float view_proj[16];
	float planes[24];
	// inputs: one 16 KiB block per 256-entity tile, rows = tx,ty,tz, pitch,yaw,roll, sx,sy,sz,
	// aabb centre xyz, aabb half xyz (float) and parent (u32); tiles never straddle a level.
	uint32_t* blocks = nullptr;    // [tiles][16][256]
	uint2* tile_prange = nullptr;  // [tiles] {first parent, parent count} of the tile (count 0: none, 0xFFFFFFFF: gather)
	// outputs
	float* world = nullptr;        // [E][16]
	float* mvp = nullptr;          // [E][16]
	uint32_t* visible = nullptr;       // [E] where the emit kernel writes (own buffer or a bound one)
	uint32_t* visible_count = nullptr; // [1] device
	uint32_t* visible_own = nullptr;
	uint32_t* visible_count_own = nullptr;
	// visibility record of a frame: header (chunk bases, total, group offsets; see
	// record_header_words in zn_scene.cu), then [tiles][8] mask words
	uint32_t* record = nullptr;            // where the kernels write (own buffer or a bound one)
	uint32_t* record_own = nullptr;        // [2][record_words]: the scene's own record, double-buffered by frame so that frame k+1's
	                                       // level kernels can write masks while frame k's emit still reads its own
	uint32_t record_slot = 0;
	bool tail_is_emit = false;             // the last kernel this scene launched was an emit: every level kernel before it has completed
	const uint32_t* frame_record = nullptr;      // record of the last frame with a view, and of the one before it
	const uint32_t* prev_frame_record = nullptr;
	bool record_bound = false;             // zn_scene_bind_visibility_record: caller-owned memory, one buffer, no frame overlap
	uint32_t record_words = 0;
	std::vector<uint32_t*> peer_records;   // copies of the record kept by peer GPUs (mapped device pointers), written by the kernels
	std::vector<uint32_t*> peer_flags;     // one flag word per rank (own included), raised to signal_value by zn_scene_expand_visible_signalled
	uint32_t signal_value = 0;
	uint32_t* tile_first_entity = nullptr; // [tiles]
	uint32_t chunk_count = 0;              // emit CTAs: 32 tiles each
	uint32_t* chunk_total = nullptr;       // [3][chunks] visible count per chunk: frame k accumulates into set k % 3, its emit clears set (k + 2) % 3
	uint32_t frame_parity = 0;
	uint32_t* level_ticket = nullptr;      // [levels + 1] monotonically increasing tickets: one counter per level, then the
	                                       // cull pass's tile counter
	uint32_t cull_ticket_base = 0;         // tickets the cull counter has handed out in earlier launches
	int max_ctas_cull = 0;                 // co-resident CTAs of scene_cull_kernel
	CUtensorMap tm_world_all, tm_mvp_all;  // all E rows (the cull pass is one launch over every level)
	std::vector<std::pair<void*, CUtensorMap>> mvp_maps;   // tensor maps of caller-owned MVP arrays (zn_scene_cull)
	const void* abort_status = nullptr;    // status word of the last zn_scene_wait_flags, consumed by the next plain expansion
	uint4* expand_entries = nullptr;       // work list of the two-launch expansion: non-empty groups of the records being expanded
	uint32_t expand_capacity = 0;
	uint32_t* expand_count = nullptr;      // [2] by launch parity
	uint32_t expand_parity = 0;
	int max_ctas_expand = 0;               // co-resident CTAs of expand_work_kernel
	bool level_fusion = true;              // zn_scene_set_level_fusion
	uint32_t top_levels = 0;               // leading levels composed by scene_top_kernel (0: none qualify)
	int expand_mode = 0;                   // ZN_EXPAND_GRID / ZN_EXPAND_WORKLIST (zn_scene_set_expand_mode)
	cudaEvent_t cull_events[3] = {nullptr, nullptr, nullptr};   // profiling: before cull, after cull, after emit
	uint32_t* host_count = nullptr;      // pinned [1]
	bool profiling = false;
	std::vector<cudaEvent_t> prof_events; // [levels+2]
	struct zn_pipe* pipe = nullptr;       // pipelined host frames (zn_scene_frame_host_submit/wait), lazily created
	bool pipe_ready = false;              // false while `pipe` is only partially built
};

struct zn_mesh_tile {
	uint32_t first_vertex;
	uint32_t vertex_count;
	uint32_t instance_lo;          // instances [instance_lo, instance_hi] own vertices of this tile
	uint32_t instance_hi;
};

struct zn_mesh {
	zn_ctx* ctx = nullptr;
	uint32_t vertex_count = 0;
	uint32_t instance_count = 0;
	uint32_t tile_count = 0;
	zn_mesh_tile* tiles = nullptr;   // device [tile_count]
	float* pos_in = nullptr;         // [V][3]
	float* nrm_in = nullptr;
	float* pos_out = nullptr;
	float* nrm_out = nullptr;
	float* models_own = nullptr;     // [I][16]
	uint32_t* instance_first_vertex = nullptr;  // device [I+1]
	float* inst_mats = nullptr;      // [I][24] model rows + normal matrix, only when some tile holds more instances than its shared-memory window
	bool multi_instance_tiles = false;
	bool needs_inst_mats = false;
	uint32_t uniform_vpi = 0;        // vertices per instance when the ranges are equal (no boundary table reads), else 0
	uint32_t smem_window = 0;        // most instances any multi-instance tile touches (<= 1024): sizes the kernel's dynamic shared memory
	const float* models = nullptr;   // models_own or a scene's world array
};
