/* zenyth_b200.h — C ABI of the B200-native scene-update path for the Zenyth engine.
 *
 * This is the drop-in boundary: plain pointers and sizes, integer status codes, no C++ or
 * torch types.  The upstream engine (TNtube/Zenyth) has no FFI and no implementation of this
 * path; its only plugin seam is Zenyth::IRenderer (Core/include/IRenderer.hpp:5-12) and the
 * frame hooks Application::OnUpdate / OnRender (Core/include/Application.hpp:38-39, driven by
 * Core/src/Application.cpp:41-52).  Each entry point below names the reference declaration it
 * stands in for or would be called from.  include/zenyth_b200.hpp layers the engine-idiom C++
 * types (Zenyth::Transform / Scene / Camera / Mesh) on top of these calls; INTEGRATION.md shows
 * the binding a maintainer adds on the reference side.
 *
 * Conventions (DESIGN.md "Conventions"): matrices are 16 floats, column-major, element (r,c) at
 * [c*4+r], translation in [12..14]  (Core/include/math/matrix.hpp:7-8,57-58); right-handed;
 * column vectors (matrix.hpp:50); angles in radians (Core/include/math/utils.hpp:5); clip depth
 * 0..1.  Entities are level-ordered: every parent index is smaller than its children's and
 * level l is the index range [level_offsets[l], level_offsets[l+1]).
 *
 * Errors: every call returns ZN_OK (0) or a ZN_ERR_* code and never throws; zn_last_error()
 * returns the calling thread's last message.  (The reference's convention is
 * std::runtime_error on set-up failure, Core/src/Application.cpp:28-30; the C++ wrapper
 * rethrows.)  Threading: one host thread per context at a time, one context per GPU.  There is
 * no CPU fallback: without a usable sm_100 device zn_ctx_create fails with ZN_ERR_NO_DEVICE.
 */
#ifndef ZENYTH_B200_H
#define ZENYTH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZN_ABI_VERSION 2

#define ZN_OK 0
#define ZN_ERR_INVALID 1   /* bad argument */
#define ZN_ERR_CUDA 2      /* a CUDA runtime / driver call failed */
#define ZN_ERR_NO_DEVICE 3 /* no CUDA device, or not compute capability 10.x */
#define ZN_ERR_STATE 4     /* call order (e.g. update before a camera was set) */
#define ZN_ERR_NOMEM 5

#define ZN_NO_PARENT 0xFFFFFFFFu

typedef struct zn_ctx zn_ctx;
typedef struct zn_scene zn_scene;
typedef struct zn_mesh zn_mesh;

/* ---- library ------------------------------------------------------------------------------ */
int zn_abi_version(void);
const char* zn_last_error(void);
int zn_device_count(int* out_count);

/* ---- context: one per GPU.  Stands where IRenderer::Init / destructor sit
 *      (Core/include/IRenderer.hpp:7-8; lifetime per Core/src/Application.cpp:32-34,54-56). ---- */
int zn_ctx_create(int device_ordinal, zn_ctx** out_ctx);
int zn_ctx_destroy(zn_ctx* ctx);
/* All work of this context is enqueued on `cuda_stream` (a cudaStream_t / CUstream passed as
 * void*); NULL restores the context's own stream. */
int zn_ctx_set_stream(zn_ctx* ctx, void* cuda_stream);
int zn_ctx_get_stream(zn_ctx* ctx, void** out_cuda_stream);
int zn_ctx_sync(zn_ctx* ctx);
int zn_ctx_sm_count(zn_ctx* ctx, int* out_sm_count);
/* Page-locked host memory for the *_host entry points (plain malloc memory also works, slower). */
int zn_host_alloc(zn_ctx* ctx, size_t bytes, void** out_ptr);
int zn_host_free(zn_ctx* ctx, void* ptr);
/* Peer memory (one process per GPU): a plain device allocation, its 64-byte CUDA IPC handle, and
 * the mapping of another rank's handle into this process (peer access over NVLink is enabled
 * lazily).  The handle bytes travel between ranks by any host channel (torch.distributed,
 * MPI, a pipe). */
int zn_device_alloc(zn_ctx* ctx, size_t bytes, void** out_ptr);
int zn_device_free(zn_ctx* ctx, void* ptr);
/* Synchronising copies between host memory and any device pointer this library handed out or was
 * given (lists of an exchange, draw arguments, caller-owned cull outputs): for hosts that do not
 * link the CUDA runtime themselves. */
int zn_device_read(zn_ctx* ctx, const void* dev_ptr, void* host_dst, size_t bytes);
int zn_device_write(zn_ctx* ctx, void* dev_ptr, const void* host_src, size_t bytes);
int zn_ipc_export(zn_ctx* ctx, void* dev_ptr, unsigned char out_handle[64]);
int zn_ipc_open(zn_ctx* ctx, const unsigned char handle[64], void** out_dev_ptr);
int zn_ipc_close(zn_ctx* ctx, void* dev_ptr);

/* ---- camera maths: the part of mat4 a Camera needs, completing the stubs
 *      mat4::look_at / perspective / orthographic / perspective_reverse_z / operator*
 *      (Core/include/math/matrix.hpp:32,52-55; stub bodies Core/src/math/matrix.cpp:55-57,107-121).
 *      Host-side, pure functions. ---- */
void zn_mat4_look_at(const float eye[3], const float center[3], const float up[3], float out[16]);
void zn_mat4_perspective(float fov_y_rad, float aspect, float near_z, float far_z, float out[16]);
void zn_mat4_orthographic(float left, float right, float bottom, float top, float near_z, float far_z, float out[16]);
void zn_mat4_perspective_reverse_z(float fov_y_rad, float aspect, float near_z, float out[16]);
void zn_mat4_mul(const float a[16], const float b[16], float out[16]);
/* Six planes (nx,ny,nz,d) from the rows of view_proj: left,right,bottom,top,near,far. */
void zn_frustum_planes(const float view_proj[16], float out_planes[24]);

/* ---- hierarchy layout (host only, no device needed): entities in creation order, each naming its
 *      parent (ZN_NO_PARENT for roots), to the order the scene calls below require — level by level,
 *      a parent always in an earlier level, the children of a level sorted by parent, siblings in
 *      their original relative order.  The reference has no Scene to take an order from; this is
 *      what a Scene::AddEntity / SetParent pair would maintain.
 *        out_order[new]            original index of the entity now at `new`     [count]
 *        out_parent[new]           its parent's NEW index, or ZN_NO_PARENT        [count], may be NULL
 *        out_level_offsets[0..L]   level l is [offsets[l], offsets[l+1])          needs L+1 <= level_capacity
 *        *out_level_count          L (also set when level_capacity was too small, so a caller can retry)
 *      Fails with ZN_ERR_INVALID on a parent index out of range, a self-parent or a cycle. ---- */
int zn_hierarchy_sort(const uint32_t* parent, uint32_t count, uint32_t* out_order, uint32_t* out_parent,
                      uint32_t* out_level_offsets, uint32_t level_capacity, uint32_t* out_level_count);

/* ---- scene: TRS -> world over the hierarchy, MVP, AABB-vs-frustum cull + ordered compaction.
 *      Called from where Application::OnUpdate(dt) and OnRender() run
 *      (Core/src/Application.cpp:47,50). ---- */
typedef struct zn_scene_desc {
	uint32_t entity_count;
	uint32_t level_count;          /* >= 1 */
	const uint32_t* level_offsets; /* [level_count+1], level_offsets[0]==0, last==entity_count;
	                                  NULL means one level of roots */
} zn_scene_desc;

int zn_scene_create(zn_ctx* ctx, const zn_scene_desc* desc, zn_scene** out_scene);
int zn_scene_destroy(zn_scene* scene);

/* Host -> device upload of entities [first, first+count).  Each array is `count` records of
 * three floats, `stride_bytes` apart (12 for packed float[3], 16 for arrays of the reference's
 * 16-byte zenyth::math::vec3, Core/include/math/vector.hpp:69-72).  NULL arrays are skipped.
 * rotation = (pitch, yaw, roll) radians as mat4::from_euler (matrix.hpp:24).
 * Buffer lifetime: these two calls return once the copies are ENQUEUED on the context stream.
 * Pageable arrays are staged by the driver before the call returns and may be reused at once;
 * page-locked arrays (zn_host_alloc) are read by the DMA engine later and must stay untouched until
 * the stream has passed the copy (zn_ctx_sync, or any synchronising read-back).
 * Either call for at most 16 entities — the handful a frame's game logic moved — is ONE kernel launch
 * that carries the values in its parameters (read before the call returns): no staging copy, no repack passes. */
int zn_scene_set_transforms(zn_scene* scene, uint32_t first, uint32_t count,
                            const float* position, const float* rotation, const float* scale, size_t stride_bytes);
int zn_scene_set_bounds(zn_scene* scene, uint32_t first, uint32_t count,
                        const float* aabb_center, const float* aabb_half_extent, size_t stride_bytes);
/* parent[i] is an entity index in an earlier level, or ZN_NO_PARENT.  Level 0 must be all roots. */
int zn_scene_set_parents(zn_scene* scene, uint32_t first, uint32_t count, const uint32_t* parent);
/* view and proj exactly as mat4::data() (matrix.hpp:60-61). */
int zn_scene_set_camera(zn_scene* scene, const float view[16], const float proj[16]);
/* Redirect the visible list + count of subsequent updates into caller-owned DEVICE memory (for
 * example a slot of a double-buffered all-gather input, so a collective can consume frame k while
 * frame k+1 is computed).  capacity (in indices) must be >= entity_count.  Both NULL restores the
 * scene's own buffers. */
int zn_scene_bind_visible_output(zn_scene* scene, void* visible_dev, uint32_t capacity, void* visible_count_dev);
/* Multi-GPU exchange at 1 bit per entity.  A frame's visible set also exists as a fixed-size
 * "visibility record" (zn_scene_buffers.visibility_record).  Ranks all-gather the records (27x
 * fewer bytes than the index lists at the bench's visible fraction, and a size known without
 * reading the count) and every rank expands them into the same ordered index lists with
 * zn_scene_expand_visible — the arithmetic of the local compaction, so the lists are identical
 * to an all-gather of the lists themselves.
 *   zn_scene_bind_visibility_record: subsequent updates write the record into caller-owned DEVICE
 *     memory (>= visibility_record_words words; e.g. a slot of a double-buffered all-gather input).
 *     NULL restores the scene's own buffer.
 *   zn_scene_expand_visible: expands `record_count` (<= 64) records laid out `record_stride_words`
 *     apart, of scenes with THIS scene's level layout, into lists_dev[r*list_stride ...] and
 *     counts_dev[r]; index_bases[r] (host array, may be NULL) is added to record r's indices;
 *     record `skip_index` is left out (-1: none).  Asynchronous on the context stream. */
int zn_scene_bind_visibility_record(zn_scene* scene, void* record_dev);
/* Fused all-gather over peer memory: every update also stores the record, tile by tile from
 * inside the compute kernels, into up to 7 additional buffers — device pointers into PEER GPUs'
 * memory, mapped through zn_ipc_open (NVLink P2P) — so after a cross-rank barrier every rank holds
 * every rank's record without a collective having moved it.  peer_count == 0 unbinds. */
int zn_scene_bind_peer_records(zn_scene* scene, void* const* peer_record_ptrs, uint32_t peer_count);
int zn_scene_expand_visible(zn_scene* scene, const void* records_dev, uint32_t record_count, uint32_t record_stride_words,
                            int skip_index, const uint32_t* index_bases, void* lists_dev, uint32_t list_stride, void* counts_dev);
/* The cross-rank completion signal of that fused all-gather, again without a collective.
 *   bind_peer_signal: flag_ptrs[q] is ONE 32-bit word in rank q's memory (this rank's cell of q's flag
 *     array; mapped through zn_ipc_open, the own cell is a plain local pointer).  flag_count == 0 unbinds.
 *   set_signal_value: the sequence number of the latest update enqueued on the stream; it only grows
 *     (compared as a signed 32-bit distance, so it may wrap).
 *   expand_visible_signalled: zn_scene_expand_visible that (1) first raises the bound flags to the
 *     signal value — every update enqueued before it is complete by then, peer stores included
 *     (one fence.sys, then relaxed system-scope stores, by one thread) — and (2) whose CTAs wait (ld.acquire.sys) until
 *     flags_dev[(r / records_per_flag) * flag_stride_words] has reached `expected` for their record
 *     r, so the stream needs no host- or NCCL-side barrier.  The wait is bounded: after timeout_ms
 *     the kernel stores 1 + r into *status_dev (a zero-initialised device word; the first failure
 *     wins), does NOT expand record r (a record whose owner never signalled may be half written) and
 *     sets counts_dev[r] = 0; the other records are expanded as usual.
 *   NOTE: zn_scene_update never raises a flag itself — the flags of frame k go up in the NEXT
 *     zn_scene_expand_visible_signalled / zn_scene_wait_flags launch on the stream, so the last frame
 *     of a run must be followed by one more such call or its peers wait for the timeout.
 *     Records [skip_index, skip_index + skip_count) are left out (a rank that owns several scenes
 *     already has their lists); records_per_flag = scenes per rank. */
int zn_scene_bind_peer_signal(zn_scene* scene, void* const* flag_ptrs, uint32_t flag_count);
/* The same signal + bounded wait as a launch of its own on the context stream (all flag_count
 * flags must reach `expected`); a plain zn_scene_expand_visible may follow it.  If the wait times
 * out (*status_dev = 1 + the flag's index), the next zn_scene_expand_visible on this scene expands
 * nothing and zeroes its counts. */
int zn_scene_wait_flags(zn_scene* scene, const void* flags_dev, uint32_t flag_stride_words, uint32_t flag_count, uint32_t expected,
                        uint32_t timeout_ms, void* status_dev);
int zn_scene_set_signal_value(zn_scene* scene, uint32_t value);
int zn_scene_expand_visible_signalled(zn_scene* scene, const void* records_dev, uint32_t record_count, uint32_t record_stride_words,
                                      int skip_index, uint32_t skip_count, const uint32_t* index_bases, void* lists_dev, uint32_t list_stride, void* counts_dev,
                                      const void* flags_dev, uint32_t flag_stride_words, uint32_t records_per_flag, uint32_t expected,
                                      uint32_t timeout_ms, void* status_dev);
/* How zn_scene_expand_visible(_signalled) and zn_exchange_expand expand records with this scene:
 *   ZN_EXPAND_GRID (default)  one launch, one warp per 4-tile group of every record: best when most groups are
 *                             non-empty (headline camera, f = 0.86: 64 us for seven 16.7M-entity records; 80 us with
 *                             the work list);
 *   ZN_EXPAND_WORKLIST        two launches chained programmatically: a classification pass (one THREAD per group)
 *                             appends the non-empty groups to a work list, a persistent grid works it off — no warp
 *                             is ever spent on an empty group: best for sparse visible sets (camera inside the scene,
 *                             f = 0.21, 70% of the chunks empty: 26 us against 36 us).
 * The lists are identical either way.  The engine knows which kind of view it renders; the library cannot know the
 * visible fraction on the host without a synchronising read-back. */
#define ZN_EXPAND_GRID 0
#define ZN_EXPAND_WORKLIST 1
int zn_scene_set_expand_mode(zn_scene* scene, int mode);
/* The leading levels of a hierarchy whose tiles fit one CTA's shared memory (at most 5 tiles of 256 entities in
 * total, at least two levels: 7 + 56 + 448 entities in the bench scene) are composed by ONE launch that keeps their
 * world matrices on chip (scene_top_kernel) instead of one launch per level.  Results are bit-identical; the switch
 * exists for A/B measurement.  Default: on. */
int zn_scene_set_level_fusion(zn_scene* scene, int enabled);
/* Added to every index written to the visible list (shard -> global numbering). */
int zn_scene_set_index_base(zn_scene* scene, uint32_t index_base);

/* One frame, asynchronous on the context stream: world[E], mvp[E], visible list + count — the
 * fused form, for the camera given to zn_scene_set_camera.  A scene that has no camera (none set
 * yet, or zn_scene_clear_camera) runs the world-only form below instead. */
int zn_scene_update(zn_scene* scene);

/* Update and cull as the two calls the reference's frame loop has room for: Scene::Update(dt) from
 * Application::OnUpdate (Core/src/Application.cpp:47) and Scene::Cull(camera) from OnRender, between
 * IRenderer::BeginFrame and EndFrame (Application.cpp:49-51).
 *   zn_scene_update_world: TRS -> world over the hierarchy only (no camera involved): reads 40 B and
 *     writes 64 B per entity; MVPs, masks and the visible list are left as they were.
 *   zn_scene_cull: one camera pass over the world matrices of the last update (fused or world-only):
 *     MVP = (proj*view)*world, AABB-vs-frustum, ordered compaction — 152 B per entity + 4 B per
 *     visible index, no sincos, no world matrix rewritten.  May be called any number of times per
 *     frame (main view, shadow views, a camera that moved between OnUpdate and OnRender); the result
 *     of each call equals a fused zn_scene_update with that camera bit for bit.  `out` redirects
 *     this pass's results into caller-owned DEVICE memory (NULL, or NULL members: the scene's own
 *     MVP array / currently bound visible list + count); the camera of zn_scene_set_camera is not
 *     changed.  The visibility record (and bound peer records) are rewritten by every pass. */
typedef struct zn_cull_output {
	void* mvp_dev;            /* float[entity_count][16], 16-byte aligned, or NULL */
	void* visible_dev;        /* uint32[visible_capacity] or NULL (then visible_count_dev must be NULL too) */
	uint32_t visible_capacity;/* >= entity_count */
	void* visible_count_dev;  /* uint32[1] */
} zn_cull_output;
int zn_scene_update_world(zn_scene* scene);
int zn_scene_cull(zn_scene* scene, const float view[16], const float proj[16], const zn_cull_output* out);
int zn_scene_clear_camera(zn_scene* scene);

/* Results.  These synchronise the stream. */
int zn_scene_visible_count(zn_scene* scene, uint32_t* out_count);
int zn_scene_read_visible(zn_scene* scene, uint32_t* out_indices, uint32_t capacity, uint32_t* out_count);
int zn_scene_read_world(zn_scene* scene, uint32_t first, uint32_t count, float* out_matrices);
int zn_scene_read_mvp(zn_scene* scene, uint32_t first, uint32_t count, float* out_matrices);
int zn_scene_read_inputs(zn_scene* scene, uint32_t first, uint32_t count,
                         float* position, float* rotation, float* scale,
                         float* aabb_center, float* aabb_half_extent, uint32_t* parent); /* packed [count][3] */

/* Device-resident results for a renderer or a collective to consume in place. */
typedef struct zn_scene_buffers {
	void* world;            /* float[E][16] */
	void* mvp;              /* float[E][16] */
	void* visible;          /* uint32[E], first *visible_count entries valid, ascending */
	void* visible_count;    /* uint32[1] */
	void* visibility_record;          /* (of the LAST update or cull pass — the scene's own record is double-buffered by
	                                     frame, so ask again after every frame —)
	                                     uint32[visibility_record_words]: the frame's visible set at 1 bit per
	                                     entity — a header of list offsets per 8192-entity chunk and the total
	                                     count, then 8 mask words per 256-entity tile.  Opaque: produced by
	                                     zn_scene_update, consumed by zn_scene_expand_visible. */
	uint32_t visibility_record_words;
	uint32_t entity_count;
} zn_scene_buffers;
int zn_scene_device_buffers(zn_scene* scene, zn_scene_buffers* out);

/* What the draw submission after IRenderer::EndFrame (Core/include/IRenderer.hpp:10) consumes,
 * built on the device from the last update: instance_mvp_dev[k] = mvp[visible[k]] for k in
 * [0, count) (float[instance_capacity][16], instance_capacity >= entity_count) and one argument
 * record at draw_args_dev in the D3D12_DRAW_INDEXED_ARGUMENTS layout — five 32-bit words
 * {index_count_per_instance, count, 0, 0, 0}.  No host round trip: the count is read on the
 * device.  Asynchronous on the context stream. */
int zn_scene_build_draw(zn_scene* scene, uint32_t index_count_per_instance, void* instance_mvp_dev, uint32_t instance_capacity,
                        void* draw_args_dev);

/* End-to-end frame with HOST buffers: uploads the packed [count][3] inputs that are not NULL for the
 * entity range [first, first+count) (NULL = keep what is resident; count == 0 = the whole scene),
 * runs the frame, downloads what is asked for (NULL = leave on the device), and returns after the
 * data has landed.  This is the call a CPU engine would swap in.  A frame that only spins its
 * entities passes `rotation` alone (12 bytes per entity instead of 36). */
typedef struct zn_frame_host_io {
	const float* position;      /* in  [count][3] or NULL */
	const float* rotation;      /* in  [count][3] or NULL */
	const float* scale;         /* in  [count][3] or NULL */
	float* world;               /* out [E][16] or NULL */
	float* mvp;                 /* out [E][16] or NULL */
	uint32_t* visible;          /* out [E] or NULL */
	uint32_t visible_count;     /* out */
	uint32_t first;             /* in  first entity the input arrays cover */
	uint32_t count;             /* in  entities they cover; 0 = entity_count - first */
} zn_frame_host_io;
int zn_scene_frame_host(zn_scene* scene, zn_frame_host_io* io);
/* The same frame, pipelined: up to two frames in flight.  submit() enqueues the upload of the
 * non-NULL input arrays (page-locked memory for overlap), ONE repack launch into the tile-major
 * blocks, the update and the download of the visible count, and returns; wait() completes the
 * OLDEST submitted frame: fills io->visible_count and, when io->visible is not NULL,
 * io->visible[0..count).  The upload of frame k+1 overlaps the update and the download of frame k
 * (PCIe is full duplex), so a steady-state frame costs max(upload, update, download) instead of
 * their sum.  The host arrays passed to submit() must stay untouched until the matching wait()
 * returns; world/mvp of zn_frame_host_io are ignored here (they stay on the device:
 * zn_scene_device_buffers). */
int zn_scene_frame_host_submit(zn_scene* scene, const zn_frame_host_io* io);
int zn_scene_frame_host_wait(zn_scene* scene, zn_frame_host_io* io);

/* Bytes one zn_scene_update moves by construction: 188 B per root entity,
 * 192 B per child, + 4 B per visible index. */
int zn_scene_algorithmic_bytes(zn_scene* scene, uint32_t visible_count, uint64_t* out_bytes);
/* Number of kernel launches one zn_scene_update enqueues. */
int zn_scene_launches_per_update(zn_scene* scene, uint32_t* out_launches);
/* Measurement support: when enabled, zn_scene_update brackets every level launch (and the
 * compaction tail) with CUDA events on the context stream; zn_scene_level_times then returns the
 * device time in milliseconds of each level kernel of the LAST update (out_ms[level_count]) and
 * of the scan+emit tail (out_tail_ms, may be NULL).  Synchronises. */
int zn_scene_set_profiling(zn_scene* scene, int enable);
int zn_scene_level_times(zn_scene* scene, float* out_ms, uint32_t capacity, float* out_tail_ms);
/* Test support: the persistent kernels claim their tiles from monotonically increasing 32-bit ticket
 * counters (one per level, one for the cull pass) that wrap after ~2^32
 * tickets — every ~74k frames for the last level of the 16M scene.  This call restarts all of them
 * at `origin` (e.g. 0xFFFFFF00) so a test can run a scene across the wrap.  Synchronises. */
int zn_scene_set_ticket_origin(zn_scene* scene, uint32_t origin);
/* With profiling enabled: device time of the LAST zn_scene_cull's cull kernel and of its emit kernel. */
int zn_scene_cull_times(zn_scene* scene, float* out_cull_ms, float* out_emit_ms);

/* Scene ingest: a minimal binary SoA dump (the reference has no on-disk format).  Little-endian:
 * "ZNSCENE1", uint32 entity_count, uint32 level_count, uint32 level_offsets[level_count+1], then
 * float32 position[E][3], rotation[E][3], scale[E][3], aabb_center[E][3], aabb_half_extent[E][3],
 * uint32 parent[E] — 64 bytes per entity.  zn_scene_load creates the scene it reads. */
int zn_scene_save(zn_scene* scene, const char* path);
int zn_scene_load(zn_ctx* ctx, const char* path, zn_scene** out_scene);

/* ---- multi-GPU host layer (one process per GPU).  Host logic only — the data path
 *      is the kernels' own peer stores; no collective library is involved.
 *
 *      Sharding.  Independent scenes are dealt out as contiguous blocks (zn_shard_scenes: 8 scenes over
 *      1/2/4/8 GPUs -> 8/4/2/1 per GPU) and numbered globally by zn_scene_index_bases (exclusive prefix
 *      sum of the entity counts -> zn_scene_set_index_base).  ONE scene is split into hierarchy-closed
 *      entity shards by root subtree (zn_shard_scene_by_roots): roots are dealt out in index order as
 *      contiguous runs balanced by subtree size, every entity follows its root and keeps its relative
 *      order, so each shard is level-ordered and sorted by parent as it stands and no parent is ever
 *      read across GPUs.  out_owner[e] = rank of entity e, out_local_index[e] = its index inside that
 *      shard, out_shard_counts[r] = entities of shard r. ---- */
int zn_shard_scenes(uint32_t num_scenes, uint32_t world_size, uint32_t rank, uint32_t* out_first, uint32_t* out_count);
int zn_scene_index_bases(const uint32_t* entity_counts, uint32_t num_scenes, uint32_t* out_bases);
int zn_shard_scene_by_roots(const uint32_t* level_offsets, uint32_t level_count, const uint32_t* parent /* may be NULL: all roots */,
                            uint32_t world_size, uint32_t* out_owner, uint32_t* out_local_index, uint32_t* out_shard_counts);

/*      The per-frame exchange: every rank ends up with every rank's compacted visible list and count.
 *      Records travel inside the compute kernels (zn_scene_bind_peer_records), completion as flags in
 *      peer memory (zn_scene_bind_peer_signal), a rank's own lists come straight from its emit kernels
 *      (zn_scene_bind_visible_output) and only the peers' records are expanded.  The object below owns
 *      the buffers (records[slots][world*S][record_words] + flags, lists[2][world*S][capacity],
 *      counts, status) and performs the binding arithmetic; see zn_exchange.cpp for the slot-reuse
 *      argument.  `allgather` is the one thing the caller supplies: a blocking HOST all-gather among
 *      the ranks (every rank contributes `bytes` bytes, receives world_size*bytes in rank order; return
 *      0 on success) — MPI_Allgather, a socket, torch.distributed.  It is used by create (twice:
 *      handles, then "everyone mapped") and destroy (twice) only, never per frame; create and destroy
 *      are therefore collective calls.  If any rank cannot allocate, export or map, EVERY rank's
 *      create fails with ZN_ERR_CUDA.
 *        per frame k:  zn_exchange_bind(x, scenes, S, k);  zn_scene_update(scenes[j]) for all j;
 *                      zn_exchange_expand(x, scenes[0], k - 1, index_bases);
 *        at the end:   zn_exchange_expand(x, scenes[0], last, ...) — also raises the last frame's flags.
 *      zn_exchange_expand is asynchronous on the context stream: the launch raises this rank's flags for
 *      the last bound frame (everything before it on the stream is complete), waits on the device for
 *      the owners' flags of `frame` and derives their lists.  Expanding one frame behind leaves a frame of
 *      slack between ranks, so the flags are normally up and nothing spins.  The launch is PROGRAMMATIC: the
 *      expansion of frame k-1 touches only exchange-owned buffers that the update of frame k does not (another
 *      record slot, the other half of the lists), so its CTAs may start beside the tail of that update; only the
 *      one thread that raises the flags waits for it, and everything enqueued after the expansion is ordered
 *      behind both (ZN_EXCHANGE_EARLY_START=0 in the environment makes it a plain launch, for A/B runs).
 *      Frames are expanded in order;
 *      a frame that was never bound (k - 1 = -1) or is already expanded is a no-op.
 *      index_bases[world*S] (host, may be NULL) is added to each record's indices.
 *      ZN_EXCHANGE_OVERLAP in desc.flags (opt-in): expansions run on the exchange's own low-priority side
 *      stream, ordered behind everything enqueued on the context stream so far, so frame k can be
 *      expanded right after its update (zn_exchange_expand(k)) while the context stream already runs
 *      update(k+1); zn_exchange_bind then makes update(k) wait for this rank's expand(k-2) (flow control,
 *      zn_exchange.cpp).  Measured gain ~1%: the fused frame leaves no idle HBM bandwidth to hide in.
 *      zn_exchange_join makes the context stream wait for the last expansion (a no-op without overlap).
 *      zn_exchange_signal raises this rank's flags for the last bound frame WITHOUT waiting for anyone
 *      (a launch of one thread behind the update) and returns when they are up.  Production code never
 *      needs it — expand signals — but ranks that share ONE GPU (tests) must never let a kernel spin on a
 *      flag another process on the same device has yet to raise: they signal, rendezvous on the host,
 *      then expand.
 *      zn_exchange_result synchronises and returns the counts of the LAST expanded frame and the
 *      device address + stride (in indices) of its lists: list r is lists[r*stride ..).  It fails with
 *      ZN_ERR_STATE if a peer never signalled within timeout_ms (that record was not expanded; the
 *      message names the rank). */
#define ZN_EXCHANGE_OVERLAP 1u
typedef struct zn_exchange zn_exchange;
typedef int (*zn_allgather_fn)(void* user, const void* send, void* recv, size_t bytes);
typedef struct zn_exchange_desc {
	uint32_t world_size;
	uint32_t rank;
	uint32_t records_per_rank;  /* S: scenes this rank owns (the same on every rank) */
	uint32_t record_words;      /* zn_scene_buffers.visibility_record_words (one level layout for all scenes) */
	uint32_t list_capacity;     /* entities per scene */
	uint32_t slots;             /* 0 = 4, the minimum */
	uint32_t timeout_ms;        /* 0 = 2000 */
	zn_allgather_fn allgather;  /* may be NULL when world_size == 1 */
	void* user;
	uint32_t flags;             /* 0, or ZN_EXCHANGE_OVERLAP */
} zn_exchange_desc;
int zn_exchange_create(zn_ctx* ctx, const zn_exchange_desc* desc, zn_exchange** out_exchange);
int zn_exchange_destroy(zn_exchange* exchange, zn_scene* const* scenes, uint32_t scene_count); /* unbinds the scenes first */
int zn_exchange_unbind(zn_exchange* exchange, zn_scene* const* scenes, uint32_t scene_count);
int zn_exchange_bind(zn_exchange* exchange, zn_scene* const* scenes, uint32_t scene_count, uint64_t frame);
int zn_exchange_expand(zn_exchange* exchange, zn_scene* scene, int64_t frame, const uint32_t* index_bases);
int zn_exchange_signal(zn_exchange* exchange, zn_scene* scene);
int zn_exchange_join(zn_exchange* exchange);
int zn_exchange_result(zn_exchange* exchange, uint32_t* out_counts, void** out_lists_dev, uint32_t* out_list_stride);
int zn_exchange_info(zn_exchange* exchange, uint32_t* out_total_records, uint32_t* out_list_capacity, uint64_t* out_peer_bytes_per_frame);

/* ---- mesh batch: bulk vertex position / normal transform by per-instance model matrices.
 *      pos' = model.transform_point(pos); nrm' = normalize(model.inverse_transpose().transform_normal(nrm))
 *      (Core/include/math/matrix.hpp:43-48).  The Renderer-side preprocessing that would sit
 *      between IRenderer::BeginFrame and EndFrame (Core/include/IRenderer.hpp:9-10). ---- */
typedef struct zn_mesh_desc {
	uint32_t vertex_count;
	uint32_t instance_count;
	const uint32_t* instance_first_vertex; /* [instance_count+1], ascending, [0]==0, last==vertex_count;
	                                          NULL means equal ranges (vertex_count % instance_count == 0) */
} zn_mesh_desc;

int zn_mesh_create(zn_ctx* ctx, const zn_mesh_desc* desc, zn_mesh** out_mesh);
int zn_mesh_destroy(zn_mesh* mesh);
int zn_mesh_set_vertices(zn_mesh* mesh, uint32_t first, uint32_t count,
                         const float* position, const float* normal, size_t stride_bytes);
int zn_mesh_set_models(zn_mesh* mesh, uint32_t first, uint32_t count, const float* model_matrices /*[count][16] host*/);
/* Use world matrices [first_entity, first_entity+instance_count) of a scene as the models (no copy). */
int zn_mesh_bind_scene_models(zn_mesh* mesh, zn_scene* scene, uint32_t first_entity);
int zn_mesh_transform(zn_mesh* mesh); /* asynchronous */
int zn_mesh_read(zn_mesh* mesh, uint32_t first, uint32_t count, float* out_position, float* out_normal); /* packed [count][3] */
int zn_mesh_read_inputs(zn_mesh* mesh, uint32_t first, uint32_t count, float* position, float* normal);
typedef struct zn_mesh_buffers {
	void* position_in;  /* float[V][3] */
	void* normal_in;
	void* position_out;
	void* normal_out;
	uint32_t vertex_count;
} zn_mesh_buffers;
int zn_mesh_device_buffers(zn_mesh* mesh, zn_mesh_buffers* out);
/* Mesh-batch ingest, the companion of zn_scene_save/load.  Little-endian: "ZNMESH01", uint32
 * vertex_count, uint32 instance_count, uint32 instance_first_vertex[instance_count+1], float32
 * position[V][3], normal[V][3], model[I][16] (the matrices in use at save time: a bound scene's
 * world matrices are written out as plain models).  zn_mesh_load creates the mesh it reads. */
int zn_mesh_save(zn_mesh* mesh, const char* path);
int zn_mesh_load(zn_ctx* ctx, const char* path, zn_mesh** out_mesh);

/* ---- synthetic inputs generated on the device (bench / test support).
 *      A stateless hash of (seed, index, field) -> float, bit-identical to the host generator. */
int zn_scene_fill_synthetic(zn_scene* scene, uint32_t seed, uint32_t fanout, float aabb_center_range);
int zn_mesh_fill_synthetic(zn_mesh* mesh, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif /* ZENYTH_B200_H */
