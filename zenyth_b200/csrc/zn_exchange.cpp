// zn_exchange.cpp — the multi-GPU host layer of the path, in C++ behind the C ABI (one process per
// GPU).  Everything here is host logic over the scene entry points of this library: no torch, no
// NCCL, no collective on the data path.
//
//   zn_shard_scenes / zn_scene_index_bases   independent scenes -> ranks, shard -> global numbering
//   zn_shard_scene_by_roots                  ONE scene -> hierarchy-closed entity shards by root subtree
//   zn_exchange_*                            the per-frame exchange of compacted visible lists, fused into
//                                            the compute kernels over NVLink peer memory (CUDA IPC)
//
// The exchange.  Every rank owns ONE device allocation
//     records [slots][world * S][record_words]    gathered visibility records, slot = frame % slots;
//                                                 S = scenes per rank, rank r's scene j is record r*S + j
//     flags   [world] (128-byte stride)           flags[r] = sequence number of the last frame rank r completed
// exports its CUDA IPC handle and maps every peer's.  A bound scene's level / cull / emit kernels
// store its visibility record straight into slot [rank] of EVERY rank's `records` while they run
// (plain st.global on mapped peer pointers).  Completion travels the same way: the first thread of
// the expansion kernel that follows an update on the stream raises flags[rank] in every rank's
// allocation (one fence.sys, relaxed system-scope stores), and the expansion's warps wait on the
// owners' flags themselves (ld.acquire.sys, bounded).  The only thing the caller supplies is a host
// all-gather for the 64-byte handles at set-up — MPI_Allgather, a socket, torch.distributed: any
// channel — and it is never called per frame.
//
// Pipeline per frame k (default, everything on the context stream, the expansion one frame behind):
//     zn_exchange_bind(k); zn_scene_update of every owned scene; zn_exchange_expand(k-1).
// The expansion launch raises this rank's flags for frame k (the update before it on the stream is
// complete) and waits on the device for the peers' frame k-1 flags — which, one frame later, are
// normally up already: the frame of slack absorbs jitter between ranks and no CTA spins.  The launch is
// programmatic (early_start in expand_visible_on): expand(k-1) shares nothing with update(k) — another
// record slot, the other half of the lists — so its CTAs start beside the tail of update(k)'s emit kernel;
// only the flag-raising thread executes griddepcontrol.wait, i.e. the flags still go up after update(k)
// has completed, and the kernels after the expansion still wait for all of it (N = 2: 0.534 -> 0.529 ms,
// N = 8: 0.590 -> 0.580 ms per frame).
// Slot reuse.  Update(k+1) on rank r stores into slot (k+1) % 4 of every peer p.  It is
// stream-ordered after r's expand(k-1), which waited for p's frame k-1 signal; p raised it in the
// launch after its update(k-1), i.e. after its expand(k-3) was enqueued — so the youngest frame a
// peer is guaranteed to have finished reading is k-3 and a slot may be reused after FOUR frames.
// Lists are double-buffered: lists[k % 2] is complete (all ranks' lists) once expand(k) has run and
// its own-rank part is overwritten by update(k+2).
//
// ZN_EXCHANGE_OVERLAP (opt-in): the expansion of frame k runs on the exchange's own low-priority SIDE
// STREAM right behind update(k) (ordered by an event) while the context stream already runs
// update(k+1); update(f) then waits for the rank's OWN expand(f-2) (cudaStreamWaitEvent), which bounds
// the drift between ranks to two frames and keeps the four slots safe by the same argument (p's
// update(f) follows p's expand(f-2), which saw my flag f-1, raised by my expand(f-2) launch — behind my
// expand(f-4) on my side stream).  Measured: it buys ~1% (N = 8: 0.594 vs 0.600 ms per frame; one GPU
// emulation, update + expansion of 7 records: 573 vs 577 us, scripts/overlap_probe.py) — the frame
// already runs at 0.99 of the HBM roofline, so there is no idle bandwidth for the expansion's writes to
// hide in: N lists on every rank cost their bytes (roofline.scaling_ceiling in bench.py).  It is kept
// as an option because a frame with less traffic per entity (zn_scene_cull passes, small scenes) does
// leave room.
#include "zn_internal.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace {

constexpr uint32_t kFlagStrideWords = 32;   // one 128-byte line per writer

} // namespace

struct zn_exchange {
	zn_ctx* ctx = nullptr;
	uint32_t world = 0, rank = 0, per_rank = 0, total = 0;
	uint32_t record_words = 0, capacity = 0, slots = 0, timeout_ms = 0;
	zn_allgather_fn allgather = nullptr;
	void* user = nullptr;
	size_t slot_bytes = 0, flags_offset = 0;
	uint8_t* base = nullptr;                 // own allocation
	std::vector<uint8_t*> peer_base;         // [world], own entry == base
	uint32_t* lists = nullptr;               // [2][total][capacity]
	uint32_t* counts = nullptr;              // [2][total]
	uint32_t* status = nullptr;              // [1]
	uint32_t* host_counts = nullptr;         // pinned [total + 1]
	int64_t issued = -1, expanded = -1;
	cudaStream_t side = nullptr;             // the expansions run here, concurrently with the next update on the context stream
	cudaEvent_t updated = nullptr;           // context stream: everything up to and including update(frame) is enqueued
	cudaEvent_t done[2] = {nullptr, nullptr};// side stream: expand(frame) complete, by frame parity
	bool overlap = false;
};

using namespace zn;

static void exchange_release(zn_exchange* x) {
	if (!x) return;
	cudaSetDevice(x->ctx->device);
	for (uint32_t r = 0; r < x->peer_base.size(); ++r)
		if (r != x->rank && x->peer_base[r]) cudaIpcCloseMemHandle(x->peer_base[r]);
	x->peer_base.clear();
	if (x->base) cudaFree(x->base);
	x->base = nullptr;
	if (x->lists) cudaFree(x->lists);
	if (x->counts) cudaFree(x->counts);
	if (x->status) cudaFree(x->status);
	if (x->host_counts) cudaFreeHost(x->host_counts);
	x->lists = x->counts = x->status = x->host_counts = nullptr;
	if (x->updated) cudaEventDestroy(x->updated);
	for (cudaEvent_t& e : x->done) { if (e) cudaEventDestroy(e); e = nullptr; }
	if (x->side) cudaStreamDestroy(x->side);
	x->updated = nullptr;
	x->side = nullptr;
}

extern "C" {

int zn_shard_scenes(uint32_t num_scenes, uint32_t world_size, uint32_t rank, uint32_t* out_first, uint32_t* out_count) {
	ZN_REQUIRE(out_first && out_count, "zn_shard_scenes: NULL argument");
	ZN_REQUIRE(world_size >= 1 && rank < world_size, "zn_shard_scenes: rank %u outside world of %u", rank, world_size);
	const uint32_t base = num_scenes / world_size, extra = num_scenes % world_size;
	*out_first = rank * base + std::min(rank, extra);
	*out_count = base + (rank < extra ? 1u : 0u);
	return ZN_OK;
}

int zn_scene_index_bases(const uint32_t* entity_counts, uint32_t num_scenes, uint32_t* out_bases) {
	ZN_REQUIRE((entity_counts && out_bases) || num_scenes == 0, "zn_scene_index_bases: NULL argument");
	uint64_t acc = 0;
	for (uint32_t k = 0; k < num_scenes; ++k) {
		out_bases[k] = uint32_t(acc);
		acc += entity_counts[k];
		ZN_REQUIRE(acc <= (1ull << 32), "zn_scene_index_bases: global entity index does not fit in 32 bits");
	}
	return ZN_OK;
}

int zn_shard_scene_by_roots(const uint32_t* level_offsets, uint32_t level_count, const uint32_t* parent, uint32_t world_size,
                            uint32_t* out_owner, uint32_t* out_local_index, uint32_t* out_shard_counts) {
	ZN_REQUIRE(level_offsets && level_count >= 1 && out_owner && out_local_index && out_shard_counts, "zn_shard_scene_by_roots: NULL argument");
	ZN_REQUIRE(world_size >= 1, "zn_shard_scene_by_roots: world_size must be >= 1");
	const uint32_t E = level_offsets[level_count];
	// root of every entity: one pass in level order (a parent always lives in an earlier level)
	std::vector<uint32_t> root(E);
	std::vector<uint32_t> roots;
	for (uint32_t l = 0; l < level_count; ++l) {
		const uint32_t lo = level_offsets[l], hi = level_offsets[l + 1];
		ZN_REQUIRE(lo <= hi && hi <= E, "zn_shard_scene_by_roots: level_offsets is not ascending");
		for (uint32_t e = lo; e < hi; ++e) {
			const uint32_t p = parent ? parent[e] : ZN_NO_PARENT;
			if (p == ZN_NO_PARENT) { root[e] = e; roots.push_back(e); continue; }
			ZN_REQUIRE(p < lo, "zn_shard_scene_by_roots: level %u: a parent index is not in an earlier level", l);
			root[e] = root[p];
		}
	}
	const size_t R = roots.size();
	ZN_REQUIRE(R >= world_size, "zn_shard_scene_by_roots: %zu root subtrees cannot be split over %u ranks", R, world_size);
	std::vector<uint64_t> size_of(E, 0);
	for (uint32_t e = 0; e < E; ++e) size_of[root[e]]++;
	// Contiguous runs of roots: root k goes to the rank its cumulative midpoint falls in, then the cut
	// points are nudged so that no rank is left empty.
	std::vector<uint32_t> owner_of_root(R);
	{
		double cum = 0.0;
		uint32_t running = 0;
		const double total = double(E);
		for (size_t k = 0; k < R; ++k) {
			const double sub = double(size_of[roots[k]]);
			cum += sub;
			const double mid = cum - sub / 2.0;
			uint32_t o = uint32_t(std::min<double>(double(int64_t(mid * double(world_size) / total)), double(world_size - 1)));
			running = std::max(running, o);
			owner_of_root[k] = running;
		}
	}
	std::vector<size_t> cuts(world_size > 1 ? world_size - 1 : 0);   // first root of ranks 1..W-1
	for (uint32_t r = 0; r + 1 < world_size; ++r)
		cuts[r] = size_t(std::lower_bound(owner_of_root.begin(), owner_of_root.end(), r + 1) - owner_of_root.begin());
	for (uint32_t r = 0; r + 1 < world_size; ++r) {
		const size_t lo = r ? cuts[r - 1] + 1 : 1;
		cuts[r] = std::min(std::max(cuts[r], lo), R - (world_size - 1 - r));
	}
	for (size_t k = 0; k < R; ++k)
		owner_of_root[k] = uint32_t(std::upper_bound(cuts.begin(), cuts.end(), k) - cuts.begin());
	std::vector<uint32_t> owner_by_entity_root(E, 0);
	for (size_t k = 0; k < R; ++k) owner_by_entity_root[roots[k]] = owner_of_root[k];
	std::fill(out_shard_counts, out_shard_counts + world_size, 0u);
	for (uint32_t e = 0; e < E; ++e) {
		const uint32_t o = owner_by_entity_root[root[e]];
		out_owner[e] = o;
		out_local_index[e] = out_shard_counts[o]++;     // scene-wide order is kept inside a shard
	}
	return ZN_OK;
}

int zn_exchange_create(zn_ctx* ctx, const zn_exchange_desc* d, zn_exchange** out) {
	ZN_REQUIRE(ctx && d && out, "zn_exchange_create: NULL argument");
	*out = nullptr;
	ZN_REQUIRE(d->world_size >= 1 && d->rank < d->world_size, "zn_exchange_create: rank %u outside world of %u", d->rank, d->world_size);
	ZN_REQUIRE(d->world_size <= 8, "zn_exchange_create: at most 8 ranks (one NVSwitch domain; the kernels hold 7 peer pointers)");
	ZN_REQUIRE(d->records_per_rank >= 1 && d->world_size * d->records_per_rank <= 64, "zn_exchange_create: at most 64 records per exchange");
	ZN_REQUIRE(d->record_words >= 1 && d->list_capacity >= 1, "zn_exchange_create: record_words and list_capacity must be >= 1");
	ZN_REQUIRE(d->slots == 0 || d->slots >= 4, "zn_exchange_create: slots must be >= 4 (a record cell is reused four frames later)");
	ZN_REQUIRE(d->allgather || d->world_size == 1, "zn_exchange_create: an all-gather callback is required for world_size > 1");
	zn_exchange* x = new zn_exchange();
	x->ctx = ctx;
	x->world = d->world_size; x->rank = d->rank; x->per_rank = d->records_per_rank; x->total = x->world * x->per_rank;
	x->record_words = d->record_words; x->capacity = d->list_capacity;
	x->slots = d->slots ? d->slots : 4; x->timeout_ms = d->timeout_ms ? d->timeout_ms : 2000;
	x->allgather = d->allgather; x->user = d->user;
	x->overlap = (d->flags & ZN_EXCHANGE_OVERLAP) != 0;
	x->slot_bytes = size_t(x->total) * x->record_words * 4;
	x->flags_offset = (x->slots * x->slot_bytes + 127) / 128 * 128;
	// Set-up is collective: a rank that cannot allocate, export or map still takes part in every
	// rendezvous below, so all ranks learn of a failure together instead of one hanging the others.
	struct Hello { unsigned char handle[64]; uint32_t ok; uint32_t pad[15]; } mine = {};
	std::vector<Hello> all(x->world);
	char problem[256] = "";
	cudaSetDevice(ctx->device);
	if (cudaMalloc(&x->base, x->flags_offset + size_t(x->world) * kFlagStrideWords * 4) != cudaSuccess ||
	    cudaMemset(x->base, 0, x->flags_offset + size_t(x->world) * kFlagStrideWords * 4) != cudaSuccess) {
		snprintf(problem, sizeof(problem), "rank %u: cannot allocate the exchange buffer: %s", x->rank, cudaGetErrorString(cudaGetLastError()));
		x->base = nullptr;
	} else if (x->world > 1) {
		cudaIpcMemHandle_t h;
		if (cudaIpcGetMemHandle(&h, x->base) != cudaSuccess)
			snprintf(problem, sizeof(problem), "rank %u: cudaIpcGetMemHandle: %s", x->rank, cudaGetErrorString(cudaGetLastError()));
		else
			std::memcpy(mine.handle, &h, 64);
	}
	mine.ok = problem[0] ? 0u : 1u;
	all[x->rank] = mine;
	if (x->world > 1 && x->allgather(x->user, &mine, all.data(), sizeof(Hello)) != 0) {
		exchange_release(x); delete x;
		return fail(ZN_ERR_STATE, "zn_exchange_create: the all-gather callback failed");
	}
	bool everyone = true;
	for (const Hello& h : all) everyone = everyone && h.ok;
	x->peer_base.assign(x->world, nullptr);
	if (everyone) {
		for (uint32_t r = 0; r < x->world && !problem[0]; ++r) {
			if (r == x->rank) { x->peer_base[r] = x->base; continue; }
			cudaIpcMemHandle_t h;
			std::memcpy(&h, all[r].handle, 64);
			void* p = nullptr;
			if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
				snprintf(problem, sizeof(problem), "rank %u: cannot map rank %u's buffer: %s", x->rank, r, cudaGetErrorString(cudaGetLastError()));
			else
				x->peer_base[r] = static_cast<uint8_t*>(p);
		}
	}
	if (everyone && !problem[0]) {
		const size_t list_words = size_t(2) * x->total * x->capacity;
		if (cudaMalloc(&x->lists, list_words * 4) != cudaSuccess || cudaMalloc(&x->counts, size_t(2) * x->total * 4) != cudaSuccess ||
		    cudaMalloc(&x->status, 4) != cudaSuccess || cudaHostAlloc(&x->host_counts, (size_t(x->total) + 1) * 4, cudaHostAllocDefault) != cudaSuccess ||
		    cudaMemset(x->counts, 0, size_t(2) * x->total * 4) != cudaSuccess || cudaMemset(x->status, 0, 4) != cudaSuccess ||
		    cudaEventCreateWithFlags(&x->updated, cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreateWithFlags(&x->done[0], cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreateWithFlags(&x->done[1], cudaEventDisableTiming) != cudaSuccess)
			snprintf(problem, sizeof(problem), "rank %u: cannot allocate the list buffers: %s", x->rank, cudaGetErrorString(cudaGetLastError()));
	}
	if (everyone && !problem[0]) {
		// the side stream runs at the LOWEST priority: the frame on the context stream goes first wherever both want an SM slot
		int least = 0, greatest = 0;
		cudaDeviceGetStreamPriorityRange(&least, &greatest);
		if (cudaStreamCreateWithPriority(&x->side, cudaStreamNonBlocking, least) != cudaSuccess)
			snprintf(problem, sizeof(problem), "rank %u: cannot create the side stream: %s", x->rank, cudaGetErrorString(cudaGetLastError()));
	}
	// second rendezvous: every rank has mapped every buffer before anyone writes — and every rank hears of any failure
	Hello second = {};
	second.ok = (everyone && !problem[0]) ? 1u : 0u;
	all[x->rank] = second;
	if (x->world > 1 && x->allgather(x->user, &second, all.data(), sizeof(Hello)) != 0) std::snprintf(problem, sizeof(problem), "rank %u: the all-gather callback failed", x->rank);
	bool good = !problem[0];
	uint32_t bad_rank = x->rank;
	for (uint32_t r = 0; r < x->world; ++r)
		if (!all[r].ok) { good = false; if (!problem[0]) bad_rank = r; }
	if (!good) {
		exchange_release(x); delete x;
		if (problem[0]) return fail(ZN_ERR_CUDA, "zn_exchange_create: peer memory unavailable: %s", problem);
		return fail(ZN_ERR_CUDA, "zn_exchange_create: peer memory unavailable: rank %u could not allocate, export or map", bad_rank);
	}
	*out = x;
	return ZN_OK;
}

int zn_exchange_unbind(zn_exchange* x, zn_scene* const* scenes, uint32_t scene_count) {
	ZN_REQUIRE(x, "zn_exchange_unbind: exchange is NULL");
	for (uint32_t j = 0; j < scene_count; ++j) {
		if (!scenes || !scenes[j]) continue;
		zn_scene_bind_peer_records(scenes[j], nullptr, 0);
		zn_scene_bind_peer_signal(scenes[j], nullptr, 0);
		zn_scene_bind_visibility_record(scenes[j], nullptr);
		zn_scene_bind_visible_output(scenes[j], nullptr, 0, nullptr);
	}
	return ZN_OK;
}

int zn_exchange_destroy(zn_exchange* x, zn_scene* const* scenes, uint32_t scene_count) {
	if (!x) return ZN_OK;
	zn_exchange_unbind(x, scenes, scene_count);
	cudaSetDevice(x->ctx->device);
	cudaStreamSynchronize(x->ctx->stream);
	if (x->side) cudaStreamSynchronize(x->side);
	// nobody may still be writing into a buffer that is about to go away: rendezvous, release, rendezvous
	uint32_t token = x->rank;
	std::vector<uint32_t> tokens(x->world);
	int rc = ZN_OK;
	if (x->world > 1 && x->allgather(x->user, &token, tokens.data(), 4) != 0) rc = fail(ZN_ERR_STATE, "zn_exchange_destroy: the all-gather callback failed");
	exchange_release(x);
	if (x->world > 1 && rc == ZN_OK && x->allgather(x->user, &token, tokens.data(), 4) != 0) rc = fail(ZN_ERR_STATE, "zn_exchange_destroy: the all-gather callback failed");
	delete x;
	return rc;
}

int zn_exchange_bind(zn_exchange* x, zn_scene* const* scenes, uint32_t scene_count, uint64_t frame) {
	ZN_REQUIRE(x && scenes, "zn_exchange_bind: NULL argument");
	ZN_REQUIRE(scene_count == x->per_rank, "zn_exchange_bind: %u scenes given, this rank owns %u records", scene_count, x->per_rank);
	ZN_REQUIRE(int64_t(frame) == x->issued + 1, "zn_exchange_bind: frames are bound in order (expected %lld)", (long long)(x->issued + 1));
	const uint32_t slot = uint32_t(frame % x->slots);
	ZN_CUDA(cudaSetDevice(x->ctx->device));
	// flow control: update(frame) may not start before this rank's own expand(frame-2) has completed
	if (x->overlap && x->expanded >= int64_t(frame) - 2 && frame >= 2) ZN_CUDA(cudaStreamWaitEvent(x->ctx->stream, x->done[frame % 2], 0));
	for (uint32_t j = 0; j < scene_count; ++j) {
		zn_scene* sc = scenes[j];
		ZN_REQUIRE(sc && sc->ctx == x->ctx, "zn_exchange_bind: scene %u is NULL or belongs to another context", j);
		// every exchanged scene shares ONE level layout: the expansion reads all records with the layout of the scene it is given
		ZN_REQUIRE(sc->record_words == x->record_words && sc->entity_count <= x->capacity,
		           "zn_exchange_bind: scene %u (%u record words, %u entities) does not match the exchange (%u record words, capacity %u)", j,
		           sc->record_words, sc->entity_count, x->record_words, x->capacity);
		const uint32_t record = x->rank * x->per_rank + j;
		const size_t off = slot * x->slot_bytes + size_t(record) * x->record_words * 4;
		void* peers[8];
		uint32_t n = 0;
		for (uint32_t r = 0; r < x->world; ++r)
			if (r != x->rank) peers[n++] = x->peer_base[r] + off;
		int rc = zn_scene_bind_visibility_record(sc, x->base + off);
		if (rc == ZN_OK) rc = zn_scene_bind_peer_records(sc, peers, n);
		const size_t list_at = (size_t(frame % 2) * x->total + record) * x->capacity;
		if (rc == ZN_OK) rc = zn_scene_bind_visible_output(sc, x->lists + list_at, x->capacity, x->counts + (frame % 2) * x->total + record);
		if (rc != ZN_OK) return rc;
	}
	// the FIRST scene carries the signal: the expansion launched with it raises this rank's flag in every rank's allocation
	void* flags[8];
	for (uint32_t r = 0; r < x->world; ++r) flags[r] = x->peer_base[r] + x->flags_offset + size_t(x->rank) * kFlagStrideWords * 4;
	int rc = zn_scene_bind_peer_signal(scenes[0], flags, x->world);
	if (rc == ZN_OK) rc = zn_scene_set_signal_value(scenes[0], uint32_t(frame + 1));
	if (rc != ZN_OK) return rc;
	x->issued = int64_t(frame);
	return ZN_OK;
}

// Measurement knob (A/B runs on one box): ZN_EXCHANGE_EARLY_START=0 launches the expansion as a normal, fully
// stream-ordered kernel instead of programmatically beside the tail of the update before it.
static bool early_start_enabled() {
	static const bool on = [] {
		const char* v = std::getenv("ZN_EXCHANGE_EARLY_START");
		return !(v && v[0] == '0');
	}();
	return on;
}

static int exchange_launch(zn_exchange* x, zn_scene* scene, int64_t frame, const uint32_t* index_bases, bool signal_only) {
	ZN_CUDA(cudaSetDevice(x->ctx->device));
	cudaStream_t stream = x->ctx->stream;
	if (x->overlap) {
		// side stream, behind everything enqueued on the context stream so far (the update of `frame` included)
		ZN_CUDA(cudaEventRecord(x->updated, x->ctx->stream));
		ZN_CUDA(cudaStreamWaitEvent(x->side, x->updated, 0));
		stream = x->side;
	}
	const uint8_t* records = x->base + size_t(frame % x->slots) * x->slot_bytes;
	const size_t half = size_t(frame % 2) * x->total;
	const uint32_t first_own = x->rank * x->per_rank;
	return expand_visible_on(stream, scene, signal_only ? "zn_exchange_signal" : "zn_exchange_expand", records, x->total, x->record_words,
	                         signal_only ? 0 : int(first_own), signal_only ? x->total : x->per_rank, index_bases, x->lists + half * x->capacity,
	                         x->capacity, x->counts + half, x->base + x->flags_offset, kFlagStrideWords, x->per_rank, uint32_t(frame + 1),
	                         x->timeout_ms, x->status, nullptr, /*early_start=*/!x->overlap && early_start_enabled());
}

int zn_exchange_signal(zn_exchange* x, zn_scene* scene) {
	ZN_REQUIRE(x && scene, "zn_exchange_signal: NULL argument");
	if (x->issued < 0) return fail(ZN_ERR_STATE, "zn_exchange_signal: no frame has been bound yet");
	const int rc = exchange_launch(x, scene, x->issued, nullptr, true);
	if (rc != ZN_OK) return rc;
	// synchronous: when this returns the flags are up (the update they announce is complete)
	ZN_CUDA(cudaStreamSynchronize(x->overlap ? x->side : x->ctx->stream));
	return ZN_OK;
}

int zn_exchange_expand(zn_exchange* x, zn_scene* scene, int64_t frame, const uint32_t* index_bases) {
	ZN_REQUIRE(x && scene, "zn_exchange_expand: NULL argument");
	if (frame < 0 || frame > x->issued || frame <= x->expanded) return ZN_OK;   // never issued, or already expanded
	ZN_REQUIRE(frame == x->expanded + 1, "zn_exchange_expand: frames are expanded in order (expected %lld)", (long long)(x->expanded + 1));
	const int rc = exchange_launch(x, scene, frame, index_bases, false);
	if (rc != ZN_OK) return rc;
	if (x->overlap) ZN_CUDA(cudaEventRecord(x->done[frame % 2], x->side));
	x->expanded = frame;
	return ZN_OK;
}

int zn_exchange_join(zn_exchange* x) {
	ZN_REQUIRE(x, "zn_exchange_join: exchange is NULL");
	if (!x->overlap || x->expanded < 0) return ZN_OK;
	ZN_CUDA(cudaSetDevice(x->ctx->device));
	ZN_CUDA(cudaStreamWaitEvent(x->ctx->stream, x->done[x->expanded % 2], 0));
	return ZN_OK;
}

int zn_exchange_result(zn_exchange* x, uint32_t* out_counts, void** out_lists_dev, uint32_t* out_list_stride) {
	ZN_REQUIRE(x, "zn_exchange_result: exchange is NULL");
	if (x->expanded < 0) return fail(ZN_ERR_STATE, "zn_exchange_result: no frame has been expanded yet");
	ZN_CUDA(cudaSetDevice(x->ctx->device));
	if (x->overlap) ZN_CUDA(cudaStreamWaitEvent(x->ctx->stream, x->done[x->expanded % 2], 0));
	const size_t half = size_t(x->expanded % 2) * x->total;
	ZN_CUDA(cudaMemcpyAsync(x->host_counts, x->counts + half, size_t(x->total) * 4, cudaMemcpyDeviceToHost, x->ctx->stream));
	ZN_CUDA(cudaMemcpyAsync(x->host_counts + x->total, x->status, 4, cudaMemcpyDeviceToHost, x->ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(x->ctx->stream));
	const uint32_t status = x->host_counts[x->total];
	if (status)
		return fail(ZN_ERR_STATE, "zn_exchange_result: rank %u: rank %u (record %u) never signalled a frame within %u ms", x->rank,
		            (status - 1) / x->per_rank, status - 1, x->timeout_ms);
	if (out_counts) std::memcpy(out_counts, x->host_counts, size_t(x->total) * 4);
	if (out_lists_dev) *out_lists_dev = x->lists + half * x->capacity;
	if (out_list_stride) *out_list_stride = x->capacity;
	return ZN_OK;
}

int zn_exchange_info(zn_exchange* x, uint32_t* out_total_records, uint32_t* out_list_capacity, uint64_t* out_peer_bytes_per_frame) {
	ZN_REQUIRE(x, "zn_exchange_info: exchange is NULL");
	if (out_total_records) *out_total_records = x->total;
	if (out_list_capacity) *out_list_capacity = x->capacity;
	if (out_peer_bytes_per_frame) *out_peer_bytes_per_frame = uint64_t(x->per_rank) * x->record_words * 4 * (x->world - 1);
	return ZN_OK;
}

} // extern "C"
