// zn_mesh.cu — K4: bulk mesh vertex / normal transform by per-instance model matrices.
//
//   pos' = model.transform_point(pos)                                   (matrix.hpp:45-46)
//   nrm' = normalize(model.inverse_transpose().transform_normal(nrm))   (matrix.hpp:43,47-48)
//
// Vertices are packed float[V][3] streams (48 algorithmic bytes per vertex in + out).  The stream
// is cut into tiles of 1024 vertices regardless of instance boundaries, so every tile is 16-byte
// aligned: thread 0 pulls the tile's positions and normals into shared memory with two
// cp.async.bulk copies (12 KiB each), every thread transforms 4 vertices in place (12-word stride:
// conflict-free float4 shared accesses) and the tile leaves as two bulk stores.  No thread issues
// a global load or store for vertex data, and HBM only sees 12 KiB bursts.
//   - a tile inside ONE instance (the common case for large meshes): thread 32 inverts that
//     instance's model matrix while the copies are in flight;
//   - a tile spanning SEVERAL instances (small meshes): while the copies are in flight, thread k
//     inverts the k-th instance's model matrix into shared memory (up to 256 instances per tile: no
//     matrix table in HBM, 64 B read per instance instead of 64 + 96 + 96) and every thread finds the
//     instance of its FIRST vertex by binary search over the tile's slice of instance_first_vertex
//     in shared memory; its other three vertices advance linearly from there.  The shared window is
//     sized per mesh (dynamic shared memory: the most instances any of its tiles touches, up to
//     1024 — one vertex per instance); only tiles that touch more than that (through runs of EMPTY
//     instances) fall back to a per-instance matrix table built by mesh_instance_mats_kernel.
#include "zn_internal.hpp"
#include "zn_math.cuh"
#include "zn_ptx.cuh"

#include <algorithm>
#include <vector>

namespace zn {

constexpr int kMeshThreads = kMeshTileVerts / 4;
constexpr uint32_t kMeshSmemMats = 1024;       // most instances of one tile whose matrices / boundaries are kept in shared memory
                                               // (a 1024-vertex tile touches more only through EMPTY instances)
constexpr int kSmemMatFloats = 21;             // rows 0..2 of the model (12) + 3x3 normal matrix (9); odd stride: conflict-free
constexpr int kInstMatFloats = 24;             // the same in global memory (tiles with > kMeshSmemMats instances), padded to 16 bytes

// N[r*3 + c] = ((model^-1)^T)(r, c) = inverse(c, r), r, c < 3.  A scene's world matrices are affine (bottom row exactly
// 0,0,0,1): for those the general 4x4 inverse folds to ~45 operations with bit-identical results
// (mat4_inverse_affine3x3); anything else — a projective model, an infinity — takes the general routine.
__device__ __forceinline__ void normal_matrix(const float* __restrict__ m, float* __restrict__ N) {
	float inv3[9];
	if (mat4_inverse_affine3x3(m, inv3)) {
#pragma unroll
		for (int r = 0; r < 3; ++r)
#pragma unroll
			for (int c = 0; c < 3; ++c) N[r * 3 + c] = inv3[c * 3 + r];
		return;
	}
	float inv[16];
	mat4_inverse(m, inv);
#pragma unroll
	for (int r = 0; r < 3; ++r)
#pragma unroll
		for (int c = 0; c < 3; ++c) N[r * 3 + c] = inv[r * 4 + c];   // transpose(inverse)(r,c) = inverse(c,r)
}

// Per-instance matrices for tiles that span several instances.
__global__ void mesh_instance_mats_kernel(const float* __restrict__ models, uint32_t instances, float* __restrict__ out) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= instances) return;
	float m[16];
	const float4* src = reinterpret_cast<const float4*>(models + size_t(k) * 16);
#pragma unroll
	for (int c = 0; c < 4; ++c) {
		const float4 col = __ldg(src + c);
		m[c * 4 + 0] = col.x; m[c * 4 + 1] = col.y; m[c * 4 + 2] = col.z; m[c * 4 + 3] = col.w;
	}
	float* o = out + size_t(k) * kInstMatFloats;
	float N[9];
	normal_matrix(m, N);
#pragma unroll
	for (int r = 0; r < 3; ++r) {
#pragma unroll
		for (int c = 0; c < 4; ++c) o[r * 4 + c] = m[c * 4 + r];
#pragma unroll
		for (int c = 0; c < 3; ++c) o[12 + r * 3 + c] = N[r * 3 + c];
	}
}

// One vertex: position by the model rows M[12], normal by N[9] then normalised (zero left as is).
__device__ __forceinline__ void transform_vertex(const float* __restrict__ M, const float* __restrict__ N, float* p, float* n) {
	const float px = p[0], py = p[1], pz = p[2];
	const float nx = n[0], ny = n[1], nz = n[2];
	float q[3], g[3];
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		q[r] = add(add(add(mul(M[r * 4 + 0], px), mul(M[r * 4 + 1], py)), mul(M[r * 4 + 2], pz)), M[r * 4 + 3]);
		g[r] = add(add(mul(N[r * 3 + 0], nx), mul(N[r * 3 + 1], ny)), mul(N[r * 3 + 2], nz));
	}
	const float len = __fsqrt_rn(add(add(mul(g[0], g[0]), mul(g[1], g[1])), mul(g[2], g[2])));
	if (len > 0.0f) { g[0] = __fdiv_rn(g[0], len); g[1] = __fdiv_rn(g[1], len); g[2] = __fdiv_rn(g[2], len); }
	p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
	n[0] = g[0]; n[1] = g[1]; n[2] = g[2];
}

// MULTI = false: every tile of the mesh lies inside one instance (the lean kernel the 64Mi-vertex
// batch runs); MULTI = true also handles tiles that span several instances.
template <bool MULTI, bool UNIFORM>
__global__ void __launch_bounds__(kMeshThreads, 4)
mesh_tile_kernel(const zn_mesh_tile* __restrict__ tiles, const float* __restrict__ models,
                 const uint32_t* __restrict__ instance_first_vertex, const float* __restrict__ inst_mats,
                 const float* __restrict__ pos_in, const float* __restrict__ nrm_in,
                 float* __restrict__ pos_out, float* __restrict__ nrm_out, uint32_t window, uint32_t uniform_vpi,
                 uint32_t total_vertices) {
	__shared__ __align__(128) float s_pos[kMeshTileVerts * 3];
	__shared__ __align__(128) float s_nrm[kMeshTileVerts * 3];
	__shared__ float s_model[12];   // rows 0..2 of the model, row-major [r*4+c]
	__shared__ float s_nm[9];       // upper 3x3 of (model^-1)^T, row-major
	extern __shared__ uint32_t s_dyn[];   // MULTI: first[window + 1], then mats[window][21]
	uint32_t* s_first = s_dyn;
	float* s_mats = reinterpret_cast<float*>(s_dyn + window + 1);
	__shared__ uint64_t s_bar;

	const int t = threadIdx.x;
	// Tiles are cut from the global vertex stream, so a tile's extent is arithmetic: the bulk copies are issued without
	// waiting for anything; only the instances a tile touches come from the table.  With equal instance ranges
	// (uniform_vpi vertices each — instanced draws of one mesh) those and every boundary are arithmetic too: neither the
	// tile table nor instance_first_vertex is read and two dependent round trips leave the tile's start-up chain
	// (4Mi x 8 vertices: 93 -> 113 G vertices/s).
	zn_mesh_tile tile;
	tile.first_vertex = blockIdx.x * uint32_t(kMeshTileVerts);
	tile.vertex_count = min(uint32_t(kMeshTileVerts), total_vertices - tile.first_vertex);
	const uint32_t floats = ((tile.vertex_count + 3u) & ~3u) * 3u;   // whole 4-vertex groups: a multiple of 16 bytes
	const size_t base = size_t(tile.first_vertex) * 3;
	if (t == 0) {
		mbar_init(&s_bar, 1);
		fence_mbar_init();
		mbar_arrive_expect_tx(&s_bar, floats * 8);
		bulk_g2s(s_pos, pos_in + base, floats * 4, &s_bar);
		bulk_g2s(s_nrm, nrm_in + base, floats * 4, &s_bar);
	}
	if (UNIFORM) {
		tile.instance_lo = tile.first_vertex / uniform_vpi;
		tile.instance_hi = (tile.first_vertex + tile.vertex_count - 1u) / uniform_vpi;
	} else {
		const uint4 td = __ldg(reinterpret_cast<const uint4*>(tiles) + blockIdx.x);   // {first_vertex, vertex_count, instance_lo, instance_hi}
		tile.instance_lo = td.z;
		tile.instance_hi = td.w;
	}
	constexpr bool uniform = UNIFORM;
	const bool single = !MULTI || tile.instance_lo == tile.instance_hi;
	const uint32_t n_inst = tile.instance_hi - tile.instance_lo + 1u;
	const bool staged = MULTI && n_inst <= window;
	const bool smem_mats = staged;

	// Model rows + normal matrix of one instance: M[r*4+c] = model(r,c), N[r*3+c] = inverse(c,r).
	auto instance_matrices = [&](const float4* col, float* M, float* N) {
		float m[16];
#pragma unroll
		for (int c = 0; c < 4; ++c) { m[c * 4 + 0] = col[c].x; m[c * 4 + 1] = col[c].y; m[c * 4 + 2] = col[c].z; m[c * 4 + 3] = col[c].w; }
		normal_matrix(m, N);
#pragma unroll
		for (int r = 0; r < 3; ++r)
#pragma unroll
			for (int c = 0; c < 4; ++c) M[r * 4 + c] = m[c * 4 + r];
	};

	if (t == 32 && single) {
		const float4* src = reinterpret_cast<const float4*>(models + size_t(tile.instance_lo) * 16);
		const float4 col[4] = {__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3)};
		instance_matrices(col, s_model, s_nm);
	}
	// ---- while the copies are in flight: instance boundaries, per-instance matrices, first-vertex search ----
	if (MULTI && !single) {
		if (staged && !uniform)
			for (uint32_t k = t; k <= n_inst; k += kMeshThreads) s_first[k] = __ldg(instance_first_vertex + tile.instance_lo + k);
		if (smem_mats)
			for (uint32_t k = t; k < n_inst; k += kMeshThreads) {
				// an instance without a vertex in this tile (empty, or all its vertices behind the tile's end) needs no matrix;
				// the model is fetched alongside the two boundaries (one round trip), the inverse only computed if needed
				const uint32_t begin = uniform ? (tile.instance_lo + k) * uniform_vpi : __ldg(instance_first_vertex + tile.instance_lo + k);
				const uint32_t end = uniform ? begin + uniform_vpi : __ldg(instance_first_vertex + tile.instance_lo + k + 1);
				const uint32_t lo = max(begin, tile.first_vertex);
				const uint32_t hi = min(end, tile.first_vertex + tile.vertex_count);
				const float4* src = reinterpret_cast<const float4*>(models + size_t(tile.instance_lo + k) * 16);
				const float4 col[4] = {__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3)};
				if (lo >= hi) continue;
				float M[12], N[9];
				instance_matrices(col, M, N);
				float* o = s_mats + k * kSmemMatFloats;
#pragma unroll
				for (int i = 0; i < 12; ++i) o[i] = M[i];
#pragma unroll
				for (int i = 0; i < 9; ++i) o[12 + i] = N[i];
			}
	}
	__syncthreads();
	auto first_of = [&](uint32_t k) {
		return uniform ? (tile.instance_lo + k) * uniform_vpi : staged ? s_first[k] : __ldg(instance_first_vertex + tile.instance_lo + k);
	};
	uint32_t inst = 0;
	if (uniform && !single) {
		inst = (tile.first_vertex + uint32_t(t) * 4) / uniform_vpi - tile.instance_lo;
		if (inst >= n_inst) inst = n_inst - 1;       // threads past the tile's last vertex
	} else if (MULTI && !single && uint32_t(t) * 4 < tile.vertex_count) {
		// instance of this thread's first vertex: the last k in [0, n_inst) with first[k] <= vertex
		const uint32_t vertex = tile.first_vertex + uint32_t(t) * 4;
		uint32_t lo = 0, hi = n_inst - 1;
		while (lo < hi) {
			const uint32_t mid = (lo + hi + 1) >> 1;
			if (first_of(mid) <= vertex) lo = mid; else hi = mid - 1;
		}
		inst = lo;
	}
	mbar_wait(&s_bar, 0);

	if (uint32_t(t) * 4 < tile.vertex_count) {
		float4* pp = reinterpret_cast<float4*>(s_pos) + t * 3;
		float4* np = reinterpret_cast<float4*>(s_nrm) + t * 3;
		float4 a = pp[0], b = pp[1], c = pp[2];
		float p[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
		a = np[0]; b = np[1]; c = np[2];
		float n[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
		float M[12], N[9];
		if (single) {
#pragma unroll
			for (int k = 0; k < 12; ++k) M[k] = s_model[k];
#pragma unroll
			for (int k = 0; k < 9; ++k) N[k] = s_nm[k];
#pragma unroll
			for (int v = 0; v < 4; ++v) transform_vertex(M, N, p + 3 * v, n + 3 * v);
		} else {
			uint32_t have = 0xFFFFFFFFu;
#pragma unroll
			for (int v = 0; v < 4; ++v) {
				const uint32_t vertex = tile.first_vertex + uint32_t(t) * 4 + v;
				while (inst + 1 < n_inst && first_of(inst + 1) <= vertex) ++inst;   // the next vertices advance from the first one's instance
				if (inst != have) {
					if (smem_mats) {
						const float* src = s_mats + inst * kSmemMatFloats;
#pragma unroll
						for (int k = 0; k < 12; ++k) M[k] = src[k];
#pragma unroll
						for (int k = 0; k < 9; ++k) N[k] = src[12 + k];
					} else {
						const float4* src = reinterpret_cast<const float4*>(inst_mats + size_t(tile.instance_lo + inst) * kInstMatFloats);
#pragma unroll
						for (int k = 0; k < 3; ++k) { const float4 r = __ldg(src + k); M[4 * k] = r.x; M[4 * k + 1] = r.y; M[4 * k + 2] = r.z; M[4 * k + 3] = r.w; }
						const float4 r3 = __ldg(src + 3), r4 = __ldg(src + 4);
						const float r5 = __ldg(reinterpret_cast<const float*>(src + 5));
						N[0] = r3.x; N[1] = r3.y; N[2] = r3.z; N[3] = r3.w; N[4] = r4.x; N[5] = r4.y; N[6] = r4.z; N[7] = r4.w; N[8] = r5;
					}
					have = inst;
				}
				transform_vertex(M, N, p + 3 * v, n + 3 * v);
			}
		}
		pp[0] = make_float4(p[0], p[1], p[2], p[3]); pp[1] = make_float4(p[4], p[5], p[6], p[7]); pp[2] = make_float4(p[8], p[9], p[10], p[11]);
		np[0] = make_float4(n[0], n[1], n[2], n[3]); np[1] = make_float4(n[4], n[5], n[6], n[7]); np[2] = make_float4(n[8], n[9], n[10], n[11]);
	}
	fence_proxy_async_smem();
	__syncthreads();
	if (t == 0) {
		bulk_s2g(pos_out + base, s_pos, floats * 4);
		bulk_s2g(nrm_out + base, s_nrm, floats * 4);
		bulk_commit();
		bulk_wait_read_all();
	}
}

// strided 3-float host records -> packed float[count][3]
__global__ void pack3_kernel(const float* __restrict__ src, uint32_t stride_f, uint32_t count, float* __restrict__ dst) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const float* s = src + size_t(i) * stride_f;
	dst[size_t(i) * 3 + 0] = s[0]; dst[size_t(i) * 3 + 1] = s[1]; dst[size_t(i) * 3 + 2] = s[2];
}

static int upload_vertices(zn_mesh* m, float* dst, uint32_t first, uint32_t count, const float* host, size_t stride_bytes) {
	if (!host || count == 0) return ZN_OK;
	zn_ctx* ctx = m->ctx;
	if (stride_bytes == 12) {
		ZN_CUDA(cudaMemcpyAsync(dst + size_t(first) * 3, host, size_t(count) * 12, cudaMemcpyHostToDevice, ctx->stream));
		return ZN_OK;
	}
	const uint32_t chunk = 1u << 22;
	const size_t need = size_t(std::min(count, chunk)) * stride_bytes;
	if (ctx->staging_bytes < need) {
		ZN_CUDA(cudaStreamSynchronize(ctx->stream));
		if (ctx->staging) ZN_CUDA(cudaFree(ctx->staging));
		ctx->staging = nullptr; ctx->staging_bytes = 0;
		ZN_CUDA(cudaMalloc(&ctx->staging, need));
		ctx->staging_bytes = need;
	}
	for (uint32_t done = 0; done < count; done += chunk) {
		const uint32_t n = std::min(chunk, count - done);
		ZN_CUDA(cudaMemcpyAsync(ctx->staging, reinterpret_cast<const uint8_t*>(host) + size_t(done) * stride_bytes,
		                        size_t(n - 1) * stride_bytes + 12, cudaMemcpyHostToDevice, ctx->stream));
		pack3_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(static_cast<const float*>(ctx->staging), uint32_t(stride_bytes / 4), n,
		                                                      dst + size_t(first + done) * 3);
		ZN_CUDA(cudaGetLastError());
	}
	return ZN_OK;
}

} // namespace zn

using namespace zn;

extern "C" {

int zn_mesh_create(zn_ctx* ctx, const zn_mesh_desc* desc, zn_mesh** out_mesh) {
	ZN_REQUIRE(ctx && desc && out_mesh, "zn_mesh_create: NULL argument");
	*out_mesh = nullptr;
	const uint32_t V = desc->vertex_count, I = desc->instance_count;
	ZN_REQUIRE(V > 0 && I > 0 && I <= V, "zn_mesh_create: need 0 < instance_count <= vertex_count (got %u, %u)", I, V);
	ZN_REQUIRE(V <= 0x50000000u, "zn_mesh_create: vertex_count %u too large", V);
	std::vector<uint32_t> first(size_t(I) + 1);
	if (desc->instance_first_vertex) {
		first.assign(desc->instance_first_vertex, desc->instance_first_vertex + I + 1);
		ZN_REQUIRE(first.front() == 0 && first.back() == V, "zn_mesh_create: instance_first_vertex must start at 0 and end at vertex_count");
		for (uint32_t k = 0; k < I; ++k) ZN_REQUIRE(first[k] <= first[k + 1], "zn_mesh_create: instance_first_vertex is not ascending at %u", k);
	} else {
		ZN_REQUIRE(V % I == 0, "zn_mesh_create: vertex_count %u is not a multiple of instance_count %u", V, I);
		for (uint32_t k = 0; k <= I; ++k) first[k] = uint32_t(uint64_t(V / I) * k);
	}
	// tiles of 1024 vertices over the whole stream; [instance_lo, instance_hi] = the instances a tile touches
	std::vector<zn_mesh_tile> tiles;
	bool multi = false, table = false;
	uint32_t window = 0;
	{
		uint32_t k = 0;
		for (uint32_t v = 0; v < V; v += kMeshTileVerts) {
			zn_mesh_tile t;
			t.first_vertex = v;
			t.vertex_count = std::min<uint32_t>(kMeshTileVerts, V - v);
			while (k + 1 < I && first[k + 1] <= v) ++k;                      // last instance starting at or before v
			t.instance_lo = k;
			uint32_t h = k;
			while (h + 1 < I && first[h + 1] <= v + t.vertex_count - 1) ++h; // last instance starting inside the tile
			t.instance_hi = h;
			multi = multi || (h != k);
			if (h != k) {
				if (h - k + 1 > kMeshSmemMats) table = true;
				else window = std::max(window, h - k + 1);
			}
			tiles.push_back(t);
		}
	}
	ZN_CUDA(cudaSetDevice(ctx->device));
	zn_mesh* m = new zn_mesh();
	m->ctx = ctx;
	m->vertex_count = V;
	m->instance_count = I;
	m->tile_count = uint32_t(tiles.size());
	auto cleanup = [&](int rc) { zn_mesh_destroy(m); return rc; };
#define ZN_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return cleanup(fail(e_ == cudaErrorMemoryAllocation ? ZN_ERR_NOMEM : ZN_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_))); } while (0)
	ZN_TRY(cudaMalloc(&m->tiles, tiles.size() * sizeof(zn_mesh_tile)));
	const size_t vpad = (size_t(V) + 3) & ~size_t(3);   // bulk copies move whole 4-vertex groups
	ZN_TRY(cudaMalloc(&m->pos_in, vpad * 12));
	ZN_TRY(cudaMalloc(&m->nrm_in, vpad * 12));
	ZN_TRY(cudaMalloc(&m->pos_out, vpad * 12));
	ZN_TRY(cudaMalloc(&m->nrm_out, vpad * 12));
	ZN_TRY(cudaMemsetAsync(m->pos_in, 0, vpad * 12, ctx->stream));
	ZN_TRY(cudaMemsetAsync(m->nrm_in, 0, vpad * 12, ctx->stream));
	ZN_TRY(cudaMalloc(&m->instance_first_vertex, (size_t(I) + 1) * 4));
	ZN_TRY(cudaMemcpyAsync(m->instance_first_vertex, first.data(), (size_t(I) + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
	m->multi_instance_tiles = multi;
	m->needs_inst_mats = table;
	// equal ranges — no table given, or a table that says so: boundaries are arithmetic, the kernel never reads the table
	bool equal_ranges = V % I == 0;
	for (uint32_t k = 0; equal_ranges && k <= I; ++k) equal_ranges = first[k] == uint32_t(uint64_t(V / I) * k);
	m->uniform_vpi = (equal_ranges && !table) ? V / I : 0u;
	m->smem_window = window;
	if (table) ZN_TRY(cudaMalloc(&m->inst_mats, size_t(I) * kInstMatFloats * 4));
	ZN_TRY(cudaMalloc(&m->models_own, size_t(I) * 64));
	ZN_TRY(cudaMemcpyAsync(m->tiles, tiles.data(), tiles.size() * sizeof(zn_mesh_tile), cudaMemcpyHostToDevice, ctx->stream));
	ZN_TRY(cudaMemsetAsync(m->models_own, 0, size_t(I) * 64, ctx->stream));
	ZN_TRY(cudaStreamSynchronize(ctx->stream));
#undef ZN_TRY
	m->models = m->models_own;
	*out_mesh = m;
	return ZN_OK;
}

int zn_mesh_destroy(zn_mesh* m) {
	if (!m) return ZN_OK;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	cudaFree(m->tiles); cudaFree(m->pos_in); cudaFree(m->nrm_in); cudaFree(m->pos_out); cudaFree(m->nrm_out); cudaFree(m->models_own);
	cudaFree(m->instance_first_vertex); cudaFree(m->inst_mats);
	delete m;
	return ZN_OK;
}

int zn_mesh_set_vertices(zn_mesh* m, uint32_t first, uint32_t count, const float* position, const float* normal, size_t stride_bytes) {
	ZN_REQUIRE(m, "zn_mesh_set_vertices: mesh is NULL");
	ZN_REQUIRE(uint64_t(first) + count <= m->vertex_count, "zn_mesh_set_vertices: range exceeds vertex_count %u", m->vertex_count);
	ZN_REQUIRE(stride_bytes >= 12 && stride_bytes % 4 == 0, "zn_mesh_set_vertices: stride_bytes %zu must be >= 12 and a multiple of 4", stride_bytes);
	ZN_CUDA(cudaSetDevice(m->ctx->device));
	int rc = upload_vertices(m, m->pos_in, first, count, position, stride_bytes);
	if (rc == ZN_OK) rc = upload_vertices(m, m->nrm_in, first, count, normal, stride_bytes);
	if (rc != ZN_OK) return rc;
	ZN_CUDA(cudaStreamSynchronize(m->ctx->stream));
	return ZN_OK;
}

int zn_mesh_set_models(zn_mesh* m, uint32_t first, uint32_t count, const float* model_matrices) {
	ZN_REQUIRE(m && (model_matrices || count == 0), "zn_mesh_set_models: NULL argument");
	ZN_REQUIRE(uint64_t(first) + count <= m->instance_count, "zn_mesh_set_models: range exceeds instance_count %u", m->instance_count);
	ZN_CUDA(cudaSetDevice(m->ctx->device));
	ZN_CUDA(cudaMemcpyAsync(m->models_own + size_t(first) * 16, model_matrices, size_t(count) * 64, cudaMemcpyHostToDevice, m->ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(m->ctx->stream));
	m->models = m->models_own;
	return ZN_OK;
}

int zn_mesh_bind_scene_models(zn_mesh* m, zn_scene* scene, uint32_t first_entity) {
	ZN_REQUIRE(m && scene, "zn_mesh_bind_scene_models: NULL argument");
	ZN_REQUIRE(m->ctx == scene->ctx, "zn_mesh_bind_scene_models: mesh and scene belong to different contexts");
	ZN_REQUIRE(uint64_t(first_entity) + m->instance_count <= scene->entity_count,
	           "zn_mesh_bind_scene_models: entities [%u, %u+%u) exceed the scene's %u", first_entity, first_entity, m->instance_count, scene->entity_count);
	m->models = scene->world + size_t(first_entity) * 16;
	return ZN_OK;
}

int zn_mesh_transform(zn_mesh* m) {
	ZN_REQUIRE(m, "zn_mesh_transform: mesh is NULL");
	ZN_CUDA(cudaSetDevice(m->ctx->device));
	if (m->needs_inst_mats)
		mesh_instance_mats_kernel<<<(m->instance_count + 127) / 128, 128, 0, m->ctx->stream>>>(m->models, m->instance_count, m->inst_mats);
	if (m->multi_instance_tiles) {
		const size_t dyn = (size_t(m->smem_window) + 1) * 4 + size_t(m->smem_window) * kSmemMatFloats * 4;
		if (m->uniform_vpi) {
			ZN_CUDA(cudaFuncSetAttribute(mesh_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn)));
			mesh_tile_kernel<true, true><<<m->tile_count, kMeshThreads, dyn, m->ctx->stream>>>(m->tiles, m->models, m->instance_first_vertex, m->inst_mats,
			                                                                                  m->pos_in, m->nrm_in, m->pos_out, m->nrm_out, m->smem_window,
			                                                                                  m->uniform_vpi, m->vertex_count);
		} else {
			ZN_CUDA(cudaFuncSetAttribute(mesh_tile_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn)));
			mesh_tile_kernel<true, false><<<m->tile_count, kMeshThreads, dyn, m->ctx->stream>>>(m->tiles, m->models, m->instance_first_vertex, m->inst_mats,
			                                                                                   m->pos_in, m->nrm_in, m->pos_out, m->nrm_out, m->smem_window,
			                                                                                   0u, m->vertex_count);
		}
	} else {
		mesh_tile_kernel<false, false><<<m->tile_count, kMeshThreads, 4, m->ctx->stream>>>(m->tiles, m->models, m->instance_first_vertex, m->inst_mats,
		                                                                           m->pos_in, m->nrm_in, m->pos_out, m->nrm_out, 0u, 0u, m->vertex_count);
	}
	ZN_CUDA(cudaGetLastError());
	return ZN_OK;
}

static int read_packed(zn_mesh* m, const float* dev, uint32_t first, uint32_t count, float* out) {
	if (!out || count == 0) return ZN_OK;
	ZN_CUDA(cudaMemcpyAsync(out, dev + size_t(first) * 3, size_t(count) * 12, cudaMemcpyDeviceToHost, m->ctx->stream));
	return ZN_OK;
}

int zn_mesh_read(zn_mesh* m, uint32_t first, uint32_t count, float* out_position, float* out_normal) {
	ZN_REQUIRE(m, "zn_mesh_read: mesh is NULL");
	ZN_REQUIRE(uint64_t(first) + count <= m->vertex_count, "zn_mesh_read: range exceeds vertex_count %u", m->vertex_count);
	ZN_CUDA(cudaSetDevice(m->ctx->device));
	int rc = read_packed(m, m->pos_out, first, count, out_position);
	if (rc == ZN_OK) rc = read_packed(m, m->nrm_out, first, count, out_normal);
	if (rc != ZN_OK) return rc;
	ZN_CUDA(cudaStreamSynchronize(m->ctx->stream));
	return ZN_OK;
}

int zn_mesh_read_inputs(zn_mesh* m, uint32_t first, uint32_t count, float* position, float* normal) {
	ZN_REQUIRE(m, "zn_mesh_read_inputs: mesh is NULL");
	ZN_REQUIRE(uint64_t(first) + count <= m->vertex_count, "zn_mesh_read_inputs: range exceeds vertex_count %u", m->vertex_count);
	ZN_CUDA(cudaSetDevice(m->ctx->device));
	int rc = read_packed(m, m->pos_in, first, count, position);
	if (rc == ZN_OK) rc = read_packed(m, m->nrm_in, first, count, normal);
	if (rc != ZN_OK) return rc;
	ZN_CUDA(cudaStreamSynchronize(m->ctx->stream));
	return ZN_OK;
}

int zn_mesh_device_buffers(zn_mesh* m, zn_mesh_buffers* out) {
	ZN_REQUIRE(m && out, "zn_mesh_device_buffers: NULL argument");
	out->position_in = m->pos_in; out->normal_in = m->nrm_in;
	out->position_out = m->pos_out; out->normal_out = m->nrm_out;
	out->vertex_count = m->vertex_count;
	return ZN_OK;
}

} // extern "C"
