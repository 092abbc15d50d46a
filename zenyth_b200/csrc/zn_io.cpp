// zn_io.cpp — scene ingest: a minimal binary SoA dump, so real content can
// replace the synthetic scenes.  The reference has no on-disk format (and no serialization at all).
//
// Layout (little-endian, no padding):
//   char     magic[8]      = "ZNSCENE1"
//   uint32   entity_count
//   uint32   level_count
//   uint32   level_offsets[level_count + 1]
//   float32  position[E][3], rotation[E][3] (pitch,yaw,roll), scale[E][3],
//            aabb_center[E][3], aabb_half_extent[E][3]
//   uint32   parent[E]      (0xFFFFFFFF = root)
// 64 bytes per entity: the packed form of what zn_scene_set_transforms / set_bounds / set_parents
// take.  Host code only: it drives the C ABI's own upload / readback entry points.
//
// Mesh batches use the same idea ("ZNMESH01"):
//   char     magic[8]      = "ZNMESH01"
//   uint32   vertex_count, instance_count
//   uint32   instance_first_vertex[instance_count + 1]
//   float32  position[V][3], normal[V][3]
//   float32  model[I][16]   (column-major mat4::data() layout; the matrices in use at save time,
//                            i.e. a bound scene's world matrices are written out as plain models)
#include "zn_internal.hpp"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <memory>

namespace {

constexpr char kMagic[8] = {'Z', 'N', 'S', 'C', 'E', 'N', 'E', '1'};
constexpr char kMeshMagic[8] = {'Z', 'N', 'M', 'E', 'S', 'H', '0', '1'};
constexpr uint32_t kChunk = 1u << 22;   // entities per I/O pass (48 MiB of float3)

struct File {
	FILE* f = nullptr;
	~File() { if (f) std::fclose(f); }
};

} // namespace

extern "C" {

int zn_scene_save(zn_scene* s, const char* path) {
	ZN_REQUIRE(s && path, "zn_scene_save: NULL argument");
	File out;
	out.f = std::fopen(path, "wb");
	if (!out.f) return zn::fail(ZN_ERR_INVALID, "zn_scene_save: cannot open %s for writing: %s", path, std::strerror(errno));
	const uint32_t E = s->entity_count, L = uint32_t(s->levels.size());
	std::vector<uint32_t> offs(L + 1);
	for (uint32_t l = 0; l < L; ++l) offs[l] = s->levels[l].begin;
	offs[L] = E;
	bool ok = std::fwrite(kMagic, 1, 8, out.f) == 8 && std::fwrite(&E, 4, 1, out.f) == 1 && std::fwrite(&L, 4, 1, out.f) == 1 &&
	          std::fwrite(offs.data(), 4, offs.size(), out.f) == offs.size();
	std::vector<float> buf(size_t(std::min(E, kChunk)) * 3);
	for (int a = 0; a < 5 && ok; ++a)
		for (uint32_t first = 0; first < E && ok; first += kChunk) {
			const uint32_t n = std::min(kChunk, E - first);
			float* dst[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
			dst[a] = buf.data();
			const int rc = zn_scene_read_inputs(s, first, n, dst[0], dst[1], dst[2], dst[3], dst[4], nullptr);
			if (rc != ZN_OK) return rc;
			ok = std::fwrite(buf.data(), 12, n, out.f) == n;
		}
	std::vector<uint32_t> par(std::min(E, kChunk));
	for (uint32_t first = 0; first < E && ok; first += kChunk) {
		const uint32_t n = std::min(kChunk, E - first);
		const int rc = zn_scene_read_inputs(s, first, n, nullptr, nullptr, nullptr, nullptr, nullptr, par.data());
		if (rc != ZN_OK) return rc;
		ok = std::fwrite(par.data(), 4, n, out.f) == n;
	}
	if (!ok) return zn::fail(ZN_ERR_INVALID, "zn_scene_save: short write to %s", path);
	return ZN_OK;
}

int zn_scene_load(zn_ctx* ctx, const char* path, zn_scene** out_scene) {
	ZN_REQUIRE(ctx && path && out_scene, "zn_scene_load: NULL argument");
	*out_scene = nullptr;
	File in;
	in.f = std::fopen(path, "rb");
	if (!in.f) return zn::fail(ZN_ERR_INVALID, "zn_scene_load: cannot open %s: %s", path, std::strerror(errno));
	char magic[8];
	uint32_t E = 0, L = 0;
	if (std::fread(magic, 1, 8, in.f) != 8 || std::memcmp(magic, kMagic, 8) != 0)
		return zn::fail(ZN_ERR_INVALID, "zn_scene_load: %s is not a ZNSCENE1 file", path);
	if (std::fread(&E, 4, 1, in.f) != 1 || std::fread(&L, 4, 1, in.f) != 1 || L == 0 || L > E)
		return zn::fail(ZN_ERR_INVALID, "zn_scene_load: %s has a bad header", path);
	std::vector<uint32_t> offs(size_t(L) + 1);
	if (std::fread(offs.data(), 4, offs.size(), in.f) != offs.size()) return zn::fail(ZN_ERR_INVALID, "zn_scene_load: %s is truncated", path);
	zn_scene_desc desc;
	desc.entity_count = E;
	desc.level_count = L;
	desc.level_offsets = offs.data();
	zn_scene* s = nullptr;
	int rc = zn_scene_create(ctx, &desc, &s);
	if (rc != ZN_OK) return rc;
	auto bail = [&](int code) { zn_scene_destroy(s); return code; };
	std::vector<float> buf(size_t(std::min(E, kChunk)) * 3);
	for (int a = 0; a < 5; ++a)
		for (uint32_t first = 0; first < E; first += kChunk) {
			const uint32_t n = std::min(kChunk, E - first);
			if (std::fread(buf.data(), 12, n, in.f) != n) return bail(zn::fail(ZN_ERR_INVALID, "zn_scene_load: %s is truncated", path));
			const float* src[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
			src[a] = buf.data();
			rc = a < 3 ? zn_scene_set_transforms(s, first, n, src[0], src[1], src[2], 12) : zn_scene_set_bounds(s, first, n, src[3], src[4], 12);
			if (rc == ZN_OK) rc = zn_ctx_sync(ctx);   // buf is reused by the next pass
			if (rc != ZN_OK) return bail(rc);
		}
	std::vector<uint32_t> par(std::min(E, kChunk));
	for (uint32_t first = 0; first < E; first += kChunk) {
		const uint32_t n = std::min(kChunk, E - first);
		if (std::fread(par.data(), 4, n, in.f) != n) return bail(zn::fail(ZN_ERR_INVALID, "zn_scene_load: %s is truncated", path));
		if ((rc = zn_scene_set_parents(s, first, n, par.data())) != ZN_OK) return bail(rc);
	}
	*out_scene = s;
	return ZN_OK;
}

int zn_mesh_save(zn_mesh* m, const char* path) {
	ZN_REQUIRE(m && path, "zn_mesh_save: NULL argument");
	File out;
	out.f = std::fopen(path, "wb");
	if (!out.f) return zn::fail(ZN_ERR_INVALID, "zn_mesh_save: cannot open %s for writing: %s", path, std::strerror(errno));
	const uint32_t V = m->vertex_count, I = m->instance_count;
	ZN_CUDA(cudaSetDevice(m->ctx->device));
	std::vector<uint32_t> first(size_t(I) + 1);
	ZN_CUDA(cudaMemcpyAsync(first.data(), m->instance_first_vertex, first.size() * 4, cudaMemcpyDeviceToHost, m->ctx->stream));
	ZN_CUDA(cudaStreamSynchronize(m->ctx->stream));
	bool ok = std::fwrite(kMeshMagic, 1, 8, out.f) == 8 && std::fwrite(&V, 4, 1, out.f) == 1 && std::fwrite(&I, 4, 1, out.f) == 1 &&
	          std::fwrite(first.data(), 4, first.size(), out.f) == first.size();
	std::vector<float> buf(size_t(std::min(V, kChunk)) * 3);
	for (int a = 0; a < 2 && ok; ++a)
		for (uint32_t at = 0; at < V && ok; at += kChunk) {
			const uint32_t n = std::min(kChunk, V - at);
			const int rc = zn_mesh_read_inputs(m, at, n, a == 0 ? buf.data() : nullptr, a == 1 ? buf.data() : nullptr);
			if (rc != ZN_OK) return rc;
			ok = std::fwrite(buf.data(), 12, n, out.f) == n;
		}
	std::vector<float> models(size_t(std::min(I, kChunk)) * 16);
	for (uint32_t at = 0; at < I && ok; at += kChunk) {
		const uint32_t n = std::min(kChunk, I - at);
		ZN_CUDA(cudaMemcpyAsync(models.data(), m->models + size_t(at) * 16, size_t(n) * 64, cudaMemcpyDeviceToHost, m->ctx->stream));
		ZN_CUDA(cudaStreamSynchronize(m->ctx->stream));
		ok = std::fwrite(models.data(), 64, n, out.f) == n;
	}
	if (!ok) return zn::fail(ZN_ERR_INVALID, "zn_mesh_save: short write to %s", path);
	return ZN_OK;
}

int zn_mesh_load(zn_ctx* ctx, const char* path, zn_mesh** out_mesh) {
	ZN_REQUIRE(ctx && path && out_mesh, "zn_mesh_load: NULL argument");
	*out_mesh = nullptr;
	File in;
	in.f = std::fopen(path, "rb");
	if (!in.f) return zn::fail(ZN_ERR_INVALID, "zn_mesh_load: cannot open %s: %s", path, std::strerror(errno));
	char magic[8];
	uint32_t V = 0, I = 0;
	if (std::fread(magic, 1, 8, in.f) != 8 || std::memcmp(magic, kMeshMagic, 8) != 0)
		return zn::fail(ZN_ERR_INVALID, "zn_mesh_load: %s is not a ZNMESH01 file", path);
	if (std::fread(&V, 4, 1, in.f) != 1 || std::fread(&I, 4, 1, in.f) != 1 || I == 0 || I > V)
		return zn::fail(ZN_ERR_INVALID, "zn_mesh_load: %s has a bad header", path);
	std::vector<uint32_t> first(size_t(I) + 1);
	if (std::fread(first.data(), 4, first.size(), in.f) != first.size()) return zn::fail(ZN_ERR_INVALID, "zn_mesh_load: %s is truncated", path);
	zn_mesh_desc desc;
	desc.vertex_count = V;
	desc.instance_count = I;
	desc.instance_first_vertex = first.data();
	zn_mesh* m = nullptr;
	int rc = zn_mesh_create(ctx, &desc, &m);   // validates the ranges
	if (rc != ZN_OK) return rc;
	auto bail = [&](int code) { zn_mesh_destroy(m); return code; };
	std::vector<float> buf(size_t(std::min(V, kChunk)) * 3);
	for (int a = 0; a < 2; ++a)
		for (uint32_t at = 0; at < V; at += kChunk) {
			const uint32_t n = std::min(kChunk, V - at);
			if (std::fread(buf.data(), 12, n, in.f) != n) return bail(zn::fail(ZN_ERR_INVALID, "zn_mesh_load: %s is truncated", path));
			// zn_mesh_set_vertices synchronises, so buf may be reused by the next pass
			if ((rc = zn_mesh_set_vertices(m, at, n, a == 0 ? buf.data() : nullptr, a == 1 ? buf.data() : nullptr, 12)) != ZN_OK) return bail(rc);
		}
	std::vector<float> models(size_t(std::min(I, kChunk)) * 16);
	for (uint32_t at = 0; at < I; at += kChunk) {
		const uint32_t n = std::min(kChunk, I - at);
		if (std::fread(models.data(), 64, n, in.f) != n) return bail(zn::fail(ZN_ERR_INVALID, "zn_mesh_load: %s is truncated", path));
		if ((rc = zn_mesh_set_models(m, at, n, models.data())) != ZN_OK) return bail(rc);
	}
	*out_mesh = m;
	return ZN_OK;
}

} // extern "C"
