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
#include "zn_internal.hpp"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <memory>

namespace {

constexpr char kMagic[8] = {'Z', 'N', 'S', 'C', 'E', 'N', 'E', '1'};
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

} // extern "C"
