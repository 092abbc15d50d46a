// zenyth_b200.hpp — engine-idiom C++ host layer over the C ABI (include/zenyth_b200.h).
//
// The upstream engine declares no Transform / Scene / Camera / Mesh; these types are
// what its frame loop would call, written in its own idiom: namespace Zenyth and PascalCase
// methods for engine objects (Core/include/Application.hpp:7-53), `Desc` aggregates with defaults
// (AppDesc, Application.hpp:9-14; WindowDesc, Window.hpp:36-41), `m_` members, [[nodiscard]],
// deleted copies on owning objects (Application.hpp:21-24), std::runtime_error on failure
// (Core/src/Application.cpp:28-30).  Maths types stay the reference's: anything with the layout
// of zenyth::math::vec3 (16 bytes, Core/include/math/vector.hpp:69-72) or mat4 (64 bytes,
// column-major, Core/include/math/matrix.hpp:66-69) passes through uncopied.
//
// Call sites in the reference's loop (Core/src/Application.cpp:41-52):
//     OnUpdate(dt)  -> scene.SetTransforms(...dirty range...); scene.Update();
//     OnRender()    -> scene.VisibleCount() / DeviceBuffers() feed the draw; mesh.Transform();
// Header-only; link against libzenyth_b200.so.
#pragma once

#include "zenyth_b200.h"

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace Zenyth {

namespace detail {
inline void Check(int code, const char* what) {
	if (code != ZN_OK) throw std::runtime_error(std::string(what) + " : " + zn_last_error());
}
} // namespace detail

// Plain 3-float and 16-float carriers for callers that do not link the reference's math library.
struct Float3 {
	float x = 0.0f, y = 0.0f, z = 0.0f;
};
struct Mat4 {
	float m[16] = {};   // column-major, element (r,c) at m[c*4+r]
	[[nodiscard]] const float* data() const noexcept { return m; }
	[[nodiscard]] float* data() noexcept { return m; }
};

// Local TRS of one entity.  rotation = (pitch, yaw, roll) radians, as mat4::from_euler
// (Core/include/math/matrix.hpp:24); local = T * Ry(yaw)*Rx(pitch)*Rz(roll) * S.
struct Transform {
	Float3 position{0.0f, 0.0f, 0.0f};
	Float3 rotation{0.0f, 0.0f, 0.0f};
	Float3 scale{1.0f, 1.0f, 1.0f};
};

struct Aabb {
	Float3 center{};
	Float3 halfExtent{};
};

struct CameraDesc {
	Float3 eye{0.0f, 0.0f, 0.0f};
	Float3 center{0.0f, 0.0f, -1.0f};
	Float3 up{0.0f, 1.0f, 0.0f};
	float fovYRad = 1.04719755f;          // 60 degrees
	float aspect = 1280.0f / 720.0f;      // Sandbox/include/SandboxApp.hpp:6
	float nearZ = 0.1f;
	float farZ = 5000.0f;
};

// View / projection pair.  Completes what mat4::look_at / mat4::perspective
// (matrix.hpp:52-53, stubbed at matrix.cpp:107-113) would have produced.
class Camera {
public:
	explicit Camera(const CameraDesc& desc = {}) : m_desc(desc) { Rebuild(); }

	[[nodiscard]] const Mat4& View() const noexcept { return m_view; }
	[[nodiscard]] const Mat4& Proj() const noexcept { return m_proj; }
	[[nodiscard]] Mat4 ViewProj() const noexcept {
		Mat4 vp;
		zn_mat4_mul(m_proj.m, m_view.m, vp.m);
		return vp;
	}
	[[nodiscard]] const CameraDesc& Desc() const noexcept { return m_desc; }

	void LookAt(const Float3& eye, const Float3& center, const Float3& up) noexcept {
		m_desc.eye = eye; m_desc.center = center; m_desc.up = up;
		Rebuild();
	}
	// What IRenderer::Resize (Core/include/IRenderer.hpp:11) feeds.
	void Resize(uint32_t width, uint32_t height) noexcept {
		if (width == 0 || height == 0) return;   // same guard as Application.cpp:70
		m_desc.aspect = static_cast<float>(width) / static_cast<float>(height);
		Rebuild();
	}

private:
	void Rebuild() noexcept {
		zn_mat4_look_at(&m_desc.eye.x, &m_desc.center.x, &m_desc.up.x, m_view.m);
		zn_mat4_perspective(m_desc.fovYRad, m_desc.aspect, m_desc.nearZ, m_desc.farZ, m_proj.m);
	}
	CameraDesc m_desc;
	Mat4 m_view, m_proj;
};

// One per GPU.  Move-only, like every owning object in the reference.
class GpuContext {
public:
	explicit GpuContext(int deviceOrdinal = 0) { detail::Check(zn_ctx_create(deviceOrdinal, &m_ctx), "GpuContext::GpuContext"); }
	~GpuContext() { if (m_ctx) zn_ctx_destroy(m_ctx); }
	GpuContext(const GpuContext&) = delete;
	GpuContext& operator=(const GpuContext&) = delete;
	GpuContext(GpuContext&& o) noexcept : m_ctx(std::exchange(o.m_ctx, nullptr)) {}
	GpuContext& operator=(GpuContext&& o) noexcept {
		if (this != &o) { if (m_ctx) zn_ctx_destroy(m_ctx); m_ctx = std::exchange(o.m_ctx, nullptr); }
		return *this;
	}

	void SetStream(void* cudaStream) { detail::Check(zn_ctx_set_stream(m_ctx, cudaStream), "GpuContext::SetStream"); }
	void Sync() { detail::Check(zn_ctx_sync(m_ctx), "GpuContext::Sync"); }
	// Synchronising device <-> host copies for hosts that do not link the CUDA runtime themselves.
	void Read(const void* device, void* host, size_t bytes) { detail::Check(zn_device_read(m_ctx, device, host, bytes), "GpuContext::Read"); }
	void Write(void* device, const void* host, size_t bytes) { detail::Check(zn_device_write(m_ctx, device, host, bytes), "GpuContext::Write"); }
	[[nodiscard]] zn_ctx* Handle() const noexcept { return m_ctx; }

private:
	zn_ctx* m_ctx = nullptr;
};

// Entities in creation order, each naming its parent (ZN_NO_PARENT for roots), to the level-ordered,
// sorted-by-parent layout Scene requires.  order[new] = original index; per-entity arrays follow it.
struct HierarchyLayout {
	std::vector<uint32_t> order;          // [count] new index -> original index
	std::vector<uint32_t> parent;         // [count] parents as NEW indices
	std::vector<uint32_t> levelOffsets;   // [levels + 1], what SceneDesc::levelOffsets takes
};
[[nodiscard]] inline HierarchyLayout SortHierarchy(const std::vector<uint32_t>& parent) {
	HierarchyLayout out;
	const uint32_t n = static_cast<uint32_t>(parent.size());
	out.order.resize(n);
	out.parent.resize(n);
	out.levelOffsets.resize(66);
	uint32_t levels = 0;
	int rc = zn_hierarchy_sort(parent.data(), n, out.order.data(), out.parent.data(), out.levelOffsets.data(),
	                           static_cast<uint32_t>(out.levelOffsets.size()), &levels);
	if (rc != ZN_OK && levels + 1 > out.levelOffsets.size()) {   // deeper than guessed: retry with the count that came back
		out.levelOffsets.resize(levels + 1);
		rc = zn_hierarchy_sort(parent.data(), n, out.order.data(), out.parent.data(), out.levelOffsets.data(),
		                       static_cast<uint32_t>(out.levelOffsets.size()), &levels);
	}
	detail::Check(rc, "SortHierarchy");
	out.levelOffsets.resize(levels + 1);
	return out;
}

struct SceneDesc {
	uint32_t entityCount = 0;
	std::vector<uint32_t> levelOffsets;   // empty = one level of roots
};

// Level-ordered entity hierarchy resident on one GPU.
class Scene {
public:
	Scene(GpuContext& ctx, const SceneDesc& desc) : m_entityCount(desc.entityCount) {
		zn_scene_desc d{};
		d.entity_count = desc.entityCount;
		d.level_count = desc.levelOffsets.empty() ? 0u : static_cast<uint32_t>(desc.levelOffsets.size() - 1);
		d.level_offsets = desc.levelOffsets.empty() ? nullptr : desc.levelOffsets.data();
		detail::Check(zn_scene_create(ctx.Handle(), &d, &m_scene), "Scene::Scene");
	}
	~Scene() { if (m_scene) zn_scene_destroy(m_scene); }
	Scene(const Scene&) = delete;
	Scene& operator=(const Scene&) = delete;
	Scene(Scene&& o) noexcept : m_scene(std::exchange(o.m_scene, nullptr)), m_entityCount(o.m_entityCount) {}

	// Scene ingest: the ZNSCENE1 dump described in zenyth_b200.h (the reference has no on-disk format).
	[[nodiscard]] static Scene LoadFile(GpuContext& ctx, const std::string& path) {
		zn_scene* h = nullptr;
		detail::Check(zn_scene_load(ctx.Handle(), path.c_str(), &h), "Scene::LoadFile");
		Scene s(h);
		s.m_entityCount = s.DeviceBuffers().entity_count;
		return s;
	}
	void SaveFile(const std::string& path) { detail::Check(zn_scene_save(m_scene, path.c_str()), "Scene::SaveFile"); }

	[[nodiscard]] uint32_t EntityCount() const noexcept { return m_entityCount; }

	// Arrays of `count` records `strideBytes` apart: 12 for packed floats, 16 for arrays of
	// zenyth::math::vec3, sizeof(Transform) for arrays of Transform (pass &t[0].position.x etc.).
	void SetTransforms(uint32_t first, uint32_t count, const float* position, const float* rotation, const float* scale, size_t strideBytes) {
		detail::Check(zn_scene_set_transforms(m_scene, first, count, position, rotation, scale, strideBytes), "Scene::SetTransforms");
	}
	void SetTransforms(uint32_t first, const std::vector<Transform>& t) {
		if (t.empty()) return;
		SetTransforms(first, static_cast<uint32_t>(t.size()), &t[0].position.x, &t[0].rotation.x, &t[0].scale.x, sizeof(Transform));
	}
	void SetTransform(uint32_t index, const Transform& t) {
		SetTransforms(index, 1, &t.position.x, &t.rotation.x, &t.scale.x, sizeof(Transform));
	}
	void SetBounds(uint32_t first, uint32_t count, const float* center, const float* halfExtent, size_t strideBytes) {
		detail::Check(zn_scene_set_bounds(m_scene, first, count, center, halfExtent, strideBytes), "Scene::SetBounds");
	}
	void SetBounds(uint32_t first, const std::vector<Aabb>& b) {
		if (b.empty()) return;
		SetBounds(first, static_cast<uint32_t>(b.size()), &b[0].center.x, &b[0].halfExtent.x, sizeof(Aabb));
	}
	void SetParents(uint32_t first, const std::vector<uint32_t>& parent) {
		detail::Check(zn_scene_set_parents(m_scene, first, static_cast<uint32_t>(parent.size()), parent.data()), "Scene::SetParents");
	}
	void SetCamera(const Camera& camera) {
		detail::Check(zn_scene_set_camera(m_scene, camera.View().m, camera.Proj().m), "Scene::SetCamera");
	}
	void SetIndexBase(uint32_t base) { detail::Check(zn_scene_set_index_base(m_scene, base), "Scene::SetIndexBase"); }
	// Record expansion strategy (multi-GPU): a warp per group (default, dense visible sets) or classification + work
	// list (sparse visible sets: a camera inside the scene).
	void SetLevelFusion(bool on) { detail::Check(zn_scene_set_level_fusion(m_scene, on ? 1 : 0), "Scene::SetLevelFusion"); }
	void SetExpandWorkList(bool on) { detail::Check(zn_scene_set_expand_mode(m_scene, on ? ZN_EXPAND_WORKLIST : ZN_EXPAND_GRID), "Scene::SetExpandWorkList"); }

	// One frame, asynchronous: world matrices over the hierarchy, MVP, visible list (the fused form,
	// for the camera of SetCamera; a scene without a camera runs UpdateWorld instead).
	void Update() { detail::Check(zn_scene_update(m_scene), "Scene::Update"); }
	// The two calls the reference's frame loop has room for (Core/src/Application.cpp:47 and :49-51):
	// UpdateWorld() from OnUpdate(dt) composes TRS -> world over the hierarchy; Cull(camera) from
	// OnRender derives MVPs and the ordered visible list for one view from those world matrices.
	// Cull may run for any number of cameras per frame; each result equals a fused Update with that
	// camera bit for bit.  CullInto redirects one pass into caller-owned device memory.
	void UpdateWorld() { detail::Check(zn_scene_update_world(m_scene), "Scene::UpdateWorld"); }
	void Cull(const Camera& camera) { detail::Check(zn_scene_cull(m_scene, camera.View().m, camera.Proj().m, nullptr), "Scene::Cull"); }
	void CullInto(const Camera& camera, const zn_cull_output& out) {
		detail::Check(zn_scene_cull(m_scene, camera.View().m, camera.Proj().m, &out), "Scene::CullInto");
	}
	void ClearCamera() { detail::Check(zn_scene_clear_camera(m_scene), "Scene::ClearCamera"); }

	// The all-in-one frame a CPU engine swaps in: uploads the packed host arrays that are not null,
	// updates, downloads what is asked for; returns the visible count after the data has landed.
	uint32_t FrameHost(const float* position, const float* rotation, const float* scale, float* world, float* mvp, uint32_t* visible) {
		zn_frame_host_io io{position, rotation, scale, world, mvp, visible, 0, 0, 0};
		detail::Check(zn_scene_frame_host(m_scene, &io), "Scene::FrameHost");
		return io.visible_count;
	}
	// The same frame pipelined, up to two in flight: the upload of frame k+1 overlaps the update and
	// the download of frame k.  The arrays must stay untouched until the matching WaitFrameHost.
	// Arrays that are null are not uploaded; the others cover entities [first, first+count) (count 0: to the end).
	void SubmitFrameHost(const float* position, const float* rotation, const float* scale, uint32_t first = 0, uint32_t count = 0) {
		const zn_frame_host_io io{position, rotation, scale, nullptr, nullptr, nullptr, 0, first, count};
		detail::Check(zn_scene_frame_host_submit(m_scene, &io), "Scene::SubmitFrameHost");
	}
	uint32_t WaitFrameHost(uint32_t* visible) {
		zn_frame_host_io io{nullptr, nullptr, nullptr, nullptr, nullptr, visible, 0, 0, 0};
		detail::Check(zn_scene_frame_host_wait(m_scene, &io), "Scene::WaitFrameHost");
		return io.visible_count;
	}

	[[nodiscard]] uint32_t VisibleCount() {
		uint32_t n = 0;
		detail::Check(zn_scene_visible_count(m_scene, &n), "Scene::VisibleCount");
		return n;
	}
	[[nodiscard]] std::vector<uint32_t> Visible() {
		std::vector<uint32_t> out(VisibleCount());
		uint32_t n = 0;
		detail::Check(zn_scene_read_visible(m_scene, out.data(), static_cast<uint32_t>(out.size()), &n), "Scene::Visible");
		out.resize(n);
		return out;
	}
	[[nodiscard]] std::vector<Mat4> World(uint32_t first, uint32_t count) {
		std::vector<Mat4> out(count);
		detail::Check(zn_scene_read_world(m_scene, first, count, count ? out[0].m : nullptr), "Scene::World");
		return out;
	}
	[[nodiscard]] std::vector<Mat4> Mvp(uint32_t first, uint32_t count) {
		std::vector<Mat4> out(count);
		detail::Check(zn_scene_read_mvp(m_scene, first, count, count ? out[0].m : nullptr), "Scene::Mvp");
		return out;
	}
	// Device-resident results for the renderer (no copy).
	[[nodiscard]] zn_scene_buffers DeviceBuffers() {
		zn_scene_buffers b{};
		detail::Check(zn_scene_device_buffers(m_scene, &b), "Scene::DeviceBuffers");
		return b;
	}
	// Visible list -> packed per-instance MVP buffer + one D3D12_DRAW_INDEXED_ARGUMENTS record, both
	// in caller-owned device memory, built on the device (no host round trip).
	void BuildDraw(uint32_t indexCountPerInstance, void* instanceMvpDevice, uint32_t instanceCapacity, void* drawArgsDevice) {
		detail::Check(zn_scene_build_draw(m_scene, indexCountPerInstance, instanceMvpDevice, instanceCapacity, drawArgsDevice), "Scene::BuildDraw");
	}
	[[nodiscard]] zn_scene* Handle() const noexcept { return m_scene; }

private:
	explicit Scene(zn_scene* adopted) noexcept : m_scene(adopted) {}
	zn_scene* m_scene = nullptr;
	uint32_t m_entityCount = 0;
};

struct MeshDesc {
	uint32_t vertexCount = 0;
	uint32_t instanceCount = 0;
	std::vector<uint32_t> instanceFirstVertex;   // empty = equal ranges
};

// A batch of instanced vertices: positions and normals transformed by per-instance model matrices.
class Mesh {
public:
	Mesh(GpuContext& ctx, const MeshDesc& desc) : m_vertexCount(desc.vertexCount) {
		zn_mesh_desc d{};
		d.vertex_count = desc.vertexCount;
		d.instance_count = desc.instanceCount;
		d.instance_first_vertex = desc.instanceFirstVertex.empty() ? nullptr : desc.instanceFirstVertex.data();
		detail::Check(zn_mesh_create(ctx.Handle(), &d, &m_mesh), "Mesh::Mesh");
	}
	~Mesh() { if (m_mesh) zn_mesh_destroy(m_mesh); }
	Mesh(const Mesh&) = delete;
	Mesh& operator=(const Mesh&) = delete;
	Mesh(Mesh&& o) noexcept : m_mesh(std::exchange(o.m_mesh, nullptr)), m_vertexCount(o.m_vertexCount) {}

	// Mesh-batch ingest: the ZNMESH01 dump described in zenyth_b200.h.
	[[nodiscard]] static Mesh LoadFile(GpuContext& ctx, const std::string& path) {
		zn_mesh* h = nullptr;
		detail::Check(zn_mesh_load(ctx.Handle(), path.c_str(), &h), "Mesh::LoadFile");
		Mesh m(h);
		m.m_vertexCount = m.DeviceBuffers().vertex_count;
		return m;
	}
	void SaveFile(const std::string& path) { detail::Check(zn_mesh_save(m_mesh, path.c_str()), "Mesh::SaveFile"); }

	void SetVertices(uint32_t first, uint32_t count, const float* position, const float* normal, size_t strideBytes) {
		detail::Check(zn_mesh_set_vertices(m_mesh, first, count, position, normal, strideBytes), "Mesh::SetVertices");
	}
	void SetModels(uint32_t first, const std::vector<Mat4>& models) {
		if (models.empty()) return;
		detail::Check(zn_mesh_set_models(m_mesh, first, static_cast<uint32_t>(models.size()), models[0].m), "Mesh::SetModels");
	}
	// Use world matrices [firstEntity, firstEntity + instanceCount) of a scene as the models.
	void BindSceneModels(Scene& scene, uint32_t firstEntity) {
		detail::Check(zn_mesh_bind_scene_models(m_mesh, scene.Handle(), firstEntity), "Mesh::BindSceneModels");
	}
	void Transform() { detail::Check(zn_mesh_transform(m_mesh), "Mesh::Transform"); }
	void Read(uint32_t first, uint32_t count, float* outPosition, float* outNormal) {
		detail::Check(zn_mesh_read(m_mesh, first, count, outPosition, outNormal), "Mesh::Read");
	}
	[[nodiscard]] zn_mesh_buffers DeviceBuffers() {
		zn_mesh_buffers b{};
		detail::Check(zn_mesh_device_buffers(m_mesh, &b), "Mesh::DeviceBuffers");
		return b;
	}
	[[nodiscard]] uint32_t VertexCount() const noexcept { return m_vertexCount; }

private:
	explicit Mesh(zn_mesh* adopted) noexcept : m_mesh(adopted) {}
	zn_mesh* m_mesh = nullptr;
	uint32_t m_vertexCount = 0;
};

// ---- multi-GPU host layer (one process per GPU) -----------------------------------
// Independent scenes over ranks: the scene ids rank `rank` owns (contiguous blocks, sizes differing
// by at most one) and every scene's global index base (-> Scene::SetIndexBase).
[[nodiscard]] inline std::vector<uint32_t> ShardScenes(uint32_t sceneCount, uint32_t worldSize, uint32_t rank) {
	uint32_t first = 0, count = 0;
	detail::Check(zn_shard_scenes(sceneCount, worldSize, rank, &first, &count), "ShardScenes");
	std::vector<uint32_t> ids(count);
	for (uint32_t k = 0; k < count; ++k) ids[k] = first + k;
	return ids;
}
[[nodiscard]] inline std::vector<uint32_t> SceneIndexBases(const std::vector<uint32_t>& entityCounts) {
	std::vector<uint32_t> bases(entityCounts.size());
	detail::Check(zn_scene_index_bases(entityCounts.data(), static_cast<uint32_t>(entityCounts.size()), bases.data()), "SceneIndexBases");
	return bases;
}

// ONE scene split into hierarchy-closed entity shards by root subtree.  A shard is a scene of its
// own: levelOffsets / parent (shard-local indices) are what SceneDesc and Scene::SetParents take,
// globalIndex[local] maps a shard-local visible index back to the scene (monotonic, so a shard's
// ascending list stays ascending).  Per-entity arrays follow globalIndex.
struct SceneShard {
	std::vector<uint32_t> globalIndex;    // [n] ascending
	std::vector<uint32_t> parent;         // [n] shard-local, ZN_NO_PARENT for roots
	std::vector<uint32_t> levelOffsets;   // [levels + 1]; levels the shard owns nothing of are dropped
};
[[nodiscard]] inline std::vector<SceneShard> ShardSceneByRoots(const std::vector<uint32_t>& levelOffsets, const std::vector<uint32_t>& parent,
                                                               uint32_t worldSize) {
	if (levelOffsets.size() < 2) throw std::runtime_error("ShardSceneByRoots : levelOffsets needs at least one level");
	const uint32_t levels = static_cast<uint32_t>(levelOffsets.size() - 1), E = levelOffsets.back();
	if (!parent.empty() && parent.size() != E) throw std::runtime_error("ShardSceneByRoots : parent must be empty or hold one entry per entity");
	std::vector<uint32_t> owner(E), local(E), counts(worldSize);
	detail::Check(zn_shard_scene_by_roots(levelOffsets.data(), levels, parent.empty() ? nullptr : parent.data(), worldSize, owner.data(), local.data(),
	                                      counts.data()), "ShardSceneByRoots");
	std::vector<SceneShard> shards(worldSize);
	std::vector<std::vector<uint32_t>> perLevel(worldSize, std::vector<uint32_t>(levels, 0u));
	for (uint32_t r = 0; r < worldSize; ++r) { shards[r].globalIndex.reserve(counts[r]); shards[r].parent.reserve(counts[r]); }
	for (uint32_t l = 0; l < levels; ++l)
		for (uint32_t e = levelOffsets[l]; e < levelOffsets[l + 1]; ++e) {
			SceneShard& s = shards[owner[e]];
			s.globalIndex.push_back(e);
			const uint32_t p = parent.empty() ? ZN_NO_PARENT : parent[e];
			s.parent.push_back(p == ZN_NO_PARENT ? ZN_NO_PARENT : local[p]);
			perLevel[owner[e]][l]++;
		}
	for (uint32_t r = 0; r < worldSize; ++r) {
		shards[r].levelOffsets.push_back(0u);
		for (uint32_t l = 0; l < levels; ++l)
			if (perLevel[r][l]) shards[r].levelOffsets.push_back(shards[r].levelOffsets.back() + perLevel[r][l]);
	}
	return shards;
}

// The per-frame exchange of compacted visible lists, fused into the compute kernels over NVLink
// peer memory (zn_exchange_*): every rank ends up with every rank's list and count; no collective
// library is involved.  The caller supplies one blocking HOST all-gather (MPI_Allgather, a socket,
// ...) used at construction and destruction only — both are therefore collective.
//     per frame k:  exchange.Bind(scenes, k);  scene.Update() for every owned scene;  exchange.Expand(k - 1);
//     at the end:   exchange.Expand(last);  auto lists = exchange.Result();   // counts + device lists of the last expanded frame
// Expanding one frame behind leaves a frame of slack between ranks (the owners' flags are normally up
// when the expansion kernel looks).  With PeerExchangeDesc::overlap the expansions run on a low-priority
// side stream and frame k may be expanded right after its own update (Expand(k)); measured gain ~1%.
struct PeerExchangeDesc {
	uint32_t worldSize = 1;
	uint32_t rank = 0;
	uint32_t recordsPerRank = 1;          // scenes this rank owns (the same on every rank)
	uint32_t timeoutMs = 2000;
	bool overlap = false;                 // ZN_EXCHANGE_OVERLAP: expansions on a side stream
	zn_allgather_fn allGather = nullptr;  // int(void* user, const void* send, void* recv, size_t bytes)
	void* user = nullptr;
	std::vector<uint32_t> indexBases;     // [worldSize * recordsPerRank] added to each record's indices; empty = none
};
class PeerExchange {
public:
	struct Lists {
		std::vector<uint32_t> counts;     // [worldSize * recordsPerRank]
		const uint32_t* device = nullptr; // list r is device[r * stride .. r * stride + counts[r])
		uint32_t stride = 0;
	};
	// `layout`: any scene with the level layout all exchanged scenes share (sizes the records and lists).
	PeerExchange(GpuContext& ctx, Scene& layout, const PeerExchangeDesc& desc) : m_indexBases(desc.indexBases) {
		const zn_scene_buffers b = layout.DeviceBuffers();
		zn_exchange_desc d{};
		d.world_size = desc.worldSize; d.rank = desc.rank; d.records_per_rank = desc.recordsPerRank;
		d.record_words = b.visibility_record_words; d.list_capacity = b.entity_count;
		d.slots = 0; d.timeout_ms = desc.timeoutMs; d.allgather = desc.allGather; d.user = desc.user;
		d.flags = desc.overlap ? ZN_EXCHANGE_OVERLAP : 0u;
		detail::Check(zn_exchange_create(ctx.Handle(), &d, &m_exchange), "PeerExchange::PeerExchange");
		m_total = desc.worldSize * desc.recordsPerRank;
		if (!m_indexBases.empty() && m_indexBases.size() != m_total) {
			zn_exchange_destroy(m_exchange, nullptr, 0);
			throw std::runtime_error("PeerExchange::PeerExchange : indexBases must hold one entry per record");
		}
	}
	~PeerExchange() { if (m_exchange) zn_exchange_destroy(m_exchange, m_bound.data(), static_cast<uint32_t>(m_bound.size())); }
	PeerExchange(const PeerExchange&) = delete;
	PeerExchange& operator=(const PeerExchange&) = delete;

	void Bind(const std::vector<Scene*>& scenes, uint64_t frame) {
		m_bound.clear();
		for (Scene* s : scenes) m_bound.push_back(s->Handle());
		detail::Check(zn_exchange_bind(m_exchange, m_bound.data(), static_cast<uint32_t>(m_bound.size()), frame), "PeerExchange::Bind");
	}
	// Raise this rank's flags for the last bound frame without waiting for anyone, and return when they are
	// up.  Only ranks that SHARE one GPU need it (signal, rendezvous on the host, then Expand), so that no
	// kernel ever spins on a flag another process on the same device has yet to raise.
	void Signal() {
		if (m_bound.empty()) throw std::runtime_error("PeerExchange::Signal : Bind first");
		detail::Check(zn_exchange_signal(m_exchange, m_bound[0]), "PeerExchange::Signal");
	}
	// Make the context stream wait for the last expansion (ordering for a consumer of the lists).
	void Join() { detail::Check(zn_exchange_join(m_exchange), "PeerExchange::Join"); }
	// Derive every PEER's list of `frame` (asynchronous; raises this rank's flags for the last bound frame
	// and waits on the device for the owners' flags of `frame`).
	void Expand(int64_t frame) {
		if (m_bound.empty()) throw std::runtime_error("PeerExchange::Expand : Bind first");
		detail::Check(zn_exchange_expand(m_exchange, m_bound[0], frame, m_indexBases.empty() ? nullptr : m_indexBases.data()), "PeerExchange::Expand");
	}
	// Counts and device lists of the last expanded frame (synchronises); throws if a peer never signalled.
	[[nodiscard]] Lists Result() {
		Lists out;
		out.counts.resize(m_total);
		void* dev = nullptr;
		detail::Check(zn_exchange_result(m_exchange, out.counts.data(), &dev, &out.stride), "PeerExchange::Result");
		out.device = static_cast<const uint32_t*>(dev);
		return out;
	}
	[[nodiscard]] zn_exchange* Handle() const noexcept { return m_exchange; }

private:
	zn_exchange* m_exchange = nullptr;
	uint32_t m_total = 0;
	std::vector<uint32_t> m_indexBases;
	std::vector<zn_scene*> m_bound;
};

// The reference's plugin seam, in its four-method shape (Core/include/IRenderer.hpp:5-12), with a
// portable handle in place of HWND.  A D3D12 (or any) renderer owns one of these beside its swap
// chain: BeginFrame runs the scene update the draw depends on, EndFrame leaves world / MVP /
// visible list on the device for the draw submission.
class SceneRenderer {
public:
	SceneRenderer(Scene& scene, Camera& camera) : m_scene(scene), m_camera(camera) {}
	void Init(void* /*nativeWindowHandle*/, uint32_t width, uint32_t height) { Resize(width, height); }
	void BeginFrame() {
		m_scene.SetCamera(m_camera);
		m_scene.Update();
	}
	// A second view of the same frame (shadow pass, a camera that moved after OnUpdate): re-cull
	// without recomposing the hierarchy.
	void CullView(const Camera& camera) { m_scene.Cull(camera); }
	void EndFrame() { m_visibleCount = m_scene.VisibleCount(); }
	void Resize(uint32_t width, uint32_t height) { m_camera.Resize(width, height); }
	[[nodiscard]] uint32_t VisibleCount() const noexcept { return m_visibleCount; }

private:
	Scene& m_scene;
	Camera& m_camera;
	uint32_t m_visibleCount = 0;
};

} // namespace Zenyth
