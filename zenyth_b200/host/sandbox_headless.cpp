// sandbox_headless.cpp — the reference's Sandbox frame loop, headless, over the GPU scene path.
//
// The reference's SandboxApp has empty OnUpdate / OnRender hooks and no scene
// (Sandbox/src/SandboxApp.cpp:7-13); its loop is Application::Run (Core/src/Application.cpp:41-52):
//     dt = Tick(); OnUpdate(dt); BeginFrame(); OnRender(); EndFrame();
// This program runs that loop shape on the stand-in scene of SURVEY.md §8(d) config 1 — a 5-level
// tree, fan-out 8, one root = 4,681 entities, camera at the Sandbox's 1280x720 viewport — with
// OnUpdate spinning the root (yaw += dt) and the SceneRenderer seam doing the GPU frame.
//
//   sandbox_headless [frames]          run on GPU 0, print ms/frame and the visible count, then check that the
//                                      scene survives SaveFile -> LoadFile -> FrameHost with the same visible list
//   sandbox_headless --expect-no-gpu   exit 0 iff creating a context throws std::runtime_error
//
// Build: g++ -std=c++17 -Iinclude zenyth_b200/host/sandbox_headless.cpp -Lzenyth_b200/lib -lzenyth_b200
#include "zenyth_b200.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

struct Lcg {
	uint32_t s;
	float Next() { s = s * 1664525u + 1013904223u; return float(s >> 8) * (1.0f / 16777216.0f); }
	float Range(float lo, float hi) { return lo + Next() * (hi - lo); }
};

class HeadlessSandbox {
public:
	HeadlessSandbox() : m_ctx(0), m_scene(m_ctx, MakeDesc()), m_camera(MakeCamera()), m_renderer(m_scene, m_camera) {
		const uint32_t n = m_scene.EntityCount();
		std::vector<Zenyth::Transform> t(n);
		std::vector<Zenyth::Aabb> b(n);
		std::vector<uint32_t> parent(n, ZN_NO_PARENT);
		Lcg rng{0x5A454E59u};
		uint32_t levelBegin = 0, prevBegin = 0;
		for (uint32_t l = 0, size = 1; l < 5; ++l, size *= 8) {
			for (uint32_t k = 0; k < size; ++k) {
				const uint32_t i = levelBegin + k;
				const float range = l == 0 ? 0.0f : 10.0f;
				t[i].position = {rng.Range(-range, range), rng.Range(-range, range), rng.Range(-range, range)};
				t[i].rotation = {rng.Range(-3.1f, 3.1f), rng.Range(-3.1f, 3.1f), rng.Range(-3.1f, 3.1f)};
				t[i].scale = {rng.Range(0.9f, 1.1f), rng.Range(0.9f, 1.1f), rng.Range(0.9f, 1.1f)};
				b[i].halfExtent = {rng.Range(0.1f, 1.0f), rng.Range(0.1f, 1.0f), rng.Range(0.1f, 1.0f)};
				if (l > 0) parent[i] = prevBegin + k / 8;
			}
			prevBegin = levelBegin;
			levelBegin += size;
		}
		m_root = t[0];
		m_scene.SetTransforms(0, t);
		m_scene.SetBounds(0, b);
		m_scene.SetParents(0, parent);
		m_renderer.Init(nullptr, 1280, 720);   // Sandbox/include/SandboxApp.hpp:6
	}

	void OnUpdate(float dt) {
		m_root.rotation.y += dt;
		m_scene.SetTransform(0, m_root);
	}
	void OnRender() {}

	double Run(int frames) {
		const auto t0 = std::chrono::steady_clock::now();
		for (int f = 0; f < frames; ++f) {
			OnUpdate(1.0f / 60.0f);
			m_renderer.BeginFrame();
			OnRender();
			m_renderer.EndFrame();
		}
		const auto t1 = std::chrono::steady_clock::now();
		return std::chrono::duration<double, std::milli>(t1 - t0).count() / frames;
	}
	uint32_t VisibleCount() const { return m_renderer.VisibleCount(); }

	// Scene ingest through the C++ layer: dump the scene, load it into a second Scene and run one
	// all-in-one host frame on it (no uploads: the file carried the inputs); the visible list must be
	// the one the live scene just produced.
	bool FileRoundTrip(const std::string& path) {
		m_scene.SaveFile(path);
		Zenyth::Scene copy = Zenyth::Scene::LoadFile(m_ctx, path);
		if (copy.EntityCount() != m_scene.EntityCount()) return false;
		copy.SetCamera(m_camera);
		std::vector<uint32_t> visible(copy.EntityCount());
		const uint32_t n = copy.FrameHost(nullptr, nullptr, nullptr, nullptr, nullptr, visible.data());
		visible.resize(n);
		return visible == m_scene.Visible();
	}
	uint32_t EntityCount() const { return m_scene.EntityCount(); }

private:
	static Zenyth::SceneDesc MakeDesc() {
		Zenyth::SceneDesc d;
		d.levelOffsets = {0, 1, 9, 73, 585, 4681};
		d.entityCount = 4681;
		return d;
	}
	static Zenyth::Camera MakeCamera() {
		Zenyth::CameraDesc c;
		c.eye = {0.0f, 0.0f, 60.0f};
		c.center = {0.0f, 0.0f, 0.0f};
		return Zenyth::Camera(c);
	}
	Zenyth::GpuContext m_ctx;
	Zenyth::Scene m_scene;
	Zenyth::Camera m_camera;
	Zenyth::SceneRenderer m_renderer;
	Zenyth::Transform m_root;
};

} // namespace

int main(int argc, char** argv) {
	if (argc > 1 && std::strcmp(argv[1], "--expect-no-gpu") == 0) {
		try {
			Zenyth::GpuContext ctx(0);
		} catch (const std::runtime_error& e) {
			std::printf("no GPU, as expected: %s\n", e.what());
			return 0;
		}
		std::printf("a context was created: a GPU is present\n");
		return 1;
	}
	const int frames = argc > 1 ? std::atoi(argv[1]) : 1000;
	try {
		HeadlessSandbox app;
		app.Run(10);
		const double ms = app.Run(frames);
		const char* tmp = std::getenv("TMPDIR");
		const std::string path = std::string(tmp ? tmp : "/tmp") + "/sandbox_headless.znscene";
		const bool roundTrip = app.FileRoundTrip(path);
		std::remove(path.c_str());
		std::printf("{\"entities\": %u, \"frames\": %d, \"ms_per_frame\": %.4f, \"visible\": %u, \"file_round_trip\": %s}\n", app.EntityCount(),
		            frames, ms, app.VisibleCount(), roundTrip ? "true" : "false");
	} catch (const std::runtime_error& e) {
		std::fprintf(stderr, "sandbox_headless: %s\n", e.what());
		return 2;
	}
	return 0;
}
