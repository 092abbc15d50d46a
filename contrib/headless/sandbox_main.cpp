// sandbox_main.cpp — SURVEY.md §8(f) N1: the reference's OWN frame loop driving the GPU scene path.
//
// Linked from: the reference's Core/src/Application.cpp and Core/src/Timer.cpp (compiled in place,
// unmodified), contrib/headless/HeadlessWindow.cpp (a windowless Zenyth::Window), and
// libzenyth_b200.so.  SandboxApp derives from the reference's Zenyth::Application
// (Core/include/Application.hpp:16-51) and fills the hooks the reference's Sandbox leaves empty
// (Sandbox/src/SandboxApp.cpp:3-22); GpuSceneRenderer implements the reference's IRenderer seam
// (Core/include/IRenderer.hpp:5-12) and is installed with Application::SetRenderer exactly as
// Sandbox/src/main.cpp:21-22 installs D3D12Renderer.  The scene is the config-1 stand-in of
// a 5-level tree, fan-out 8, 4,681 entities, 1280x720 viewport.
//
// The path is called where SURVEY.md §8(b) says it would be: Scene::UpdateWorld() from OnUpdate(dt)
// (Core/src/Application.cpp:47) and Scene::Cull(camera) from OnRender(), between the renderer's
// BeginFrame and EndFrame (Application.cpp:49-51).  At shutdown the last frame is repeated through
// the fused Scene::Update() with the same camera and must give the same visible list.
#include "pch.hpp"
#include "Application.hpp"

#include "zenyth_b200.hpp"

#include <cstdio>

namespace {

struct Lcg {
	uint32_t s;
	float Next() { s = s * 1664525u + 1013904223u; return float(s >> 8) * (1.0f / 16777216.0f); }
	float Range(float lo, float hi) { return lo + Next() * (hi - lo); }
};

struct SharedState {
	Zenyth::GpuContext gpu{0};
	std::unique_ptr<Zenyth::Scene> scene;
	Zenyth::Camera camera;
	uint32_t visible = 0;
	uint32_t frames = 0;
	uint32_t resizes = 0;
};

// The reference's plugin seam: Init once, BeginFrame/EndFrame around OnRender, Resize from events.
class GpuSceneRenderer : public Zenyth::IRenderer {
public:
	explicit GpuSceneRenderer(SharedState& st) : m_st(st) {}
	void Init(HWND, uint32_t width, uint32_t height) override { m_st.camera.Resize(width, height); }
	void BeginFrame() override {}            // nothing to prepare: the world matrices were composed in OnUpdate
	void EndFrame() override {
		m_st.visible = m_st.scene->VisibleCount();
		++m_st.frames;
	}
	void Resize(uint32_t width, uint32_t height) override {
		m_st.camera.Resize(width, height);
		++m_st.resizes;
	}

private:
	SharedState& m_st;
};

class SandboxApp : public Zenyth::Application {
public:
	explicit SandboxApp(SharedState& st) : Application({.title = L"Sandbox", .width = 1280, .height = 720}), m_st(st) {}

protected:
	void OnInit() override {
		Zenyth::SceneDesc desc;
		desc.levelOffsets = {0, 1, 9, 73, 585, 4681};
		desc.entityCount = 4681;
		m_st.scene = std::make_unique<Zenyth::Scene>(m_st.gpu, desc);
		std::vector<Zenyth::Transform> t(desc.entityCount);
		std::vector<Zenyth::Aabb> b(desc.entityCount);
		std::vector<uint32_t> parent(desc.entityCount, ZN_NO_PARENT);
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
		m_st.scene->SetTransforms(0, t);
		m_st.scene->SetBounds(0, b);
		m_st.scene->SetParents(0, parent);
		Zenyth::CameraDesc c;
		c.eye = {0.0f, 0.0f, 60.0f};
		c.center = {0.0f, 0.0f, 0.0f};
		c.aspect = m_st.camera.Desc().aspect;   // set by IRenderer::Init from the window size
		m_st.camera = Zenyth::Camera(c);
	}
	void OnUpdate(float dt) override {
		m_root.rotation.y += dt;                // yaw += dt
		m_st.scene->SetTransform(0, m_root);
		m_st.scene->UpdateWorld();              // Scene::Update(dt): TRS -> world over the hierarchy
	}
	void OnRender() override {
		m_st.scene->Cull(m_st.camera);          // Scene::Cull(camera): MVP + frustum test + ordered visible list
	}
	void OnShutdown() override {
		m_totalTime = GetTimer().TotalTime();
		// the same frame through the fused call: identical list
		const std::vector<uint32_t> split = m_st.scene->Visible();
		m_st.scene->SetCamera(m_st.camera);
		m_st.scene->Update();
		m_splitEqualsFused = split == m_st.scene->Visible();
	}

public:
	double m_totalTime = 0.0;
	bool m_splitEqualsFused = false;

private:
	SharedState& m_st;
	Zenyth::Transform m_root;
};

} // namespace

int main() {
	try {
		SharedState st;
		SandboxApp app(st);
		app.SetRenderer(std::make_unique<GpuSceneRenderer>(st));   // Sandbox/src/main.cpp:21-22
		app.Run();                                                  // Core/src/Application.cpp:26-57
		std::printf("{\"entities\": 4681, \"frames\": %u, \"resizes\": %u, \"visible\": %u, \"ms_per_frame\": %.4f, \"aspect\": %.4f, "
		            "\"split_equals_fused\": %s}\n",
		            st.frames, st.resizes, st.visible, st.frames ? app.m_totalTime * 1e3 / st.frames : 0.0, st.camera.Desc().aspect,
		            app.m_splitEqualsFused ? "true" : "false");
	} catch (const std::runtime_error& e) {
		std::fprintf(stderr, "sandbox_reference_loop: %s\n", e.what());
		return 2;
	}
	return 0;
}
