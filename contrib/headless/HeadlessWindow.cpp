// HeadlessWindow.cpp — a windowless implementation of the reference's Zenyth::Window
// (declared in Core/include/Window.hpp:43-80; the Win32 one is Core/src/Window.cpp).  It creates no
// native window: PumpMessages() reports "still open" for a fixed number of frames
// (ZENYTH_HEADLESS_FRAMES, default 120), emits one WindowResize event on the way — so
// Application::OnEvent's resize path (Core/src/Application.cpp:70-73) runs — and then a
// WindowClose, which makes Application::Run leave its loop exactly as a user closing the window
// would (Application.cpp:66-68).
#include "pch.hpp"
#include "Window.hpp"

#include <cstdlib>

namespace Zenyth {

uint32_t Window::s_windowCount = 0;

namespace {
int g_frames_left = 120;
int g_pumped = 0;
}

Window::Window(const WindowDesc& desc) : m_width(desc.width), m_height(desc.height), m_resizable(desc.resizable) {
	if (const char* env = std::getenv("ZENYTH_HEADLESS_FRAMES")) g_frames_left = std::max(1, std::atoi(env));
	g_pumped = 0;
	++s_windowCount;
}

Window::~Window() { --s_windowCount; }

bool Window::PumpMessages() const {
	++g_pumped;
	if (g_pumped == 10 && m_resizable) {
		Event e{};
		e.type = EventType::WindowResize;
		e.resize = {1600u, 800u};   // aspect 2.0: distinguishable from the 1280x720 start
		EmitEvent(e);
	}
	if (g_pumped > g_frames_left) {
		Event e{};
		e.type = EventType::WindowClose;
		EmitEvent(e);
	}
	return true;
}

void Window::EmitEvent(const Event& e) const {
	if (m_callback) m_callback(e);
}

void Window::RegisterWindowClass() const {}
void Window::CreateNativeWindow(const WindowDesc&) {}
LRESULT Window::WndProcStatic(HWND, UINT, WPARAM, LPARAM) { return 0; }
LRESULT Window::WndProc(HWND, UINT, WPARAM, LPARAM) { return 0; }

} // namespace Zenyth
