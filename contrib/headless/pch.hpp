// contrib/headless/pch.hpp — shadows the reference's Core/include/pch.hpp (which needs <windows.h>)
// so that the reference's OWN Application.cpp and Timer.cpp compile on Linux, in place and
// unmodified.  It supplies only the Win32 vocabulary those two files and the
// reference's headers spell: handle and integer typedefs, CALLBACK, and the QueryPerformance*
// pair (Core/src/Timer.cpp:8-44) on top of clock_gettime.  No reference source is copied.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <ctime>
#include <exception>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <immintrin.h>

namespace std { using ::sqrtf; }

using HWND = void*;
using HINSTANCE = void*;
using UINT = unsigned int;
using WPARAM = std::uintptr_t;
using LPARAM = std::intptr_t;
using LRESULT = std::intptr_t;
using BOOL = int;
#define CALLBACK
union LARGE_INTEGER {
	long long QuadPart;
};
inline BOOL QueryPerformanceFrequency(LARGE_INTEGER* f) { f->QuadPart = 1000000000LL; return 1; }
inline BOOL QueryPerformanceCounter(LARGE_INTEGER* c) {
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	c->QuadPart = static_cast<long long>(ts.tv_sec) * 1000000000LL + ts.tv_nsec;
	return 1;
}
