// ORACLE BUILD SHIM — test infrastructure, not product code.
//
// Shadows /root/reference/Core/include/pch.hpp (which pulls in <windows.h>, <wrl.h>, ...;
// Core/include/pch.hpp:8-9,31-32) so that the reference's own math translation units
// (Core/src/math/vector.cpp, Core/src/math/matrix.cpp) compile on Linux, in place, without
// copying or editing them.  This directory must precede Core/include on the -I line.
//
// The math headers silently rely on the PCH for these includes (SURVEY.md §2 "source hazards"):
//   vector.hpp:67,71 uses __m128; constants.hpp:5 uses std::numeric_limits; utils.hpp:7 uses
//   std::min/max; matrix.hpp:15 uses uint32_t; vector.cpp:37,90,127 call std::sqrtf, which
//   libstdc++ 13 does not declare.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <immintrin.h>
namespace std { using ::sqrtf; }
