// Stand-in for the header CMake generates from cmake/in/build_settings.h.in (reference CMakeLists.txt:14-19):
// single precision (EDYN_DOUBLE_PRECISION undefined), profiling disabled, asserts off.
#ifndef EDYN_BUILD_SETTINGS_H
#define EDYN_BUILD_SETTINGS_H
#define EDYN_DISABLE_PROFILING
#define EDYN_DISABLE_ASSERT
#endif
