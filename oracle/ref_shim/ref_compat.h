// Force-included (-include) when compiling the reference's sources on Linux (test infrastructure only):
// neutralises the MSVC-only keywords the sources use and pulls in the headers MSVC includes implicitly.
#pragma once
#define __declspec(x)
#define __stdcall
#define __int64 long long
#ifdef __cplusplus
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <string>
#include <functional>
#endif
