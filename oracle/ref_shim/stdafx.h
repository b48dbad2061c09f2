// Build shim for compiling the reference on Linux (test infrastructure only).
// The reference's utils.h:17 includes the Win32 precompiled header; all it needs from it is BYTE.
#pragma once
typedef unsigned char BYTE;
