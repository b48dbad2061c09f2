// Build shim (test infrastructure only): the reference's simpleimage.h:4 includes <windows.h>
// for BITMAPINFO; nothing on the hot path uses it.
#pragma once
struct BITMAPINFO { int unused; };
