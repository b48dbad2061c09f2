// Link stubs (test infrastructure only) for the Win32/debug-IO helpers the reference's
// depthprocessing.cpp references but the hot path never needs: SimpleImage (simpleimage.cpp uses
// Win32 BITMAPINFO/ATL/libjpeg/libpng) and the writePGM debug dumps (pgm.cpp).
#include "simpleimage.h"
#include "pgm.h"
#include <cstdlib>
#include <cstring>

SimpleImage::SimpleImage(const SimpleImage &) : width(-1), height(-1), data_ptr(NULL), bytes_per_pixel(1), bip(NULL) {}
SimpleImage &SimpleImage::operator=(const SimpleImage &) { return *this; }
SimpleImage::~SimpleImage() { if (data_ptr) free(data_ptr); }
void SimpleImage::create(int w, int h, int bpp, unsigned char *data) {
	if (data_ptr) free(data_ptr);
	width = w; height = h; bytes_per_pixel = bpp;
	data_ptr = (unsigned char*)calloc((size_t)w * h * bpp, 1);
	if (data) memcpy(data_ptr, data, (size_t)w * h * bpp);
}
int SimpleImage::writeToFile(const char *, FileType) { return 0; }

bool writePGM(const char *, int, int, unsigned char *) { return true; }
bool writePGM(const char *, int, int, int *) { return true; }
bool writePGM(const char *, int, int, double *) { return true; }
bool writePGM(const char *, int, int, double *, unsigned char *) { return true; }
bool writePGM(const char *, int, int, float *) { return true; }
