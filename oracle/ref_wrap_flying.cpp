// Flat C entry point over the reference's KinectCapture::filterFlyingPixels (test infrastructure only).  The member
// function's body is the reference's own text, src/LiveScanClient/kinectCapture.cpp:132-174, extracted in-stream by
// oracle/Makefile (sed -n) and compiled behind ref_shim/kinect_capture_stub.h; this file only calls it.
#include "kinect_capture_stub.h"

extern "C" void ref_filter_flying_pixels(unsigned short *depth, int w, int h, int k, float thr, int max_non_fitting)
{
	KinectCapture c;
	c.nDepthFrameWidth = w;
	c.nDepthFrameHeight = h;
	c.pDepth = depth;
	c.filterFlyingPixels(k, thr, max_non_fitting);
}
