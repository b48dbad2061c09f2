// Test infrastructure only.  The three members KinectCapture::filterFlyingPixels (src/LiveScanClient/kinectCapture.cpp:132-174)
// touches, declared as ICapture (include/LiveScanClient/iCapture.h:49,58) declares them, so that those 43 lines of the
// reference compile on their own: kinectCapture.h itself needs the closed-source Kinect SDK's Kinect.h.
#pragma once
#include <cstdlib>
#include <vector>
typedef unsigned short UINT16;
class KinectCapture {
public:
	int nDepthFrameHeight, nDepthFrameWidth;
	UINT16 *pDepth;
	void filterFlyingPixels(int neighbourhoodSize, float thr, int maxNonFittingNeighbours);
};
