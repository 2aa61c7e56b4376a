// TEST INFRASTRUCTURE ONLY (oracle build).
// Minimal stand-in for the OpenCV types that /root/reference/src/NeRFRenderer.h:46-68 names in two image
// helper functions that the ray-batch hot path never calls.  It exists only so that the reference's renderer
// template can be instantiated by oracle/ref_bindings.cpp in a container without C++ OpenCV.
#pragma once
#include <cstring>
#include <vector>

#define CV_8U 0
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))

namespace cv {
struct Mat {
	unsigned char* data = nullptr;
	int rows = 0, cols = 0, type_ = 0;
	std::vector<unsigned char> own;
	Mat() {}
	Mat(int r, int c, int type, void* p) : data(static_cast<unsigned char*>(p)), rows(r), cols(c), type_(type) {}
	int channels() const { return (type_ >> 3) + 1; }
	bool isContinuous() const { return true; }
	Mat clone() const { Mat m; copyTo(m); return m; }
	void copyTo(Mat& dst) const
	{
		dst.rows = rows; dst.cols = cols; dst.type_ = type_;
		dst.own.assign(data, data + static_cast<size_t>(rows) * cols * channels());
		dst.data = dst.own.data();
	}
};
}
