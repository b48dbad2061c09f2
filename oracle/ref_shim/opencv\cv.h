// Mini "cv" shim (test infrastructure only; never linked into the product).
//
// The reference's icp.cpp:16 does `#include "opencv\cv.h"`; on Linux that resolves to a file whose
// name literally contains a backslash. OpenCV 3.2.0 (include/opencv2/core/version.hpp:53-55 in the
// reference) ships there as headers + Windows import libs only, so it cannot be linked here.
// This header restates just the cv::Mat arithmetic icp.cpp:80-83,138-168 touches, following the
// published OpenCV 3.2 behaviour for CV_32F:
//   * Mat - Mat, Mat += Mat            : element-wise fp32
//   * reduce(.., 0, CV_REDUCE_AVG)      : sequential fp32 column sums, then * (float)(1.0/rows)
//   * A * B, no transposes, inner dim 3 and result width 3 (gemm small-matrix path): plain fp32,
//     left-to-right  a0*b0 + a1*b1 + a2*b2
//   * every other product (A.t()*B, A*B.t()): gemm general path, fp64 accumulators, rounded once
//   * SVD of 3x3                        : one-sided Jacobi, singular values descending, u and vt
//   * determinant(3x3 CV_32F)           : fp32 expression, returned as double
// Nothing in the reference tree pins these semantics with a test => "parity unpinned" for this
// boundary; the parity tolerances (1e-5 in R, 1e-4 m in t) are what absorb the residual.
#pragma once
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>
#include <algorithm>

#define CV_32F 5
#define CV_REDUCE_AVG 1

namespace cv {

class Mat {
public:
	int rows = 0, cols = 0;
	unsigned char *data = nullptr;
	size_t stride = 0;                 // in floats
	std::shared_ptr<std::vector<float>> owner;   // empty when wrapping user memory
	bool transposed_view = false;      // set by t(): lazy transpose flag consumed by operator*

	Mat() {}
	Mat(int r, int c, int /*type*/) { create(r, c); }
	Mat(int r, int c, int /*type*/, void *user) : rows(r), cols(c), data((unsigned char*)user), stride(c) {}

	void create(int r, int c) {
		owner = std::make_shared<std::vector<float>>((size_t)r * c, 0.0f);
		rows = r; cols = c; stride = c; data = (unsigned char*)owner->data(); transposed_view = false;
	}
	float *ptr(int r) const { return (float*)data + (size_t)r * stride; }
	template <typename T> T &at(int r, int c) const { return ((T*)data)[(size_t)r * stride + c]; }

	Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.stride = stride; m.data = (unsigned char*)ptr(r); m.owner = owner; return m; }

	// Lazy transpose: only ever used as an operand of operator* in icp.cpp.
	Mat t() const { Mat m(*this); m.transposed_view = !transposed_view; return m; }

	Mat &operator+=(const Mat &b) {
		for (int r = 0; r < rows; r++) { float *d = ptr(r); const float *s = b.ptr(b.rows == 1 ? 0 : r); for (int c = 0; c < cols; c++) d[c] = d[c] + s[c]; }
		return *this;
	}
	static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < std::min(r, c); i++) m.at<float>(i, i) = 1.0f; return m; }
};

inline Mat operator-(const Mat &a, const Mat &b) {
	Mat d(a.rows, a.cols, CV_32F);
	for (int r = 0; r < a.rows; r++) { const float *pa = a.ptr(r), *pb = b.ptr(r); float *pd = d.ptr(r); for (int c = 0; c < a.cols; c++) pd[c] = pa[c] - pb[c]; }
	return d;
}

inline Mat operator*(const Mat &a, const Mat &b) {
	const bool ta = a.transposed_view, tb = b.transposed_view;
	const int ar = ta ? a.cols : a.rows, ac = ta ? a.rows : a.cols;
	const int bc = tb ? b.rows : b.cols;
	Mat d(ar, bc, CV_32F);
	if (!ta && !tb && ac == 3 && bc == 3) {
		// small-matrix path: fp32, left to right
		const float *b0 = b.ptr(0), *b1 = b.ptr(1), *b2 = b.ptr(2);
		for (int i = 0; i < ar; i++) {
			const float *pa = a.ptr(i); float *pd = d.ptr(i);
			float t0 = pa[0] * b0[0] + pa[1] * b1[0] + pa[2] * b2[0];
			float t1 = pa[0] * b0[1] + pa[1] * b1[1] + pa[2] * b2[1];
			float t2 = pa[0] * b0[2] + pa[1] * b1[2] + pa[2] * b2[2];
			pd[0] = t0; pd[1] = t1; pd[2] = t2;
		}
		return d;
	}
	// general path: fp64 accumulators
	std::vector<double> acc((size_t)ar * bc, 0.0);
	for (int k = 0; k < ac; k++)
		for (int i = 0; i < ar; i++) {
			const double av = ta ? (double)a.at<float>(k, i) : (double)a.at<float>(i, k);
			for (int j = 0; j < bc; j++) {
				const double bv = tb ? (double)b.at<float>(j, k) : (double)b.at<float>(k, j);
				acc[(size_t)i * bc + j] += av * bv;
			}
		}
	for (int i = 0; i < ar; i++) for (int j = 0; j < bc; j++) d.at<float>(i, j) = (float)acc[(size_t)i * bc + j];
	return d;
}

inline void reduce(const Mat &src, Mat &dst, int /*dim=0*/, int /*CV_REDUCE_AVG*/) {
	std::vector<float> buf(src.cols);
	for (int c = 0; c < src.cols; c++) buf[c] = src.ptr(0)[c];
	for (int r = 1; r < src.rows; r++) { const float *p = src.ptr(r); for (int c = 0; c < src.cols; c++) buf[c] = buf[c] + p[c]; }
	dst.create(1, src.cols);
	const float scale = (float)(1.0 / src.rows);
	for (int c = 0; c < src.cols; c++) dst.at<float>(0, c) = buf[c] * scale;
}

inline double determinant(const Mat &m) {
	const float *a = m.ptr(0), *b = m.ptr(1), *c = m.ptr(2);
	float d = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
	return d;
}

// One-sided (Hestenes) Jacobi SVD, fp32 storage with fp64 dot products, singular values sorted
// descending; A = u * diag(w) * vt.
class SVD {
public:
	Mat u, w, vt;
	SVD &operator()(const Mat &src) {
		const int n = 3;
		float At[3][3], Vt[3][3]; double W[3];
		for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { At[i][j] = src.at<float>(j, i); Vt[i][j] = (i == j) ? 1.0f : 0.0f; }
		for (int i = 0; i < n; i++) { double sd = 0; for (int k = 0; k < n; k++) sd += (double)At[i][k] * At[i][k]; W[i] = sd; }
		const float eps = 1.1920929e-07f * 2;
		for (int iter = 0; iter < 30; iter++) {
			bool changed = false;
			for (int i = 0; i < n - 1; i++) for (int j = i + 1; j < n; j++) {
				float *Ai = At[i], *Aj = At[j];
				double a = W[i], p = 0, b = W[j];
				for (int k = 0; k < n; k++) p += (double)Ai[k] * Aj[k];
				if (std::abs(p) <= eps * std::sqrt(a * b)) continue;
				p *= 2;
				double beta = a - b, gamma = hypot(p, beta);
				float c, s;
				if (beta < 0) { double delta = (gamma - beta) * 0.5; s = (float)std::sqrt(delta / gamma); c = (float)(p / (gamma * s * 2)); }
				else { c = (float)std::sqrt((gamma + beta) / (gamma * 2)); s = (float)(p / (gamma * c * 2)); }
				a = b = 0;
				for (int k = 0; k < n; k++) { float t0 = c * Ai[k] + s * Aj[k]; float t1 = -s * Ai[k] + c * Aj[k]; Ai[k] = t0; Aj[k] = t1; a += (double)t0 * t0; b += (double)t1 * t1; }
				W[i] = a; W[j] = b;
				changed = true;
				float *Vi = Vt[i], *Vj = Vt[j];
				for (int k = 0; k < n; k++) { float t0 = c * Vi[k] + s * Vj[k]; float t1 = -s * Vi[k] + c * Vj[k]; Vi[k] = t0; Vj[k] = t1; }
			}
			if (!changed) break;
		}
		for (int i = 0; i < n; i++) { double sd = 0; for (int k = 0; k < n; k++) sd += (double)At[i][k] * At[i][k]; W[i] = std::sqrt(sd); }
		for (int i = 0; i < n - 1; i++) {
			int j = i; for (int k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
			if (i != j) { std::swap(W[i], W[j]); for (int k = 0; k < n; k++) { std::swap(At[i][k], At[j][k]); std::swap(Vt[i][k], Vt[j][k]); } }
		}
		u.create(3, 3); vt.create(3, 3); w.create(3, 1);
		for (int i = 0; i < n; i++) {
			w.at<float>(i, 0) = (float)W[i];
			double sd = W[i];
			// rows of At are u-columns scaled by w; a (near-)zero singular value cannot happen for the
			// well-conditioned cross-covariances ICP produces, but keep the result finite anyway.
			float inv = sd > 0 ? (float)(1.0 / sd) : 0.0f;
			for (int k = 0; k < n; k++) { u.at<float>(k, i) = At[i][k] * inv; vt.at<float>(i, k) = Vt[i][k]; }
		}
		return *this;
	}
};

}  // namespace cv
