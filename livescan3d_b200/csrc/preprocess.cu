// preprocess.cu — the per-frame passes that run on the raw depth/colour maps just before the vertex path
// (SURVEY.md §8f "next" rows N1, N2), on the device.
//
// Reference behaviour reproduced (paths relative to the LiveScan3D tree):
//   depthMapAndColorRadialCorrection       src/NativeUtils/depthprocessing.cpp:191-261   (export :1794-1815)
//   KinectCapture::filterFlyingPixels      src/LiveScanClient/kinectCapture.cpp:132-174
//
// Radial correction is a forward warp with "last source in raster order wins" followed by an IN-PLACE raster-order
// hole fill, i.e. a sequential recurrence in the reference.  Here:
//   k_rad_scatter   atomicMax of (source raster index + 1) per destination pixel        == last writer in raster order
//   k_rad_gather    destination pulls depth + colour from its winner; holes are marked pending
//   k_rad_round     every pending hole is finalised as soon as its outcome is certain: either its four raster-earlier
//                   neighbours (NW, N, NE, W) are final — then the reference's own neighbour loop is evaluated on final
//                   values — or even if every unresolved one of them were filled it could not collect more than 4
//                   neighbours, so it stays 0.  Values are published with "write, fence, flag", so any interleaving of
//                   threads gives the sequential result.  A few grid-wide rounds settle all but the genuine fill cascades;
//   k_rad_chain     one block per sensor iterates the remaining (short) worklist to its fixed point;
//   k_rad_writeback results back into the caller's buffers (the reference works in place).
#include "ls3d_common.cuh"
#include "ls3d_internal.h"
#include "../../include/ls3d.h"

#include <algorithm>
#include <climits>
#include <cstring>
#include <vector>

namespace ls3d {

struct PreSensor {
	int w, h;
	long long pix_begin;               // first pixel of this sensor in the packed frame
	float cx, cy, fx, fy, r2, r4, r6;  // IntrinsicCameraParameters, depthprocessing.h:90-98
	int pad;
};

enum : unsigned char { kRadDone = 1, kRadHole = 2 };
constexpr int kRadRounds = 3;            // grid-wide rounds before the per-sensor worklists
constexpr int kRadChainThreads = 1024;

__device__ __forceinline__ unsigned ld_vol_u8(const unsigned char *p) {
	unsigned v;
	asm volatile("ld.volatile.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ unsigned ld_vol_u16(const unsigned short *p) {
	unsigned short v;
	asm volatile("ld.volatile.global.u16 %0, [%1];" : "=h"(v) : "l"(p));
	return v;
}

// C's (int) of a float the way x86 does it: out of range / NaN -> INT_MIN ("integer indefinite"), which then fails the >= 0 test
__device__ __forceinline__ int c_float_to_int(float f) {
	return (f > -2147483904.0f && f < 2147483648.0f) ? __float2int_rz(f) : INT_MIN;
}

__global__ void __launch_bounds__(256) k_rad_scatter(const uint8_t *__restrict__ depth, const PreSensor *__restrict__ sd, int *winner) {
	const PreSensor s = sd[blockIdx.y];
	const unsigned short *dm = reinterpret_cast<const unsigned short *>(depth) + s.pix_begin;
	const int px = s.w * s.h;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < px; p += gridDim.x * blockDim.x) {
		if (dm[p] == 0) continue;
		const int y = p / s.w, x = p - y * s.w;
		// depthprocessing.cpp:206-212, fp32, evaluation order of the C expressions, no contraction
		const float u = __fdiv_rn(__fsub_rn((float)x, s.cx), s.fx);
		const float v = __fdiv_rn(__fsub_rn((float)y, s.cy), s.fy);
		const float r = __fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v));
		float d = __fsub_rn(1.0f, __fmul_rn(s.r2, r));
		d = __fsub_rn(d, __fmul_rn(__fmul_rn(s.r4, r), r));
		d = __fsub_rn(d, __fmul_rn(__fmul_rn(__fmul_rn(s.r6, r), r), r));
		const int xc = c_float_to_int(__fadd_rn(__fmul_rn(__fmul_rn(u, d), s.fx), s.cx));
		const int yc = c_float_to_int(__fadd_rn(__fmul_rn(__fmul_rn(v, d), s.fy), s.cy));
		if (xc >= 0 && yc >= 0 && xc < s.w && yc < s.h) atomicMax(&winner[s.pix_begin + xc + (long long)yc * s.w], p + 1);
	}
}

__global__ void __launch_bounds__(256) k_rad_gather(const uint8_t *__restrict__ depth, const uint8_t *__restrict__ colors, const PreSensor *__restrict__ sd,
	const int *__restrict__ winner, unsigned short *__restrict__ fdepth, uint8_t *__restrict__ fcolors, unsigned char *__restrict__ state)
{
	const PreSensor s = sd[blockIdx.y];
	const unsigned short *dm = reinterpret_cast<const unsigned short *>(depth) + s.pix_begin;
	const uint8_t *cm = colors + 3 * s.pix_begin;
	const int px = s.w * s.h;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < px; p += gridDim.x * blockDim.x) {
		const long long gp = s.pix_begin + p;
		const int src = winner[gp] - 1;
		unsigned short d = 0;
		uint8_t c0 = 0, c1 = 0, c2 = 0;
		if (src >= 0) { d = dm[src]; c0 = cm[3 * (size_t)src]; c1 = cm[3 * (size_t)src + 1]; c2 = cm[3 * (size_t)src + 2]; }
		fdepth[gp] = d;
		fcolors[3 * gp] = c0; fcolors[3 * gp + 1] = c1; fcolors[3 * gp + 2] = c2;
		const int y = p / s.w, x = p - y * s.w;
		const bool interior = x >= 1 && x < s.w - 1 && y >= 1 && y < s.h - 1;
		state[gp] = (interior && d == 0) ? kRadHole : kRadDone;           // holes start pending
	}
}

// Try to finalise the pending hole at pixel p of sensor s.  Returns true when it is final now.
__device__ bool rad_try_resolve(const PreSensor &s, int p, unsigned short *fdepth, uint8_t *fcolors, unsigned char *state) {
	const long long gp = s.pix_begin + p;
	const int w = s.w;
	const int nb[8] = {-w - 1, -w, -w + 1, -1, 1, w - 1, w, w + 1};      // depthprocessing.cpp:226
	unsigned st[8];
	bool all_done = true;
#pragma unroll
	for (int i = 0; i < 8; i++) st[i] = ld_vol_u8(state + gp + nb[i]);
#pragma unroll
	for (int i = 0; i < 4; i++) all_done = all_done && (st[i] & kRadDone);
	__threadfence();                                                       // values of neighbours seen as final are read after their flags
	int val[8];
#pragma unroll
	for (int i = 0; i < 8; i++) {
		// raster-earlier neighbours: their current (final, if flagged) value; raster-later neighbours: the warped value, which for an
		// original hole is 0 whatever has been filled into it since (the reference has not reached it yet)
		if (i < 4) val[i] = (st[i] & kRadDone) ? (int)ld_vol_u16(fdepth + gp + nb[i]) : -1;       // -1: not known yet
		else val[i] = (st[i] & kRadHole) ? 0 : (int)ld_vol_u16(fdepth + gp + nb[i]);
	}
	if (!all_done) {
		int possible = 0;
#pragma unroll
		for (int i = 0; i < 8; i++) possible += val[i] != 0 ? 1 : 0;       // unknown (-1) counts as "might be filled"
		if (possible > 4) return false;
		__threadfence();
		state[gp] = kRadHole | kRadDone;                                   // cannot reach n > 4: stays 0 (depthprocessing.cpp:249)
		return true;
	}
	int n = 0, sum = 0, sr = 0, sg = 0, sb = 0, prev = -1;
#pragma unroll
	for (int i = 0; i < 8; i++) {
		if (val[i] > 0 && (prev == -1 || abs(val[i] - prev) < 30)) {           // :239
			prev = val[i]; n++; sum += val[i];
			const uint8_t *c = fcolors + 3 * (gp + nb[i]);
			sr += (int)ld_vol_u8(c); sg += (int)ld_vol_u8(c + 1); sb += (int)ld_vol_u8(c + 2);
		}
	}
	if (n > 4) {
		fcolors[3 * gp] = (uint8_t)(sr / n); fcolors[3 * gp + 1] = (uint8_t)(sg / n); fcolors[3 * gp + 2] = (uint8_t)(sb / n);
		fdepth[gp] = (unsigned short)(sum / n);
	}
	__threadfence();
	state[gp] = kRadHole | kRadDone;
	return true;
}

// one grid-wide round; the last one queues what is still pending (worklist of sensor s: items + s.pix_begin, count[s])
__global__ void __launch_bounds__(256) k_rad_round(const PreSensor *__restrict__ sd, unsigned short *fdepth, uint8_t *fcolors, unsigned char *state,
	int queue, int *items, int *count)
{
	const PreSensor s = sd[blockIdx.y];
	const int px = s.w * s.h;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < px; p += gridDim.x * blockDim.x) {
		if (ld_vol_u8(state + s.pix_begin + p) & kRadDone) continue;
		if (!rad_try_resolve(s, p, fdepth, fcolors, state) && queue) items[s.pix_begin + atomicAdd(&count[blockIdx.y], 1)] = p;
	}
}

// the remaining fill cascades of one sensor, to their fixed point.  The raster-first pending item can always be resolved, so
// every round makes progress; the round limit only guards against a broken invariant (flagged in err).
__global__ void __launch_bounds__(kRadChainThreads) k_rad_chain(const PreSensor *__restrict__ sd, unsigned short *fdepth, uint8_t *fcolors, unsigned char *state,
	int *items, const int *__restrict__ count, int *err)
{
	__shared__ int s_left, s_prev;
	const PreSensor s = sd[blockIdx.x];
	int *mine = items + s.pix_begin;
	const int n = count[blockIdx.x];
	if (n == 0) return;
	if (threadIdx.x == 0) { s_left = n; s_prev = n + 1; }
	__syncthreads();
	for (int round = 0; round <= n; round++) {
		for (int j = threadIdx.x; j < n; j += blockDim.x) {
			const int p = mine[j];
			if (p >= 0 && rad_try_resolve(s, p, fdepth, fcolors, state)) { mine[j] = -1; atomicSub(&s_left, 1); }
		}
		__syncthreads();
		const int left = s_left, prev = s_prev;
		__syncthreads();
		if (left == 0) return;
		if (left == prev) { if (threadIdx.x == 0) atomicOr(err, 16); return; }
		if (threadIdx.x == 0) s_prev = left;
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256) k_rad_writeback(uint8_t *__restrict__ depth, uint8_t *__restrict__ colors, long long total_px,
	const unsigned short *__restrict__ fdepth, const uint8_t *__restrict__ fcolors)
{
	unsigned short *dm = reinterpret_cast<unsigned short *>(depth);
	for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total_px; p += (long long)gridDim.x * blockDim.x) {
		dm[p] = fdepth[p];
		colors[3 * p] = fcolors[3 * p]; colors[3 * p + 1] = fcolors[3 * p + 1]; colors[3 * p + 2] = fcolors[3 * p + 2];
	}
}

// ------------------------------------------------------------------------------------------------------
// flying-pixel filter (N2)
// ------------------------------------------------------------------------------------------------------
// KinectCapture::filterFlyingPixels, kinectCapture.cpp:132-174: a pixel is zeroed when more than nNeighbours/2 of the
// (2k+1)^2-1 pixels around it differ from it by more than thr (the caller's maxNonFittingNeighbours is overwritten, :150);
// the neighbour offsets are x*width + y with x,y in [-k,k] (:141-147), i.e. a square window either way; decisions are made
// on the unmodified image and applied afterwards (:169-172), so the stencil is order-free.  `thr` is a float compared with
// an int difference (:163).
__global__ void __launch_bounds__(256) k_flying_pixels(const unsigned short *__restrict__ in, unsigned short *__restrict__ out, int w, int h, int k, float thr) {
	const int n_nb = (2 * k + 1) * (2 * k + 1) - 1;
	const int max_bad = n_nb / 2;
	const int px = w * h;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < px; p += gridDim.x * blockDim.x) {
		const int y = p / w, x = p - y * w;
		unsigned short v = in[p];
		if (y >= k && y < h - k && x >= k && x < w - k) {
			const int val = v;
			int bad = 0;
			for (int a = -k; a <= k; a++)
				for (int b = -k; b <= k; b++) {
					if (a == 0 && b == 0) continue;
					const int diff = abs((int)in[p + a * w + b] - val);
					bad += (float)diff > thr ? 1 : 0;
				}
			if (bad > max_bad) v = 0;
		}
		out[p] = v;
	}
}

}  // namespace ls3d

using namespace ls3d;

// ======================================================================================================
// host side
// ======================================================================================================
namespace {

struct PreCtx {
	std::vector<int> w, h;
	long long total_px = 0;
	DevBuf sd, winner, fdepth, fcolors, state, items, count, in_depth, in_colors, tmp;
	PreSensor *pin_sd = nullptr;
	int *pin_err = nullptr;
	int sm_count = 148;
};

PreCtx *g_pre = nullptr;

PreCtx *pre_ctx(int n_maps, const int *widths, const int *heights) {
	if (g_pre && (int)g_pre->w.size() == n_maps && !memcmp(g_pre->w.data(), widths, sizeof(int) * n_maps) && !memcmp(g_pre->h.data(), heights, sizeof(int) * n_maps)) return g_pre;
	if (g_pre) {
		DevBuf *bufs[] = {&g_pre->sd, &g_pre->winner, &g_pre->fdepth, &g_pre->fcolors, &g_pre->state, &g_pre->items, &g_pre->count, &g_pre->in_depth, &g_pre->in_colors, &g_pre->tmp};
		for (DevBuf *b : bufs) b->release();
		if (g_pre->pin_sd) cudaFreeHost(g_pre->pin_sd);
		if (g_pre->pin_err) cudaFreeHost(g_pre->pin_err);
		delete g_pre;
		g_pre = nullptr;
	}
	PreCtx *c = new PreCtx();
	int dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev);
	long long acc = 0;
	for (int i = 0; i < n_maps; i++) {
		if (widths[i] <= 0 || heights[i] <= 0 || (long long)widths[i] * heights[i] >= (1ll << 30)) { set_error("radial correction: map %d has unsupported size %dx%d", i, widths[i], heights[i]); delete c; return nullptr; }
		acc += (long long)widths[i] * heights[i];
	}
	if (acc >= (1ll << 31)) { set_error("radial correction: frame too large"); delete c; return nullptr; }
	c->w.assign(widths, widths + n_maps);
	c->h.assign(heights, heights + n_maps);
	c->total_px = acc;
	const size_t n = (size_t)acc;
	bool ok = c->sd.reserve(sizeof(PreSensor) * n_maps, "alloc descriptors") && c->winner.reserve(4 * n, "alloc warp winners") && c->fdepth.reserve(2 * n, "alloc warped depth") &&
		c->fcolors.reserve(3 * n, "alloc warped colours") && c->state.reserve(n, "alloc hole states") && c->items.reserve(4 * n, "alloc worklists") &&
		c->count.reserve(4 * (size_t)n_maps + 4, "alloc worklist counts");
	ok = ok && cuda_ok(cudaHostAlloc((void **)&c->pin_sd, sizeof(PreSensor) * n_maps, cudaHostAllocDefault), "alloc pinned descriptors");
	ok = ok && cuda_ok(cudaHostAlloc((void **)&c->pin_err, 64, cudaHostAllocDefault), "alloc pinned status");
	if (!ok) { delete c; return nullptr; }
	g_pre = c;
	return c;
}

// enqueue the whole correction on st for device-resident packed buffers (in place); err word = count[n_maps]
int radial_enqueue(PreCtx *c, int n_maps, uint8_t *d_depth, uint8_t *d_colors, const float *intr_params, cudaStream_t st) {
	long long acc = 0;
	int max_px = 1;
	for (int i = 0; i < n_maps; i++) {
		PreSensor &s = c->pin_sd[i];
		memset(&s, 0, sizeof(s));
		s.w = c->w[i]; s.h = c->h[i]; s.pix_begin = acc;
		const float *ip = intr_params + 7 * i;
		s.cx = ip[0]; s.cy = ip[1]; s.fx = ip[2]; s.fy = ip[3]; s.r2 = ip[4]; s.r4 = ip[5]; s.r6 = ip[6];
		acc += (long long)s.w * s.h;
		max_px = std::max(max_px, s.w * s.h);
	}
	bool ok = cuda_ok(cudaMemcpyAsync(c->sd.p, c->pin_sd, sizeof(PreSensor) * n_maps, cudaMemcpyHostToDevice, st), "upload descriptors") &&
		cuda_ok(cudaMemsetAsync(c->winner.p, 0, 4 * (size_t)c->total_px, st), "clear winners") &&
		cuda_ok(cudaMemsetAsync(c->count.p, 0, 4 * (size_t)n_maps + 4, st), "clear worklist counts");
	if (!ok) return -1;
	const dim3 grid((unsigned)std::max(1, std::min((max_px + 255) / 256, c->sm_count * 8 / std::max(1, std::min(n_maps, 8)) + 1)), (unsigned)n_maps);
	const PreSensor *sd = c->sd.as<PreSensor>();
	int *count = c->count.as<int>();
	k_rad_scatter<<<grid, 256, 0, st>>>(d_depth, sd, c->winner.as<int>());
	k_rad_gather<<<grid, 256, 0, st>>>(d_depth, d_colors, sd, c->winner.as<int>(), c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>(), c->state.as<unsigned char>());
	for (int r = 0; r < kRadRounds; r++)
		k_rad_round<<<grid, 256, 0, st>>>(sd, c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>(), c->state.as<unsigned char>(), r == kRadRounds - 1 ? 1 : 0, c->items.as<int>(), count);
	k_rad_chain<<<n_maps, kRadChainThreads, 0, st>>>(sd, c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>(), c->state.as<unsigned char>(), c->items.as<int>(), count, count + n_maps);
	k_rad_writeback<<<(unsigned)std::max<long long>(1, std::min<long long>((c->total_px + 255) / 256, (long long)c->sm_count * 8)), 256, 0, st>>>(d_depth, d_colors, c->total_px,
		c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>());
	count_launch(4 + kRadRounds);
	return cuda_ok(cudaGetLastError(), "radial correction kernels") ? 4 + kRadRounds : -1;
}

}  // namespace

// Device-resident entry: corrects the packed depth / colour buffers of n_maps sensors IN PLACE on `stream` (no synchronisation).
extern "C" int ls3d_radial_correction_device(int n_maps, void *d_depth_maps, void *d_depth_colors, const int *widths, const int *heights, const float *intr_params, void *stream) {
	clear_error();
	if (n_maps <= 0 || !d_depth_maps || !d_depth_colors || !widths || !heights || !intr_params) { set_error("ls3d_radial_correction_device: bad arguments"); return -1; }
	if (!ensure_device()) return -1;
	std::lock_guard<std::mutex> lk(api_mutex());
	PreCtx *c = pre_ctx(n_maps, widths, heights);
	if (!c) return -1;
	return radial_enqueue(c, n_maps, (uint8_t *)d_depth_maps, (uint8_t *)d_depth_colors, intr_params, (cudaStream_t)stream);
}

// Replaces depthMapAndColorSetRadialCorrection (include/NativeUtils/depthprocessing.h:111, src/NativeUtils/depthprocessing.cpp:1794-1815;
// C# binding LiveScanServer/KinectServer.cs:51-53): host buffers, corrected in place.
extern "C" void depthMapAndColorSetRadialCorrection(int n_maps, unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights, float *intr_params) {
	clear_error();
	if (n_maps <= 0) return;
	if (!depth_maps || !depth_colors || !widths || !heights || !intr_params) { set_error("depthMapAndColorSetRadialCorrection: null argument"); return; }
	if (!ensure_device()) return;
	std::lock_guard<std::mutex> lk(api_mutex());
	cudaStream_t st = api_stream();
	if (!st) return;
	PreCtx *c = pre_ctx(n_maps, widths, heights);
	if (!c) return;
	const size_t n = (size_t)c->total_px;
	if (!c->in_depth.reserve(2 * n, "alloc depth input") || !c->in_colors.reserve(3 * n, "alloc colour input")) return;
	bool ok = cuda_ok(cudaMemcpyAsync(c->in_depth.p, depth_maps, 2 * n, cudaMemcpyHostToDevice, st), "upload depth") &&
		cuda_ok(cudaMemcpyAsync(c->in_colors.p, depth_colors, 3 * n, cudaMemcpyHostToDevice, st), "upload colours");
	if (!ok || radial_enqueue(c, n_maps, c->in_depth.as<uint8_t>(), c->in_colors.as<uint8_t>(), intr_params, st) < 0) return;
	// results go to pinned staging first: the caller's arrays are only overwritten once the whole call has succeeded
	ok = cuda_ok(cudaMemcpyAsync(c->pin_err, c->count.as<int>() + n_maps, sizeof(int), cudaMemcpyDeviceToHost, st), "read status") &&
		cuda_ok(cudaStreamSynchronize(st), "radial correction");
	if (!ok) return;
	if (*c->pin_err) { set_error("radial correction: device status flags 0x%x", *c->pin_err); return; }
	ok = cuda_ok(cudaMemcpyAsync(depth_maps, c->in_depth.p, 2 * n, cudaMemcpyDeviceToHost, st), "read depth") &&
		cuda_ok(cudaMemcpyAsync(depth_colors, c->in_colors.p, 3 * n, cudaMemcpyDeviceToHost, st), "read colours") &&
		cuda_ok(cudaStreamSynchronize(st), "radial correction read-back");
	(void)ok;
}

// KinectCapture::filterFlyingPixels (kinectCapture.cpp:132-174) on one host depth image, in place.  maxNonFittingNeighbours is accepted and
// ignored exactly as in the reference (:150 overwrites it with nNeighbours / 2).  Returns the number of kernels launched (1) or -1.
extern "C" int ls3d_filter_flying_pixels(unsigned short *depth, int width, int height, int neighbourhoodSize, float thr, int maxNonFittingNeighbours) {
	(void)maxNonFittingNeighbours;
	clear_error();
	if (!depth || width <= 0 || height <= 0 || neighbourhoodSize < 0 || (long long)width * height >= (1ll << 30)) { set_error("ls3d_filter_flying_pixels: bad arguments"); return -1; }
	if (!ensure_device()) return -1;
	std::lock_guard<std::mutex> lk(api_mutex());
	cudaStream_t st = api_stream();
	if (!st) return -1;
	static DevBuf in, out;
	const size_t n = (size_t)width * height;
	if (!in.reserve(2 * n, "alloc depth") || !out.reserve(2 * n, "alloc filtered depth")) return -1;
	int sm = 148, dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
	if (!cuda_ok(cudaMemcpyAsync(in.p, depth, 2 * n, cudaMemcpyHostToDevice, st), "upload depth")) return -1;
	k_flying_pixels<<<(unsigned)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)sm * 8)), 256, 0, st>>>(in.as<unsigned short>(), out.as<unsigned short>(), width, height, neighbourhoodSize, thr);
	count_launch(1);
	bool ok = cuda_ok(cudaGetLastError(), "k_flying_pixels") && cuda_ok(cudaMemcpyAsync(depth, out.p, 2 * n, cudaMemcpyDeviceToHost, st), "read depth") &&
		cuda_ok(cudaStreamSynchronize(st), "flying pixel filter");
	return ok ? 1 : -1;
}

// Device-resident variant: in -> out (distinct device buffers of width*height u16), enqueued on `stream`.
extern "C" int ls3d_filter_flying_pixels_device(const void *d_in, void *d_out, int width, int height, int neighbourhoodSize, float thr, void *stream) {
	clear_error();
	if (!d_in || !d_out || d_in == d_out || width <= 0 || height <= 0 || neighbourhoodSize < 0) { set_error("ls3d_filter_flying_pixels_device: bad arguments (in and out must be distinct)"); return -1; }
	if (!ensure_device()) return -1;
	int sm = 148, dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
	const size_t n = (size_t)width * height;
	k_flying_pixels<<<(unsigned)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)sm * 8)), 256, 0, (cudaStream_t)stream>>>((const unsigned short *)d_in, (unsigned short *)d_out, width, height, neighbourhoodSize, thr);
	count_launch(1);
	return cuda_ok(cudaGetLastError(), "k_flying_pixels") ? 1 : -1;
}
