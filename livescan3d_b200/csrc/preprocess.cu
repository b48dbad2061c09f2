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
//                   threads gives the sequential result.  One grid-wide round settles the vast majority of holes;
//   k_rad_fixpoint  what hangs on other holes (fill cascades, the thin hole curves of the warp) is settled by fixpoint iteration: a
//                   cluster of eight blocks per sensor re-evaluates the pending holes from each other's tentative values (64-bit
//                   words in shared / distributed shared memory) until a pass changes nothing — the recurrence is over a DAG, so
//                   the fixpoint is unique and is the reference's result; averages of 5-8 values make it contract in 15-30 passes;
//   k_rad_chains    the chain-by-chain kernel of rounds 1-2 (pending pixels listed in raster order and resolved chunk by chunk, each
//                   entry waiting for the earlier entries it reads): still runs for a sensor too dense for k_rad_fixpoint's shared
//                   memory, or with LS3D_RADIAL_CHAINS=1; k_rad_wavefront, the lockstep version before it, behind LS3D_RADIAL_WAVEFRONT=1;
//   k_rad_writeback results back into the caller's buffers (the reference works in place).
#include "ls3d_common.cuh"
#include "ls3d_internal.h"
#include "../../include/ls3d.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
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
#ifndef LS3D_RAD_ROUNDS
#define LS3D_RAD_ROUNDS 1
#endif
constexpr int kRadRounds = LS3D_RAD_ROUNDS;   // grid-wide rounds before the chain kernel
constexpr int kRadChainThreads = 1024;

// loads that must observe other threads' stores: device scope for the grid-wide round (other SMs write), block scope for the
// wavefront (one block per sensor: every writer is on this SM, so the loads may hit its L1 and the fences stay local)
template <bool kCta> __device__ __forceinline__ unsigned ld_vol_u8(const unsigned char *p) {
	unsigned v;
	if (kCta) asm volatile("ld.relaxed.cta.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	else asm volatile("ld.volatile.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
template <bool kCta> __device__ __forceinline__ unsigned ld_vol_u16(const unsigned short *p) {
	unsigned short v;
	if (kCta) asm volatile("ld.relaxed.cta.global.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
	else asm volatile("ld.volatile.global.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
	return v;
}
// In the wavefront the ordering comes from its own step structure (__syncwarp inside a warp, fence + progress counter between warps),
// so the per-pixel fences are only needed in the grid-wide round.
template <bool kCta> __device__ __forceinline__ void rad_fence() { if (!kCta) __threadfence(); }

// C's (int) of a float the way x86 does it: out of range / NaN -> INT_MIN ("integer indefinite"), which then fails the >= 0 test
__device__ __forceinline__ int c_float_to_int(float f) {
	return (f > -2147483904.0f && f < 2147483648.0f) ? __float2int_rz(f) : INT_MIN;
}

// The correction is a chain of short kernels: from the gather on each is launched programmatically behind its predecessor (pdl_enter).
__global__ void __launch_bounds__(256) k_rad_scatter(const uint8_t *__restrict__ depth, const PreSensor *__restrict__ sd, int *winner) {
	pdl_trigger();
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
	pdl_enter();
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
template <bool kCta>
__device__ __forceinline__ bool rad_try_resolve(const PreSensor &s, int p, unsigned short *fdepth, uint8_t *fcolors, unsigned char *state) {
	const long long gp = s.pix_begin + p;
	const int w = s.w;
	const int nb[8] = {-w - 1, -w, -w + 1, -1, 1, w - 1, w, w + 1};      // depthprocessing.cpp:226
	unsigned st[8];
	int raw[8];
	unsigned col[8][3];
	bool all_done = true;
	if (kCta) {
		// wavefront: the raster-earlier neighbours are final by construction, and every address is known up front, so all 40 loads are
		// issued together (one memory round trip instead of a dependent chain through the consistency loop below).  Earlier neighbours
		// may have been written by other threads of this block since the gather: block-scope strong loads; the rest cannot have changed.
#pragma unroll
		for (int i = 0; i < 8; i++) {
			st[i] = (unsigned)state[gp + nb[i]] | kRadDone;                // only the immutable hole bit matters here
			const uint8_t *c = fcolors + 3 * (gp + nb[i]);
			if (i < 4) {
				raw[i] = (int)ld_vol_u16<true>(fdepth + gp + nb[i]);
				col[i][0] = ld_vol_u8<true>(c); col[i][1] = ld_vol_u8<true>(c + 1); col[i][2] = ld_vol_u8<true>(c + 2);
			} else {
				raw[i] = (int)fdepth[gp + nb[i]];
				col[i][0] = c[0]; col[i][1] = c[1]; col[i][2] = c[2];
			}
		}
	} else {
#pragma unroll
		for (int i = 0; i < 8; i++) st[i] = ld_vol_u8<false>(state + gp + nb[i]);
#pragma unroll
		for (int i = 0; i < 4; i++) all_done = all_done && (st[i] & kRadDone);
		__threadfence();                                                   // values of neighbours seen as final are read after their flags
#pragma unroll
		for (int i = 0; i < 8; i++) {
			raw[i] = (int)ld_vol_u16<false>(fdepth + gp + nb[i]);
			const uint8_t *c = fcolors + 3 * (gp + nb[i]);
			col[i][0] = ld_vol_u8<false>(c); col[i][1] = ld_vol_u8<false>(c + 1); col[i][2] = ld_vol_u8<false>(c + 2);
		}
	}
	int val[8];
#pragma unroll
	for (int i = 0; i < 8; i++) {
		// raster-earlier neighbours: their final value (unknown while not flagged); raster-later neighbours: the warped value, which for an
		// original hole is 0 whatever has been filled into it since (the reference has not reached it yet)
		if (i < 4) val[i] = (st[i] & kRadDone) ? raw[i] : -1;
		else val[i] = (st[i] & kRadHole) ? 0 : raw[i];
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
			sr += (int)col[i][0]; sg += (int)col[i][1]; sb += (int)col[i][2];
		}
	}
	if (n > 4) {
		fcolors[3 * gp] = (uint8_t)(sr / n); fcolors[3 * gp + 1] = (uint8_t)(sg / n); fcolors[3 * gp + 2] = (uint8_t)(sb / n);
		fdepth[gp] = (unsigned short)(sum / n);
	}
	rad_fence<kCta>();
	state[gp] = kRadHole | kRadDone;
	return true;
}

// one grid-wide round: settles every hole whose outcome does not hang on another undecided hole (the vast majority)
__global__ void __launch_bounds__(256) k_rad_round(const PreSensor *__restrict__ sd, unsigned short *fdepth, uint8_t *fcolors, unsigned char *state)
{
	pdl_enter();
	const PreSensor s = sd[blockIdx.y];
	const int px = s.w * s.h;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < px; p += gridDim.x * blockDim.x) {
		if (state[s.pix_begin + p] & kRadDone) continue;                   // its own flag: nobody else writes it (a plain, coalesced load for the 96 % that are no holes)
		rad_try_resolve<false>(s, p, fdepth, fcolors, state);
	}
}

// The same round under a SNAPSHOT rule: a hole is settled here only from what was final before the kernel started — none of its four
// raster-earlier neighbours is a hole (or it cannot reach five neighbours whatever they turn out to be).  Nothing it reads can be
// written concurrently, so there are no fences and no volatile loads (k_rad_round's write-fence-flag protocol lets it also settle
// holes whose earlier neighbours happen to be finished by another thread in time: 37 us against this kernel's few; the holes it
// leaves — the thin curves of the warp, every link of which has an earlier hole for a neighbour — are what k_rad_fixpoint is for).
__global__ void __launch_bounds__(256) k_rad_first(const PreSensor *__restrict__ sd, unsigned short *fdepth, uint8_t *fcolors, unsigned char *state)
{
	pdl_enter();
	const PreSensor s = sd[blockIdx.y];
	const int px = s.w * s.h, w = s.w;
	const int nb[8] = {-w - 1, -w, -w + 1, -1, 1, w - 1, w, w + 1};      // depthprocessing.cpp:226
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < px; p += gridDim.x * blockDim.x) {
		const long long gp = s.pix_begin + p;
		if (state[gp] & kRadDone) continue;
		int val[8];
		bool earlier_hole = false;
#pragma unroll
		for (int i = 0; i < 8; i++) {
			const bool hole = (state[gp + nb[i]] & kRadHole) != 0;          // immutable
			if (i < 4) { earlier_hole = earlier_hole || hole; val[i] = hole ? -1 : (int)fdepth[gp + nb[i]]; }
			else val[i] = hole ? 0 : (int)fdepth[gp + nb[i]];                // the reference has not reached it yet: still 0
		}
		if (earlier_hole) {
			int possible = 0;
#pragma unroll
			for (int i = 0; i < 8; i++) possible += val[i] != 0 ? 1 : 0;    // unknown (-1) counts as "might be filled"
			if (possible <= 4) state[gp] = kRadHole | kRadDone;             // cannot reach n > 4: stays 0 (depthprocessing.cpp:249)
			continue;
		}
		int n = 0, sum = 0, sr = 0, sg = 0, sb = 0, prev = -1;
#pragma unroll
		for (int i = 0; i < 8; i++) {
			if (val[i] > 0 && (prev == -1 || abs(val[i] - prev) < 30)) {       // :239
				const uint8_t *c = fcolors + 3 * (gp + nb[i]);
				prev = val[i]; n++; sum += val[i];
				sr += (int)c[0]; sg += (int)c[1]; sb += (int)c[2];
			}
		}
		if (n > 4) {
			fcolors[3 * gp] = (uint8_t)(sr / n); fcolors[3 * gp + 1] = (uint8_t)(sg / n); fcolors[3 * gp + 2] = (uint8_t)(sb / n);
			fdepth[gp] = (unsigned short)(sum / n);
		}
		state[gp] = kRadHole | kRadDone;
	}
}

// What the grid-wide round left pending — holes that wait on other holes: the genuine fill cascades, and borders of invalid
// regions that could only be decided once their neighbours were — is finished by a skewed wavefront, one block per sensor:
// thread r owns image row r and visits pixel x = t - 2r at step t, so the pixels a hole reads as "already processed"
// (NW, N, NE, W) were finished at earlier steps by construction.  Inside a warp the skew is kept by lockstep
// (__syncwarp per step); between warps lane 0 waits on the progress counter of the row above (the previous warp's last row).
// w + 2h steps whatever the dependency structure, instead of one block-wide round per link of the longest chain; a per-row
// bitmap of the pending pixels in shared memory lets a warp skip, without touching memory, every step at which none of its rows has work.
__global__ void __launch_bounds__(kRadChainThreads) k_rad_wavefront(const PreSensor *__restrict__ sd, unsigned short *fdepth, uint8_t *fcolors, unsigned char *state, int words_per_row, int *err)
{
	extern __shared__ unsigned pend[];                          // [row of the band][words_per_row]: bit x = pixel x of that row is still pending
	__shared__ volatile int prog[kRadChainThreads / 32];       // pixels finished in the last row of each warp
	const PreSensor s = sd[blockIdx.x];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int rows_per_band = (int)blockDim.x;
	const int wpr = (s.w + 31) >> 5;                            // <= words_per_row
	for (int band = 0; band * rows_per_band < s.h; band++) {
		if (threadIdx.x < kRadChainThreads / 32) prog[threadIdx.x] = 0;
		// the pending map of this band's rows (static from here on: only this kernel resolves pixels now); a warp packs its own 32 rows
		unsigned my_any = 0;
		for (int j = 0; j < 32; j++) {
			const int yy = band * rows_per_band + warp * 32 + j;
			for (int x0 = 0; x0 < wpr * 32; x0 += 32) {
				const int x = x0 + lane;
				bool pnd = false;
				if (yy >= 1 && yy < s.h - 1 && x >= 1 && x < s.w - 1) pnd = !(state[s.pix_begin + (long long)yy * s.w + x] & kRadDone);
				const unsigned bits = __ballot_sync(kFull, pnd);
				if (lane == 0) pend[(warp * 32 + j) * words_per_row + (x0 >> 5)] = bits;
				if (j == lane) my_any |= bits;
			}
		}
		__syncthreads();                                        // also: the previous band's rows are complete
		const int y = band * rows_per_band + threadIdx.x;
		if (band * rows_per_band + warp * 32 < s.h) {
			const unsigned *mine = pend + threadIdx.x * words_per_row;
			const int steps = s.w + 2 * 31;
			for (int t = 0; t < steps; t++) {
				const int x = t - 2 * lane;
				const bool todo = my_any && x >= 0 && x < s.w && ((mine[x >> 5] >> (x & 31)) & 1u);
				if (__any_sync(kFull, todo)) {
					if (warp > 0 && lane == 0 && todo) {
						// NW, N, NE of my pixel t are pixels t-1, t, t+1 of the row above (the previous warp's last row).  Only one that was
						// still pending when the band started can change under me; everything else in that row was final before this kernel
						// began, so there is nothing to wait for.  (No gain on the synthetic bench frame: what the round leaves pending there are
						// the thin hole curves of the forward warp, which cross the rows — genuine chains of ~h links, ~1 us per link.)
						const unsigned *above = mine - words_per_row;
						bool dep = false;
#pragma unroll
						for (int dx = -1; dx <= 1; dx++) {
							const int xx = t + dx;
							if (xx >= 0 && xx < s.w) dep = dep || ((above[xx >> 5] >> (xx & 31)) & 1u);
						}
						if (dep) {
							const int need = min(t + 2, s.w);
							while (prog[warp - 1] < need) __nanosleep(32);
						}
					}
					__syncwarp();
					if (todo && !rad_try_resolve<true>(s, y * s.w + x, fdepth, fcolors, state)) atomicOr(err, 16);
					__syncwarp();
				}
				// publish the last row's progress every 8 pixels (and at the row's end): one fence per 8 steps instead of one per step
				if (lane == 31 && x >= 0 && x < s.w && ((x & 7) == 7 || x == s.w - 1)) { __threadfence_block(); prog[warp] = x + 1; }
			}
		}
		__syncthreads();
	}
}

// What the grid-wide round left pending, finished chain by chain instead of step by step.  On warped Kinect frames the pending pixels
// are thin hole curves that cross the rows: genuine dependency chains of ~h links, so what matters is the latency of ONE link, and a
// link through global memory and a lockstep step costs ~1 us (k_rad_wavefront above: w + 2h steps, 1.6 ms for 512x424).  Here one
// block per sensor (a) lists its pending pixels in raster order (ordered compaction; pixel -> list index map on the side), (b) walks
// the list in chunks of 1024 entries, one thread per entry: everything an entry needs that cannot change any more — its four
// raster-later neighbours, and the raster-earlier ones that are not pending entries of the same chunk — is loaded up front; the
// thread then only waits for the (at most four) earlier entries of its own chunk, whose results travel through shared memory as
// single 64-bit words (value + final bit, so no fence): a link costs one spin iteration, ~0.2 us.  Entries of earlier chunks are
// final in global memory before a chunk starts (block barrier), so chains may cross chunk boundaries anywhere.
// (Round 2 tried the obvious alternative — one block per sensor sweeping the rows with a cp.async ring of eight rows in shared
// memory, every horizontal run of pending pixels walked by one thread, one barrier per row — and measured 1.16 ms against this
// kernel's 0.51: the pending curves are near-horizontal over long stretches, a row costs its longest run, and rows that are
// independent of each other in the dependency graph still wait for each other in a sweep.  The list keeps only true dependencies.)
constexpr int kChainThreads = 1024, kChainAdmit = 4;
__global__ void __launch_bounds__(kChainThreads) k_rad_chains(const PreSensor *__restrict__ sd, unsigned short *fdepth, uint8_t *fcolors, unsigned char *state,
	int *__restrict__ list, int *__restrict__ pidx, int *err, const int *__restrict__ only_if = nullptr)
{
	pdl_enter();
	if (only_if && !only_if[blockIdx.x]) return;              // behind k_rad_fixpoint: only the sensors it left alone
	__shared__ unsigned long long s_val[kChainThreads];          // entry i of the chunk: depth | r << 16 | g << 24 | b << 32 | final << 63
	__shared__ unsigned s_w[32];
	__shared__ unsigned s_total;
	__shared__ unsigned s_ndone;                                   // warps of the current chunk that have finished all their entries
	const PreSensor s = sd[blockIdx.x];
	const int px = s.w * s.h, w = s.w;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	int *lst = list + s.pix_begin, *pix = pidx + s.pix_begin;
	// ---- (a) pending list, raster order: 16 consecutive pixels per thread and pass ----
	unsigned n_pending = 0;
	for (int base = 0; base < px; base += kChainThreads * 16) {
		const int p0 = base + tid * 16;
		unsigned m = 0;
#pragma unroll
		for (int j = 0; j < 16; j++)
			if (p0 + j < px && !(state[s.pix_begin + p0 + j] & kRadDone)) m |= 1u << j;
		const unsigned cnt = __popc(m), incl = warp_incl_scan(cnt, lane);
		if (lane == 31) s_w[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			const unsigned v = s_w[lane], sc = warp_incl_scan(v, lane);
			s_w[lane] = sc - v;
			if (lane == 31) s_total = sc;
		}
		__syncthreads();
		unsigned o = n_pending + s_w[warp] + incl - cnt;
		while (m) {
			const int j = __ffs(m) - 1;
			m &= m - 1;
			lst[o] = p0 + j;
			pix[p0 + j] = (int)o;
			o++;
		}
		n_pending += s_total;
		__syncthreads();
	}
	// ---- (b) chunks of the list ----
	const int nb[8] = {-w - 1, -w, -w + 1, -1, 1, w - 1, w, w + 1};      // depthprocessing.cpp:226
	for (unsigned cb = 0; cb < n_pending; cb += kChainThreads) {
		((volatile unsigned long long *)s_val)[tid] = 0ull;
		if (tid == 0) s_ndone = 0;
		__syncthreads();                         // the list, the index map and every earlier chunk's results are visible from here on
		const bool has = cb + tid < n_pending;
		long long gp = 0;
		int val[8], lidx[4];
		unsigned col[8][3];
		bool in_chunk[4] = {false, false, false, false};
		if (has) {
			const int p = lst[cb + tid];
			gp = s.pix_begin + p;
#pragma unroll
			for (int i = 0; i < 8; i++) {
				const long long q = gp + nb[i];
				const uint8_t *c = fcolors + 3 * q;
				if (i < 4) {
					const unsigned st = ld_vol_u8<true>(state + q);
					in_chunk[i] = (st & kRadHole) && !(st & kRadDone);         // pending and raster-earlier: an entry of this chunk
					lidx[i] = 0;
					if (in_chunk[i]) {
						lidx[i] = pix[p + nb[i]] - (int)cb;
						if (lidx[i] < 0 || lidx[i] >= tid) { atomicOr(err, 32); in_chunk[i] = false; }
						val[i] = 0;
					} else {
						val[i] = (int)ld_vol_u16<true>(fdepth + q);              // final: as warped, or as filled by the round / an earlier chunk
					}
					col[i][0] = ld_vol_u8<true>(c); col[i][1] = ld_vol_u8<true>(c + 1); col[i][2] = ld_vol_u8<true>(c + 2);
				} else {
					// raster-later: the warped value, which for an original hole is 0 whatever has been filled into it since
					val[i] = (state[q] & kRadHole) ? 0 : (int)fdepth[q];
					col[i][0] = c[0]; col[i][1] = c[1]; col[i][2] = c[2];
				}
			}
		}
		__syncthreads();                         // nobody publishes before everybody has taken its snapshot: no entry of this chunk is final in it
		bool done = !has;
		// Results sweep through the chunk as a front (an entry only depends on entries at most a row earlier): a warp far behind the
		// front has nothing to poll for and would only take issue slots from the warps that resolve — it sleeps until all but
		// kChainAdmit of the warps before it are done, and only then polls, tightly: a link should cost one loop iteration
		// (measured: no admission 0.67 ms, window 1 / 4 / 8 warps 0.90 / 0.54 / 0.55 ms; any sleeping inside the poll loop costs more)
		while ((int)*(volatile unsigned *)&s_ndone < warp - kChainAdmit) __nanosleep(400);
		for (;;) {
			if (!done) {
				bool ready = true;
				unsigned long long pv[4] = {0, 0, 0, 0};
#pragma unroll
				for (int i = 0; i < 4; i++)
					if (in_chunk[i]) { pv[i] = ((volatile unsigned long long *)s_val)[lidx[i]]; ready = ready && (pv[i] >> 63); }
				if (ready) {
#pragma unroll
					for (int i = 0; i < 4; i++)
						if (in_chunk[i]) {
							val[i] = (int)(pv[i] & 0xffffull);
							col[i][0] = (unsigned)(pv[i] >> 16) & 0xffu; col[i][1] = (unsigned)(pv[i] >> 24) & 0xffu; col[i][2] = (unsigned)(pv[i] >> 32) & 0xffu;
						}
					int n = 0, sum = 0, sr = 0, sg = 0, sb = 0, prev = -1;
#pragma unroll
					for (int i = 0; i < 8; i++) {
						if (val[i] > 0 && (prev == -1 || abs(val[i] - prev) < 30)) {       // depthprocessing.cpp:239
							prev = val[i]; n++; sum += val[i];
							sr += (int)col[i][0]; sg += (int)col[i][1]; sb += (int)col[i][2];
						}
					}
					unsigned long long r = 1ull << 63;                                   // stays 0 (:249)
					if (n > 4) {
						// x / n for 5 <= n <= 8 and x < 2^20 (eight depths / colour bytes) as a multiply: floor(x * ceil(2^32 / n) / 2^32) is exact there
						// (the error of the scaled reciprocal, < x / 2^32 < 2^-12, cannot carry a quotient with fractional part <= 7/8 over the next integer)
						const unsigned rcp = n == 5 ? 858993460u : n == 6 ? 715827883u : n == 7 ? 613566757u : 536870912u;
						const unsigned d = __umulhi((unsigned)sum, rcp), cr = __umulhi((unsigned)sr, rcp) & 0xffu, cg = __umulhi((unsigned)sg, rcp) & 0xffu, cbl = __umulhi((unsigned)sb, rcp) & 0xffu;
						r |= (unsigned long long)(d & 0xffffu) | ((unsigned long long)cr << 16) | ((unsigned long long)cg << 24) | ((unsigned long long)cbl << 32);
						fcolors[3 * gp] = (uint8_t)cr; fcolors[3 * gp + 1] = (uint8_t)cg; fcolors[3 * gp + 2] = (uint8_t)cbl;
						fdepth[gp] = (unsigned short)d;
					}
					((volatile unsigned long long *)s_val)[tid] = r;
					state[gp] = kRadHole | kRadDone;
					done = true;
				}
			}
			if (__all_sync(kFull, done)) break;
		}
		if (lane == 0) atomicAdd(&s_ndone, 1u);
		__syncthreads();
	}
}

// What the grid-wide round left pending, finished by FIXPOINT ITERATION instead of along the chains (round 2).  The fill of a hole is
// a function of its eight neighbours; the four raster-earlier ones may be pending holes themselves, which is what makes the reference's
// raster-order loop a recurrence.  But the recurrence is over a DAG (a hole only reads raster-earlier holes), so the assignment "every
// pending hole holds exactly what the fill rule gives for the values its neighbours hold" has ONE solution — the reference's result —
// and any iteration that ends in a full pass without a change has found it.  So: every pending hole is evaluated at once from the
// tentative values (unfilled = 0 to start with), again and again until a pass changes nothing.  A fill is an average of five to
// eight values, so a wrong guess upstream reaches the next link divided by n >= 5 and dies out within a few links: the thin hole curves
// of a warped Kinect frame (chains of ~h links, 0.43 ms in k_rad_chains at ~0.3 us per link) settle in 15-30 passes.  Only a chain
// whose fill DECISIONS hang on one another (each link has exactly four valid fixed neighbours) needs one pass per link, which is what
// the chain kernel costs anyway.  After the first pass a hole is re-evaluated only if a raster-earlier neighbour changed in the
// previous pass (two alternating "changed" bits beside the value); it may read a value another thread rewrites in the same pass —
// harmless: whoever changes flags itself, and its readers run again next pass, after the barrier.
//
// A pass must be CHEAP (there are tens of them, one after the other): a cluster of eight 1024-thread blocks per sensor, each owning
// the pending holes of an eighth of the image, at most two per thread.  A hole's tentative value is one 64-bit word in its owner's
// shared memory (depth | r | g | b | changed bits: one atomic word, no torn colours); what never changes during the kernel — the
// four raster-later neighbours, the raster-earlier ones that are not pending — is fetched once, the former parked in shared memory,
// the latter in registers together with the references (block, index) to the pending ones.  So a pass is at most four shared-memory
// reads (distributed shared memory for a neighbour in the block before), ~100 instructions and a __syncthreads.  Passes are LOCAL to
// a block until it sees no change; only then does the cluster meet and the blocks learn from each other's flags whether anyone
// changed since the last meeting.  The blocks count their passes independently, so a hole with a neighbour in another block cannot
// use that neighbour's "changed" bits: it is simply re-evaluated in the first pass after every meeting.  The last meeting follows a
// round in which no block changed anything: every such hole was evaluated in it from values that no longer moved.
// A sensor with more than 2048 pending holes in one eighth (never seen on warped frames) is left to k_rad_chains (bail flag).
constexpr int kFixThreads = 1024, kFixCluster = 8, kFixPerThread = 2, kFixCap = kFixThreads * kFixPerThread, kFixRowCap = 1024;
constexpr unsigned long long kFixRef = 1ull << 63;

__device__ __forceinline__ unsigned long long rad_record(const unsigned short *fdepth, const uint8_t *fcolors, long long q) {
	const uint8_t *c = fcolors + 3 * q;
	return (unsigned long long)fdepth[q] | ((unsigned long long)c[0] << 16) | ((unsigned long long)c[1] << 24) | ((unsigned long long)c[2] << 32);
}

__global__ void __cluster_dims__(kFixCluster, 1, 1) __launch_bounds__(kFixThreads) k_rad_fixpoint(const PreSensor *__restrict__ sd, unsigned short *fdepth, uint8_t *fcolors,
	unsigned char *state, int *__restrict__ list, int *pidx, int *err, int *bail)
{
	pdl_enter();
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	extern __shared__ __align__(16) unsigned long long s_fix[];   // 80 KB, opt-in
	unsigned long long *s_T = s_fix;                              // [kFixCap] tentative value of this block's pending hole e
	unsigned long long (*s_L)[4] = reinterpret_cast<unsigned long long (*)[4]>(s_fix + kFixCap);      // [kFixCap][4] its four raster-later neighbours (immutable)
	unsigned char (*s_dirty)[kFixCap] = reinterpret_cast<unsigned char (*)[kFixCap]>(s_fix + 5 * kFixCap);    // [2][kFixCap] "a neighbour you read changed", by pass parity
	__shared__ unsigned s_w[32];
	__shared__ unsigned s_total;
	__shared__ int s_flag[2], s_over;
	const unsigned rank = cluster.block_rank();
	const int sensor = blockIdx.x / kFixCluster;
	const PreSensor s = sd[sensor];
	const int px = s.w * s.h, w = s.w;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	// ---- bands of whole rows with about the same number of pending holes each (the forward warp leaves most of its holes near the
	// image border: equal row counts gave the first and last block four times the holes of the middle ones) ----
	__shared__ unsigned s_rows[kFixRowCap];                    // pending holes per row of this block's EQUAL share of the rows
	__shared__ unsigned s_pre[kFixThreads + 1];
	__shared__ int s_bound[kFixCluster + 1];
	__shared__ unsigned s_carry;
	const int h = s.h;
	const int rows_per = (h + kFixCluster - 1) / kFixCluster;
	if (rows_per > kFixRowCap) {                                // > 8192 rows: not an image this path is for
		if (rank == 0 && tid == 0) bail[sensor] = 1;
		return;                                                 // uniform over the cluster: nobody waits at a barrier
	}
	for (int j = warp; j < rows_per; j += kFixThreads / 32) {
		const int y = (int)rank * rows_per + j;
		unsigned c = 0;
		if (y < h) {
			const unsigned char *rp = state + s.pix_begin + (long long)y * w;
			if ((w & 3) == 0 && (((uintptr_t)rp) & 3) == 0) {
				// four flags per load: a pending pixel has bit 0 (kRadDone) clear
				for (int x = 4 * lane; x < w; x += 128) c += 4u - (unsigned)__popc(*reinterpret_cast<const unsigned *>(rp + x) & 0x01010101u);
			} else {
				for (int x = lane; x < w; x += 32) c += (rp[x] & kRadDone) ? 0u : 1u;
			}
		}
		c = warp_sum(c);
		if (lane == 0) s_rows[j] = c;
	}
	if (tid <= kFixCluster) s_bound[tid] = tid == 0 ? 0 : h;
	if (tid == 0) s_carry = 0;
	cluster.sync();
	{
		// exclusive prefix over all h rows (every block computes the same), 1024 rows per turn
		unsigned total = 0;
		for (int y = tid; y < h; y += kFixThreads) total += *cluster.map_shared_rank(&s_rows[y % rows_per], (unsigned)(y / rows_per));
		total = warp_sum(total);
		if (lane == 0) atomicAdd(&s_carry, total);
		__syncthreads();
		total = s_carry;
		__syncthreads();
		unsigned carry = 0;
		for (int base = 0; base < h; base += kFixThreads) {
			const int y = base + tid;
			const unsigned c = y < h ? *cluster.map_shared_rank(&s_rows[y % rows_per], (unsigned)(y / rows_per)) : 0u;
			const unsigned incl = warp_incl_scan(c, lane);
			if (lane == 31) s_w[warp] = incl;
			__syncthreads();
			if (warp == 0) {
				const unsigned v = s_w[lane], sc = warp_incl_scan(v, lane);
				s_w[lane] = sc - v;
				if (lane == 31) s_total = sc;
			}
			__syncthreads();
			const unsigned excl = carry + s_w[warp] + incl - c;
			s_pre[tid] = excl;                                   // s_pre[kFixThreads]: the previous turn's last row (not used for y = 0)
			__syncthreads();
			if (y < h) {
				const unsigned before = tid ? s_pre[tid - 1] : s_pre[kFixThreads];
#pragma unroll
				for (int r = 1; r < kFixCluster; r++) {
					const unsigned target = (unsigned)((unsigned long long)total * r / kFixCluster);
					if (excl >= target && (y == 0 || before < target)) s_bound[r] = y;       // first row whose prefix reaches the target
				}
			}
			carry += s_total;
			__syncthreads();
			if (tid == kFixThreads - 1) s_pre[kFixThreads] = excl;
		}
	}
	__syncthreads();
	auto range_lo = [&](unsigned r) { return s_bound[r] * w; };
	// this block's pixels [r0, r1) and its slice of the list buffer
	const int r0 = range_lo(rank), r1 = range_lo(rank + 1);
	int *lst = list + s.pix_begin + r0;
	unsigned n_pending = 0;
	for (int base = r0; base < r1; base += kFixThreads * 16) {
		const int p0 = base + tid * 16;
		unsigned m = 0;
		const unsigned char *sp = state + s.pix_begin + p0;
		if (p0 + 16 <= r1 && (((uintptr_t)sp) & 15) == 0) {
			// 16 flags in one load; bit 0 of every byte is kRadDone
			const uint4 f = *reinterpret_cast<const uint4 *>(sp);
			const unsigned wd[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
			for (int j = 0; j < 16; j++)
				if (!((wd[j >> 2] >> (8 * (j & 3))) & kRadDone)) m |= 1u << j;
		} else {
#pragma unroll
			for (int j = 0; j < 16; j++)
				if (p0 + j < r1 && !(sp[j] & kRadDone)) m |= 1u << j;
		}
		const unsigned cnt = __popc(m), incl = warp_incl_scan(cnt, lane);
		if (lane == 31) s_w[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			const unsigned v = s_w[lane], sc = warp_incl_scan(v, lane);
			s_w[lane] = sc - v;
			if (lane == 31) s_total = sc;
		}
		__syncthreads();
		unsigned o = n_pending + s_w[warp] + incl - cnt;
		while (m) {
			const int j = __ffs(m) - 1;
			m &= m - 1;
			lst[o] = p0 + j;
			pidx[s.pix_begin + p0 + j] = (int)o;
			o++;
		}
		n_pending += s_total;
		__syncthreads();
	}
	if (tid == 0) s_over = n_pending > (unsigned)kFixCap ? 1 : 0;
	__threadfence();                                          // the index map before the meeting (other blocks read it through L2)
	cluster.sync();
	{
		int over = 0;
#pragma unroll
		for (int r = 0; r < kFixCluster; r++) over |= *cluster.map_shared_rank(&s_over, r);
		if (over) {
			if (rank == 0 && tid == 0) bail[sensor] = 1;
			cluster.sync();                                   // nobody leaves while its flag may still be read
			return;
		}
	}
	const int nb[8] = {-w - 1, -w, -w + 1, -1, 1, w - 1, w, w + 1};      // depthprocessing.cpp:226
	// ---- per hole: what cannot change, and where the pending neighbours live ----
	unsigned long long E[kFixPerThread][4];                   // raster-earlier neighbours: a fixed record, or kFixRef | owner << 16 | index
	unsigned short succ[kFixPerThread][4];                    // the raster-later neighbours that are pending holes of THIS block (0xffff: none): who to wake when this hole changes
	bool remote[kFixPerThread];
#pragma unroll
	for (int j = 0; j < kFixPerThread; j++) {
		const unsigned e = tid + j * kFixThreads;
		remote[j] = false;
#pragma unroll
		for (int i = 0; i < 4; i++) { E[j][i] = 0; succ[j][i] = 0xffffu; }
		if (e < kFixCap) { s_dirty[0][e] = 0; s_dirty[1][e] = 0; }
		if (e < n_pending) {
			const int p = lst[e];
			const long long gp = s.pix_begin + p;
#pragma unroll
			for (int i = 0; i < 4; i++) {
				const long long q = gp + nb[i];
				const unsigned st = state[q];
				if ((st & kRadHole) && !(st & kRadDone)) {
					unsigned owner = rank;
					while (p + nb[i] < range_lo(owner)) owner--;
					E[j][i] = kFixRef | ((unsigned long long)owner << 16) | (unsigned long long)(unsigned)ld_volatile_u32(reinterpret_cast<const unsigned *>(pidx + q));
					remote[j] = remote[j] || owner != rank;
				} else {
					E[j][i] = rad_record(fdepth, fcolors, q);         // final before this kernel started
				}
			}
#pragma unroll
			for (int i = 4; i < 8; i++) {
				const long long q = gp + nb[i];
				// raster-later: the warped value, which for an original hole is 0 whatever has been filled into it since
				const unsigned st = state[q];
				s_L[e][i - 4] = (st & kRadHole) ? 0ull : rad_record(fdepth, fcolors, q);
				if ((st & kRadHole) && !(st & kRadDone) && p + nb[i] < r1) succ[j][i - 4] = (unsigned short)ld_volatile_u32(reinterpret_cast<const unsigned *>(pidx + q));
			}
			s_T[e] = 0ull;                                        // unfilled: depth 0, colour 0 (:249)
		}
	}
	cluster.sync();                                           // every block's tentative values exist before anyone reads them
	unsigned pass = 0;
	bool converged = false;
	for (int round = 0; round <= px + 1; round++) {
		int any_round = 0;
		for (bool first = true;; first = false) {
			const unsigned cur = pass & 1u;
			int changed = 0;
#pragma unroll
			for (int j = 0; j < kFixPerThread; j++) {
				const unsigned e = tid + j * kFixThreads;
				if (e >= n_pending) continue;
				// evaluated when a hole it reads (in this block) changed in the previous pass; in the very first pass; and, for a hole
				// with a neighbour in the block before, in the first pass after every meeting
				const bool woken = ((volatile unsigned char *)s_dirty[cur])[e] != 0;
				if (!(woken || pass == 0 || (first && remote[j]))) continue;
				if (woken) s_dirty[cur][e] = 0;
				unsigned long long rec[4];
#pragma unroll
				for (int i = 0; i < 4; i++) {
					rec[i] = E[j][i];
					if (rec[i] & kFixRef) {
						const unsigned owner = (unsigned)(rec[i] >> 16) & 0xffu, idx = (unsigned)rec[i] & 0xffffu;
						rec[i] = owner == rank ? ((volatile unsigned long long *)s_T)[idx] : *(volatile unsigned long long *)cluster.map_shared_rank(&s_T[idx], owner);
					}
				}
				int n = 0, sum = 0, sr = 0, sg = 0, sb = 0, prev = -1;
#pragma unroll
				for (int i = 0; i < 8; i++) {
					const unsigned long long r = i < 4 ? rec[i] : s_L[e][i - 4];
					const int v = (int)(r & 0xffffull);
					if (v > 0 && (prev == -1 || abs(v - prev) < 30)) {                 // depthprocessing.cpp:239
						prev = v; n++; sum += v;
						sr += (int)((r >> 16) & 0xffull); sg += (int)((r >> 24) & 0xffull); sb += (int)((r >> 32) & 0xffull);
					}
				}
				unsigned long long out = 0ull;                                         // n <= 4: stays as warped, depth 0 and colour 0 (:249)
				if (n > 4) {
					// x / n for 5 <= n <= 8 and x < 2^20 as a multiply, exact there (see k_rad_chains)
					const unsigned rcp = n == 5 ? 858993460u : n == 6 ? 715827883u : n == 7 ? 613566757u : 536870912u;
					out = (unsigned long long)(__umulhi((unsigned)sum, rcp) & 0xffffu) | ((unsigned long long)(__umulhi((unsigned)sr, rcp) & 0xffu) << 16) |
						((unsigned long long)(__umulhi((unsigned)sg, rcp) & 0xffu) << 24) | ((unsigned long long)(__umulhi((unsigned)sb, rcp) & 0xffu) << 32);
				}
				if (out != ((volatile unsigned long long *)s_T)[e]) {
					((volatile unsigned long long *)s_T)[e] = out;
					changed = 1;
#pragma unroll
					for (int i = 0; i < 4; i++)
						if (succ[j][i] != 0xffffu) s_dirty[cur ^ 1u][succ[j][i]] = 1;       // its readers run in the next pass
				}
			}
			pass++;
			const int any = __syncthreads_or(changed);
			any_round |= any;
			if (!any) break;
		}
		if (tid == 0) s_flag[round & 1] = any_round;
		cluster.sync();
		int all = 0;
#pragma unroll
		for (int r = 0; r < kFixCluster; r++) all |= *cluster.map_shared_rank(&s_flag[round & 1], r);
		if (!all) { converged = true; break; }
	}
	if (!converged && tid == 0) atomicOr(err, 64);
	// ---- the settled values into the image ----
#pragma unroll
	for (int j = 0; j < kFixPerThread; j++) {
		const unsigned e = tid + j * kFixThreads;
		if (e < n_pending) {
			const long long gp = s.pix_begin + lst[e];
			const unsigned long long v = s_T[e];
			if (v & 0xffffull) {
				fdepth[gp] = (unsigned short)(v & 0xffffull);
				fcolors[3 * gp] = (uint8_t)(v >> 16); fcolors[3 * gp + 1] = (uint8_t)(v >> 24); fcolors[3 * gp + 2] = (uint8_t)(v >> 32);
			}
			state[gp] = kRadHole | kRadDone;
		}
	}
	cluster.sync();                                           // nobody leaves while its shared memory may still be read
}

__global__ void __launch_bounds__(256) k_rad_writeback(uint8_t *__restrict__ depth, uint8_t *__restrict__ colors, long long total_px,
	const unsigned short *__restrict__ fdepth, const uint8_t *__restrict__ fcolors)
{
	pdl_enter();
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
	cudaEvent_t ev_staged = nullptr;     // recorded after the descriptor upload out of pin_sd; waited for before pin_sd is rewritten
	cudaEvent_t ev_done = nullptr;       // end of the previous correction: the scratch below is shared by every caller / stream
	int *pin_err = nullptr;
	int sm_count = 148;
	std::vector<PreSensor> last_sd;     // what the device copy of the descriptors holds (empty: nothing uploaded yet)
};

PreCtx *g_pre = nullptr;

PreCtx *pre_ctx(int n_maps, const int *widths, const int *heights) {
	if (g_pre && (int)g_pre->w.size() == n_maps && !memcmp(g_pre->w.data(), widths, sizeof(int) * n_maps) && !memcmp(g_pre->h.data(), heights, sizeof(int) * n_maps)) return g_pre;
	if (g_pre) {
		DevBuf *bufs[] = {&g_pre->sd, &g_pre->winner, &g_pre->fdepth, &g_pre->fcolors, &g_pre->state, &g_pre->items, &g_pre->count, &g_pre->in_depth, &g_pre->in_colors, &g_pre->tmp};
		for (DevBuf *b : bufs) b->release();
		if (g_pre->pin_sd) cudaFreeHost(g_pre->pin_sd);
		if (g_pre->ev_staged) cudaEventDestroy(g_pre->ev_staged);
		if (g_pre->ev_done) cudaEventDestroy(g_pre->ev_done);
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
		c->fcolors.reserve(3 * n, "alloc warped colours") && c->state.reserve(n, "alloc hole states") && c->items.reserve(4 * n, "alloc pending lists") &&
		c->count.reserve(8 * (size_t)n_maps + 8, "alloc worklist counts");
	ok = ok && cuda_ok(cudaHostAlloc((void **)&c->pin_sd, sizeof(PreSensor) * n_maps, cudaHostAllocDefault), "alloc pinned descriptors");
	ok = ok && cuda_ok(cudaHostAlloc((void **)&c->pin_err, 64, cudaHostAllocDefault), "alloc pinned status");
	ok = ok && cuda_ok(cudaEventCreateWithFlags(&c->ev_staged, cudaEventDisableTiming), "create staging event") &&
		cuda_ok(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming), "create completion event");
	if (!ok) { delete c; return nullptr; }
	g_pre = c;
	return c;
}

// enqueue the whole correction on st for device-resident packed buffers (in place); err word = count[n_maps]
int radial_enqueue(PreCtx *c, int n_maps, uint8_t *d_depth, uint8_t *d_colors, const float *intr_params, cudaStream_t st) {
	static const bool rad_pdl = getenv("LS3D_RADIAL_PDL") ? atoi(getenv("LS3D_RADIAL_PDL")) != 0 : true;
	long long acc = 0;
	int max_px = 1;
	// the scratch buffers are one set per process: a correction enqueued on ANOTHER stream must not start before the previous one
	// has finished
	if (!cuda_ok(cudaStreamWaitEvent(st, c->ev_done, 0), "order after the previous correction")) return -1;
	std::vector<PreSensor> now((size_t)n_maps);
	for (int i = 0; i < n_maps; i++) {
		PreSensor &s = now[i];
		memset(&s, 0, sizeof(s));
		s.w = c->w[i]; s.h = c->h[i]; s.pix_begin = acc;
		const float *ip = intr_params + 7 * i;
		s.cx = ip[0]; s.cy = ip[1]; s.fx = ip[2]; s.fy = ip[3]; s.r2 = ip[4]; s.r4 = ip[5]; s.r6 = ip[6];
		acc += (long long)s.w * s.h;
		max_px = std::max(max_px, s.w * s.h);
	}
	bool ok = true;
	if (c->last_sd.size() != now.size() || memcmp(c->last_sd.data(), now.data(), sizeof(PreSensor) * now.size()) != 0) {
		// new intrinsics (normally once per rig): the previous upload must have left the pinned block before it is rewritten
		c->last_sd.clear();
		if (!cuda_ok(cudaEventSynchronize(c->ev_staged), "wait for the previous descriptor upload")) return -1;
		memcpy(c->pin_sd, now.data(), sizeof(PreSensor) * now.size());
		ok = cuda_ok(cudaMemcpyAsync(c->sd.p, c->pin_sd, sizeof(PreSensor) * n_maps, cudaMemcpyHostToDevice, st), "upload descriptors") &&
			cuda_ok(cudaEventRecord(c->ev_staged, st), "record descriptor upload");
		if (ok) c->last_sd = now;
	}
	ok = ok && cuda_ok(cudaMemsetAsync(c->winner.p, 0, 4 * (size_t)c->total_px, st), "clear winners") &&
		cuda_ok(cudaMemsetAsync(c->count.p, 0, 8 * (size_t)n_maps + 8, st), "clear worklist counts");
	if (!ok) return -1;
	const dim3 grid((unsigned)std::max(1, std::min((max_px + 255) / 256, c->sm_count * 8 / std::max(1, std::min(n_maps, 8)) + 1)), (unsigned)n_maps);
	const PreSensor *sd = c->sd.as<PreSensor>();
	int *count = c->count.as<int>();
	k_rad_scatter<<<grid, 256, 0, st>>>(d_depth, sd, c->winner.as<int>());
	launch_chain(rad_pdl, k_rad_gather, grid, 256, 0, st, d_depth, d_colors, sd, c->winner.as<int>(), c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>(), c->state.as<unsigned char>());
	static const int env_snapshot = getenv("LS3D_RADIAL_SNAPSHOT") ? atoi(getenv("LS3D_RADIAL_SNAPSHOT")) : 1;         // 0: the fenced round of rounds 1-2 (A/B)
	for (int r = 0; r < kRadRounds; r++) {
		if (env_snapshot && r == 0) launch_chain(rad_pdl, k_rad_first, dim3((unsigned)((max_px + 255) / 256), (unsigned)n_maps), 256, 0, st, sd, c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>(), c->state.as<unsigned char>());
		else launch_chain(rad_pdl, k_rad_round, grid, 256, 0, st, sd, c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>(), c->state.as<unsigned char>());
	}
	static const int env_wavefront = getenv("LS3D_RADIAL_WAVEFRONT") ? atoi(getenv("LS3D_RADIAL_WAVEFRONT")) : 0;      // 1: the round-1 lockstep wavefront (A/B)
	static const int env_chains = getenv("LS3D_RADIAL_CHAINS") ? atoi(getenv("LS3D_RADIAL_CHAINS")) : 0;               // 1: the chain kernel of round 1/2 instead of the fixpoint iteration (A/B)
	if (env_wavefront) {
		int max_h = 1, max_w = 1;
		for (int i = 0; i < n_maps; i++) { max_h = std::max(max_h, c->h[i]); max_w = std::max(max_w, c->w[i]); }
		const int words_per_row = (max_w + 31) / 32;
		int rows = std::min(kRadChainThreads, (max_h + 31) / 32 * 32);
		const int smem_budget = 200 * 1024;                           // of the 227 KB a block may use
		while (rows > 32 && (size_t)rows * words_per_row * 4 > (size_t)smem_budget) rows -= 32;
		const size_t wf_smem = (size_t)rows * words_per_row * 4;
		if (wf_smem > (size_t)smem_budget) { set_error("radial correction: image width %d too large for the wavefront's pending map", max_w); return -1; }
		if (!cuda_ok(cudaFuncSetAttribute(k_rad_wavefront, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_budget), "wavefront shared memory")) return -1;
		k_rad_wavefront<<<n_maps, rows, wf_smem, st>>>(sd, c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>(), c->state.as<unsigned char>(), words_per_row, count + n_maps);
	} else if (env_chains == 0) {
		// the winner map is free after the gather: it becomes the pixel -> list index map; count[n_maps + 1 + i] = sensor i was too dense
		constexpr size_t fix_smem = sizeof(unsigned long long) * kFixCap * 5 + 2 * kFixCap;
		static bool fix_attr = false;
		if (!fix_attr) {
			if (!cuda_ok(cudaFuncSetAttribute(k_rad_fixpoint, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fix_smem), "hole fill shared memory")) return -1;
			fix_attr = true;
		}
		launch_chain(rad_pdl, k_rad_fixpoint, dim3((unsigned)(n_maps * kFixCluster)), kFixThreads, fix_smem, st, sd, c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>(), c->state.as<unsigned char>(), c->items.as<int>(),
			c->winner.as<int>(), count + n_maps, count + n_maps + 1);
		launch_chain(rad_pdl, k_rad_chains, dim3((unsigned)n_maps), kChainThreads, 0, st, sd, c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>(), c->state.as<unsigned char>(), c->items.as<int>(), c->winner.as<int>(), count + n_maps,
			count + n_maps + 1);
	} else {
		// the winner map is free after the gather: it becomes the pixel -> list index map
		launch_chain(rad_pdl, k_rad_chains, dim3((unsigned)n_maps), kChainThreads, 0, st, sd, c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>(), c->state.as<unsigned char>(), c->items.as<int>(), c->winner.as<int>(), count + n_maps, nullptr);
	}
	launch_chain(rad_pdl, k_rad_writeback, dim3((unsigned)std::max<long long>(1, std::min<long long>((c->total_px + 255) / 256, (long long)c->sm_count * 8))), 256, 0, st, d_depth, d_colors, c->total_px,
		c->fdepth.as<unsigned short>(), c->fcolors.as<uint8_t>());
	count_launch(4 + kRadRounds);
	if (!cuda_ok(cudaEventRecord(c->ev_done, st), "record end of correction")) return -1;
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
	// The caller's arrays are written only after every kernel has finished and the device status word has been read back clean:
	// a failure anywhere before that leaves them untouched (ls3d.h).  The read-back itself then goes straight into them (staging
	// it through another page-locked block and a host copy cost 0.5 ms per call); should one of the two transfers fail — a lost
	// device, nothing the arithmetic can cause — the error text says so and the buffers must be considered undefined.
	ok = cuda_ok(cudaMemcpyAsync(c->pin_err, c->count.as<int>() + n_maps, sizeof(int), cudaMemcpyDeviceToHost, st), "read status") &&
		cuda_ok(cudaStreamSynchronize(st), "radial correction");
	if (!ok) return;
	if (*c->pin_err) { set_error("radial correction: device status flags 0x%x", *c->pin_err); return; }
	ok = cuda_ok(cudaMemcpyAsync(depth_maps, c->in_depth.p, 2 * n, cudaMemcpyDeviceToHost, st), "read depth") &&
		cuda_ok(cudaMemcpyAsync(depth_colors, c->in_colors.p, 3 * n, cudaMemcpyDeviceToHost, st), "read colours");
	ok = cuda_ok(cudaStreamSynchronize(st), "radial correction read-back") && ok;
	if (!ok) set_error("radial correction: the read-back into the caller's buffers failed (%s); their contents are undefined", ls3d_last_error());
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
