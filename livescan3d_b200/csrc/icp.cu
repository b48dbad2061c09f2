// icp.cu — the NativeUtils ICP refinement loop on the device, plus its host-buffer entry points.
//
// Reference behaviour reproduced (paths relative to the LiveScan3D tree):
//   ICP                       src/NativeUtils/icp.cpp:75-177
//   FindClosestPointForEach   src/NativeUtils/icp.cpp:18-32   (nanoflann kd-tree, exact 1-NN, include/nanoflann.h)
//   one-to-one dedupe         src/NativeUtils/icp.cpp:95-126  (smaller d2 wins, later source index wins ties)
//   GetStandardDeviation / RejectOutlierMatches   src/NativeUtils/icp.cpp:34-73  (sigma of SQUARED distances, keep d2 <= 2.5 sigma)
//   Kabsch block              src/NativeUtils/icp.cpp:138-168 (T = mean(p-q); M = sum (q+T) p^T about the origin; Rk = U V^T)
//
// Device design:
//   * the target grid is built ONCE per call (the reference rebuilds its kd-tree every iteration): a dense
//     G^3 grid in Morton order.  Morton order makes every octree node a contiguous range of the cell-start
//     array, so node occupancy is start[end]-start[begin] and no pyramid has to be stored.
//   * k_icp_match: exact nearest neighbour = home cell, then the 26 neighbours whose box can still hold a closer
//     point, then — only if the 3x3x3 block cannot prove the answer — a near-first depth-first walk of the
//     implicit octree with the current best as the pruning bound (exact for arbitrarily distant points).
//     One-to-one dedupe is a 64-bit atomicMin of (bits(d2) << 32 | ~i) per target point.
//   * k_icp_stats / k_icp_sums: warp-shuffle + block reductions in fp64 with a fixed-order last-block final
//     pass (deterministic); the 3x3 SVD (one-sided Jacobi) runs on the device in the next kernel's prologue.
//   * iterations are stream-ordered launches (or one CUDA graph): no host round trip inside ICP.
#include "ls3d_common.cuh"
#include "ls3d_internal.h"
#include "../../include/ls3d.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ls3d {

constexpr unsigned long long kSlotEmpty = 0x7FFFFFFFFFFFFFFFull;   // positive as int64: MIN-reducible by NCCL/torch as signed
constexpr int kTraceCap = 64;

struct IcpGrid {
	float ox, oy, oz;     // origin (bbox min of the target)
	float h, inv_h;       // cell edge and its reciprocal
	int G, levels;        // cells per axis (power of two), log2(G)
	int pad;
};

struct IcpState {
	float R[9], t[3];       // accumulated pose (ls3d_icp_Rt points here)
	int iters_applied, err, pad0, pad1;
	unsigned ticket_stats, ticket_sums, n_work, blocks_done;   // n_work: packet counter of the match stage
	unsigned n_packets, n_heavy, n_light, sched_valid;         // packets of the current source ordering; longest-first schedule (k_icp_stats)
	float xf[12];           // the update the next match kernel applies: T[3] then Rk[9] (written by the solve step)
	unsigned hq_n, hq_head, hq_done, light_done;   // heavy packets of the current match stage: queued / handed out / blocks finished; the packet kernel has run dry
	unsigned cls_n[8], cls_fill[8];            // packets per cost class of the match stage just run / placed so far by the scheduler (k_icp_stats)
	unsigned red_bar, red_epoch, red_pad0, red_pad1;   // k_icp_reduce: arrivals at its grid barriers; launches so far (never reset: the cross-rank flags count on)
};

// cost class of a packet for the longest-first schedule: 0 = over budget (block-wide stage), then 7 classes of 8 visits, most expensive first
__device__ __forceinline__ unsigned pk_cost_class(unsigned cost, unsigned budget) { return cost > budget ? 0u : 7u - min(6u, cost >> 3); }

struct IcpBox { unsigned mn[3], mx[3]; };

__device__ __forceinline__ unsigned spread3(unsigned v) {     // 10 bits -> every third bit
	v &= 0x3ffu;
	v = (v | (v << 16)) & 0x030000FFu;
	v = (v | (v << 8)) & 0x0300F00Fu;
	v = (v | (v << 4)) & 0x030C30C3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}
__device__ __forceinline__ unsigned morton3(unsigned x, unsigned y, unsigned z) { return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2); }
__device__ __forceinline__ unsigned compact3(unsigned v) {   // inverse of spread3
	v &= 0x09249249u;
	v = (v ^ (v >> 2)) & 0x030C30C3u;
	v = (v ^ (v >> 4)) & 0x0300F00Fu;
	v = (v ^ (v >> 8)) & 0x030000FFu;
	v = (v ^ (v >> 16)) & 0x000003FFu;
	return v;
}

// Octree node record (one u64 per node of every level, and per occupied cell): the bounding box of the node's POINTS,
// quantised to 8 bits per bound inside the node's cube (rounded outward, so the decoded box always contains the
// points), plus the 8-bit child-occupancy mask.  bytes: 0..2 min xyz, 3..5 max xyz, 6 child mask.
// Tight boxes are what keeps far queries cheap: a wall tangent to the search ball passes the cube test in ~2d/h cells
// but the box test in a handful.
__device__ __forceinline__ unsigned long long box_encode(const float mn[3], const float mx[3], float ox, float oy, float oz, float size, unsigned mask) {
	const float sc = 255.0f / size;
	const float o[3] = {ox, oy, oz};
	unsigned long long r = (unsigned long long)mask << 48;
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const float lo = fminf(fmaxf(floorf((mn[a] - o[a]) * sc - 0.02f), 0.0f), 255.0f);
		const float hi = fminf(fmaxf(ceilf((mx[a] - o[a]) * sc + 0.02f), 0.0f), 255.0f);
		r |= (unsigned long long)(unsigned)lo << (8 * a);
		r |= (unsigned long long)(unsigned)hi << (8 * (3 + a));
	}
	return r;
}
__device__ __forceinline__ void box_decode(unsigned long long r, float ox, float oy, float oz, float size, float mn[3], float mx[3]) {
	const float q = size * (1.0f / 255.0f);
	mn[0] = ox + (float)(unsigned)(r & 0xff) * q;         mn[1] = oy + (float)(unsigned)((r >> 8) & 0xff) * q;  mn[2] = oz + (float)(unsigned)((r >> 16) & 0xff) * q;
	mx[0] = ox + (float)(unsigned)((r >> 24) & 0xff) * q; mx[1] = oy + (float)(unsigned)((r >> 32) & 0xff) * q; mx[2] = oz + (float)(unsigned)((r >> 40) & 0xff) * q;
}
// conservative squared distance from the origin-relative query to a decoded record box
__device__ __forceinline__ float rec_lb2(unsigned long long r, float rx, float ry, float rz, float ox, float oy, float oz, float size, float slack) {
	float mn[3], mx[3];
	box_decode(r, ox, oy, oz, size, mn, mx);
	const float dx = fmaxf(fmaxf(mn[0] - rx, rx - mx[0]) - slack, 0.0f);
	const float dy = fmaxf(fmaxf(mn[1] - ry, ry - mx[1]) - slack, 0.0f);
	const float dz = fmaxf(fmaxf(mn[2] - rz, rz - mx[2]) - slack, 0.0f);
	return (dx * dx + dy * dy + dz * dz) * 0.99999f;
}

// first entry of level l (1..L) in the child-mask array: sum_{j<l} (G >> j)^3 with G = 2^L
__device__ __host__ __forceinline__ unsigned icp_mask_off(int L, int l) { return ((1u << (3 * L)) - (1u << (3 * (L - l + 1)))) / 7u; }
// the same as a table for the search kernels (the level is warp-uniform: one constant-bank load instead of shifts and a division by 7)
constexpr unsigned icp_off_c(int L, int l) { return (l >= 1 && l <= L + 1) ? (unsigned)(((1ull << (3 * L)) - (1ull << (3 * (L - l + 1)))) / 7ull) : 0u; }
#define LS3D_OFF_ROW(L) {0u, icp_off_c(L, 1), icp_off_c(L, 2), icp_off_c(L, 3), icp_off_c(L, 4), icp_off_c(L, 5), icp_off_c(L, 6), icp_off_c(L, 7), icp_off_c(L, 8), icp_off_c(L, 9)}
__constant__ unsigned c_icp_off[10][10] = {LS3D_OFF_ROW(0), LS3D_OFF_ROW(1), LS3D_OFF_ROW(2), LS3D_OFF_ROW(3), LS3D_OFF_ROW(4), LS3D_OFF_ROW(5), LS3D_OFF_ROW(6), LS3D_OFF_ROW(7), LS3D_OFF_ROW(8), LS3D_OFF_ROW(9)};

__device__ __forceinline__ int icp_cell(float rel, float inv_h, int G) {
	float u = floorf(rel * inv_h);
	u = fminf(fmaxf(u, 0.0f), (float)(G - 1));
	return (int)u;
}

// ------------------------------------------------------------------------------------------------------
// target grid build
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_icp_bbox(const float *__restrict__ v, int n, IcpBox *box) {
	float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
		for (int a = 0; a < 3; a++) {
			const float x = v[3 * (size_t)i + a];
			if (isfinite(x)) { mn[a] = fminf(mn[a], x); mx[a] = fmaxf(mx[a], x); }
		}
	}
#pragma unroll
	for (int a = 0; a < 3; a++) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			mn[a] = fminf(mn[a], __shfl_xor_sync(kFull, mn[a], o));
			mx[a] = fmaxf(mx[a], __shfl_xor_sync(kFull, mx[a], o));
		}
	}
	// one set of atomics per block (one per warp was 40 k same-address atomics for a 213 k-point target)
	__shared__ unsigned s_mn[3], s_mx[3];
	if (threadIdx.x < 3) { s_mn[threadIdx.x] = 0xffffffffu; s_mx[threadIdx.x] = 0u; }
	__syncthreads();
	if ((threadIdx.x & 31) == 0) {
#pragma unroll
		for (int a = 0; a < 3; a++) { atomicMin(&s_mn[a], f2ord(mn[a])); atomicMax(&s_mx[a], f2ord(mx[a])); }
	}
	__syncthreads();
	if (threadIdx.x < 3) { atomicMin(&box->mn[threadIdx.x], s_mn[threadIdx.x]); atomicMax(&box->mx[threadIdx.x], s_mx[threadIdx.x]); }
}


__global__ void k_icp_grid_params(const IcpBox *box, IcpGrid *grid, int G, int levels) {
	double lo[3], ext = 0;
	for (int a = 0; a < 3; a++) {
		double l = (double)ord2f(box->mn[a]), h = (double)ord2f(box->mx[a]);
		if (!(l <= h)) { l = 0; h = 0; }
		lo[a] = l;
		if (h - l > ext) ext = h - l;
	}
	double h = ext / (double)G * 1.0001;
	if (!(h > 1e-30)) h = 1.0;
	grid->ox = (float)lo[0]; grid->oy = (float)lo[1]; grid->oz = (float)lo[2];
	grid->h = (float)h;
	grid->inv_h = (float)(1.0 / h);
	grid->G = G;
	grid->levels = levels;
}

// level 0: one record per cell, indexed by the cell's Morton code: the tight box of its points with bit 48 set ("occupied"),
// 0 for an empty cell
__global__ void __launch_bounds__(256) k_icp_cellbox(const IcpGrid *__restrict__ grid, const unsigned *__restrict__ cell_start, const float4 *__restrict__ sorted,
	unsigned long long *__restrict__ cellrec, unsigned ncells)
{
	const IcpGrid g = *grid;
	for (unsigned m = blockIdx.x * blockDim.x + threadIdx.x; m < ncells; m += gridDim.x * blockDim.x) {
		const unsigned s = cell_start[m], e = cell_start[m + 1];
		unsigned long long rec = 0;
		if (s != e) {
			float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
			for (unsigned p = s; p < e; p++) {
				const float4 c = sorted[p];
				const float r[3] = {c.x - g.ox, c.y - g.oy, c.z - g.oz};
#pragma unroll
				for (int a = 0; a < 3; a++) { mn[a] = fminf(mn[a], r[a]); mx[a] = fmaxf(mx[a], r[a]); }
			}
			const float nx = (float)compact3(m), ny = (float)compact3(m >> 1), nz = (float)compact3(m >> 2);
			rec = box_encode(mn, mx, nx * g.h, ny * g.h, nz * g.h, g.h, 1u);
		}
		cellrec[m] = rec;
	}
}

// one octree node of level l from its 8 children (cells for l == 1)
__device__ __forceinline__ void icp_make_node(const IcpGrid &g, int l, unsigned mp, const unsigned *__restrict__ cell_start,
	const unsigned long long *__restrict__ cellbox, unsigned long long *nodes)
{
	const int L = g.levels;
	const float csize = g.h * (float)(1u << (l - 1));
	float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
	unsigned mask = 0;
	for (unsigned c = 0; c < 8; c++) {
		const unsigned cmp = (mp << 3) | c;
		unsigned long long r;
		r = l == 1 ? cellbox[cmp] : nodes[icp_mask_off(L, l - 1) + cmp];
		if (((r >> 48) & 0xff) == 0) continue;        // empty cell / node
		mask |= 1u << c;
		float cmn[3], cmx[3];
		box_decode(r, (float)compact3(cmp) * csize, (float)compact3(cmp >> 1) * csize, (float)compact3(cmp >> 2) * csize, csize, cmn, cmx);
#pragma unroll
		for (int a = 0; a < 3; a++) { mn[a] = fminf(mn[a], cmn[a]); mx[a] = fmaxf(mx[a], cmx[a]); }
	}
	const float size = csize * 2.0f;
	unsigned long long rec = 0;
	if (mask) rec = box_encode(mn, mx, (float)compact3(mp) * size, (float)compact3(mp >> 1) * size, (float)compact3(mp >> 2) * size, size, mask);
	nodes[icp_mask_off(L, l) + mp] = rec;
}

__global__ void __launch_bounds__(256) k_icp_nodebox(const IcpGrid *__restrict__ grid, int l, const unsigned *__restrict__ cell_start,
	const unsigned long long *__restrict__ cellbox, unsigned long long *nodes)
{
	const IcpGrid g = *grid;
	const unsigned dim = (unsigned)g.G >> l, n = dim * dim * dim;
	for (unsigned mp = blockIdx.x * blockDim.x + threadIdx.x; mp < n; mp += gridDim.x * blockDim.x) icp_make_node(g, l, mp, cell_start, cellbox, nodes);
}

// levels l_first..L in one block (at most 512 nodes per level), level by level
__global__ void __launch_bounds__(512) k_icp_nodebox_top(const IcpGrid *__restrict__ grid, int l_first, const unsigned *__restrict__ cell_start,
	const unsigned long long *__restrict__ cellbox, unsigned long long *nodes)
{
	const IcpGrid g = *grid;
	for (int l = l_first; l <= g.levels; l++) {
		const unsigned dim = (unsigned)g.G >> l, n = dim * dim * dim;
		for (unsigned mp = threadIdx.x; mp < n; mp += blockDim.x) icp_make_node(g, l, mp, cell_start, cellbox, nodes);
		__threadfence_block();
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256) k_icp_count(const float *__restrict__ v, int n, const IcpGrid *__restrict__ grid,
	unsigned *cell_count, unsigned *__restrict__ cell_of, unsigned *__restrict__ rank_of)
{
	const IcpGrid g = *grid;
	const int lane = threadIdx.x & 31;
	for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += gridDim.x * blockDim.x) {
		const int i = i0 + lane;
		const bool active = i < n;
		unsigned m = 0xffffffffu - (unsigned)lane;      // unique tags for idle lanes
		if (active) {
			const float x = v[3 * (size_t)i], y = v[3 * (size_t)i + 1], z = v[3 * (size_t)i + 2];
			m = morton3(icp_cell(x - g.ox, g.inv_h, g.G), icp_cell(y - g.oy, g.inv_h, g.G), icp_cell(z - g.oz, g.inv_h, g.G));
		}
		const unsigned grp = __match_any_sync(kFull, m);
		const int leader = __ffs(grp) - 1;
		unsigned base = 0;
		if (active && lane == leader) base = atomicAdd(&cell_count[m], (unsigned)__popc(grp));
		base = __shfl_sync(kFull, base, leader);
		if (active) { cell_of[i] = m; rank_of[i] = base + __popc(grp & ((1u << lane) - 1u)); }
	}
}

// in-place exclusive scan of n unsigned values (single pass, decoupled look-back)
__global__ void __launch_bounds__(kScanThreads) k_exclusive_scan(unsigned *data, int n, unsigned *tile_counter, unsigned long long *status, int *err) {
	__shared__ unsigned sm[16];
	__shared__ int s_tile;
	const int ntiles = (n + kTile - 1) / kTile;
	const int tid = threadIdx.x;
	for (;;) {
		if (tid == 0) s_tile = (int)atomicAdd(tile_counter, 1u);
		__syncthreads();
		const int tile = s_tile;
		if (tile >= ntiles) break;
		const int i0 = tile * kTile + tid * 8;
		unsigned v[8];
		if (i0 + 8 <= n) {
			const uint4 a = *reinterpret_cast<const uint4 *>(data + i0), b = *reinterpret_cast<const uint4 *>(data + i0 + 4);
			v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
		} else {
#pragma unroll
			for (int j = 0; j < 8; j++) v[j] = (i0 + j < n) ? data[i0 + j] : 0u;
		}
		unsigned cnt = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) cnt += v[j];
		unsigned total, base;
		const unsigned off = tile_scan(cnt, sm, status, tile, err, &total, &base);
		unsigned run = base + off;
#pragma unroll
		for (int j = 0; j < 8; j++) { const unsigned x = v[j]; v[j] = run; run += x; }
		if (i0 + 8 <= n) {
			*reinterpret_cast<uint4 *>(data + i0) = make_uint4(v[0], v[1], v[2], v[3]);
			*reinterpret_cast<uint4 *>(data + i0 + 4) = make_uint4(v[4], v[5], v[6], v[7]);
		} else {
#pragma unroll
			for (int j = 0; j < 8; j++) if (i0 + j < n) data[i0 + j] = v[j];
		}
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256) k_icp_scatter(const float *__restrict__ v, int n, const unsigned *__restrict__ cell_of,
	const unsigned *__restrict__ rank_of, const unsigned *__restrict__ cell_start, float4 *__restrict__ sorted)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const unsigned pos = cell_start[cell_of[i]] + rank_of[i];
		sorted[pos] = make_float4(v[3 * (size_t)i], v[3 * (size_t)i + 1], v[3 * (size_t)i + 2], __int_as_float(i));
	}
}

__global__ void __launch_bounds__(256) k_fill_u64(unsigned long long *p, int n, unsigned long long v) {
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

// ------------------------------------------------------------------------------------------------------
// 3x3 helpers (OpenCV 3.2 CV_32F semantics as the oracle restates them, oracle/ls3d_oracle.cpp svd3/mul33)
// ------------------------------------------------------------------------------------------------------
__device__ void mul33(const float *A, const float *B, float *D) {      // fp32, left to right, no FMA
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
			D[i * 3 + j] = __fadd_rn(__fadd_rn(__fmul_rn(A[i * 3 + 0], B[0 * 3 + j]), __fmul_rn(A[i * 3 + 1], B[1 * 3 + j])), __fmul_rn(A[i * 3 + 2], B[2 * 3 + j]));
}

__device__ float det33(const float *m) {
	const float a = __fsub_rn(__fmul_rn(m[4], m[8]), __fmul_rn(m[5], m[7]));
	const float b = __fsub_rn(__fmul_rn(m[3], m[8]), __fmul_rn(m[5], m[6]));
	const float c = __fsub_rn(__fmul_rn(m[3], m[7]), __fmul_rn(m[4], m[6]));
	return __fadd_rn(__fsub_rn(__fmul_rn(m[0], a), __fmul_rn(m[1], b)), __fmul_rn(m[2], c));
}

// One-sided (Hestenes) Jacobi SVD of a 3x3: fp32 storage, fp64 dot products, singular values descending.
// M = U diag(w) Vt.
__device__ void svd33(const float *M, float *U, float *Vt) {
	float A[3][3], V[3][3];
	double W[3];
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) { A[i][j] = M[j * 3 + i]; V[i][j] = (i == j) ? 1.0f : 0.0f; }     // rows of A = columns of M
	for (int i = 0; i < 3; i++) { double s = 0; for (int k = 0; k < 3; k++) s += (double)A[i][k] * (double)A[i][k]; W[i] = s; }
	const double eps = 2.0 * 1.1920929e-07;
	for (int sweep = 0; sweep < 30; sweep++) {
		bool changed = false;
		for (int i = 0; i < 2; i++)
			for (int j = i + 1; j < 3; j++) {
				double a = W[i], b = W[j], p = 0;
				for (int k = 0; k < 3; k++) p += (double)A[i][k] * (double)A[j][k];
				if (fabs(p) <= eps * sqrt(a * b)) continue;
				p *= 2;
				const double beta = a - b, gamma = hypot(p, beta);
				float c, s;
				if (beta < 0) { const double delta = (gamma - beta) * 0.5; s = (float)sqrt(delta / gamma); c = (float)(p / (gamma * s * 2)); }
				else { c = (float)sqrt((gamma + beta) / (gamma * 2)); s = (float)(p / (gamma * c * 2)); }
				a = b = 0;
				for (int k = 0; k < 3; k++) {
					const float t0 = __fadd_rn(__fmul_rn(c, A[i][k]), __fmul_rn(s, A[j][k]));
					const float t1 = __fadd_rn(__fmul_rn(-s, A[i][k]), __fmul_rn(c, A[j][k]));
					A[i][k] = t0; A[j][k] = t1;
					a += (double)t0 * (double)t0; b += (double)t1 * (double)t1;
				}
				W[i] = a; W[j] = b;
				changed = true;
				for (int k = 0; k < 3; k++) {
					const float t0 = __fadd_rn(__fmul_rn(c, V[i][k]), __fmul_rn(s, V[j][k]));
					const float t1 = __fadd_rn(__fmul_rn(-s, V[i][k]), __fmul_rn(c, V[j][k]));
					V[i][k] = t0; V[j][k] = t1;
				}
			}
		if (!changed) break;
	}
	for (int i = 0; i < 3; i++) { double s = 0; for (int k = 0; k < 3; k++) s += (double)A[i][k] * (double)A[i][k]; W[i] = sqrt(s); }
	for (int i = 0; i < 2; i++) {
		int j = i;
		for (int k = i + 1; k < 3; k++) if (W[j] < W[k]) j = k;
		if (i != j) {
			const double tw = W[i]; W[i] = W[j]; W[j] = tw;
			for (int k = 0; k < 3; k++) { float t = A[i][k]; A[i][k] = A[j][k]; A[j][k] = t; t = V[i][k]; V[i][k] = V[j][k]; V[j][k] = t; }
		}
	}
	for (int i = 0; i < 3; i++) {
		const float inv = W[i] > 0 ? (float)(1.0 / W[i]) : 0.0f;
		for (int k = 0; k < 3; k++) { U[k * 3 + i] = __fmul_rn(A[i][k], inv); Vt[i * 3 + k] = V[i][k]; }
	}
}

// (T, Rk) of the iteration whose 16 correspondence sums are in `sums`; executed by one thread per block
// (every block derives the same values from the same inputs).  Returns false when there were no matches.
__device__ bool icp_solve(const double *sums, float *T, float *Rk) {
	const double m = sums[0];
	if (!(m >= 0.5)) {
		T[0] = T[1] = T[2] = 0.0f;
		for (int i = 0; i < 9; i++) Rk[i] = (i % 4 == 0) ? 1.0f : 0.0f;
		return false;
	}
	const float scale = (float)(1.0 / m);                       // cv::reduce AVG: column sum * (1/rows)
	for (int a = 0; a < 3; a++) T[a] = __fmul_rn((float)sums[1 + a], scale);
	float M[9];
	for (int a = 0; a < 3; a++)
		for (int b = 0; b < 3; b++) M[a * 3 + b] = (float)(sums[7 + 3 * a + b] + (double)T[a] * sums[4 + b]);   // sum (q+T)_a p_b
	float U[9], Vt[9];
	svd33(M, U, Vt);
	mul33(U, Vt, Rk);
	if ((double)det33(Rk) < 0) {                                 // icp.cpp:156-163
		const float D[9] = {1, 0, 0, 0, 1, 0, 0, 0, -1};
		float UD[9];
		mul33(U, D, UD);
		mul33(UD, Vt, Rk);
	}
	return true;
}

// pose accumulation of icp.cpp:167-168, by exactly one thread per iteration
__device__ void icp_accumulate(IcpState *st, const float *T, const float *Rk, bool solved, Ls3dIcpTrace *trace, int trace_idx, const double *sums) {
	for (int j = 0; j < 3; j++) {
		double s = 0;
		for (int k = 0; k < 3; k++) s += (double)T[k] * (double)st->R[j * 3 + k];     // tempT * matR^T: general gemm path, fp64 accumulate
		st->t[j] = __fadd_rn(st->t[j], (float)s);
	}
	float nr[9];
	mul33(st->R, Rk, nr);
	for (int i = 0; i < 9; i++) st->R[i] = nr[i];
	st->iters_applied += 1;
	if (!solved) st->err |= kErrNoMatches;
	if (trace && trace_idx >= 0 && trace_idx < kTraceCap) {
		trace[trace_idx].n_accepted = (int)(sums[0] + 0.5);
		for (int a = 0; a < 3; a++) trace[trace_idx].T[a] = T[a];
		for (int a = 0; a < 9; a++) trace[trace_idx].Rk[a] = Rk[a];
	}
}

// ------------------------------------------------------------------------------------------------------
// nearest neighbour
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned octant_permute(unsigned m, unsigned near) {      // bit c -> bit (c ^ near)
	if (near & 1u) m = ((m & 0xAAu) >> 1) | ((m & 0x55u) << 1);
	if (near & 2u) m = ((m & 0xCCu) >> 2) | ((m & 0x33u) << 2);
	if (near & 4u) m = ((m & 0xF0u) >> 4) | ((m & 0x0Fu) << 4);
	return m;
}

// apply (T, Rk) the way icp.cpp:143-165 does: fp32 add, then row-vector times matrix, left to right
__device__ __forceinline__ void apply_xform(float &x, float &y, float &z, const float *T, const float *Rk) {
	const float a0 = __fadd_rn(x, T[0]), a1 = __fadd_rn(y, T[1]), a2 = __fadd_rn(z, T[2]);
	x = __fadd_rn(__fadd_rn(__fmul_rn(a0, Rk[0]), __fmul_rn(a1, Rk[3])), __fmul_rn(a2, Rk[6]));
	y = __fadd_rn(__fadd_rn(__fmul_rn(a0, Rk[1]), __fmul_rn(a1, Rk[4])), __fmul_rn(a2, Rk[7]));
	z = __fadd_rn(__fadd_rn(__fmul_rn(a0, Rk[2]), __fmul_rn(a1, Rk[5])), __fmul_rn(a2, Rk[8]));
}

// Where the dedupe slot of a target point lives: with several ranks the slot array is sharded by target chunk (k_icp_reduce's
// chunks: `chunk` points each, `per` chunks per rank) and ptr[r] is rank r's array, mapped into this process (NVLink).
struct SlotMap { int world, chunk, per, pad; unsigned long long *ptr[8]; };

__device__ __forceinline__ void nn_commit(int i, float d2, int idx, const SlotMap &sm, int *__restrict__ nn_idx, float *__restrict__ nn_d2) {
	nn_idx[i] = idx;
	nn_d2[i] = idx >= 0 ? d2 : 0.0f;
	if (idx >= 0) {
		// one-to-one dedupe (icp.cpp:95-126): smallest d2 wins the target point, the LATER source index wins ties
		const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
		unsigned long long *slots = sm.ptr[0];
		if (sm.world > 1) slots = sm.ptr[min(sm.world - 1, (idx / sm.chunk) / sm.per)];      // the owner's array: a 64-bit atomic over NVLink
		atomicMin(&slots[idx], key);      // remote ones are ordered system-wide by the block's fence at the end of the match kernels
	}
}
// ------------------------------------------------------------------------------------------------------
// packet nearest neighbour: one warp = 32 spatially adjacent queries walking the octree together
// ------------------------------------------------------------------------------------------------------
// A thread-per-query walk spends most of its issue slots idle: neighbouring lanes take different paths of very
// different length (ncu, round 1: 9-12 of 32 lanes active).  Here the source points are Morton-sorted once per call
// (icp_build_order), so a warp's 32 queries sit within about one grid cell of each other and need nearly the same nodes.
// The warp walks ONE path — the union of what its lanes need: every branch below is warp-uniform (ballots), the node
// records are loaded once per warp (lanes 0-7 fetch the 8 children of a node in one request and park them in shared
// memory, lanes 8-16 the cell ranges), candidate points are fetched 32 at a time by one coalesced request and broadcast
// from shared memory, and only the lower-bound arithmetic and the distance tests are per lane.
// Exactness: a node is skipped only when EVERY lane's conservative lower bound exceeds that lane's best; the climb stops only
// when every lane's ball fits inside the subtree already searched (distance to the subtree's cube faces; faces on the grid
// boundary have nothing behind them), so the result is exact at any distance like nanoflann's; ties resolve to the smallest
// target index.  `best` starts from last iteration's neighbour (a real candidate, so exactness is untouched).
constexpr int kPkWarps = 8;
#ifndef LS3D_PK_DONE
#define LS3D_PK_DONE 4
#endif
constexpr int kPkDone = LS3D_PK_DONE;             // home cells scanned up front (and skipped by the walk)

struct PkWarp {
	float4 lo[8][8], hi[8][8];      // [level t-1][child]: point boxes of the children of the node being iterated at level t, decoded
	                                // (origin-relative, already widened by the slack): lane 8a+c decodes axis a of child c, every lane reads all
	float px[32], py[32], pz[32];   // staged candidate points, one plane per coordinate: an LDS.64 yields the same coordinate of two candidates
	int pi[32];                     // ... and their original indices
	uint2 rng[8];                   // point ranges of the 8 cells under the current level-1 node
	unsigned char rem[8];           // remaining (octant-permuted) child masks per level
};

// d2/idx: best real candidate so far.  bnd: pruning bound = min(d2, upper bounds proven from non-empty boxes): a non-empty box
// holds a point no farther than its farthest corner, so the nearest neighbour is at most that far even before any point of
// it has been seen — far queries start pruning at once instead of walking with an infinite bound (first iteration: no seed).
struct PkLane { float qx, qy, qz, rx, ry, rz, d2, bnd; int idx; bool valid; };

// What is the same for the whole packet: the bounding box of its queries (origin-relative) for the one-test-per-child coarse cull,
// and the reference lane's query, which fixes the near-first visiting order (warp-uniform values).
struct PkPacket { float qlo[3], qhi[3], refx, refy, refz; };

// packed fp32 (sm_100 FADD2 / FMUL2 / FFMA2): both halves rounded exactly like the scalar _rn sequence.  The adds are issued as
// fma(a, 1.0f, b) with the 1.0f from a kernel argument, so ptxas cannot contract mul + add into one FFMA2 (which would drop the
// product's rounding; it does that even under --fmad=false — seen in SASS of the organized neighbour count, frame.cu).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// all lanes test the points sorted[s, e) against their own query; ties resolve to the smallest target index.
// Optionally a real function call (arguments and result in registers; -DLS3D_PK_SCAN_CALL=1): the packet kernels scan from
// several places and the loop, unrolled, is a third of their code.
#ifndef LS3D_PK_SCAN_CALL
#define LS3D_PK_SCAN_CALL 0            // measured: the call (185 us / iteration) loses to the inlined copies (170 us) although the latter is 42 KB of code
#endif
#if LS3D_PK_SCAN_CALL
__device__ __noinline__
#else
__device__ __forceinline__
#endif
unsigned long long pk_scan_fn(const float4 *__restrict__ sorted, unsigned s, unsigned e, float qx, float qy, float qz,
	float best_d2, int best_idx, PkWarp *shp, f32x2 one2)
{
	PkWarp &sh = *shp;
	const int lane = threadIdx.x & 31;
	const f32x2 q2x = pk2(qx, qx), q2y = pk2(qy, qy), q2z = pk2(qz, qz);
	// the next batch of 32 points is requested before the current one is tested, so its L2 round trip overlaps the arithmetic
	float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
	if (s + (unsigned)lane < e) c = __ldg(sorted + s + lane);
	for (unsigned base = s; base < e; base += 32) {
		const unsigned n = min(32u, e - base);
		__syncwarp();
		if ((unsigned)lane < n) {
			sh.px[lane] = c.x; sh.py[lane] = c.y; sh.pz[lane] = c.z; sh.pi[lane] = __float_as_int(c.w);
		} else if ((unsigned)lane == n) {
			// pad an odd count: a candidate infinitely far away (d2 = inf never wins)
			sh.px[lane] = INFINITY; sh.py[lane] = INFINITY; sh.pz[lane] = INFINITY; sh.pi[lane] = 0x7fffffff;
		}
		if (base + 32u + (unsigned)lane < e) c = __ldg(sorted + base + 32u + lane);
		__syncwarp();
		const unsigned np = (n + 1) >> 1;
#pragma unroll 2
		for (unsigned j = 0; j < np; j++) {
			const f32x2 X = *reinterpret_cast<const f32x2 *>(sh.px + 2 * j), Y = *reinterpret_cast<const f32x2 *>(sh.py + 2 * j), Z = *reinterpret_cast<const f32x2 *>(sh.pz + 2 * j);
			const int2 I = *reinterpret_cast<const int2 *>(sh.pi + 2 * j);
			const f32x2 d0 = sub2(q2x, X), d1 = sub2(q2y, Y), d2 = sub2(q2z, Z);
			const f32x2 sq = fma2(fma2(mul2(d0, d0), one2, mul2(d1, d1)), one2, mul2(d2, d2));      // == dist2_ref on both halves
			float da, db;
			upk2(sq, da, db);
			// after the first few candidates almost nothing improves on the best: one warp-uniform branch skips the tie logic
			if (__any_sync(kFull, fminf(da, db) <= best_d2)) {
				if (da < best_d2 || (da == best_d2 && I.x < best_idx)) { best_d2 = da; best_idx = I.x; }
				if (db < best_d2 || (db == best_d2 && I.y < best_idx)) { best_d2 = db; best_idx = I.y; }
			}
		}
	}
	return ((unsigned long long)__float_as_uint(best_d2) << 32) | (unsigned)best_idx;
}
__device__ __forceinline__ void pk_scan(const float4 *__restrict__ sorted, unsigned s, unsigned e, PkLane &q, PkWarp &sh, f32x2 one2) {
	const unsigned long long r = pk_scan_fn(sorted, s, e, q.qx, q.qy, q.qz, q.d2, q.idx, &sh, one2);
	q.d2 = __uint_as_float((unsigned)(r >> 32));
	q.idx = (int)(unsigned)r;
	q.bnd = fminf(q.bnd, q.d2);
}

// conservative squared distance from the origin-relative query to a decoded (slack-widened) box
__device__ __forceinline__ float pk_lb2(const float4 lo, const float4 hi, float rx, float ry, float rz) {
	const float dx = fmaxf(fmaxf(lo.x - rx, rx - hi.x), 0.0f), dy = fmaxf(fmaxf(lo.y - ry, ry - hi.y), 0.0f), dz = fmaxf(fmaxf(lo.z - rz, rz - hi.z), 0.0f);
	return (dx * dx + dy * dy + dz * dz) * 0.99999f;
}

// The 8 child records of node (t; ump) -> sh.lo/hi[t-1] (and their point ranges -> sh.rng when the children are cells).
// One memory round trip; lane 8a + c (a < 3) fetches child c's record and decodes its axis a, lanes 24-31 fetch the cell ranges.
// Returns the mask of non-empty children; `coarse` = those of them whose box is within sqrt(maxbnd) of the PACKET's query box
// (a superset of what any lane needs: one box-box test per child, in parallel over the children, instead of 8 per lane).
__device__ __forceinline__ unsigned pk_load_children(const IcpGrid &g, const unsigned *__restrict__ cell_start, const unsigned long long *__restrict__ nodes,
	const unsigned long long *__restrict__ cellrec, PkWarp &sh, int t, unsigned ux, unsigned uy, unsigned uz, unsigned ump, float slack,
	const PkPacket &pkt, float maxbnd, unsigned &coarse)
{
	const int lane = threadIdx.x & 31;
	const unsigned c = (unsigned)lane & 7u;
	const int a = lane >> 3;
	unsigned long long rec = 0;
	float gap2 = 0.0f;
	__syncwarp();                                                        // earlier readers of sh.lo/hi / sh.rng are done
	if (a < 3) {
		const unsigned cmp = (ump << 3) | c;
		rec = t == 1 ? __ldg(cellrec + cmp) : __ldg(nodes + c_icp_off[g.levels][t - 1] + cmp);      // 0 = empty
		const float half = g.h * (float)(1u << (t - 1));                 // child edge
		const float qs = half * (1.0f / 255.0f);
		const unsigned ua = a == 0 ? ux : (a == 1 ? uy : uz);
		const float o = (float)((ua << 1) | ((c >> a) & 1u)) * half;     // child origin along this axis
		const float mn = o + (float)(unsigned)((rec >> (8 * a)) & 0xff) * qs - slack;
		const float mx = o + (float)(unsigned)((rec >> (24 + 8 * a)) & 0xff) * qs + slack;
		reinterpret_cast<float *>(&sh.lo[t - 1][c])[a] = mn;
		reinterpret_cast<float *>(&sh.hi[t - 1][c])[a] = mx;
		const float ql = a == 0 ? pkt.qlo[0] : (a == 1 ? pkt.qlo[1] : pkt.qlo[2]);
		const float qh = a == 0 ? pkt.qhi[0] : (a == 1 ? pkt.qhi[1] : pkt.qhi[2]);
		const float gap = fmaxf(fmaxf(mn - qh, ql - mx), 0.0f);
		gap2 = gap * gap;
	} else if (t == 1) {
		// the 8 cells are consecutive in Morton order: cell_start[(ump << 3) + c .. + 1] is cell c's range
		const unsigned v0 = __ldg(cell_start + (ump << 3) + c), v1 = __ldg(cell_start + (ump << 3) + c + 1u);
		sh.rng[c] = make_uint2(v0, v1);
	}
	const float g2 = (gap2 + __shfl_down_sync(kFull, gap2, 8)) + __shfl_down_sync(kFull, gap2, 16);
	const bool ex = lane < 8 && rec != 0ull;
	const unsigned exist = __ballot_sync(kFull, ex);
	coarse = __ballot_sync(kFull, ex && g2 * 0.99999f <= maxbnd);
	__syncwarp();
	return exist;
}

// A first real candidate for packets in which some lane has none (first iteration, home cell empty): greedy descent from
// the root, at every level into the non-empty child nearest to that lane, and a scan of the cell it ends in.  L node
// visits buy every lane of the packet a bound close to its final one, so the exact walk that follows prunes like a seeded one.
__device__ __forceinline__ void pk_greedy_seed(const IcpGrid &g, const unsigned *__restrict__ cell_start, const unsigned long long *__restrict__ nodes,
	const unsigned long long *__restrict__ cellrec, const float4 *__restrict__ sorted, PkWarp &sh, PkLane &q, const PkPacket &pkt, int who, float slack,
	f32x2 one2, unsigned &steps, unsigned &scanned)
{
	unsigned ux = 0, uy = 0, uz = 0, ump = 0;
	for (int t = g.levels; t >= 1; t--) {
		unsigned coarse;
		unsigned exist = pk_load_children(g, cell_start, nodes, cellrec, sh, t, ux, uy, uz, ump, slack, pkt, INFINITY, coarse);
		if (!exist) return;                                               // empty target: cannot happen after ls3d_icp_set_target
		float best_lb = INFINITY;
		unsigned best_c = (unsigned)__ffs(exist) - 1u;
		while (exist) {
			const unsigned c = (unsigned)__ffs(exist) - 1u;
			exist &= exist - 1u;
			const float lb = pk_lb2(sh.lo[t - 1][c], sh.hi[t - 1][c], q.rx, q.ry, q.rz);
			if (lb < best_lb) { best_lb = lb; best_c = c; }
		}
		const unsigned c = __shfl_sync(kFull, best_c, who);
		steps++;
		if (t == 1) {
			const uint2 r = sh.rng[c];
			scanned += r.y - r.x;
			pk_scan(sorted, r.x, r.y, q, sh, one2);
			return;
		}
		ux = (ux << 1) | (c & 1u); uy = (uy << 1) | ((c >> 1) & 1u); uz = (uz << 1) | (c >> 2); ump = (ump << 3) | c;
	}
}

// Leave in sh.rem[t-1] the octant-permuted mask of the children in `allowed` that may hold something a lane still needs: the
// coarse packet-level cull only — every child is tested again, per lane and against the bounds of that moment, when the walk
// pops it (pk_walk), so nothing is lost by not doing the 8 per-lane tests here as well.
__device__ __forceinline__ void pk_enter(const IcpGrid &g, const unsigned *__restrict__ cell_start, const unsigned long long *__restrict__ nodes,
	const unsigned long long *__restrict__ cellrec, PkWarp &sh, int t, unsigned ux, unsigned uy, unsigned uz, unsigned ump, unsigned allowed,
	PkLane &q, const PkPacket &pkt, unsigned &nearpack, float slack)
{
	const int lane = threadIdx.x & 31;
	// the largest bound of any lane (bounds are non-negative, +inf included: their bit patterns order like unsigned integers)
	const float maxbnd = __uint_as_float(__reduce_max_sync(kFull, q.valid ? __float_as_uint(q.bnd) : 0u));
	unsigned coarse;
	unsigned exist = pk_load_children(g, cell_start, nodes, cellrec, sh, t, ux, uy, uz, ump, slack, pkt, maxbnd, coarse) & allowed;
	const float half = g.h * (float)(1u << (t - 1));                     // child edge
	if (maxbnd == INFINITY) {
		// some lane has no bound at all yet (degenerate inputs only: the greedy seed gives every lane a candidate): every
		// non-empty child proves one
		unsigned ex = exist;
		while (ex) {
			const unsigned c = (unsigned)__ffs(ex) - 1u;
			ex &= ex - 1u;
			// farthest corner of the (widened) box: some point of this non-empty box is at most that far
			const float4 lo = sh.lo[t - 1][c], hi = sh.hi[t - 1][c];
			const float fx = fmaxf(q.rx - lo.x, hi.x - q.rx), fy = fmaxf(q.ry - lo.y, hi.y - q.ry), fz = fmaxf(q.rz - lo.z, hi.z - q.rz);
			q.bnd = fminf(q.bnd, (fx * fx + fy * fy + fz * fz) * 1.00001f);
		}
	}
	const unsigned need = coarse & allowed;
	// near-first visiting order: the octant of the reference lane's query inside this node (warp-uniform arithmetic)
	const unsigned near = (pkt.refx >= (float)(2 * ux + 1) * half ? 1u : 0u) | (pkt.refy >= (float)(2 * uy + 1) * half ? 2u : 0u) | (pkt.refz >= (float)(2 * uz + 1) * half ? 4u : 0u);
	nearpack = (nearpack & ~(7u << (3 * (t - 1)))) | (near << (3 * (t - 1)));
	if (lane == 0) sh.rem[t - 1] = (unsigned char)octant_permute(need, near);
	__syncwarp();
}

struct PkDone { unsigned m[kPkDone]; };

// The exact search below / around one node, as ONE loop (one copy of pk_enter and one of pk_scan in the instruction stream: the
// first version inlined walk and climb separately and the kernel outgrew the instruction cache — ncu: 19 % of the stall samples
// were instruction fetch).
//   * depth-first, near-first walk of the subtree under (lvl; nx,ny,nz,nmp): every child popped is tested per lane against the
//     bounds of that moment; cells are scanned, inner nodes entered;
//   * climb (when `climb`): once the subtree is finished and some lane's ball still reaches outside its cube, the parent is
//     entered with the finished child masked out, and so on, level by level, until every ball fits (faces on the grid boundary
//     have nothing behind them) or the root is done.
// Returns early when `steps` exceeds `budget` (the caller queues the packet for the block-wide stage).
__device__ __forceinline__ void pk_search(const IcpGrid &g, const unsigned *__restrict__ cell_start, const unsigned long long *__restrict__ nodes,
	const unsigned long long *__restrict__ cellrec, const float4 *__restrict__ sorted, PkWarp &sh, int lvl, unsigned nx, unsigned ny, unsigned nz, unsigned nmp,
	PkLane &q, const PkPacket &pkt, float slack, const PkDone &done, f32x2 one2, unsigned &steps, unsigned &scanned, unsigned budget, bool climb)
{
	const int lane = threadIdx.x & 31;
	const int L = g.levels;
	int top = lvl, t = lvl;
	unsigned ux = nx, uy = ny, uz = nz, ump = nmp;          // the node whose children are being iterated (level t)
	unsigned allowed = 0xffu, nearpack = 0;
	bool enter = lvl > 0;
	for (;;) {
		if (steps > budget) return;
		if (enter) {
			pk_enter(g, cell_start, nodes, cellrec, sh, t, ux, uy, uz, ump, allowed, q, pkt, nearpack, slack);
			enter = false;
		}
		const unsigned rm = t > 0 ? sh.rem[t - 1] : 0u;
		if (rm == 0) {
			if (t < top) { t++; ux >>= 1; uy >>= 1; uz >>= 1; ump >>= 3; continue; }
			// the subtree under (top; ux,uy,uz) is finished
			if (!climb || top >= L) return;
			const float size = g.h * (float)(1u << top);
			const unsigned dim = (unsigned)g.G >> top;
			float rho = INFINITY;
			if (ux > 0) rho = fminf(rho, q.rx - (float)ux * size);
			if (ux + 1 < dim) rho = fminf(rho, (float)(ux + 1) * size - q.rx);
			if (uy > 0) rho = fminf(rho, q.ry - (float)uy * size);
			if (uy + 1 < dim) rho = fminf(rho, (float)(uy + 1) * size - q.ry);
			if (uz > 0) rho = fminf(rho, q.rz - (float)uz * size);
			if (uz + 1 < dim) rho = fminf(rho, (float)(uz + 1) * size - q.rz);
			rho = fmaxf(rho - slack, 0.0f);
			if (!__any_sync(kFull, q.valid && !(q.d2 <= rho * rho * 0.99999f))) return;
			allowed = 0xffu & ~(1u << (ump & 7u));              // the child just finished
			ux >>= 1; uy >>= 1; uz >>= 1; ump >>= 3;
			top++; t = top;
			enter = true;
			continue;
		}
		const unsigned cp = (unsigned)__ffs(rm) - 1u;
		__syncwarp();
		if (lane == 0) sh.rem[t - 1] = (unsigned char)(rm & (rm - 1u));
		__syncwarp();
		const unsigned child = cp ^ ((nearpack >> (3 * (t - 1))) & 7u);
		const unsigned cmp = (ump << 3) | child;
		if (t == 1) {
			bool skip = false;
#pragma unroll
			for (int d = 0; d < kPkDone; d++) skip |= done.m[d] == cmp;     // a home cell: already scanned
			if (skip) continue;
		}
		// the per-lane test, against the bounds as they are now
		const float lb = pk_lb2(sh.lo[t - 1][child], sh.hi[t - 1][child], q.rx, q.ry, q.rz);
		if (!__any_sync(kFull, q.valid && lb <= q.bnd)) continue;
		steps++;
		if (t == 1) {
			const uint2 r = sh.rng[child];
			scanned += r.y - r.x;
			pk_scan(sorted, r.x, r.y, q, sh, one2);
		} else {
			t--;
			ux = (ux << 1) | (child & 1u); uy = (uy << 1) | ((child >> 1) & 1u); uz = (uz << 1) | (child >> 2); ump = cmp;
			allowed = 0xffu;
			enter = true;
		}
	}
}

// order[k], k in [0, i_end - i_begin): the slice's source indices in Morton order of their (initial) home cells.
// apply != 0: first apply the update left in state->xf by the solve step to the points of the slice (points outside the
// slice are transformed by k_icp_apply).
#ifndef LS3D_PK_MINBLOCKS
#define LS3D_PK_MINBLOCKS 3
#endif
// A warp gives up on a packet after this many node / cell visits and queues it for k_icp_match_heavy, where a whole block
// searches it.  A warp's instruction stream is serial: the few packets that need 100+ visits (queries far from the target whose
// search balls graze a lot of surface) used to run for the whole length of the kernel while every other warp had long finished
// (ncu, round 2: SMSPs active 56 % of the kernel's duration).  Packets that were over budget last iteration go straight there.
#ifndef LS3D_PK_BUDGET
#define LS3D_PK_BUDGET 64
#endif
constexpr unsigned kPkBudget = LS3D_PK_BUDGET;
#ifndef LS3D_PK_BUDGET0_PCT
#define LS3D_PK_BUDGET0_PCT 75u         // first match stage of a call, in % of the budget.  Measured on the bench pair: 200 -> 1.79 ms per
#endif                                  // call, 400 -> 1.59, 100 -> 1.54, 50 -> 1.60; 62 / 75 / 88 -> 1.50-1.55 (run-to-run spread of the same size)

// this lane's query of packet `pd`: position (after the pending update when apply != 0, which is also written back), home cell,
// and the previous nearest neighbour as the first candidate
__device__ __forceinline__ int pk_load_lane(const uint2 pd, const unsigned *__restrict__ order, float *__restrict__ verts2, const float *__restrict__ verts1,
	const int *__restrict__ nn_idx, const IcpGrid &g, const float *__restrict__ xf /* T[3] Rk[9], or NULL */, PkLane &q, unsigned &mp, unsigned &hx, unsigned &hy, unsigned &hz)
{
	const int lane = threadIdx.x & 31;
	const int i = (unsigned)lane < pd.y ? (int)__ldg(order + pd.x + lane) : -1;
	q.valid = false; q.d2 = INFINITY; q.bnd = INFINITY; q.idx = -1;
	q.qx = q.qy = q.qz = q.rx = q.ry = q.rz = 0.0f;
	mp = 0; hx = 0; hy = 0; hz = 0;
	if (i >= 0) {
		float x = verts2[3 * (size_t)i], y = verts2[3 * (size_t)i + 1], z = verts2[3 * (size_t)i + 2];
		const int prev = nn_idx[i];
		if (xf) {
			// the update is re-read per packet (12 cached, warp-uniform loads) instead of living in 12 registers for the whole kernel
			float T[3], Rk[9];
#pragma unroll
			for (int a = 0; a < 3; a++) T[a] = xf[a];
#pragma unroll
			for (int a = 0; a < 9; a++) Rk[a] = xf[3 + a];
			apply_xform(x, y, z, T, Rk);
			verts2[3 * (size_t)i] = x; verts2[3 * (size_t)i + 1] = y; verts2[3 * (size_t)i + 2] = z;
		}
		q.qx = x; q.qy = y; q.qz = z;
		q.rx = x - g.ox; q.ry = y - g.oy; q.rz = z - g.oz;
		q.valid = isfinite(q.rx) && isfinite(q.ry) && isfinite(q.rz);
		if (q.valid) {
			if (prev >= 0) {
				const float d2 = dist2_ref(x, y, z, verts1[3 * (size_t)prev], verts1[3 * (size_t)prev + 1], verts1[3 * (size_t)prev + 2]);
				if (d2 == d2) { q.d2 = d2; q.bnd = d2; q.idx = prev; }
			}
			hx = (unsigned)icp_cell(q.rx, g.inv_h, g.G); hy = (unsigned)icp_cell(q.ry, g.inv_h, g.G); hz = (unsigned)icp_cell(q.rz, g.inv_h, g.G);
			mp = morton3(hx, hy, hz);
		}
	}
	return i;
}

__device__ __forceinline__ void pk_packet_box(const PkLane &q, int ref, PkPacket &pkt) {
	pkt.refx = __shfl_sync(kFull, q.rx, ref); pkt.refy = __shfl_sync(kFull, q.ry, ref); pkt.refz = __shfl_sync(kFull, q.rz, ref);
	pkt.qlo[0] = ord2f(__reduce_min_sync(kFull, q.valid ? f2ord(q.rx) : 0xffffffffu)); pkt.qhi[0] = ord2f(__reduce_max_sync(kFull, q.valid ? f2ord(q.rx) : 0u));
	pkt.qlo[1] = ord2f(__reduce_min_sync(kFull, q.valid ? f2ord(q.ry) : 0xffffffffu)); pkt.qhi[1] = ord2f(__reduce_max_sync(kFull, q.valid ? f2ord(q.ry) : 0u));
	pkt.qlo[2] = ord2f(__reduce_min_sync(kFull, q.valid ? f2ord(q.rz) : 0xffffffffu)); pkt.qhi[2] = ord2f(__reduce_max_sync(kFull, q.valid ? f2ord(q.rz) : 0u));
}

__global__ void __launch_bounds__(kPkWarps * 32, LS3D_PK_MINBLOCKS) k_icp_match_packet(float *__restrict__ verts2, int i_begin, int i_end, int apply,
	const unsigned *__restrict__ order, const uint2 *__restrict__ desc, const unsigned *__restrict__ sched, unsigned *__restrict__ pk_cost,
	const IcpGrid *__restrict__ grid, const unsigned *__restrict__ cell_start, const unsigned long long *__restrict__ nodes,
	const unsigned long long *__restrict__ cellrec, const float4 *__restrict__ sorted, const float *__restrict__ verts1, SlotMap slotmap,
	IcpState *state, int *__restrict__ nn_idx, float *__restrict__ nn_d2, unsigned *__restrict__ dbg, float one, unsigned *__restrict__ heavy_list, unsigned budget)
{
	__shared__ __align__(16) PkWarp s_pk[kPkWarps];
	// the block-wide stage (k_icp_match_heavy, a programmatic dependent launch) may start as soon as SM resources free up: it
	// consumes the queue while this kernel's last packets are still running
	asm volatile("griddepcontrol.launch_dependents;");
	// this kernel itself may have been launched early (programmatic dependent launch behind the previous iteration's reduction):
	// nothing that kernel wrote is read before it has completed
	asm volatile("griddepcontrol.wait;" ::: "memory");
	PkWarp &sh = s_pk[threadIdx.x >> 5];
	const float *xf = apply ? state->xf : nullptr;
	const IcpGrid g = *grid;
	const int lane = threadIdx.x & 31;
	const float slack = 1e-3f * g.h;
	const f32x2 one2 = pk2(one, one);
	const int n_packets = (int)state->n_packets;
	const bool use_sched = state->sched_valid != 0;
	// first match stage of a call: no seeds and no cost history yet, walks are ~40 % longer — more of them go to the block-wide stage
	const unsigned walk_budget = (use_sched || budget >= 0x3fffffffu) ? budget : (unsigned)((unsigned long long)budget * LS3D_PK_BUDGET0_PCT / 100u);
	for (;;) {
		int pk = 0;
		if (lane == 0) {
			pk = (int)atomicAdd(&state->n_work, 1u);
			// longest first: the packets that walked furthest last iteration (k_icp_stats' schedule) start in the first wave
			if (pk < n_packets && use_sched) pk = (int)sched[pk];
		}
		pk = __shfl_sync(kFull, pk, 0);
		if (pk >= n_packets) break;
		const uint2 pd = __ldg(desc + pk);
		PkLane q;
		unsigned mp, hx, hy, hz;
		const int i = pk_load_lane(pd, order, verts2, verts1, nn_idx, g, xf, q, mp, hx, hy, hz);
		const unsigned vmask = __ballot_sync(kFull, q.valid);
		unsigned steps = 0, scanned = 0;
		// over budget last iteration: straight to the block-wide stage (the points are transformed, the old neighbour stays the seed)
		bool heavy = use_sched && vmask && __ldg(pk_cost + pk) > budget;
		if (vmask && !heavy) {
			const int ref = __ffs(vmask) - 1;
			const unsigned mp_ref = __shfl_sync(kFull, mp, ref);
			const unsigned rhx = __shfl_sync(kFull, hx, ref), rhy = __shfl_sync(kFull, hy, ref), rhz = __shfl_sync(kFull, hz, ref);
			PkPacket pkt;
			pk_packet_box(q, ref, pkt);
			// ---- the home cells first: every lane gets a finite bound before the walk starts ----
			PkDone done;
#pragma unroll
			for (int d = 0; d < kPkDone; d++) done.m[d] = 0xffffffffu;
			{
				unsigned todo = vmask;
#pragma unroll
				for (int d = 0; d < kPkDone; d++) {
					if (todo) {
						const unsigned m = __shfl_sync(kFull, mp, __ffs(todo) - 1);
						todo &= ~__ballot_sync(kFull, q.valid && mp == m);
						done.m[d] = m;
						const unsigned s = __ldg(cell_start + m), e = __ldg(cell_start + m + 1);
						if (s != e) { steps++; scanned += e - s; pk_scan(sorted, s, e, q, sh, one2); }
					}
				}
			}
			{
				const unsigned unseeded = __ballot_sync(kFull, q.valid && q.idx < 0);
				if (unseeded) pk_greedy_seed(g, cell_start, nodes, cellrec, sorted, sh, q, pkt, __ffs(unseeded) - 1, slack, one2, steps, scanned);
			}
			const unsigned diff = __reduce_or_sync(kFull, q.valid ? (mp ^ mp_ref) : 0u);
			int lvl = diff ? (31 - __clz((int)diff)) / 3 + 1 : 0;      // lowest level whose node holds every lane's home cell
			// ---- the subtree all home cells share (level 0: the one home cell, done above), then its siblings level by level ----
			pk_search(g, cell_start, nodes, cellrec, sorted, sh, lvl, rhx >> lvl, rhy >> lvl, rhz >> lvl, mp_ref >> (3 * lvl), q, pkt, slack, done, one2, steps, scanned, walk_budget, true);
			heavy = steps > walk_budget;
		}
		if (heavy) {
			// the best candidate so far is a real point: it seeds the block-wide search (nothing is committed to the dedupe slots yet)
			if (i >= 0 && q.valid && q.idx >= 0) nn_idx[i] = q.idx;
			__threadfence();                                              // seeds and transformed points before the queue entry
			__syncwarp();
			if (lane == 0) st_volatile_u32(heavy_list + atomicAdd(&state->hq_n, 1u), (unsigned)pk);
		} else {
			if (i >= 0) {
				nn_commit(i, q.d2, q.valid ? q.idx : -1, slotmap, nn_idx, nn_d2);
				if (dbg) { dbg[3 * (size_t)i] = steps; dbg[3 * (size_t)i + 1] = scanned; dbg[3 * (size_t)i + 2] = 0u; }
			}
			if (lane == 0) { pk_cost[pk] = steps; atomicAdd(&state->cls_n[pk_cost_class(steps, budget)], 1u); }
		}
		__syncwarp();
	}
	// the last block to run dry re-arms the packet counter for the next match stage
	if (slotmap.world > 1) __threadfence_system(); else __threadfence();      // this block's results (incl. keys sent to peers) before its arrival
	__syncthreads();
	if (threadIdx.x == 0 && atomicAdd(&state->blocks_done, 1u) == gridDim.x - 1) {
		state->blocks_done = 0;
		state->n_work = 0;
		state->n_heavy = 0;
		state->n_light = 0;
		__threadfence();
		st_volatile_u32(&state->light_done, 1u);                      // the queue is final
	}
}

// ------------------------------------------------------------------------------------------------------
// block-wide search of the packets the warps gave up on
// ------------------------------------------------------------------------------------------------------
// One block = one heavy packet at a time; every warp holds the packet's 32 queries (lane i = query i) and the warps share the
// tree: a level-synchronous top-down sweep from the root — the frontier of nodes that some lane still needs is split over the
// warps, each tests the children of its nodes per lane against the bounds and appends the survivors to the next level's
// frontier — then the surviving cells are scanned, again split over the warps, nearest first within what a warp draws.  The
// lanes' best candidates are merged through 64-bit shared-memory atomicMin keys (bits(d2) << 32 | index: smaller distance, then
// smaller target index, exactly the packet kernel's tie rule), which every warp re-reads before each test, so a neighbour
// found by one warp prunes the others at once.  The seed is the best candidate the packet kernel left in nn_idx (a real point),
// hence the bounds are near-final from the start and the sweep touches little more than the depth-first walk would have.
// Exactness is the packet kernel's: a node is dropped only when every lane's conservative lower bound exceeds that lane's bound.
constexpr unsigned kHvEmpty = 0xffffffffu;  // queue entry not (yet) written
constexpr int kHvFrontCap = 1024;      // nodes per level; beyond that a warp walks the child depth-first on the spot
constexpr int kHvCellCap = 1024;

struct HvBlock {
	PkWarp pkw[kPkWarps];
	unsigned long long key[32];               // per query: bits(d2) << 32 | target index
	unsigned front[2][kHvFrontCap];           // Morton codes of the frontier nodes of the current / next level
	unsigned cells[kHvCellCap];               // Morton codes of the cells to scan
	uint2 cell_rng[kHvCellCap];
	float4 cell_lo[kHvCellCap], cell_hi[kHvCellCap];
	unsigned n_front[2], n_cells, next_cell, steps, scanned;
	int pk_id;
};

__device__ __forceinline__ void hv_refresh(const HvBlock &hb, PkLane &q) {
	const unsigned long long k = hb.key[threadIdx.x & 31];
	const float d = __uint_as_float((unsigned)(k >> 32));
	if (d < q.d2 || (d == q.d2 && (int)(unsigned)k < q.idx)) { q.d2 = d; q.idx = (int)(unsigned)k; }
	q.bnd = fminf(q.bnd, q.d2);
}
__device__ __forceinline__ void hv_publish(HvBlock &hb, const PkLane &q) {
	if (q.valid && q.idx >= 0) {
		const unsigned long long k = ((unsigned long long)__float_as_uint(q.d2) << 32) | (unsigned)q.idx;
		if (k < hb.key[threadIdx.x & 31]) atomicMin(&hb.key[threadIdx.x & 31], k);
	}
}

#ifndef LS3D_HV_MINBLOCKS
#define LS3D_HV_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(kPkWarps * 32, LS3D_HV_MINBLOCKS) k_icp_match_heavy(const float *__restrict__ verts2, const unsigned *__restrict__ order, const uint2 *__restrict__ desc,
	unsigned *__restrict__ pk_cost, const IcpGrid *__restrict__ grid, const unsigned *__restrict__ cell_start, const unsigned long long *__restrict__ nodes,
	const unsigned long long *__restrict__ cellrec, const float4 *__restrict__ sorted, const float *__restrict__ verts1, SlotMap slotmap,
	IcpState *state, int *__restrict__ nn_idx, float *__restrict__ nn_d2, unsigned *__restrict__ dbg, float one, unsigned *heavy_list, unsigned budget)
{
	extern __shared__ __align__(16) unsigned char hv_smem[];
	HvBlock &hb = *reinterpret_cast<HvBlock *>(hv_smem);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	PkWarp &sh = hb.pkw[warp];
	const IcpGrid g = *grid;
	const int L = g.levels;
	const float slack = 1e-3f * g.h;
	const f32x2 one2 = pk2(one, one);
	for (;;) {
		__syncthreads();                                                 // the previous packet's shared state is no longer read
		if (threadIdx.x == 0) {
			// entry h of the queue: wait until the packet kernel has published it, or has run dry without doing so
			const unsigned h = atomicAdd(&state->hq_head, 1u);
			unsigned e = ld_volatile_u32(heavy_list + h);
			for (unsigned spins = 0; e == kHvEmpty; spins++) {
				if (ld_volatile_u32(&state->light_done)) { __threadfence(); e = ld_volatile_u32(heavy_list + h); break; }
				if (spins > (1u << 21)) { atomicOr(&state->err, kErrScanSpin); break; }      // ~0.5 s: the protocol is broken; never hang the GPU
				__nanosleep(200);
				e = ld_volatile_u32(heavy_list + h);
			}
			if (e != kHvEmpty) st_volatile_u32(heavy_list + h, kHvEmpty);      // the queue is all-empty again for the next match stage
			hb.pk_id = e != kHvEmpty ? (int)e : -1;
			hb.n_front[0] = 1; hb.n_front[1] = 0; hb.front[0][0] = 0u;     // the root
			hb.n_cells = 0; hb.next_cell = 0; hb.steps = 0; hb.scanned = 0;
		}
		if (threadIdx.x < 32) hb.key[threadIdx.x] = ~0ull;
		__syncthreads();
		const int pk = hb.pk_id;
		if (pk < 0) break;
		const uint2 pd = __ldg(desc + pk);
		PkLane q;
		unsigned mp, hx, hy, hz;
		const int i = pk_load_lane(pd, order, const_cast<float *>(verts2), verts1, nn_idx, g, nullptr, q, mp, hx, hy, hz);
		const unsigned vmask = __ballot_sync(kFull, q.valid);
		PkPacket pkt;
		pk_packet_box(q, vmask ? __ffs(vmask) - 1 : 0, pkt);
		if (warp == 0) hv_publish(hb, q);
		unsigned steps = 0, scanned = 0;
		PkDone done;
#pragma unroll
		for (int d = 0; d < kPkDone; d++) done.m[d] = 0xffffffffu;
		// ---- level-synchronous sweep: nodes of level t in front[cur], their surviving children into front[cur ^ 1] (cells: the cell list) ----
		int cur = 0;
		for (int t = L; t >= 1; t--) {
			const unsigned nf = hb.n_front[cur];
			for (unsigned f = (unsigned)warp; f < nf; f += kPkWarps) {
				const unsigned ump = hb.front[cur][f];
				const unsigned ux = compact3(ump), uy = compact3(ump >> 1), uz = compact3(ump >> 2);
				hv_refresh(hb, q);
				unsigned coarse;
				unsigned exist = pk_load_children(g, cell_start, nodes, cellrec, sh, t, ux, uy, uz, ump, slack, pkt, INFINITY, coarse);
				steps++;
				if (__any_sync(kFull, q.valid && q.bnd == INFINITY)) {
					// degenerate seeds only: a non-empty child box proves a bound (its farthest corner)
					unsigned ex = exist;
					while (ex) {
						const unsigned c = (unsigned)__ffs(ex) - 1u;
						ex &= ex - 1u;
						const float4 lo = sh.lo[t - 1][c], hi = sh.hi[t - 1][c];
						const float fx = fmaxf(q.rx - lo.x, hi.x - q.rx), fy = fmaxf(q.ry - lo.y, hi.y - q.ry), fz = fmaxf(q.rz - lo.z, hi.z - q.rz);
						q.bnd = fminf(q.bnd, (fx * fx + fy * fy + fz * fz) * 1.00001f);
					}
				}
				unsigned need = 0;
				while (exist) {
					const unsigned c = (unsigned)__ffs(exist) - 1u;
					exist &= exist - 1u;
					const float lb = pk_lb2(sh.lo[t - 1][c], sh.hi[t - 1][c], q.rx, q.ry, q.rz);
					if (__any_sync(kFull, q.valid && lb <= q.bnd)) need |= 1u << c;
				}
				const unsigned nn = (unsigned)__popc(need);
				if (!nn) continue;
				unsigned base = 0;
				if (lane == 0) base = t > 1 ? atomicAdd(&hb.n_front[cur ^ 1], nn) : atomicAdd(&hb.n_cells, nn);
				base = __shfl_sync(kFull, base, 0);
				const unsigned cap = t > 1 ? (unsigned)kHvFrontCap : (unsigned)kHvCellCap;
				unsigned k = 0;
				while (need) {
					const unsigned c = (unsigned)__ffs(need) - 1u;
					need &= need - 1u;
					const unsigned cmp = (ump << 3) | c;
					if (base + k < cap) {
						if (t > 1) { if (lane == 0) hb.front[cur ^ 1][base + k] = cmp; }
						else if (lane == 0) { hb.cells[base + k] = cmp; hb.cell_rng[base + k] = sh.rng[c]; hb.cell_lo[base + k] = sh.lo[0][c]; hb.cell_hi[base + k] = sh.hi[0][c]; }
					} else if (t > 1) {
						// no room in the frontier: this warp walks the child depth-first right here
						const unsigned cx = (ux << 1) | (c & 1u), cy = (uy << 1) | ((c >> 1) & 1u), cz = (uz << 1) | (c >> 2);
						pk_search(g, cell_start, nodes, cellrec, sorted, sh, t - 1, cx, cy, cz, cmp, q, pkt, slack, done, one2, steps, scanned, 0xffffffffu, false);
						hv_publish(hb, q);
						// the walk reuses sh.lo/hi of the levels below t only: this node's remaining children are not disturbed
					} else {
						const uint2 r = sh.rng[c];
						scanned += r.y - r.x;
						pk_scan(sorted, r.x, r.y, q, sh, one2);
						hv_publish(hb, q);
					}
					k++;
				}
			}
			__syncthreads();
			if (threadIdx.x == 0) { hb.n_front[cur] = 0; if (hb.n_front[cur ^ 1] > (unsigned)kHvFrontCap) hb.n_front[cur ^ 1] = kHvFrontCap; }
			cur ^= 1;
			__syncthreads();
		}
		// ---- the surviving cells, handed out one at a time ----
		const unsigned nc = min(hb.n_cells, (unsigned)kHvCellCap);
		for (;;) {
			unsigned k = 0;
			if (lane == 0) k = atomicAdd(&hb.next_cell, 1u);
			k = __shfl_sync(kFull, k, 0);
			if (k >= nc) break;
			hv_refresh(hb, q);
			const float lb = pk_lb2(hb.cell_lo[k], hb.cell_hi[k], q.rx, q.ry, q.rz);
			if (!__any_sync(kFull, q.valid && lb <= q.bnd)) continue;
			const uint2 r = hb.cell_rng[k];
			steps++;
			scanned += r.y - r.x;
			pk_scan(sorted, r.x, r.y, q, sh, one2);
			hv_publish(hb, q);
		}
		if (lane == 0) { atomicAdd(&hb.steps, steps); atomicAdd(&hb.scanned, scanned); }
		__syncthreads();
		if (warp == 0) {
			hv_refresh(hb, q);
			if (i >= 0) {
				nn_commit(i, q.d2, q.valid ? q.idx : -1, slotmap, nn_idx, nn_d2);
				if (dbg) { dbg[3 * (size_t)i] = hb.steps; dbg[3 * (size_t)i + 1] = hb.scanned; dbg[3 * (size_t)i + 2] = 1u; }
			}
			// stays with the block-wide stage for the rest of the call — except after the first match stage, whose warps gave up earlier
			// than they will from now on (LS3D_PK_BUDGET0_PCT) and without seeds: those packets get one more try in the first wave
			if (lane == 0) {
				const bool first_stage = ld_volatile_u32(&state->sched_valid) == 0u && budget < 0x3fffffffu;
				pk_cost[pk] = first_stage ? budget : max(hb.steps, budget + 1u);
				atomicAdd(&state->cls_n[first_stage ? pk_cost_class(budget, budget) : 0u], 1u);
			}
		}
	}
	// the last block re-arms the queue for the next match stage
	if (slotmap.world > 1) __threadfence_system();
	if (threadIdx.x == 0 && atomicAdd(&state->hq_done, 1u) == gridDim.x - 1) {
		state->hq_done = 0;
		state->hq_head = 0;
		state->hq_n = 0;
		state->light_done = 0;
	}
}

// Packets = runs of at most 32 consecutive entries of the Morton order that never leave one level-kPkRunLevel node: a packet
// that straddled distant nodes (isolated points — flying pixels — sort next to each other but lie far apart) would drag one
// warp through 32 unrelated searches in series; cut at node boundaries, such points become many small packets that run in
// parallel on warps that would otherwise idle.
constexpr int kPkRunLevel = 2;
__global__ void __launch_bounds__(256) k_icp_packets(const unsigned *__restrict__ src_start, unsigned n_runs, uint2 *__restrict__ desc, IcpState *state) {
	for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += gridDim.x * blockDim.x) {
		const unsigned s = src_start[r << (3 * kPkRunLevel)], e = src_start[(r + 1) << (3 * kPkRunLevel)];
		if (s == e) continue;
		const unsigned n = e - s, np = (n + 31) >> 5;
		const unsigned base = atomicAdd(&state->n_packets, np);
		for (unsigned j = 0; j < np; j++) desc[base + j] = make_uint2(s + 32 * j, min(32u, n - 32 * j));
	}
}

// (T, Rk) applied to the source points in [a0, a1) and [b0, b1): the final update of a call, and the points outside a rank's slice
__global__ void __launch_bounds__(256) k_icp_apply(float *__restrict__ verts2, int a0, int a1, int b0, int b1, const IcpState *state) {
	float T[3], Rk[9];
#pragma unroll
	for (int a = 0; a < 3; a++) T[a] = state->xf[a];
#pragma unroll
	for (int a = 0; a < 9; a++) Rk[a] = state->xf[3 + a];
	const int na = a1 - a0, nb = b1 - b0;
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < na + nb; j += gridDim.x * blockDim.x) {
		const int i = j < na ? a0 + j : b0 + (j - na);
		float x = verts2[3 * (size_t)i], y = verts2[3 * (size_t)i + 1], z = verts2[3 * (size_t)i + 2];
		apply_xform(x, y, z, T, Rk);
		verts2[3 * (size_t)i] = x; verts2[3 * (size_t)i + 1] = y; verts2[3 * (size_t)i + 2] = z;
	}
}

// Morton order of the slice's source points on the TARGET grid (counting sort: count -> exclusive scan -> scatter)
__global__ void __launch_bounds__(256) k_icp_order_scatter(int i_begin, int n_slice, const unsigned *__restrict__ cell_of, const unsigned *__restrict__ rank_of,
	const unsigned *__restrict__ src_start, unsigned *__restrict__ order)
{
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_slice; j += gridDim.x * blockDim.x)
		order[src_start[cell_of[j]] + rank_of[j]] = (unsigned)(i_begin + j);
}

// ------------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------------
// the whole reduction of an iteration in one launch
// ------------------------------------------------------------------------------------------------------
// GetStandardDeviation + RejectOutlierMatches + the Kabsch sums + the solve step (icp.cpp:34-73, 138-168) over the dedupe slots,
// written as the reference writes it: the MEAN of the matched squared distances first (rounded to fp32, divided in fp32 like
// `mean /= data.size()`), then the sum of (double)(fp32(d - mean))^2, an fp32 division and square root — the two-pass form, not
// E[x^2] - mean^2 — then the 2.5 sigma gate and the 16 correspondence sums.  The three sums are fp64 and order-fixed where the
// reference runs fp32 accumulators down the match list in first-occurrence order; that order is a property of the reference's
// serial loop and is not reproduced (tests/test_gpu_icp_fuzz.py bounds what it costs: nothing, in accepted counts).
//
// Canonical chunked reduction.  The target range is cut into C = red_chunks(n1) equal chunks; one block reduces one chunk in a
// fixed thread order and every total is the chunk partials folded in chunk order.  Nothing in that depends on how many GPUs
// share the work: with `world` ranks each owns a contiguous run of chunks (and of dedupe slots: the match kernels atomicMin
// straight into the owner's slots over NVLink), stores its chunk partials into every rank's partial table with peer stores and
// raises a flag there; every rank then folds the same C numbers in the same order, so poses are bit-identical on 1, 2, 4 or 8
// GPUs and no collective library call sits between the three passes.  Grid barriers are arrival counters (all blocks are
// resident: at most 2 per SM); spins are bounded and raise kErrScanSpin rather than hang.
constexpr int kRedChunksMax = 296;
__host__ __device__ __forceinline__ int red_chunks(int n1) { const int c = (n1 + 255) / 256; return c < 1 ? 1 : (c > kRedChunksMax ? kRedChunksMax : c); }
__host__ __device__ __forceinline__ int red_chunk_size(int n1) { const int C = red_chunks(n1); return (n1 + C - 1) / C; }
constexpr int kRedPhaseStride = kRedChunksMax * 16;     // doubles per pass in a partial table

struct IcpPeers {
	int world, rank;
	double *part[8];           // every rank's partial table [3][kRedChunksMax][16] (part[rank] is the local one)
	unsigned *flag[8];         // every rank's flag words [8]: flag[r][src] is raised on rank r by rank src
	unsigned long long *slots[8];   // every rank's dedupe slots (only the owner's range of each is ever used)
};

__device__ __forceinline__ unsigned ld_volatile_sys_u32(const unsigned *p) { unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

// all local blocks have stored (and fenced) their pass-`phase` partials; the last one to arrive raises this rank's flag on every peer
__device__ __forceinline__ void red_arrive(IcpState *st, unsigned nblocks, unsigned phase, unsigned epoch, const IcpPeers &pe) {
	// release: everything this block has written so far (partials into the peers' tables, slot resets) before its arrival counts.
	// One fence by the arriving thread after the block barrier is cumulative over the block's writes; a system-scope fence in
	// every thread (first version) cost ~10 us per pass.
	__syncthreads();
	if (threadIdx.x == 0) {
		if (pe.world > 1) __threadfence_system(); else __threadfence();
		const unsigned arrived = atomicAdd(&st->red_bar, 1u) + 1u;
		if (pe.world > 1 && arrived == phase * nblocks) {
			__threadfence_system();
			for (int r = 0; r < pe.world; r++) if (r != pe.rank) st_volatile_u32(pe.flag[r] + pe.rank, epoch * 4u + phase);
		}
	}
}
// ... and every block of every rank has arrived
__device__ __forceinline__ void red_wait(IcpState *st, unsigned nblocks, unsigned phase, unsigned epoch, const IcpPeers &pe) {
	if (threadIdx.x == 0) {
		unsigned spins = 0;
		while (ld_volatile_u32(&st->red_bar) < phase * nblocks) if (++spins > (1u << 24)) { atomicOr(&st->err, kErrScanSpin); break; }
		for (int r = 0; r < pe.world; r++) {
			if (r == pe.rank) continue;
			spins = 0;
			while (ld_volatile_sys_u32(pe.flag[pe.rank] + r) < epoch * 4u + phase) if (++spins > (1u << 24)) { atomicOr(&st->err, kErrScanSpin); break; }
		}
		__threadfence();
	}
	__syncthreads();
}

// the block's NV values (already reduced over the block into smem[0..NV)) -> chunk c of pass `phase` in every rank's table
template <int NV>
__device__ __forceinline__ void red_store(const double *smem, int c, int phase, const IcpPeers &pe) {
	if (threadIdx.x < NV && c >= 0) {
		const double v = smem[threadIdx.x];
		for (int r = 0; r < pe.world; r++) pe.part[r][(size_t)(phase - 1) * kRedPhaseStride + (size_t)c * 16 + threadIdx.x] = v;
	}
}

// fixed-order block reduction of NV doubles per thread -> smem[0..NV) (256 threads)
template <int NV>
__device__ __forceinline__ void red_block(double *v, double *smem /* [8][NV] + [NV] */) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int a = 0; a < NV; a++) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v[a] += __shfl_xor_sync(kFull, v[a], o);
	}
	__syncthreads();                                  // smem may still be read from the previous pass
	if (lane == 0) {
#pragma unroll
		for (int a = 0; a < NV; a++) smem[NV + warp * NV + a] = v[a];
	}
	__syncthreads();
	if (threadIdx.x < NV) {
		double s = 0;
		for (int w = 0; w < 8; w++) s += smem[NV + w * NV + threadIdx.x];
		smem[threadIdx.x] = s;
	}
	__syncthreads();
}

// The C chunk partials of pass `phase` folded in a fixed order (every block, every rank: the same numbers in the same order, hence
// the same bits).  The order: chunk c belongs to column l = c % 32 and row w = (c / 32) % 8; a (row, column) cell adds its chunks
// in increasing c, a column adds its 8 cells in row order, the 32 columns are combined by a 5-step xor butterfly.  All 256
// threads load in parallel (one cell each, at most two chunks for C <= 296); the result lands in out[0..NV).
template <int NV>
__device__ __forceinline__ void red_total(const double *part_local, int C, int phase, double *cells /* [NV][256] */, double *out) {
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	{
		double acc[NV];
#pragma unroll
		for (int a = 0; a < NV; a++) acc[a] = 0;
		for (int c = t; c < C; c += 256) {
#pragma unroll
			for (int a = 0; a < NV; a++) acc[a] += __ldcg(part_local + (size_t)(phase - 1) * kRedPhaseStride + (size_t)c * 16 + a);
		}
#pragma unroll
		for (int a = 0; a < NV; a++) cells[a * 256 + t] = acc[a];          // component-major: conflict-free both ways
	}
	__syncthreads();
	for (int a = warp; a < NV; a += 8) {          // warp `a` (and a + 8) folds component a: lane = column
		double col = 0;
#pragma unroll
		for (int w = 0; w < 8; w++) col += cells[a * 256 + w * 32 + lane];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) col += __shfl_xor_sync(kFull, col, o);
		if (lane == 0) out[a] = col;
	}
	__syncthreads();
}

#ifdef LS3D_RED_TIMING
__device__ unsigned long long g_red_t[16];
__device__ __forceinline__ void red_mark(int k) { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_red_t[k] = t; } }
extern "C" __global__ void k_red_timing_dump(unsigned long long *out) { if (threadIdx.x < 16) out[threadIdx.x] = g_red_t[threadIdx.x]; }
#else
__device__ __forceinline__ void red_mark(int) {}
#endif
__global__ void __launch_bounds__(256, 2) k_icp_reduce(unsigned long long *__restrict__ slots, int n1, int c_begin, const float *__restrict__ verts1, const float *__restrict__ verts2,
	IcpState *state, Ls3dIcpTrace *trace, int trace_idx, const unsigned *__restrict__ pk_cost, unsigned *__restrict__ sched, unsigned budget, IcpPeers pe)
{
	__shared__ double smem[16 + 8 * 16];
	__shared__ double s_cells[256 * 16];
	asm volatile("griddepcontrol.launch_dependents;");                // the next match kernel may become resident (it waits for us to complete)
	red_mark(0);
	asm volatile("griddepcontrol.wait;" ::: "memory");                // the match stage (packet + block-wide kernels) is complete
	red_mark(1);
	const int C = red_chunks(n1), S = red_chunk_size(n1);
	const int c = c_begin < 0 ? -1 : c_begin + (int)blockIdx.x;      // -1: this rank owns no chunk (fewer chunks than ranks); it still takes part in the exchanges
	const unsigned nblocks = gridDim.x;
	const unsigned epoch = ld_volatile_u32(&state->red_epoch);
	double *part_local = pe.part[pe.rank];
	// every rank's match kernels have finished (their atomicMin's into our slots are performed) before the slots are read
	if (pe.world > 1) {
		if (threadIdx.x == 0) {
			if (blockIdx.x == 0) { __threadfence_system(); for (int r = 0; r < pe.world; r++) if (r != pe.rank) st_volatile_u32(pe.flag[r] + pe.rank, epoch * 4u); }
			for (int r = 0; r < pe.world; r++) {
				if (r == pe.rank) continue;
				unsigned spins = 0;
				while (ld_volatile_sys_u32(pe.flag[pe.rank] + r) < epoch * 4u) if (++spins > (1u << 24)) { atomicOr(&state->err, kErrScanSpin); break; }
			}
			__threadfence();
		}
		__syncthreads();
	}
	red_mark(2);
	const int j0 = c < 0 ? 0 : min(n1, c * S), j1 = c < 0 ? 0 : min(n1, j0 + S);

	// the chunk's slots are read once: a thread's first kRedCache keys stay in registers for all three passes (a 2 x 213 k pair has
	// 3 per thread), what lies beyond is re-read from L2
	constexpr int kRedCache = 4;
	unsigned long long kc[kRedCache];
#pragma unroll
	for (int u = 0; u < kRedCache; u++) { const int j = j0 + (int)threadIdx.x + 256 * u; kc[u] = j < j1 ? __ldcg(slots + j) : kSlotEmpty; }

	// ---- pass 1: matched count and sum of d2 -> mean ----
	double v[16];
#pragma unroll
	for (int a = 0; a < 16; a++) v[a] = 0;
#pragma unroll
	for (int u = 0; u < kRedCache; u++) if (kc[u] != kSlotEmpty) { v[0] += 1.0; v[1] += (double)__uint_as_float((unsigned)(kc[u] >> 32)); }
	for (int j = j0 + (int)threadIdx.x + 256 * kRedCache; j < j1; j += 256) {
		const unsigned long long key = __ldcg(slots + j);
		if (key != kSlotEmpty) { v[0] += 1.0; v[1] += (double)__uint_as_float((unsigned)(key >> 32)); }
	}
	red_block<2>(v, smem);
	red_store<2>(smem, c, 1, pe);
	red_mark(3);
	red_arrive(state, nblocks, 1, epoch, pe);
	// next iteration's packet schedule, built while the first barrier fills (the match stage of this iteration is complete): most
	// expensive class first, so the match kernel's tail is one cheap packet per warp instead of whatever happened to be drawn
	// last.  One atomic per class per warp (the lanes of a class are ranked by ballot).
	{
		const unsigned np = state->n_packets;
		unsigned off[8];
		unsigned run = 0;
#pragma unroll
		for (int k = 0; k < 8; k++) { off[k] = run; run += state->cls_n[k]; }
		const int lane = threadIdx.x & 31;
		for (unsigned p0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; p0 < np; p0 += gridDim.x * blockDim.x) {
			const unsigned p = p0 + lane;
			const unsigned k = p < np ? pk_cost_class(pk_cost[p], budget) : 8u;
			const unsigned grp = __match_any_sync(kFull, k);
			const int leader = __ffs(grp) - 1;
			unsigned base = 0;
			if (lane == leader && k < 8u) base = atomicAdd(&state->cls_fill[k], (unsigned)__popc(grp));
			base = __shfl_sync(kFull, base, leader);
			if (k < 8u) {
				unsigned o = 0;
#pragma unroll
				for (int kk = 0; kk < 8; kk++) if (k == (unsigned)kk) o = off[kk];
				const unsigned pos = o + base + __popc(grp & ((1u << lane) - 1u));
				if (pos < np) sched[pos] = p;
			}
		}
		if (blockIdx.x == 0 && threadIdx.x == 0) state->sched_valid = run == np ? 1u : 0u;     // every packet was classed exactly once
	}
	red_wait(state, nblocks, 1, epoch, pe);
	red_mark(4);
	red_total<2>(part_local, C, 1, s_cells, smem);
	red_mark(5);
	const double cnt = smem[0];
	const float nf = (float)cnt;                                     // (float)data.size()
	const float meanf = cnt >= 0.5 ? __fdiv_rn((float)smem[1], nf) : 0.0f;

	// ---- pass 2: sum of squared deviations -> sigma (icp.cpp:43-51) ----
	v[0] = 0;
#pragma unroll
	for (int u = 0; u < kRedCache; u++) if (kc[u] != kSlotEmpty) { const double dv = (double)__fsub_rn(__uint_as_float((unsigned)(kc[u] >> 32)), meanf); v[0] += dv * dv; }
	for (int j = j0 + (int)threadIdx.x + 256 * kRedCache; j < j1; j += 256) {
		const unsigned long long key = __ldcg(slots + j);
		if (key != kSlotEmpty) { const double dv = (double)__fsub_rn(__uint_as_float((unsigned)(key >> 32)), meanf); v[0] += dv * dv; }
	}
	red_block<1>(v, smem);
	red_store<1>(smem, c, 2, pe);
	red_mark(6);
	red_arrive(state, nblocks, 2, epoch, pe);
	red_wait(state, nblocks, 2, epoch, pe);
	red_mark(7);
	red_total<1>(part_local, C, 2, s_cells, smem);
	red_mark(8);
	const float sigma = cnt >= 0.5 ? sqrtf(__fdiv_rn((float)smem[0], nf)) : 0.0f;
	const float thr = __fmul_rn(2.5f, sigma);

	// ---- pass 3: 2.5 sigma gate + the 16 correspondence sums; every slot of the chunk is reset ----
	//   sums: [0] accepted count, [1..3] sum (p - q) (fp32 differences), [4..6] sum p, [7..15] sum q_a p_b
#pragma unroll
	for (int a = 0; a < 16; a++) v[a] = 0;
	auto take = [&](int j, unsigned long long key) {
		if (key == kSlotEmpty) return;
		slots[j] = kSlotEmpty;
		const float d2 = __uint_as_float((unsigned)(key >> 32));
		if (d2 > thr) return;
		const unsigned i = 0xFFFFFFFFu - (unsigned)key;
		const float px = verts1[3 * (size_t)j], py = verts1[3 * (size_t)j + 1], pz = verts1[3 * (size_t)j + 2];
		const float qx = verts2[3 * (size_t)i], qy = verts2[3 * (size_t)i + 1], qz = verts2[3 * (size_t)i + 2];
		v[0] += 1.0;
		v[1] += (double)__fsub_rn(px, qx); v[2] += (double)__fsub_rn(py, qy); v[3] += (double)__fsub_rn(pz, qz);
		v[4] += (double)px; v[5] += (double)py; v[6] += (double)pz;
		v[7] += (double)qx * (double)px; v[8] += (double)qx * (double)py; v[9] += (double)qx * (double)pz;
		v[10] += (double)qy * (double)px; v[11] += (double)qy * (double)py; v[12] += (double)qy * (double)pz;
		v[13] += (double)qz * (double)px; v[14] += (double)qz * (double)py; v[15] += (double)qz * (double)pz;
	};
#pragma unroll
	for (int u = 0; u < kRedCache; u++) take(j0 + (int)threadIdx.x + 256 * u, kc[u]);
	for (int j = j0 + (int)threadIdx.x + 256 * kRedCache; j < j1; j += 256) take(j, __ldcg(slots + j));
	red_block<16>(v, smem);
	red_store<16>(smem, c, 3, pe);
	red_mark(9);
	red_arrive(state, nblocks, 3, epoch, pe);
	if (blockIdx.x != 0) return;
	red_wait(state, nblocks, 3, epoch, pe);
	red_mark(10);
	// ---- block 0 of every rank: fold, solve, accumulate (identical arithmetic on identical numbers everywhere) ----
	red_total<16>(part_local, C, 3, s_cells, smem);
	red_mark(11);
	if (threadIdx.x == 0) {
		if (trace && trace_idx >= 0 && trace_idx < kTraceCap) { trace[trace_idx].n_matched = (int)(cnt + 0.5); trace[trace_idx].sigma = sigma; }
		double sums[16];
		for (int a = 0; a < 16; a++) sums[a] = smem[a];
		float T[3], Rk[9];
		const bool solved = icp_solve(sums, T, Rk);
		for (int a = 0; a < 3; a++) state->xf[a] = T[a];
		for (int a = 0; a < 9; a++) state->xf[3 + a] = Rk[a];
		icp_accumulate(state, T, Rk, solved, trace, trace_idx, sums);
		for (int k = 0; k < 8; k++) { state->cls_n[k] = 0; state->cls_fill[k] = 0; }
		state->red_bar = 0;
		state->red_epoch = epoch + 1u;
	}
	red_mark(12);
}
#ifdef LS3D_RED_TIMING
extern "C" int ls3d_debug_red_timing(unsigned long long *host16) {
	unsigned long long *d = nullptr;
	if (cudaMalloc(&d, 128) != cudaSuccess) return -1;
	k_red_timing_dump<<<1, 32>>>(d);
	const cudaError_t e = cudaMemcpy(host16, d, 128, cudaMemcpyDeviceToHost);
	cudaFree(d);
	return e == cudaSuccess ? 0 : -1;
}
#endif

// cross-rank rendezvous (sharded ICP, one block): every rank has reached this point of its stream.  ls3d_icp_set_source ends with
// it, so no rank's first match kernel can put keys into a peer's dedupe slots before that peer has re-initialised them.
__global__ void k_icp_peer_sync(IcpState *state, IcpPeers pe) {
	if (threadIdx.x != 0 || pe.world <= 1) return;
	const unsigned epoch = ld_volatile_u32(&state->red_epoch);
	__threadfence_system();
	for (int r = 0; r < pe.world; r++) if (r != pe.rank) st_volatile_u32(pe.flag[r] + pe.rank, epoch * 4u);
	for (int r = 0; r < pe.world; r++) {
		if (r == pe.rank) continue;
		unsigned spins = 0;
		while (ld_volatile_sys_u32(pe.flag[pe.rank] + r) < epoch * 4u) if (++spins > (1u << 26)) { atomicOr(&state->err, kErrScanSpin); break; }
	}
	__threadfence();
	state->red_epoch = epoch + 1u;
}

struct Pose12 { float v[12]; };     // R[9] then t[3], passed by value (no staging buffer to race on)

__global__ void k_icp_init_state(IcpState *st, Pose12 p) {
	for (int i = 0; i < 9; i++) st->R[i] = p.v[i];
	for (int i = 0; i < 3; i++) st->t[i] = p.v[9 + i];
	st->iters_applied = 0;
	st->err = 0;
	st->ticket_stats = 0;
	st->ticket_sums = 0;
	st->n_work = 0;
	st->blocks_done = 0;
	st->n_packets = 0;
	st->n_heavy = 0;
	st->n_light = 0;
	st->sched_valid = 0;
	st->hq_n = 0; st->hq_head = 0; st->hq_done = 0; st->light_done = 0;
	for (int i = 0; i < 8; i++) { st->cls_n[i] = 0; st->cls_fill[i] = 0; }
	st->red_bar = 0;                                            // red_epoch is deliberately left alone
	for (int i = 0; i < 12; i++) st->xf[i] = (i == 3 || i == 7 || i == 11) ? 1.0f : 0.0f;
}

}  // namespace ls3d

using namespace ls3d;

// ======================================================================================================
// ICP context
// ======================================================================================================
struct Ls3dIcp {
	int n1_max = 0, n2_max = 0;
	int n1 = 0, n2 = 0, i_begin = 0, i_end = 0;
	int G = 0, levels = 0;
	int iter = 0;                 // iterations whose match stage has been enqueued
	bool pending = false;         // sums of the last iteration not yet applied
	unsigned *dbg = nullptr;      // optional per-query work statistics (3 u32 per source point), see ls3d_icp_set_debug
	int sm_count = 148;
	const float *d_verts1 = nullptr;
	float *d_verts2 = nullptr;
	DevBuf grid, box, state, cell_start, nodes, cellbox, cell_of, rank_of, sorted, slots, scan_status, nn_idx, nn_d2, work, trace, small;
	DevBuf src_start, src_cell, src_rank, pk_desc, pk_cost, pk_sched, pk_heavy;   // Morton ordering of the source slice (work = the order itself), its packets, their cost and schedule
	bool order_valid = false;
	DevBuf red_part, red_flag;    // k_icp_reduce: partial table [3][kRedChunksMax][16] f64 and cross-rank flag words (both mapped by the peers when sharded)
	int world = 1, rank = 0;      // sharded ICP (ls3d_icp_set_peers): ranks sharing this call, and which one this is
	double *peer_part[8] = {};
	unsigned *peer_flag[8] = {};
	unsigned long long *peer_slots[8] = {};
	DevBuf own_v1, own_v2;        // device copies for the host-buffer API
	cudaStream_t up = nullptr;    // host-buffer API: the source cloud uploads here while the target is built
	cudaEvent_t ev_up = nullptr, ev_go = nullptr;
	float *pin = nullptr;         // pinned read-back: Rt[12] + status[4]
	cudaGraphExec_t graph = nullptr;
	int graph_iters = 0, graph_n1 = 0, graph_n2 = 0, graph_ib = 0, graph_ie = 0;
	const void *graph_v1 = nullptr, *graph_v2 = nullptr;
	int cand_iters = 0, cand_n1 = 0, cand_n2 = 0, cand_ib = 0, cand_ie = 0;      // the key of the previous ls3d_icp_run (graph built when it repeats)
	const void *cand_v1 = nullptr, *cand_v2 = nullptr;
};

static void icp_free(Ls3dIcp *c) {
	if (!c) return;
	DevBuf *bufs[] = {&c->grid, &c->box, &c->state, &c->cell_start, &c->nodes, &c->cellbox, &c->cell_of, &c->rank_of, &c->sorted, &c->slots,
		&c->scan_status, &c->nn_idx, &c->nn_d2, &c->work, &c->trace, &c->small, &c->own_v1, &c->own_v2, &c->src_start, &c->src_cell, &c->src_rank, &c->pk_desc, &c->pk_cost, &c->pk_sched, &c->pk_heavy, &c->red_part, &c->red_flag};
	for (DevBuf *b : bufs) b->release();
	if (c->pin) cudaFreeHost(c->pin);
	if (c->graph) cudaGraphExecDestroy(c->graph);
	if (c->ev_up) cudaEventDestroy(c->ev_up);
	if (c->ev_go) cudaEventDestroy(c->ev_go);
	if (c->up) cudaStreamDestroy(c->up);
	delete c;
}

extern "C" Ls3dIcp *ls3d_icp_create(int n1_max, int n2_max) {
	clear_error();
	if (!ensure_device()) return nullptr;
	if (n1_max <= 0 || n2_max < 0) { set_error("ls3d_icp_create: sizes must be positive"); return nullptr; }
	Ls3dIcp *c = new Ls3dIcp();
	int dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev);
	c->n1_max = n1_max;
	c->n2_max = n2_max;
	const size_t n1 = (size_t)n1_max, n2 = (size_t)std::max(n2_max, 1);
	bool ok = c->grid.reserve(sizeof(IcpGrid), "alloc grid") && c->box.reserve(sizeof(IcpBox), "alloc bbox") && c->state.reserve(sizeof(IcpState), "alloc state") &&
		c->cell_of.reserve(4 * n1, "alloc cells") && c->rank_of.reserve(4 * n1, "alloc ranks") && c->sorted.reserve(16 * n1, "alloc sorted target") &&
		c->slots.reserve(8 * n1, "alloc slots") &&
		c->nn_idx.reserve(4 * n2, "alloc nn index") && c->nn_d2.reserve(4 * n2, "alloc nn dist") && c->work.reserve(4 * n2 + 256, "alloc work list") &&
		c->trace.reserve(sizeof(Ls3dIcpTrace) * kTraceCap, "alloc trace") && c->small.reserve(256, "alloc small") &&
		c->src_cell.reserve(4 * n2, "alloc source cells") && c->src_rank.reserve(4 * n2, "alloc source ranks") &&
		c->red_part.reserve(sizeof(double) * 3 * kRedPhaseStride, "alloc reduction partials") && c->red_flag.reserve(256, "alloc reduction flags");
	ok = ok && cuda_ok(cudaHostAlloc((void **)&c->pin, 256, cudaHostAllocDefault), "alloc pinned read-back");
	if (!ok) { icp_free(c); return nullptr; }
	cudaMemset(c->trace.p, 0, sizeof(Ls3dIcpTrace) * kTraceCap);
	cudaMemset(c->state.p, 0, sizeof(IcpState));
	{
		// the cross-rank flag words start at 0: the first rendezvous must wait for a value they do not hold yet
		const unsigned one = 1u;
		cudaMemcpy(&c->state.as<IcpState>()->red_epoch, &one, sizeof(one), cudaMemcpyHostToDevice);
	}
	cudaMemset(c->red_flag.p, 0, 256);
	cudaMemset(c->red_part.p, 0, sizeof(double) * 3 * kRedPhaseStride);
	return c;
}

extern "C" void ls3d_icp_destroy(Ls3dIcp *c) { icp_free(c); }

static int pt_blocks(const Ls3dIcp *c, int n) { return std::max(1, std::min((n + 255) / 256, c->sm_count * 8)); }

extern "C" int ls3d_icp_set_target(Ls3dIcp *c, const void *d_verts1, int n1, void *stream) {
	if (!c || !d_verts1) { set_error("ls3d_icp_set_target: null argument"); return -1; }
	if (n1 <= 0 || n1 > c->n1_max) { set_error("ls3d_icp_set_target: n1=%d outside 1..%d (the reference throws on an empty target, nanoflann.h:904)", n1, c->n1_max); return -1; }
	cudaStream_t st = (cudaStream_t)stream;
	c->d_verts1 = (const float *)d_verts1;
	c->n1 = n1;
	c->order_valid = false;
	c->levels = n1 <= 32768 ? 6 : (n1 <= 1000000 ? 7 : 8);
	c->G = 1 << c->levels;
	const size_t cells = (size_t)c->G * c->G * c->G;
	const int scan_n = (int)cells + 1;
	const int scan_tiles = (scan_n + kTile - 1) / kTile;
	unsigned mask_total = 0;
	for (int l = 1; l <= c->levels; l++) { const unsigned n = (unsigned)(c->G >> l); mask_total += n * n * n; }
	if (!c->cell_start.reserve(4 * (cells + 8), "alloc cell starts") || !c->scan_status.reserve(8 * (size_t)(scan_tiles + 1), "alloc scan status") ||
		!c->nodes.reserve(8 * (size_t)(mask_total + 16), "alloc octree nodes") || !c->cellbox.reserve(8 * (cells + 8), "alloc cell records") ||
		!c->src_start.reserve(4 * (cells + 8), "alloc source cell starts") ||
		!c->pk_desc.reserve(8 * ((size_t)c->n2_max / 32 + (cells >> (3 * kPkRunLevel)) + 8), "alloc packet table") ||
		!c->pk_cost.reserve(4 * ((size_t)c->n2_max / 32 + (cells >> (3 * kPkRunLevel)) + 8), "alloc packet costs") ||
		!c->pk_sched.reserve(4 * ((size_t)c->n2_max / 32 + (cells >> (3 * kPkRunLevel)) + 8), "alloc packet schedule") ||
		!c->pk_heavy.reserve(4 * ((size_t)c->n2_max / 32 + (cells >> (3 * kPkRunLevel)) + 8), "alloc heavy packet queue")) return -1;
	IcpBox hb;
	for (int a = 0; a < 3; a++) { hb.mn[a] = 0xffffffffu; hb.mx[a] = 0u; }
	bool ok = cuda_ok(cudaMemcpyAsync(c->box.p, &hb, sizeof(hb), cudaMemcpyHostToDevice, st), "init bbox") &&
		cuda_ok(cudaMemsetAsync(c->cell_start.p, 0, 4 * (cells + 8), st), "clear cells") &&
		cuda_ok(cudaMemsetAsync(c->scan_status.p, 0, 8 * (size_t)(scan_tiles + 1), st), "clear scan status") &&
		cuda_ok(cudaMemsetAsync(c->small.p, 0, 256, st), "clear counters");
	if (!ok) return -1;
	unsigned *scan_counter = c->small.as<unsigned>();
	int *scan_err = c->small.as<int>() + 1;
	const int nb = pt_blocks(c, n1);
	k_icp_bbox<<<nb, 256, 0, st>>>(c->d_verts1, n1, c->box.as<IcpBox>());
	k_icp_grid_params<<<1, 1, 0, st>>>(c->box.as<IcpBox>(), c->grid.as<IcpGrid>(), c->G, c->levels);
	k_icp_count<<<nb, 256, 0, st>>>(c->d_verts1, n1, c->grid.as<IcpGrid>(), c->cell_start.as<unsigned>(), c->cell_of.as<unsigned>(), c->rank_of.as<unsigned>());
	k_exclusive_scan<<<std::min(scan_tiles, c->sm_count * 8), kScanThreads, 0, st>>>(c->cell_start.as<unsigned>(), scan_n, scan_counter, c->scan_status.as<unsigned long long>(), scan_err);
	k_icp_scatter<<<nb, 256, 0, st>>>(c->d_verts1, n1, c->cell_of.as<unsigned>(), c->rank_of.as<unsigned>(), c->cell_start.as<unsigned>(), c->sorted.as<float4>());
	k_icp_cellbox<<<std::min((unsigned)((cells + 255) / 256), (unsigned)c->sm_count * 16), 256, 0, st>>>(c->grid.as<IcpGrid>(), c->cell_start.as<unsigned>(), c->sorted.as<float4>(),
		c->cellbox.as<unsigned long long>(), (unsigned)cells);
	int launches = 7;
	int l = 1;
	for (; l <= c->levels; l++) {
		const unsigned dim = (unsigned)c->G >> l, nn = dim * dim * dim;
		if (nn <= 512) break;
		k_icp_nodebox<<<std::min((nn + 255) / 256, (unsigned)c->sm_count * 8), 256, 0, st>>>(c->grid.as<IcpGrid>(), l, c->cell_start.as<unsigned>(), c->cellbox.as<unsigned long long>(), c->nodes.as<unsigned long long>());
		launches++;
	}
	if (l <= c->levels) {
		k_icp_nodebox_top<<<1, 512, 0, st>>>(c->grid.as<IcpGrid>(), l, c->cell_start.as<unsigned>(), c->cellbox.as<unsigned long long>(), c->nodes.as<unsigned long long>());
		launches++;
	}
	k_fill_u64<<<nb, 256, 0, st>>>(c->slots.as<unsigned long long>(), n1, kSlotEmpty);
	if (!cuda_ok(cudaMemsetAsync(c->pk_heavy.p, 0xff, c->pk_heavy.cap, st), "clear heavy packet queue")) return -1;      // all entries kHvEmpty
	count_launch(launches);
	return cuda_ok(cudaGetLastError(), "target grid kernels") ? 0 : -1;
}

static IcpPeers icp_peers(Ls3dIcp *c);

extern "C" int ls3d_icp_set_source(Ls3dIcp *c, void *d_verts2, int n2, int i_begin, int i_end, const float *R0, const float *t0, void *stream) {
	if (!c || (!d_verts2 && n2 > 0) || !R0 || !t0) { set_error("ls3d_icp_set_source: null argument"); return -1; }
	if (n2 < 0 || n2 > c->n2_max || i_begin < 0 || i_end > n2 || i_begin > i_end) { set_error("ls3d_icp_set_source: bad sizes n2=%d slice [%d,%d) capacity %d", n2, i_begin, i_end, c->n2_max); return -1; }
	cudaStream_t st = (cudaStream_t)stream;
	c->d_verts2 = (float *)d_verts2;
	c->n2 = n2;
	c->i_begin = i_begin;
	c->i_end = i_end;
	c->iter = 0;
	c->pending = false;
	c->order_valid = false;
	if (n2 > 0 && !cuda_ok(cudaMemsetAsync(c->nn_idx.p, 0xff, 4 * (size_t)n2, st), "reset nn index")) return -1;     // no previous neighbour yet
	Pose12 pose;
	memcpy(pose.v, R0, 9 * sizeof(float));
	memcpy(pose.v + 9, t0, 3 * sizeof(float));
	k_icp_init_state<<<1, 1, 0, st>>>(c->state.as<IcpState>(), pose);
	count_launch(1);
	if (c->world > 1) { k_icp_peer_sync<<<1, 32, 0, st>>>(c->state.as<IcpState>(), icp_peers(c)); count_launch(1); }
	return cuda_ok(cudaGetLastError(), "k_icp_init_state") ? 0 : -1;
}

// Morton order of the slice's source points on the target grid (once per ls3d_icp_set_source; the source only moves by a
// small rigid transform afterwards, so the packets stay compact).  Returns kernels launched or -1.
static int icp_build_order(Ls3dIcp *c, cudaStream_t st) {
	const int n_slice = c->i_end - c->i_begin;
	if (n_slice <= 0) { c->order_valid = true; return 0; }
	const size_t cells = (size_t)c->G * c->G * c->G;
	const int scan_n = (int)cells + 1;
	const int scan_tiles = (scan_n + kTile - 1) / kTile;
	if (c->src_start.cap < 4 * (cells + 8)) { set_error("ICP: source ordering scratch not allocated (ls3d_icp_set_target first)"); return -1; }   // never allocate here: this may run under stream capture
	bool ok = cuda_ok(cudaMemsetAsync(c->src_start.p, 0, 4 * (cells + 8), st), "clear source cells") &&
		cuda_ok(cudaMemsetAsync(c->scan_status.p, 0, 8 * (size_t)(scan_tiles + 1), st), "clear scan status") &&
		cuda_ok(cudaMemsetAsync(c->small.p, 0, 256, st), "clear counters");
	if (!ok) return -1;
	const int nb = std::max(1, std::min((n_slice + 255) / 256, c->sm_count * 8));
	k_icp_count<<<nb, 256, 0, st>>>(c->d_verts2 + 3 * (size_t)c->i_begin, n_slice, c->grid.as<IcpGrid>(), c->src_start.as<unsigned>(), c->src_cell.as<unsigned>(), c->src_rank.as<unsigned>());
	k_exclusive_scan<<<std::min(scan_tiles, c->sm_count * 8), kScanThreads, 0, st>>>(c->src_start.as<unsigned>(), scan_n, c->small.as<unsigned>(), c->scan_status.as<unsigned long long>(), c->small.as<int>() + 1);
	k_icp_order_scatter<<<nb, 256, 0, st>>>(c->i_begin, n_slice, c->src_cell.as<unsigned>(), c->src_rank.as<unsigned>(), c->src_start.as<unsigned>(), c->work.as<unsigned>());
	const unsigned n_runs = (unsigned)(cells >> (3 * kPkRunLevel));
	if (!cuda_ok(cudaMemsetAsync(&c->state.as<IcpState>()->n_packets, 0, sizeof(unsigned), st), "clear packet count")) return -1;
	k_icp_packets<<<std::max(1u, std::min((n_runs + 255) / 256, (unsigned)c->sm_count * 8)), 256, 0, st>>>(c->src_start.as<unsigned>(), n_runs, c->pk_desc.as<uint2>(), c->state.as<IcpState>());
	count_launch(4);
	if (!cuda_ok(cudaGetLastError(), "source ordering kernels")) return -1;
	c->order_valid = true;
	return 4;
}

static SlotMap icp_slot_map(const Ls3dIcp *c) {
	SlotMap m = {};
	m.world = c->world;
	m.chunk = std::max(1, red_chunk_size(c->n1));
	m.per = std::max(1, (red_chunks(c->n1) + c->world - 1) / c->world);
	if (c->world > 1) for (int r = 0; r < c->world; r++) m.ptr[r] = c->peer_slots[r];
	else m.ptr[0] = const_cast<Ls3dIcp *>(c)->slots.as<unsigned long long>();
	return m;
}

static IcpPeers icp_peers(Ls3dIcp *c) {
	IcpPeers pe = {};
	pe.world = c->world;
	pe.rank = c->rank;
	if (c->world > 1) {
		for (int r = 0; r < c->world; r++) { pe.part[r] = c->peer_part[r]; pe.flag[r] = c->peer_flag[r]; pe.slots[r] = c->peer_slots[r]; }
	} else {
		pe.part[0] = c->red_part.as<double>(); pe.flag[0] = c->red_flag.as<unsigned>(); pe.slots[0] = c->slots.as<unsigned long long>();
	}
	return pe;
}

// Sharded ICP: `world` ranks (one process per GPU) run the same call on replicated clouds, each searching its slice of the source
// (ls3d_icp_set_source's [i_begin, i_end)) and owning a contiguous run of the target's reduction chunks and dedupe slots.
// slots / part / flag: for every rank r the device pointer, valid in THIS process, of rank r's ls3d_icp_slots / ls3d_icp_red_part /
// ls3d_icp_red_flag (CUDA-IPC mappings of the peers' allocations; entry `rank` is the local one).  world = 1 switches sharding off.
extern "C" int ls3d_icp_set_peers(Ls3dIcp *c, int world, int rank, void *const *slots, void *const *part, void *const *flag) {
	if (!c) { set_error("ls3d_icp_set_peers: null context"); return -1; }
	if (world <= 1) { c->world = 1; c->rank = 0; return 0; }
	if (world > 8 || rank < 0 || rank >= world || !slots || !part || !flag) { set_error("ls3d_icp_set_peers: world must be 1..8, rank inside it, pointer tables non-null"); return -1; }
	for (int r = 0; r < world; r++) {
		if (!slots[r] || !part[r] || !flag[r]) { set_error("ls3d_icp_set_peers: missing mapping for rank %d", r); return -1; }
		c->peer_slots[r] = (unsigned long long *)slots[r]; c->peer_part[r] = (double *)part[r]; c->peer_flag[r] = (unsigned *)flag[r];
	}
	if (c->peer_slots[rank] != c->slots.as<unsigned long long>() || c->peer_part[rank] != c->red_part.as<double>() || c->peer_flag[rank] != c->red_flag.as<unsigned>()) {
		set_error("ls3d_icp_set_peers: entry %d must be this context's own buffers", rank); return -1;
	}
	c->world = world;
	c->rank = rank;
	if (c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; }
	return 0;
}
extern "C" double *ls3d_icp_red_part(Ls3dIcp *c) { return c ? c->red_part.as<double>() : nullptr; }
extern "C" unsigned *ls3d_icp_red_flag(Ls3dIcp *c) { return c ? c->red_flag.as<unsigned>() : nullptr; }

static bool icp_use_pdl() {
	static const int v = getenv("LS3D_ICP_PDL") ? atoi(getenv("LS3D_ICP_PDL")) : 1;
	return v != 0;
}

// kernel launch with the programmatic-stream-serialization attribute: the kernel may become resident while its predecessor in the
// stream is still running; it orders itself with griddepcontrol.wait (or, for the block-wide match stage, with its queue protocol)
template <typename... KArgs, typename... Args>
static bool launch_dep(const char *what, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(block);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = icp_use_pdl() ? 1 : 0;
	return cuda_ok(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...), what);
}

// How many node / cell visits a warp spends on one packet before it hands it to the block-wide stage.  That stage exists for the
// TAIL: with ~2 packets per resident warp (2 x 213 k points) the few 100+-visit packets decide when the kernel ends, and a block
// finishes them 5x sooner.  It is the less efficient searcher, though (8 warps share one packet), so with many packets per warp —
// where the tail is a negligible share — the warps keep everything (measured at 2 x 2 M points: 1.12 ms / iteration with a budget
// of 64, 0.90 with 128, 0.84 without).
static unsigned icp_budget(const Ls3dIcp *c) {
	static const long long env = getenv("LS3D_PK_BUDGET") ? atoll(getenv("LS3D_PK_BUDGET")) : -1;      // tuning aid
	if (env >= 0) return (unsigned)std::min<long long>(env, 0x3fffffff);
	const long long n_packets = ((long long)c->i_end - c->i_begin + 31) / 32;
	const long long resident_warps = (long long)c->sm_count * 3 * kPkWarps;
	if (n_packets <= 4 * resident_warps) return kPkBudget;
	if (n_packets <= 8 * resident_warps) return 2 * kPkBudget;
	return 0x3fffffffu;
}

static int icp_launch_match(Ls3dIcp *c, int apply, int search, cudaStream_t st) {
	const int n_slice = c->i_end - c->i_begin;
	if (search && n_slice > 0) {
		if (!c->order_valid && icp_build_order(c, st) < 0) return -1;
		const int n_packets = (n_slice + 31) / 32;        // at least; the exact number (node-bounded runs) lives on the device
		// persistent grids: every block is resident from the start (the block-wide stage is a programmatic dependent launch that may
		// only begin once every block of the packet kernel has started)
		static int occ_light = 0, occ_heavy = 0;
		if (!occ_light) {
			if (!cuda_ok(cudaFuncSetAttribute(k_icp_match_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HvBlock)), "heavy stage shared memory") ||
				!cuda_ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_light, k_icp_match_packet, kPkWarps * 32, 0), "occupancy query") ||
				!cuda_ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_heavy, k_icp_match_heavy, kPkWarps * 32, sizeof(HvBlock)), "occupancy query")) return -1;
			occ_light = std::max(occ_light, 1);
			occ_heavy = std::max(occ_heavy, 1);
			if (getenv("LS3D_PK_OCC")) occ_light = std::max(1, std::min(occ_light, atoi(getenv("LS3D_PK_OCC"))));      // tuning aid: leave SM room for the block-wide stage
		}
		const int nb = std::max(1, std::min((n_packets + kPkWarps - 1) / kPkWarps + 8, c->sm_count * occ_light));
		const unsigned budget = icp_budget(c);
		if (!launch_dep("launch packet match", k_icp_match_packet, (unsigned)nb, kPkWarps * 32, 0, st, c->d_verts2, c->i_begin, c->i_end, apply, c->work.as<unsigned>(), c->pk_desc.as<uint2>(),
			c->pk_sched.as<unsigned>(), c->pk_cost.as<unsigned>(), c->grid.as<IcpGrid>(), c->cell_start.as<unsigned>(), c->nodes.as<unsigned long long>(), c->cellbox.as<unsigned long long>(),
			c->sorted.as<float4>(), c->d_verts1, icp_slot_map(c), c->state.as<IcpState>(), c->nn_idx.as<int>(), c->nn_d2.as<float>(), c->dbg, 1.0f, c->pk_heavy.as<unsigned>(), budget)) return -1;
		// the packets the warps give up on: one block each, consumed while the packet kernel's tail is still running
		if (!launch_dep("launch heavy stage", k_icp_match_heavy, (unsigned)std::max(1, std::min(n_packets, c->sm_count * occ_heavy)), kPkWarps * 32, sizeof(HvBlock), st,
			c->d_verts2, c->work.as<unsigned>(), c->pk_desc.as<uint2>(), c->pk_cost.as<unsigned>(), c->grid.as<IcpGrid>(), c->cell_start.as<unsigned>(), c->nodes.as<unsigned long long>(),
			c->cellbox.as<unsigned long long>(), c->sorted.as<float4>(), c->d_verts1, icp_slot_map(c), c->state.as<IcpState>(), c->nn_idx.as<int>(), c->nn_d2.as<float>(), c->dbg, 1.0f,
			c->pk_heavy.as<unsigned>(), budget)) return -1;
		count_launch(2);
	}
	// the points the packet kernel did not touch: everything when there is no search, otherwise what lies outside the slice
	int a0 = 0, a1 = 0, b0 = 0, b1 = 0;
	if (apply) {
		if (search && n_slice > 0) { a0 = 0; a1 = c->i_begin; b0 = c->i_end; b1 = c->n2; }
		else { a0 = 0; a1 = c->n2; }
		const int n_rest = (a1 - a0) + (b1 - b0);
		if (n_rest > 0) {
			k_icp_apply<<<pt_blocks(c, n_rest), 256, 0, st>>>(c->d_verts2, a0, a1, b0, b1, c->state.as<IcpState>());
			count_launch(1);
		}
	}
	return cuda_ok(cudaGetLastError(), "k_icp_match") ? 0 : -1;
}

extern "C" int ls3d_icp_match(Ls3dIcp *c, void *stream) {
	if (!c || !c->d_verts1 || !c->d_verts2) { set_error("ls3d_icp_match: target/source not set"); return -1; }
	const int r = icp_launch_match(c, c->pending ? 1 : 0, 1, (cudaStream_t)stream);
	c->pending = false;
	c->iter++;
	return r;
}

// mean / sigma / gate / correspondence sums / solve of the iteration just matched, in one launch (k_icp_reduce); with peers set, the
// ranks exchange their chunk partials inside the kernel
extern "C" int ls3d_icp_reduce(Ls3dIcp *c, void *stream) {
	if (!c || !c->d_verts1 || !c->d_verts2) { set_error("ls3d_icp_reduce: target/source not set"); return -1; }
	const int C = red_chunks(c->n1);
	const int per = (C + c->world - 1) / c->world;
	const int c0 = std::min(C, c->rank * per), c1 = std::min(C, (c->rank + 1) * per);
	if (!launch_dep("launch reduction", k_icp_reduce, (unsigned)std::max(1, c1 - c0), 256u, 0, (cudaStream_t)stream, c->slots.as<unsigned long long>(), c->n1, c1 > c0 ? c0 : -1, c->d_verts1,
		(const float *)c->d_verts2, c->state.as<IcpState>(), c->trace.as<Ls3dIcpTrace>(), c->iter - 1, c->pk_cost.as<unsigned>(), c->pk_sched.as<unsigned>(), icp_budget(c), icp_peers(c))) return -1;
	count_launch(1);
	c->pending = true;
	return cuda_ok(cudaGetLastError(), "k_icp_reduce") ? 0 : -1;
}

extern "C" int ls3d_icp_finish(Ls3dIcp *c, void *stream) {
	if (!c) { set_error("ls3d_icp_finish: null context"); return -1; }
	if (!c->pending) return 0;
	const int r = icp_launch_match(c, 1, 0, (cudaStream_t)stream);
	c->pending = false;
	return r;
}

static int icp_enqueue_all(Ls3dIcp *c, int maxIter, cudaStream_t st) {
	for (int it = 0; it < maxIter; it++) {
		if (ls3d_icp_match(c, st) < 0 || ls3d_icp_reduce(c, st) < 0) return -1;
	}
	return ls3d_icp_finish(c, st);
}

extern "C" int ls3d_icp_run(Ls3dIcp *c, int maxIter, void *stream) {
	if (!c || !c->d_verts1 || !c->d_verts2) { set_error("ls3d_icp_run: target/source not set"); return -1; }
	if (maxIter <= 0) return 0;
	cudaStream_t st = (cudaStream_t)stream;
	if (c->iter != 0 || c->pending) { set_error("ls3d_icp_run: call ls3d_icp_set_source first"); return -1; }
	// One CUDA graph per (shape, buffers, iteration count): 4*maxIter+1 kernel nodes replayed with one launch.
	const bool reuse = c->graph && c->graph_iters == maxIter && c->graph_n1 == c->n1 && c->graph_n2 == c->n2 && c->graph_v1 == c->d_verts1 && c->graph_v2 == c->d_verts2 &&
		c->graph_ib == c->i_begin && c->graph_ie == c->i_end;
	// Capturing and instantiating a graph only pays when the same call comes back (a server refining the same buffers, the bench).
	// A key seen for the first time runs as plain stream-ordered launches and is remembered; the graph is built when it repeats.
	// (The refine schedule concatenates a fresh target for each of its 16 calls: those never repeat.)
	const bool again = c->cand_iters == maxIter && c->cand_n1 == c->n1 && c->cand_n2 == c->n2 && c->cand_v1 == c->d_verts1 && c->cand_v2 == c->d_verts2 &&
		c->cand_ib == c->i_begin && c->cand_ie == c->i_end;
	if (!reuse && !again) {
		c->cand_iters = maxIter; c->cand_n1 = c->n1; c->cand_n2 = c->n2; c->cand_v1 = c->d_verts1; c->cand_v2 = c->d_verts2; c->cand_ib = c->i_begin; c->cand_ie = c->i_end;
		return icp_enqueue_all(c, maxIter, st);
	}
	if (!reuse) {
		if (c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; }
		cudaGraph_t graph = nullptr;
		if (cuda_ok(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal), "begin graph capture")) {
			const long long before = g_launches.load();
			const int r = icp_enqueue_all(c, maxIter, st);
			const cudaError_t e = cudaStreamEndCapture(st, &graph);
			g_launches.store(before);
			if (r == 0 && e == cudaSuccess && graph && cudaGraphInstantiate(&c->graph, graph, 0) == cudaSuccess) {
				c->graph_iters = maxIter; c->graph_n1 = c->n1; c->graph_n2 = c->n2; c->graph_v1 = c->d_verts1; c->graph_v2 = c->d_verts2;
				c->graph_ib = c->i_begin; c->graph_ie = c->i_end;
			} else {
				c->graph = nullptr;
				cudaGetLastError();
			}
			if (graph) cudaGraphDestroy(graph);
		} else {
			cudaGetLastError();
		}
		c->iter = 0;
		c->pending = false;
	}
	if (c->graph) {
		if (!cuda_ok(cudaGraphLaunch(c->graph, st), "launch ICP graph")) return -1;
		count_launch(3 * maxIter + 1 + 4);      // per iteration: match, stats, sums; once: source ordering (4) + final apply
		c->iter = maxIter;
		c->pending = false;
		return 0;
	}
	return icp_enqueue_all(c, maxIter, st);     // capture unavailable: plain stream-ordered launches
}

// Work statistics for tuning: d_stats (device, 3 u32 per source point, or NULL to switch off) receives, for the LAST match
// stage, the octree child steps, the candidate points scanned and (resume level + 1, 0 = finished in the first kernel).
extern "C" void ls3d_icp_set_debug(Ls3dIcp *c, void *d_stats) { if (c) c->dbg = (unsigned *)d_stats; }

extern "C" long long *ls3d_icp_slots(Ls3dIcp *c) { return c ? c->slots.as<long long>() : nullptr; }
extern "C" float *ls3d_icp_Rt(Ls3dIcp *c) { return c ? c->state.as<float>() : nullptr; }
extern "C" const int *ls3d_icp_nn_index(Ls3dIcp *c) { return c ? c->nn_idx.as<int>() : nullptr; }
extern "C" const float *ls3d_icp_nn_dist(Ls3dIcp *c) { return c ? c->nn_d2.as<float>() : nullptr; }
extern "C" Ls3dIcpTrace *ls3d_icp_trace_buf(Ls3dIcp *c) { return c ? c->trace.as<Ls3dIcpTrace>() : nullptr; }
extern "C" const int *ls3d_icp_status(Ls3dIcp *c) { return c ? reinterpret_cast<const int *>(c->state.as<float>() + 12) : nullptr; }

// ======================================================================================================
// Host-buffer (drop-in) entry points
// ======================================================================================================
static Ls3dIcp *g_icp = nullptr;

static Ls3dIcp *cached_icp(int n1, int n2) {
	if (g_icp && g_icp->n1_max >= n1 && g_icp->n2_max >= n2) return g_icp;
	int c1 = g_icp ? g_icp->n1_max : 0, c2 = g_icp ? g_icp->n2_max : 0;
	if (g_icp) { icp_free(g_icp); g_icp = nullptr; }
	c1 = std::max(c1, n1 + n1 / 8);
	c2 = std::max(c2, n2 + n2 / 8);
	g_icp = ls3d_icp_create(std::max(c1, 1), std::max(c2, 1));
	return g_icp;
}

extern "C" float ls3d_icp_trace(Point3f *verts1, Point3f *verts2, int nVerts1, int nVerts2, float *R, float *t, int maxIter, Ls3dIcpTrace *trace) {
	clear_error();
	const float ret = 1.0f;                                      // icp.cpp:85,176: `error` is never updated
	if (!verts1 || !verts2 || !R || !t) { set_error("ICP: null argument"); return ret; }
	if (maxIter <= 0) return ret;                                // the loop body never runs; nothing changes
	if (nVerts1 <= 0) { set_error("ICP: empty target cloud (the reference throws here, nanoflann.h:904)"); return ret; }
	if (nVerts2 <= 0) { set_error("ICP: empty source cloud"); return ret; }
	if (!ensure_device()) return ret;
	std::lock_guard<std::mutex> lk(api_mutex());
	cudaStream_t st = api_stream();
	if (!st) return ret;
	Ls3dIcp *c = cached_icp(nVerts1, nVerts2);
	if (!c) return ret;
	if (!c->own_v1.reserve(12 * (size_t)c->n1_max, "alloc target copy") || !c->own_v2.reserve(12 * (size_t)c->n2_max, "alloc source copy")) return ret;
	// the source goes up on a second stream while the target octree is being built (the build only needs the target)
	if (!c->up && (!cuda_ok(cudaStreamCreateWithFlags(&c->up, cudaStreamNonBlocking), "create upload stream") ||
		!cuda_ok(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming), "create event") || !cuda_ok(cudaEventCreateWithFlags(&c->ev_go, cudaEventDisableTiming), "create event"))) return ret;
	struct UpGuard { cudaStream_t s; ~UpGuard() { cudaStreamSynchronize(s); } } guard{c->up};       // the caller's buffer must not be read after we return
	bool ok = cuda_ok(cudaMemcpyAsync(c->own_v1.p, verts1, 12 * (size_t)nVerts1, cudaMemcpyHostToDevice, st), "upload target") &&
		cuda_ok(cudaEventRecord(c->ev_go, st), "order uploads") && cuda_ok(cudaStreamWaitEvent(c->up, c->ev_go, 0), "order uploads") &&      // target first on the wire
		cuda_ok(cudaMemcpyAsync(c->own_v2.p, verts2, 12 * (size_t)nVerts2, cudaMemcpyHostToDevice, c->up), "upload source") &&
		cuda_ok(cudaEventRecord(c->ev_up, c->up), "record source upload");
	if (!ok) return ret;
	if (ls3d_icp_set_target(c, c->own_v1.p, nVerts1, st) < 0) return ret;
	if (!cuda_ok(cudaStreamWaitEvent(st, c->ev_up, 0), "wait for the source upload")) return ret;
	if (ls3d_icp_set_source(c, c->own_v2.p, nVerts2, 0, nVerts2, R, t, st) < 0) return ret;
	if (ls3d_icp_run(c, maxIter, st) < 0) return ret;
	ok = cuda_ok(cudaMemcpyAsync(c->pin, c->state.p, sizeof(float) * 12 + sizeof(int) * 4, cudaMemcpyDeviceToHost, st), "read pose") &&
		cuda_ok(cudaMemcpyAsync(verts2, c->own_v2.p, 12 * (size_t)nVerts2, cudaMemcpyDeviceToHost, st), "read source");
	if (ok && trace) ok = cuda_ok(cudaMemcpyAsync(trace, c->trace.p, sizeof(Ls3dIcpTrace) * (size_t)std::min(maxIter, kTraceCap), cudaMemcpyDeviceToHost, st), "read trace");
	ok = ok && cuda_ok(cudaStreamSynchronize(st), "ICP");
	if (!ok) return ret;
	memcpy(R, c->pin, 9 * sizeof(float));
	memcpy(t, c->pin + 9, 3 * sizeof(float));
	const int *status = reinterpret_cast<const int *>(c->pin + 12);
	if (status[1]) set_error("ICP: device status flags 0x%x (%s)", status[1], (status[1] & kErrNoMatches) ? "an iteration had no accepted correspondences" : "internal");
	return ret;
}

extern "C" float ICP(Point3f *verts1, Point3f *verts2, int nVerts1, int nVerts2, float *R, float *t, int maxIter) {
	return ls3d_icp_trace(verts1, verts2, nVerts1, nVerts2, R, t, maxIter, nullptr);
}

extern "C" int ls3d_find_closest(const Point3f *verts1, int nVerts1, const Point3f *verts2, int nVerts2, unsigned long long *indices, float *distances) {
	clear_error();
	if (!verts1 || !verts2 || !indices || !distances || nVerts1 <= 0 || nVerts2 < 0) { set_error("ls3d_find_closest: bad arguments"); return -1; }
	if (nVerts2 == 0) return 0;
	if (!ensure_device()) return -1;
	std::lock_guard<std::mutex> lk(api_mutex());
	cudaStream_t st = api_stream();
	if (!st) return -1;
	Ls3dIcp *c = cached_icp(nVerts1, nVerts2);
	if (!c) return -1;
	if (!c->own_v1.reserve(12 * (size_t)c->n1_max, "alloc target copy") || !c->own_v2.reserve(12 * (size_t)c->n2_max, "alloc source copy")) return -1;
	const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z[3] = {0, 0, 0};
	bool ok = cuda_ok(cudaMemcpyAsync(c->own_v1.p, verts1, 12 * (size_t)nVerts1, cudaMemcpyHostToDevice, st), "upload target") &&
		cuda_ok(cudaMemcpyAsync(c->own_v2.p, verts2, 12 * (size_t)nVerts2, cudaMemcpyHostToDevice, st), "upload source");
	if (!ok) return -1;
	if (ls3d_icp_set_target(c, c->own_v1.p, nVerts1, st) < 0 || ls3d_icp_set_source(c, c->own_v2.p, nVerts2, 0, nVerts2, I, z, st) < 0) return -1;
	if (ls3d_icp_match(c, st) < 0) return -1;
	std::vector<int> idx(nVerts2);
	ok = cuda_ok(cudaMemcpyAsync(idx.data(), c->nn_idx.p, 4 * (size_t)nVerts2, cudaMemcpyDeviceToHost, st), "read nn index") &&
		cuda_ok(cudaMemcpyAsync(distances, c->nn_d2.p, 4 * (size_t)nVerts2, cudaMemcpyDeviceToHost, st), "read nn dist") &&
		cuda_ok(cudaStreamSynchronize(st), "find closest");
	if (!ok) return -1;
	for (int i = 0; i < nVerts2; i++) indices[i] = (unsigned long long)(long long)idx[i];
	return 0;
}
