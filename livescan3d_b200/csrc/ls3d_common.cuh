// ls3d_common.cuh — shared device/host helpers for libls3d_b200.so (sm_100a only).
//
// Conventions used by every kernel in this library:
//   * parity-critical fp32 arithmetic is written with __fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn so nvcc can never
//     contract it into FMAs: the reference (g++ -O2 -ffp-contract=off) evaluates a*b + c*d with separate
//     roundings, and cull / filter masks must be bit-exact (BASELINE.json north_star).
//   * 16-byte VertexC4ubV3f records move as uint4 (one LDG.128 / STG.128 per record).
//   * order-preserving compaction uses a single-pass decoupled look-back scan over 2048-element tiles
//     (tile ids handed out by an atomic counter, so a tile's predecessors are always running or done).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace ls3d {

constexpr int kTile = 2048;          // elements per compaction tile
constexpr int kScanThreads = 256;    // threads per compaction block (8 elements per thread)
constexpr unsigned kFull = 0xffffffffu;

// error bits accumulated in device status words
enum : int {
	kErrScanSpin = 1,        // look-back spin limit hit (should never happen)
	kErrCellOverflow = 2,    // more than 2^24-1 points in one voxel
	kErrNoMatches = 4,       // an ICP iteration had zero accepted correspondences
	kErrProbeLimit = 8,      // hash probe sequence exceeded the table (should never happen)
};

struct alignas(16) SensorDesc {
	int w, h, px, tile_begin;           // tile_begin: first compaction tile of this sensor (entry n_maps is the sentinel)
	long long depth_off, color_off;     // byte offsets into the packed depth / colour buffers
	long long pix_begin;                // first global pixel index of this sensor
	float cx, cy, fx, fy;               // IntrinsicCameraParameters (depthprocessing.h:90-98)
	float t[3];                         // WorldTranformation.t (depthprocessing.h:50-63)
	float R[9];                         // WorldTranformation.R, row-major
	float gox, goy, goz, ginv_h;        // voxel-hash origin and 1/cell edge for the neighbour-count filter
	unsigned tbl_off, tbl_mask;         // this sensor's region of the voxel hash table (power-of-two capacity)
	float org_rp;                       // organized neighbour count: camera-space radius r' (0 = path not applicable)
	int ray_off;                        // this sensor's ray table: xn[w] then yn[h] floats at rays + ray_off
};

// small device-resident control block, zeroed at the start of every run
struct FrameCtl {
	unsigned tile_counter_a;   // map/cull compaction
	unsigned tile_counter_b;   // filter compaction
	unsigned cursor;           // sorted-array range allocator
	unsigned work_counter;     // neighbour-count work stealing
	int n_final, n_culled, err, n_kept;   // n_kept: survivors counted by the neighbour-count kernel (known before compaction)
	int n_triangles;           // triangles emitted by the triangle stage (0 when it did not run)
	int pad[3];
};

// Programmatic dependent launch (launch_chain in ls3d_internal.h).  A chain of short kernels pays a launch latency per link when each
// waits for its predecessor to drain before it is even scheduled; launched programmatically, a kernel's blocks may move into the SMs
// once every block of the predecessor has STARTED (pdl_trigger), and the first thing they do is wait for that grid — and with it
// every earlier one — to complete and flush (griddepcontrol.wait).  Same ordering as a plain stream, minus the launch latencies;
// both instructions are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
	asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
	unsigned v;
	asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned *p, unsigned v) {
	asm volatile("st.volatile.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
	return v;
}
__device__ __forceinline__ unsigned warp_incl_scan(unsigned v, int lane) {
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		unsigned n = __shfl_up_sync(kFull, v, o);
		if (lane >= o) v += n;
	}
	return v;
}

// Decoupled look-back (Merrill & Garland): called by ALL 32 lanes of warp 0 of the block that owns `tile`.
// status[tile] = (flag << 32) | value with flag 0 = nothing yet, 1 = tile aggregate, 2 = inclusive prefix.
// Returns the exclusive prefix of `tile`.
// All tiles of these small grids are co-resident and publish their aggregates at about the same time, so a tile
// usually has to walk all the way back; the walk is therefore 256 predecessors wide per round trip (8 independent
// loads per lane in flight) instead of the textbook 32, which cuts the dependent L2 round trips 8-fold (ncu:
// 54 % of the map kernel's stall samples were the block barrier behind this walk).
// A bounded spin protects the GPU box from a hang if the protocol were ever broken; it raises kErrScanSpin.
__device__ __forceinline__ unsigned lookback_exclusive(unsigned long long *status, int tile, unsigned aggregate, int *err) {
	const int lane = threadIdx.x & 31;
	if (tile == 0) {
		if (lane == 0) st_volatile_u64(status, (2ull << 32) | aggregate);
		return 0u;
	}
	if (lane == 0) st_volatile_u64(status + tile, (1ull << 32) | aggregate);
	constexpr int kWide = 8;
	unsigned excl = 0;
	int idx = tile - 1;
	bool done = false;
	while (!done) {
		unsigned long long w[kWide];
		int spins = 0;
		for (;;) {
			bool missing = false;
#pragma unroll
			for (int u = 0; u < kWide; u++) {
				const int my = idx - lane - 32 * u;
				w[u] = my >= 0 ? ld_volatile_u64(status + my) : (2ull << 32);     // before tile 0: a zero prefix
				missing |= (w[u] >> 32) == 0;
			}
			if (!__any_sync(kFull, missing)) break;
			if (++spins > (1 << 22)) { if (lane == 0) atomicOr(err, kErrScanSpin); break; }
		}
#pragma unroll
		for (int u = 0; u < kWide; u++) {
			if (done) break;
			const unsigned flag = (unsigned)(w[u] >> 32), val = (unsigned)w[u];
			const unsigned pmask = __ballot_sync(kFull, flag == 2);
			if (pmask) {
				const int first = __ffs(pmask) - 1;
				excl += warp_sum(lane <= first ? val : 0u);
				done = true;
			} else {
				excl += warp_sum(val);
			}
		}
		idx -= 32 * kWide;
	}
	if (lane == 0) st_volatile_u64(status + tile, (2ull << 32) | (unsigned long long)(excl + aggregate));
	return excl;
}

// Block-wide exclusive scan of one count per thread (kScanThreads threads) + look-back for the tile base.
// Returns this thread's exclusive offset inside the tile; *tile_total and *tile_base are block-uniform.
// smem: at least 16 unsigned.
__device__ __forceinline__ unsigned tile_scan(unsigned cnt, unsigned *smem, unsigned long long *status, int tile, int *err,
	unsigned *tile_total, unsigned *tile_base) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned incl = warp_incl_scan(cnt, lane);
	if (lane == 31) smem[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		unsigned v = lane < (kScanThreads / 32) ? smem[lane] : 0u;
		const unsigned s = warp_incl_scan(v, lane);
		const unsigned total = __shfl_sync(kFull, s, kScanThreads / 32 - 1);
		const unsigned base = lookback_exclusive(status, tile, total, err);
		if (lane < (kScanThreads / 32)) smem[lane] = s - v;      // exclusive warp offsets
		if (lane == 0) { smem[8] = total; smem[9] = base; }
	}
	__syncthreads();
	*tile_total = smem[8];
	*tile_base = smem[9];
	return smem[warp] + incl - cnt;
}

// fp32 squared distance exactly as PointCloud::kdtree_distance evaluates it (icp.h:40-47, filter.h:38-45):
// d0*d0 + d1*d1 + d2*d2, every product and sum rounded separately.
__device__ __forceinline__ float dist2_ref(float qx, float qy, float qz, float px, float py, float pz) {
	const float d0 = __fsub_rn(qx, px), d1 = __fsub_rn(qy, py), d2 = __fsub_rn(qz, pz);
	return __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
}

// order-preserving float <-> unsigned map for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned f2ord(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __host__ __forceinline__ float ord2f(unsigned u) {
	u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
	return __uint_as_float(u);
#else
	float f; memcpy(&f, &u, 4); return f;
#endif
}

}  // namespace ls3d
