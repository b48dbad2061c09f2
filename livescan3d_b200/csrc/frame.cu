// frame.cu — the per-frame path on the device: depth -> camera space -> world transform -> bbox cull ->
// neighbour-count outlier filter -> multi-sensor merge, plus the host-buffer entry points that keep the
// reference's signatures.
//
// Reference behaviour reproduced (paths relative to the LiveScan3D tree):
//   createVertices            src/NativeUtils/depthprocessing.cpp:122-187  (+ RotatePoint :109-120)
//   formMesh (vertex part)    src/NativeUtils/depthprocessing.cpp:1578-1629
//   generateVerticesFromDepthMap / generateMeshFromDepthMaps   :1631-1657 / :1715-1792
//   filter / KNNeighbors      src/LiveScanClient/filter.cpp:36-81 / :19-34
//
// Kernels (all sensors of a run are batched into every launch; "g" is a point's index in the culled cloud):
//   K1o k_organized_count the filter's verdict per PIXEL for clouds that come from depth images (see its header): 32x16 pixel tiles,
//                          world positions of tile + halo staged in shared memory (two pixels per packed-fp32 evaluation), packed
//                          distance tests; leaves keep bytes and per-tile survivor counts for K1.  The normal frame is
//                          k_zero_control -> K1o -> K1, chained as programmatic dependent launches (launch_chain): each kernel's
//                          blocks move in while its predecessor's last blocks run and wait for it before they consume its output
//   K1 k_map_cull_compact  per 2048-pixel tile: u16 depth + RGB in, pinhole map, +t, R*, strict cull, stable
//                          compaction through a decoupled look-back scan, 16-byte records out (HBM-bound)
//   K2 k_voxel_insert      per point: its run of 8 voxels along x (30-bit key) found or claimed in a per-sensor open-addressing
//                          hash of 64-byte slots, run total and voxel count incremented (warp-aggregated); slot and rank per point
//   K3 k_bucket_alloc      one contiguous range of the sorted array per occupied run (block scan over the elected points, one
//                          atomic per block); voxel counts -> inclusive prefix sums in the slot
//   K4 k_cell_scatter      counting-sort scatter of (x,y,z,g) into run / voxel order
//   K5 k_neighbour_count   one lane per query, input order: 9 row look-ups (one 32-byte load each; 18 when the row straddles two
//                          runs), then a private cursor over the contiguous candidate ranges, two candidates per 32-byte load,
//                          leaving at k (the filter keeps i  <=>  its k-th NN distance <= thr  <=>  count >= k); k_voxel_cleanup
//                          zeroes the slots the run's points touched (no memset of the table)
//   K6 k_filter_compact    stable compaction of the survivors (look-back scan again); sensor order is the
//                          merge order, so the merged cloud is produced here with no extra copy; can store
//                          to several destinations (peer-mapped buffers) for the multi-GPU merge
#include "ls3d_common.cuh"
#include "ls3d_internal.h"
#include "../../include/ls3d.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ls3d {

constexpr int kCellBits = 11;                 // voxel coordinate bits per axis: (z, y, x >> 3) of a run of 8 voxels is a 30-bit key
constexpr int kCellMax = (1 << kCellBits) - 1;
constexpr unsigned long long kCountMask = 0xFFFFFFull;      // points per run of voxels (and per call)
constexpr int kCountWarps = 4;                // warps per block in K5
constexpr int kMaxPeers = 8;

struct PeerDst { int n; uint4 *ptr[kMaxPeers]; };

// hash of a run (bx = x >> 3, y, z): one multiplicative term per axis, so the nine rows around a voxel cost two XORs each once the
// three y- and three z-terms are known; the final shift-xor spreads the high bits into the table index
__device__ __forceinline__ unsigned hash_x(int bx) { return (unsigned)bx * 0x9E3779B1u; }
__device__ __forceinline__ unsigned hash_y(int y) { return (unsigned)y * 0x85EBCA77u; }
__device__ __forceinline__ unsigned hash_z(int z) { return (unsigned)z * 0xC2B2AE3Du; }
__device__ __forceinline__ unsigned hash_mix(unsigned h) { return h ^ (h >> 15); }
__device__ __forceinline__ int cell_coord(float x, float o, float inv_h) {
	// monotone in x: fp32 subtract, fp32 multiply by a positive constant, floor, clamp
	float u = floorf(__fmul_rn(__fsub_rn(x, o), inv_h));
	u = fminf(fmaxf(u, 0.0f), (float)kCellMax);
	return (int)u;
}

// ------------------------------------------------------------------------------------------------------
// the per-pixel arithmetic, in the reference's evaluation order (createVertices, depthprocessing.cpp:148-162,
// RotatePoint :109-120).  One definition shared by every kernel that needs a pixel's world position, so the
// organized neighbour count sees bit-identical coordinates to the ones the cloud carries.
// ------------------------------------------------------------------------------------------------------
struct PixelXform {
	float cx, cy, fx, fy, t0, t1, t2, r0, r1, r2, r3, r4, r5, r6, r7, r8;
	float minX, minY, minZ, maxX, maxY, maxZ;
};

__device__ __forceinline__ PixelXform load_xform(const SensorDesc *__restrict__ sd, int s, const float *b) {
	PixelXform m;
	m.cx = sd[s].cx; m.cy = sd[s].cy; m.fx = sd[s].fx; m.fy = sd[s].fy;
	m.t0 = sd[s].t[0]; m.t1 = sd[s].t[1]; m.t2 = sd[s].t[2];
	m.r0 = sd[s].R[0]; m.r1 = sd[s].R[1]; m.r2 = sd[s].R[2]; m.r3 = sd[s].R[3]; m.r4 = sd[s].R[4];
	m.r5 = sd[s].R[5]; m.r6 = sd[s].R[6]; m.r7 = sd[s].R[7]; m.r8 = sd[s].R[8];
	m.minX = b[0]; m.minY = b[1]; m.minZ = b[2]; m.maxX = b[3]; m.maxY = b[4]; m.maxZ = b[5];
	return m;
}

// v / 1000.0f, correctly rounded, for every integer v in [0, 65535] — i.e. bit-identical to the reference's `val / 1000.0f` on a u16
// depth (depthprocessing.cpp:149) — in three instructions instead of the ~10 of a general IEEE division: one Newton correction of
// v * fl(1/1000) with the exact FMA remainder.  Not a general identity: it is checked exhaustively over all 65536 inputs, on the
// CPU when this was written and on the device by ls3d_selftest() (tests/test_gpu_parity.py).
__device__ __forceinline__ float div1000_u16(float v) {
	const float r = 1.0f / 1000.0f;
	const float q0 = __fmul_rn(v, r);
	const float rem = __fmaf_rn(-q0, 1000.0f, v);
	return __fmaf_rn(rem, r, q0);
}

__global__ void k_selftest_div1000(int *mismatches) {
	const int d = blockIdx.x * blockDim.x + threadIdx.x;
	if (d < 65536 && __float_as_uint(div1000_u16((float)d)) != __float_as_uint(__fdiv_rn((float)d, 1000.0f))) atomicAdd(mismatches, 1);
}

// returns true when the pixel yields a vertex (non-zero depth and inside the strict cull box).
// xn = (x - cx) / fx and yn = (cy - y) / fy come from the per-sensor RAY TABLE (w + h floats, computed once per
// parameter set on the host with the same two IEEE fp32 operations the reference performs per pixel,
// depthprocessing.cpp:151-152), so only Z = d / 1000 is divided per pixel and the result is bit-identical.
__device__ __forceinline__ bool map_pixel(const PixelXform &m, float xn, float yn, unsigned d, float &wx, float &wy, float &wz) {
	if (d == 0) return false;
	const float val = (float)d;
	float Z = div1000_u16(val);
	float X = __fmul_rn(xn, Z);
	float Y = __fmul_rn(yn, Z);
	X = __fadd_rn(X, m.t0); Y = __fadd_rn(Y, m.t1); Z = __fadd_rn(Z, m.t2);
	wx = __fadd_rn(__fadd_rn(__fmul_rn(X, m.r0), __fmul_rn(Y, m.r1)), __fmul_rn(Z, m.r2));
	wy = __fadd_rn(__fadd_rn(__fmul_rn(X, m.r3), __fmul_rn(Y, m.r4)), __fmul_rn(Z, m.r5));
	wz = __fadd_rn(__fadd_rn(__fmul_rn(X, m.r6), __fmul_rn(Y, m.r7)), __fmul_rn(Z, m.r8));
	return !(wx < m.minX || wx > m.maxX || wy < m.minY || wy > m.maxY || wz < m.minZ || wz > m.maxZ);
}

struct Bounds6 { float v[6]; };

// ---- packed-fp32 variant (sm_100 FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per issue slot) ----
// The tile is kept as three float planes, so one LDS.64 yields the same coordinate of two horizontally adjacent candidates;
// the query is broadcast into both halves and every step of d0*d0 + d1*d1 + d2*d2 runs on the pair: 3 FADD2 (differences),
// 3 FMUL2 (squares), 2 adds.  Each half is rounded exactly like the scalar sequence, so the masks stay bit-exact.  The adds are
// issued as fma(a, 1.0f, b) == fl(a + b) with the 1.0f taken from a kernel argument: ptxas contracts a packed mul.rn + add.rn
// into FFMA2 even under --fmad=false (checked in SASS), which would skip the rounding of the product; it cannot do that to an
// fma whose multiplier it does not know.  Pairs start at even columns, so a window of reach HX is covered by HX+1 pairs (one
// extra real candidate on one side, which keeps the count exact); the halo is staged one column wider for it.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// map_pixel for two horizontally adjacent pixels of one row at once: dpair = their two u16 depths (low half = left pixel), xn2 their two
// ray-table entries.  Every half is rounded exactly like map_pixel's scalar sequence (the two FMAs of div1000_u16 are FMAs there too;
// -q0 * 1000 == q0 * -1000 exactly), so the positions are bit-identical; a sum that follows a product is again fma(a, 1.0f, b) with the
// opaque 1.0f.  Returns bit 0 / bit 1 = the left / right pixel yields a vertex.  The halo staging of the organized count runs on this:
// a third of the instructions per staged pixel.
__device__ __forceinline__ unsigned map_pixel_pair_xy(const PixelXform &m, f32x2 xn2, f32x2 yn2, unsigned dpair, f32x2 one2, f32x2 &wx, f32x2 &wy, f32x2 &wz) {
	const float r = 1.0f / 1000.0f;
	const f32x2 v = pk((float)(dpair & 0xffffu), (float)(dpair >> 16));
	const f32x2 q0 = mul2(v, pk(r, r));
	const f32x2 rem = fma2(q0, pk(-1000.0f, -1000.0f), v);
	f32x2 Z = fma2(rem, pk(r, r), q0);
	f32x2 X = mul2(xn2, Z), Y = mul2(yn2, Z);
	X = fma2(X, one2, pk(m.t0, m.t0)); Y = fma2(Y, one2, pk(m.t1, m.t1)); Z = fma2(Z, one2, pk(m.t2, m.t2));
	wx = fma2(fma2(mul2(X, pk(m.r0, m.r0)), one2, mul2(Y, pk(m.r1, m.r1))), one2, mul2(Z, pk(m.r2, m.r2)));
	wy = fma2(fma2(mul2(X, pk(m.r3, m.r3)), one2, mul2(Y, pk(m.r4, m.r4))), one2, mul2(Z, pk(m.r5, m.r5)));
	wz = fma2(fma2(mul2(X, pk(m.r6, m.r6)), one2, mul2(Y, pk(m.r7, m.r7))), one2, mul2(Z, pk(m.r8, m.r8)));
	float x0, x1, y0, y1, z0, z1;
	upk(wx, x0, x1); upk(wy, y0, y1); upk(wz, z0, z1);
	unsigned ok = 0;
	if ((dpair & 0xffffu) && !(x0 < m.minX || x0 > m.maxX || y0 < m.minY || y0 > m.maxY || z0 < m.minZ || z0 > m.maxZ)) ok |= 1u;
	if ((dpair >> 16) && !(x1 < m.minX || x1 > m.maxX || y1 < m.minY || y1 > m.maxY || z1 < m.minZ || z1 > m.maxZ)) ok |= 2u;
	return ok;
}
__device__ __forceinline__ unsigned map_pixel_pair(const PixelXform &m, f32x2 xn2, float yn, unsigned dpair, f32x2 one2, f32x2 &wx, f32x2 &wy, f32x2 &wz) {
	return map_pixel_pair_xy(m, xn2, pk(yn, yn), dpair, one2, wx, wy, wz);
}


__device__ __forceinline__ unsigned color_line_mask(int k) {          // threads (8 pixels = 24 bytes each) of a warp whose bytes touch line k of its 768
	const int lo = 128 * k / 24, hi = (128 * k + 127) / 24;
	return ((2u << hi) - 1u) & ~((1u << lo) - 1u);
}

// ------------------------------------------------------------------------------------------------------
// K1: map + world transform + cull (+ optional per-pixel keep mask) + stable compaction
// ------------------------------------------------------------------------------------------------------
// Two passes over the thread's 8 pixels: validity first (one bit each), then — once the tile's base is known
// from the look-back scan — the coordinates are recomputed and the 16-byte records stored straight to their
// final place.  Recomputing (~40 instructions per pixel) is cheaper than holding 24 coordinates across the scan:
// the round-1 profile showed 74-80 registers, 3 blocks/SM, two waves and 57 % barrier stalls for the staged
// version; this one keeps every tile of an 8-sensor frame resident in one wave.
// kKeepMask: AND the organized neighbour count's per-pixel mask into the validity test.  In that mode the count kernel
// has already left every tile's survivor count in tile_count[], so a tile's base is a plain sum over its predecessors
// (no inter-block dependency at all) and tiles are assigned statically; otherwise the base comes from the look-back scan.
#ifndef LS3D_MAP_MINBLOCKS
#define LS3D_MAP_MINBLOCKS 6
#endif
#ifndef LS3D_MAP_PAIRS
#define LS3D_MAP_PAIRS 1
#endif
template <bool kWriteD2V, bool kKeepMask>
__global__ void __launch_bounds__(kScanThreads, LS3D_MAP_MINBLOCKS) k_map_cull_compact(
	const uint8_t *__restrict__ depth, const uint8_t *__restrict__ colors, const SensorDesc *__restrict__ sd, const unsigned short *__restrict__ tile_sensor,
	const float *__restrict__ rays, int s_first, int s_end, Bounds6 bnd, FrameCtl *ctl, unsigned long long *status, int *culled_starts,
	uint4 *__restrict__ out, const int *__restrict__ d_out_offset, int *__restrict__ d2v, const uint8_t *__restrict__ keep_px,
	const unsigned *__restrict__ tile_count, PeerDst peers, int tile_lo, int tile_hi, float one)
{
	__shared__ unsigned sm[16];
	__shared__ int s_tile;
	// look-back variant: the world positions pass 1 computes are parked here (thread-private slots, [coordinate][pixel][thread]:
	// conflict-free, no barrier) and read back by pass 2 — 48 shared-memory accesses instead of ~360 instructions of recomputation
	// per thread, without the 24 live registers that made the first staged version lose a resident block
	__shared__ float s_pos[kKeepMask ? 1 : 3][kKeepMask ? 1 : 8][kKeepMask ? 1 : kScanThreads];
	// multi-GPU merge (peers.n > 0; launched with kTile * 16 bytes of dynamic shared memory): a tile's records are staged here
	// and then copied to every rank's buffer by whole warps — 512 contiguous bytes per store instruction; a thread storing its
	// own few records one by one over NVLink ran at a fifth of the link rate
	extern __shared__ __align__(16) uint4 s_stage[];
	// the tile's 6144 colour bytes, fetched as whole 128-byte lines by 16-byte lanes (only the lines under a survivor): when the
	// colours sit in the caller's page-locked host buffer every request crosses PCIe, and a request per 128-byte line moves four
	// times the bytes per outstanding read of the 8-byte-per-thread gather this replaces (7 GB/s measured for that one)
	__shared__ __align__(16) uint4 s_col[kTile * 3 / 16];

	pdl_trigger();             // a kernel chained behind this one (k_voxel_insert, k_triangles) may move in; it waits for our completion
	const int tile0 = sd[s_first].tile_begin;
	const int ntiles = sd[s_end].tile_begin - tile0;
	// the output offset may come from the kernel right before us (the multi-GPU count exchange): with a keep mask it is read behind the wait
	int out_off = (!kKeepMask && d_out_offset) ? *d_out_offset : 0;
	const int tid = threadIdx.x;
	// kKeepMask only: this launch places tiles [tile_lo, tile_hi) of the run (bases still count from the run's first tile), so
	// the host path can merge sensor by sensor as the colours arrive; the look-back variant always walks the whole run
	int static_tile = tile_lo + blockIdx.x;
	bool waited = false;

	for (;;) {
		int tile;
		if (kKeepMask) {
			tile = static_tile;
			static_tile += gridDim.x;
		} else {
			if (tid == 0) s_tile = (int)atomicAdd(&ctl->tile_counter_a, 1u);
			__syncthreads();
			tile = s_tile;
		}
		if (tile >= (kKeepMask ? tile_hi : ntiles)) break;
		const int s = tile_sensor[tile + tile0];
		const int w = sd[s].w, px = sd[s].px;
		const int p0 = (tile + tile0 - sd[s].tile_begin) * kTile + tid * 8;
		const int rem = px - p0;
		const PixelXform m = load_xform(sd, s, bnd.v);

		// ---- 8 u16 depths: one LDG.128 (kept packed in 4 registers), issued before the keep mask is looked at: an input, not something the
		// count kernel wrote, so with a programmatic launch it is already on its way while that kernel's last blocks finish ----
		uint4 dq = make_uint4(0u, 0u, 0u, 0u);
		const uint8_t *dp = depth + sd[s].depth_off + 2ll * p0;
		{
			if (rem >= 8 && (((uintptr_t)dp) & 15) == 0) {
				dq = __ldg(reinterpret_cast<const uint4 *>(dp));
			} else {
				unsigned t[8];
#pragma unroll
				for (int j = 0; j < 8; j++) t[j] = (j < rem) ? (unsigned)__ldg(reinterpret_cast<const unsigned short *>(dp) + j) : 0u;
				dq = make_uint4(t[0] | (t[1] << 16), t[2] | (t[3] << 16), t[4] | (t[5] << 16), t[6] | (t[7] << 16));
			}
		}
		if (kKeepMask && !waited) {
			// launched programmatically behind k_organized_count (launch_map): everything above ran beside its tail; its keep bytes and
			// tile counts are visible from here on.  A no-op for a plain launch.
			pdl_wait();
			waited = true;
			if (d_out_offset) out_off = *d_out_offset;
		}
		// ---- the organized count's verdicts for the 8 pixels: one LDG.64, bytes (0/1) -> bits by a multiply ----
		unsigned keepm = 0xffu;
		if (kKeepMask) {
			const uint8_t *kp = keep_px + sd[s].pix_begin + p0;
			keepm = 0;
			if (rem >= 8 && (((uintptr_t)kp) & 7) == 0) {
				const uint2 f = __ldg(reinterpret_cast<const uint2 *>(kp));
				constexpr unsigned kGather = (1u << 24) | (1u << 17) | (1u << 10) | (1u << 3);      // byte j (0 or 1) -> bit 24 + j of the product
				keepm = ((f.x * kGather) >> 24 & 0xfu) | ((f.y * kGather) >> 20 & 0xf0u);
			} else {
#pragma unroll
				for (int j = 0; j < 8; j++) if (j < rem && __ldg(kp + j)) keepm |= 1u << j;
			}
		}
		auto depth_of = [&](int j) -> unsigned {
			const unsigned wsel = (j < 4) ? ((j < 2) ? dq.x : dq.y) : ((j < 6) ? dq.z : dq.w);
			return (j & 1) ? (wsel >> 16) : (wsel & 0xffffu);
		};

		// ---- pass 1 (rolled: registers, not instructions, are what this kernel is short of): which pixels yield a vertex ----
		const int y0 = p0 / w, x0 = p0 - y0 * w;
		const float *xray = rays + sd[s].ray_off, *yray = xray + w;
		unsigned valid = 0;
		if (kKeepMask) {
			valid = keepm;            // the count kernel only keeps pixels that have a vertex
		} else if (LS3D_MAP_PAIRS && !(w & 1) && w >= 8 && rem >= 8) {
			// even width: four packed pair evaluations (see pass 2 of the keep-mask variant), bit-identical to map_pixel
			const f32x2 one2 = pk(one, one);
#pragma unroll 1
			for (int q = 0; q < 4; q++) {
				int x = x0 + 2 * q, y = y0;
				if (x >= w) { x -= w; y++; }
				const float2 xn = __ldg(reinterpret_cast<const float2 *>(xray + x));
				f32x2 ax, ay, az;
				const unsigned ok = map_pixel_pair(m, pk(xn.x, xn.y), __ldg(yray + y), q == 0 ? dq.x : q == 1 ? dq.y : q == 2 ? dq.z : dq.w, one2, ax, ay, az);
				float a0, a1, b0, b1, g0, g1;
				upk(ax, a0, a1); upk(ay, b0, b1); upk(az, g0, g1);
				if (ok & 1u) { s_pos[0][2 * q][tid] = a0; s_pos[1][2 * q][tid] = b0; s_pos[2][2 * q][tid] = g0; }
				if (ok & 2u) { s_pos[0][2 * q + 1][tid] = a1; s_pos[1][2 * q + 1][tid] = b1; s_pos[2][2 * q + 1][tid] = g1; }
				valid |= ok << (2 * q);
			}
		} else if (rem > 0) {
			int x = x0, y = y0;
			float yn = __ldg(yray + y);
#pragma unroll 1
			for (int j = 0; j < 8; j++) {
				float wx, wy, wz;
				if (map_pixel(m, __ldg(xray + x), yn, depth_of(j), wx, wy, wz)) {
					valid |= 1u << j;
					if (!kKeepMask) { s_pos[0][j][tid] = wx; s_pos[1][j][tid] = wy; s_pos[2][j][tid] = wz; }
				}
				if (++x == w) { x = 0; y++; yn = __ldg(yray + min(y, sd[s].h - 1)); }
			}
		}

		// ---- colours of a full, 16-byte aligned tile: line k (of the warp's six) is wanted when a thread whose 24 bytes touch it emits ----
		const uint8_t *ctile = colors + sd[s].color_off + 3ll * (p0 - tid * 8);
		const bool cvec = px - (p0 - tid * 8) >= kTile && (((uintptr_t)ctile) & 15) == 0;          // block-uniform
		if (cvec) {
			const int lane = tid & 31, warp = tid >> 5;
			const unsigned act = __ballot_sync(kFull, valid != 0);
			const uint4 *src = reinterpret_cast<const uint4 *>(ctile) + warp * 48;
			uint4 *dstc = s_col + warp * 48;
			__syncwarp();                          // the previous tile's readers of this warp's 768 bytes are done
			if (act & color_line_mask(lane >> 3)) dstc[lane] = __ldg(src + lane);
			if (lane < 16 && (act & color_line_mask(4 + (lane >> 3)))) dstc[32 + lane] = __ldg(src + 32 + lane);
		}

		// ---- stable compaction ----
		const unsigned cnt = __popc(valid);
		unsigned total, base, off;
		if (kKeepMask) {
			// base = survivors of all earlier tiles: 256 threads sum tile_count[tile0 .. tile0+tile) (no look-back, no waiting)
			unsigned part = 0;
			for (int t = tid; t < tile; t += kScanThreads) part += __ldg(tile_count + tile0 + t);
			const int lane = tid & 31, warp = tid >> 5;
			const unsigned incl = warp_incl_scan(cnt, lane);
			part = warp_sum(part);
			__syncthreads();                       // sm[] may still be read by the previous iteration
			if (lane == 31) sm[warp] = incl;
			if (lane == 0) sm[8 + warp] = part;
			__syncthreads();
			unsigned woff = 0, b = 0, tot = 0;
#pragma unroll
			for (int i = 0; i < kScanThreads / 32; i++) { b += sm[8 + i]; tot += sm[i]; if (i < warp) woff += sm[i]; }
			base = b; total = tot; off = woff + incl - cnt;
		} else {
			off = tile_scan(cnt, sm, status, tile, &ctl->err, &total, &base);
		}

		// ---- pass 2: recompute and store the records (24 colour bytes: three 8-byte reads of the staged lines) ----
		__syncwarp();
		if (valid) {
			unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
			const uint8_t *cp = colors + sd[s].color_off + 3ll * p0;
			if (cvec) {
				const uint2 *q = reinterpret_cast<const uint2 *>(s_col) + 3 * tid;
				const uint2 a = q[0], b = q[1], c = q[2];
				c0 = a.x; c1 = a.y; c2 = b.x; c3 = b.y; c4 = c.x; c5 = c.y;
			} else if (rem >= 8 && (((uintptr_t)cp) & 7) == 0) {
				const uint2 a = __ldg(reinterpret_cast<const uint2 *>(cp));
				const uint2 b = __ldg(reinterpret_cast<const uint2 *>(cp) + 1);
				const uint2 c = __ldg(reinterpret_cast<const uint2 *>(cp) + 2);
				c0 = a.x; c1 = a.y; c2 = b.x; c3 = b.y; c4 = c.x; c5 = c.y;
			} else {
				unsigned cw[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
				for (int j = 0; j < 8; j++)
					if ((valid >> j) & 1) {
#pragma unroll
						for (int c = 0; c < 3; c++) {
							const int bi = 3 * j + c;
							cw[bi >> 2] |= (unsigned)__ldg(cp + bi) << (8 * (bi & 3));
						}
					}
				c0 = cw[0]; c1 = cw[1]; c2 = cw[2]; c3 = cw[3]; c4 = cw[4]; c5 = cw[5];
			}
			size_t o = (size_t)out_off + base + off;
			// bytes 3j, 3j+1, 3j+2 of the 24-byte colour block -> R,G,B,255, then the record goes to its final place (or to the staging tile)
			auto emit = [&](int j, float wx, float wy, float wz) {
				const int bi = 3 * j, wi = bi >> 2;
				const unsigned lo = wi == 0 ? c0 : wi == 1 ? c1 : wi == 2 ? c2 : wi == 3 ? c3 : wi == 4 ? c4 : c5;
				const unsigned hi = wi == 0 ? c1 : wi == 1 ? c2 : wi == 2 ? c3 : wi == 3 ? c4 : c5;
				const unsigned rgb = __funnelshift_r(lo, hi, 8 * (bi & 3)) & 0xffffffu;
				const uint4 rec = make_uint4(rgb | 0xff000000u, __float_as_uint(wx), __float_as_uint(wy), __float_as_uint(wz));
				if (peers.n == 0) out[o] = rec;
				else s_stage[o - ((size_t)out_off + base)] = rec;
				o++;
			};
			if (LS3D_MAP_PAIRS && kKeepMask && !(w & 1) && w >= 8 && rem >= 8) {
				// even width: pixels 2q, 2q+1 of the thread share a row and an 8-byte aligned ray pair -> four packed pair evaluations,
				// unrolled (constant j: the depth and colour selections are one instruction each), bit-identical to map_pixel
				const f32x2 one2 = pk(one, one);
#pragma unroll
				for (int q = 0; q < 4; q++) {
					if (valid & (3u << (2 * q))) {
						int x = x0 + 2 * q, y = y0;
						if (x >= w) { x -= w; y++; }
						const float2 xn = __ldg(reinterpret_cast<const float2 *>(xray + x));
						f32x2 ax, ay, az;
						map_pixel_pair(m, pk(xn.x, xn.y), __ldg(yray + y), q == 0 ? dq.x : q == 1 ? dq.y : q == 2 ? dq.z : dq.w, one2, ax, ay, az);
						float a0, a1, b0, b1, g0, g1;
						upk(ax, a0, a1); upk(ay, b0, b1); upk(az, g0, g1);
						if (valid & (1u << (2 * q))) emit(2 * q, a0, b0, g0);
						if (valid & (2u << (2 * q))) emit(2 * q + 1, a1, b1, g1);
					}
				}
			} else {
				unsigned rest = valid;
#pragma unroll 1
				while (rest) {
					const int j = __ffs(rest) - 1;
					rest &= rest - 1;
					float wx, wy, wz;
					if (kKeepMask) {
						int x = x0 + j, y = y0;
						while (x >= w) { x -= w; y++; }
						map_pixel(m, __ldg(xray + x), __ldg(yray + y), depth_of(j), wx, wy, wz);
					} else {
						wx = s_pos[0][j][tid]; wy = s_pos[1][j][tid]; wz = s_pos[2][j][tid];
					}
					emit(j, wx, wy, wz);
				}
			}
		}
		if (peers.n > 0) {
			__syncthreads();
			for (int p = 0; p < peers.n; p++) {
				uint4 *dst = peers.ptr[p] + (size_t)out_off + base;
				for (unsigned i = tid; i < total; i += kScanThreads) dst[i] = s_stage[i];
			}
			__syncthreads();
		}
		if (kWriteD2V) {
			int *dst = d2v + sd[s].pix_begin + p0;
			unsigned o = base + off;
#pragma unroll
			for (int j = 0; j < 8; j++)
				if (j < rem) dst[j] = ((valid >> j) & 1) ? (int)(o++) : -1;
		}
		if (tid == 0) {
			if (tile + tile0 == sd[s].tile_begin) culled_starts[s] = (int)base;
			if (tile == ntiles - 1) { culled_starts[s_end] = (int)(base + total); ctl->n_culled = (int)(base + total); ctl->n_final = (int)(base + total); }
		}
		if (!kKeepMask) __syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------------
// K1t: triangles over the depth grid
// ------------------------------------------------------------------------------------------------------
// MeshGenerator::checkTriangleConstraints (meshGenerator.cpp:14-62) on pixel indices a, b, c of one depth image: all three
// depths non-zero, and every edge either flat (|dv| < thr) or continuing the gradient of the pixel one step beyond
// either end.  thr is the reference's double-precision expression, truncated (:26) — true IEEE division by 3.0.
__device__ __forceinline__ bool triangle_ok(const unsigned short *__restrict__ d, int a, int b, int c) {
	const int at[3] = {a, b, c};
	const int v[3] = {(int)__ldg(d + a), (int)__ldg(d + b), (int)__ldg(d + c)};
	if (v[0] == 0 || v[1] == 0 || v[2] == 0) return false;
	const int thr = (int)__dadd_rn(__dmul_rn(__ddiv_rn((double)(v[0] + v[1] + v[2]), 3.0), 0.00272), 7.273);
#pragma unroll
	for (int e = 0; e < 3; e++) {
		const int i1 = e, i2 = (e + 1) % 3;
		const int v1 = v[i1], v2 = v[i2];
		if (abs(v1 - v2) < thr) continue;
		const int step = at[i2] - at[i1];
		const int fwd = (int)__ldg(d + at[i2] + step);
		if (fwd != 0 && abs(v2 - v1 - (fwd - v2)) < thr) continue;
		const int bwd = (int)__ldg(d + at[i1] - step);
		if (bwd != 0 && abs(v2 - v1 - (v1 - bwd)) < thr) continue;
		return false;
	}
	return true;
}

// generateTrianglesGradients (meshGenerator.cpp:76-181) for every sensor of the run + formMesh's index rebasing and
// sensor-order concatenation (depthprocessing.cpp:1611-1626).  d2v holds GLOBAL vertex indices (K1 with kWriteD2V), so no
// rebasing is left to do.  Per 2048-pixel tile: the depth rows the tests can touch (2 rows above to 1 row below the tile) are
// staged in shared memory by coalesced loads, so the ~30 u16 reads per pixel never leave the SM; thread t handles pixels
// t, t+256, ... of the tile (coalesced map reads and triangle stores); a 4-bit emit mask per pixel first, then — with the
// tile's base from the look-back scan and the 64 (pass, warp) segment offsets — the index triples are stored in raster order
// (the reference's band threads concatenate to exactly that).
__device__ __forceinline__ bool triangle_ok_s(const unsigned short *d, int a, int b, int c) {      // triangle_ok on staged depths
	const int at[3] = {a, b, c};
	const int v[3] = {(int)d[a], (int)d[b], (int)d[c]};
	if (v[0] == 0 || v[1] == 0 || v[2] == 0) return false;
	const int thr = (int)__dadd_rn(__dmul_rn(__ddiv_rn((double)(v[0] + v[1] + v[2]), 3.0), 0.00272), 7.273);
#pragma unroll
	for (int e = 0; e < 3; e++) {
		const int i1 = e, i2 = (e + 1) % 3;
		const int v1 = v[i1], v2 = v[i2];
		if (abs(v1 - v2) < thr) continue;
		const int step = at[i2] - at[i1];
		const int fwd = (int)d[at[i2] + step];
		if (fwd != 0 && abs(v2 - v1 - (fwd - v2)) < thr) continue;
		const int bwd = (int)d[at[i1] - step];
		if (bwd != 0 && abs(v2 - v1 - (v1 - bwd)) < thr) continue;
		return false;
	}
	return true;
}

__global__ void __launch_bounds__(kScanThreads) k_triangles(const uint8_t *__restrict__ depth, const SensorDesc *__restrict__ sd,
	const unsigned short *__restrict__ tile_sensor, const int *__restrict__ d2v, int s_first, int s_end, FrameCtl *ctl,
	unsigned long long *status, int *tri_starts, int *__restrict__ tri)
{
	pdl_enter();
	extern __shared__ unsigned short s_depth[];        // kTile + 3 * max_w + 4 values
	__shared__ unsigned s_seg[64], s_base, s_total;
	__shared__ int s_tile;
	const int tile0 = sd[s_first].tile_begin;
	const int ntiles = sd[s_end].tile_begin - tile0;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (;;) {
		if (tid == 0) s_tile = (int)atomicAdd(&ctl->tile_counter_b, 1u);
		__syncthreads();
		const int tile = s_tile;
		if (tile >= ntiles) break;
		const int s = tile_sensor[tile + tile0];
		const int w = sd[s].w, h = sd[s].h, px = sd[s].px;
		const int tstart = (tile + tile0 - sd[s].tile_begin) * kTile;
		const unsigned short *dimg = reinterpret_cast<const unsigned short *>(depth + sd[s].depth_off);
		const int *map = d2v + sd[s].pix_begin;
		const int lo = max(0, tstart - 2 * w - 1), hi = min(px, tstart + kTile + w + 3);
		for (int i = lo + tid; i < hi; i += kScanThreads) s_depth[i - lo] = __ldg(dimg + i);
		__syncthreads();
		const unsigned short *D = s_depth - lo;            // D[p] = depth of pixel p for every p the tests below can touch

		unsigned emit = 0;                                  // 4 bits per pass
		unsigned long long lane_off = 0;                    // 8 bits per pass: this lane's offset inside its (pass, warp) segment
#pragma unroll 1
		for (int j = 0; j < 8; j++) {
			const int p = tstart + j * kScanThreads + tid;
			unsigned ok = 0;
			if (p < px) {
				const int y = p / w, x = p - y * w;
				if (y >= 2 && y < h - 2 && x >= 1 && x < w - 2 && __ldg(map + p) != -1) {
					if (triangle_ok_s(D, p, p - w, p + 1)) ok |= 1u;
					if (triangle_ok_s(D, p + 1, p - w, p - w + 1)) ok |= 2u;
					if (!ok) {
						if (triangle_ok_s(D, p, p - w, p - w + 1)) ok |= 4u;
						if (triangle_ok_s(D, p, p - w + 1, p + 1)) ok |= 8u;
					}
					if (ok) {
						// a triangle is emitted only when all three corners own vertices (meshGenerator.cpp:131-133); corners per triangle: triangles_shifts (:100-103)
						const bool m_e = __ldg(map + p + 1) != -1, m_n = __ldg(map + p - w) != -1, m_ne = __ldg(map + p - w + 1) != -1;
						if (!(m_e && m_n)) ok &= ~1u;
						if (!(m_e && m_ne && m_n)) ok &= ~2u;
						if (!(m_ne && m_n)) ok &= ~4u;
						if (!(m_e && m_ne)) ok &= ~8u;
					}
				}
			}
			const unsigned cnt = __popc(ok);
			const unsigned incl = warp_incl_scan(cnt, lane);
			if (lane == 31) s_seg[j * 8 + warp] = incl;
			emit |= ok << (4 * j);
			lane_off |= (unsigned long long)(incl - cnt) << (8 * j);
		}
		__syncthreads();
		if (warp == 0) {
			const unsigned a = s_seg[2 * lane], b = s_seg[2 * lane + 1];
			const unsigned sc = warp_incl_scan(a + b, lane);
			const unsigned total = __shfl_sync(kFull, sc, 31);
			const unsigned base = lookback_exclusive(status, tile, total, &ctl->err);
			s_seg[2 * lane] = sc - a - b;
			s_seg[2 * lane + 1] = sc - b;
			if (lane == 0) { s_base = base; s_total = total; }
		}
		__syncthreads();
		const unsigned base = s_base;
#pragma unroll 1
		for (int j = 0; j < 8; j++) {
			unsigned ok = (emit >> (4 * j)) & 0xfu;
			if (!ok) continue;
			const int p = tstart + j * kScanThreads + tid;
			int *o = tri + 3 * (size_t)(base + s_seg[j * 8 + warp] + (unsigned)((lane_off >> (8 * j)) & 0xffu));
			const int m_p = __ldg(map + p), m_e = __ldg(map + p + 1), m_n = __ldg(map + p - w), m_ne = __ldg(map + p - w + 1);
			if (ok & 1u) { o[0] = m_e; o[1] = m_n; o[2] = m_p; o += 3; }
			if (ok & 2u) { o[0] = m_e; o[1] = m_ne; o[2] = m_n; o += 3; }
			if (ok & 4u) { o[0] = m_p; o[1] = m_ne; o[2] = m_n; o += 3; }
			if (ok & 8u) { o[0] = m_p; o[1] = m_e; o[2] = m_ne; o += 3; }
		}
		if (tid == 0) {
			if (tile + tile0 == sd[s].tile_begin) tri_starts[s] = (int)base;
			if (tile == ntiles - 1) { tri_starts[s_end] = (int)(base + s_total); ctl->n_triangles = (int)(base + s_total); }
		}
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------------
// K1o: neighbour count on the ORGANIZED cloud (the depth image is a uniform grid in pixel space)
// ------------------------------------------------------------------------------------------------------
// Every point of a sensor's cloud comes from one pixel, so the points that can lie within maxDist of a pixel's
// point are confined to a window around that pixel: two camera-space points closer than r' project at most
//     |du| <= fx r' sqrt(1 + xn^2) / (Z - r'),   |dv| <= fy r' sqrt(1 + yn^2) / (Z - r')
// apart (xn, yn = the pixel's normalised ray, Z its camera depth; r' = (maxDist + fp32 slop) / sigma_min(R) is set
// on the host so that "world fp32 d2 <= thr" implies "camera distance <= r'").  A block stages the world positions
// of a 32x8 pixel tile plus an 8-pixel halo in shared memory (NaN where there is no vertex) and every thread counts
// d2 <= thr over its window, leaving at k; pixels whose window exceeds the halo (very near depth, large radii)
// walk the window in global memory instead, recomputing candidates from the depth image.  Same fp32 expressions,
// same count >= k decision as the voxel-hash path and the reference — only the candidate enumeration differs.
#ifndef LS3D_ORG_PPT
#define LS3D_ORG_PPT 2
#endif
constexpr int kOrgTW = 32, kOrgRows = 8, kOrgPPT = LS3D_ORG_PPT, kOrgTH = kOrgRows * kOrgPPT, kOrgHalo = 8;   // 32x16 pixel tile, 256 threads, 2 pixels (rows ly, ly+8) per thread
constexpr int kOrgSW = kOrgTW + 2 * kOrgHalo, kOrgSH = kOrgTH + 2 * kOrgHalo;

// count{d2 <= thr} over rows ci + dy*kOrgSW (dy = 0, -1, +1, -2, ... up to +-rv: centre rows first, so an inlier reaches k
// after a few rows), columns -HX..+HX fully unrolled: one LDS.128 per candidate at an immediate offset, no loop bookkeeping.
// HX is the warp's column reach (>= every in-halo lane's own reach; extra candidates are real points, so counting them
// is still exact).
template <int HX>
__device__ __forceinline__ int org_count_rows(const float4 *__restrict__ tile, int ci, int rv, float qx, float qy, float qz, int k, float thr) {
	int cnt = 0;
#pragma unroll 1
	for (int j = 0; j <= 2 * rv && cnt < k; j++) {
		const int dy = (j & 1) ? -((j + 1) >> 1) : (j >> 1);
		const float4 *row = tile + ci + dy * kOrgSW - HX;
#pragma unroll
		for (int dx = 0; dx <= 2 * HX; dx++) {
			const float4 c = row[dx];
			cnt += dist2_ref(qx, qy, qz, c.x, c.y, c.z) <= thr ? 1 : 0;
		}
	}
	return cnt;
}

__device__ __forceinline__ int org_count_dispatch(int hx, const float4 *__restrict__ tile, int ci, int rv, float qx, float qy, float qz, int k, float thr) {
	switch (hx) {
	case 0: return org_count_rows<0>(tile, ci, rv, qx, qy, qz, k, thr);
	case 1: return org_count_rows<1>(tile, ci, rv, qx, qy, qz, k, thr);
	case 2: return org_count_rows<2>(tile, ci, rv, qx, qy, qz, k, thr);
	case 3: return org_count_rows<3>(tile, ci, rv, qx, qy, qz, k, thr);
	case 4: return org_count_rows<4>(tile, ci, rv, qx, qy, qz, k, thr);
	case 5: return org_count_rows<5>(tile, ci, rv, qx, qy, qz, k, thr);
	case 6: return org_count_rows<6>(tile, ci, rv, qx, qy, qz, k, thr);
	case 7: return org_count_rows<7>(tile, ci, rv, qx, qy, qz, k, thr);
	default: return org_count_rows<8>(tile, ci, rv, qx, qy, qz, k, thr);
	}
}

#ifndef LS3D_ORG_PACKED
#define LS3D_ORG_PACKED 1
#endif
#ifndef LS3D_ORG_ONEARRAY
#define LS3D_ORG_ONEARRAY 1
#endif
#ifndef LS3D_ORG_OWNPAIR
#define LS3D_ORG_OWNPAIR 1
#endif
#ifndef LS3D_ORG_PAIRSTAGE
#define LS3D_ORG_PAIRSTAGE 1
#endif
#ifndef LS3D_ORG_MINBLOCKS
#define LS3D_ORG_MINBLOCKS 5
#endif
#ifndef LS3D_ORG_ROWS2
#define LS3D_ORG_ROWS2 1
#endif
#ifndef LS3D_ORG_NOBAR1
#define LS3D_ORG_NOBAR1 1
#endif
template <int HX>
__device__ __forceinline__ int org_count_rows2(const float *__restrict__ tx, const float *__restrict__ ty, const float *__restrict__ tz,
	int ci, int rv, float qx, float qy, float qz, int k, float thr, f32x2 one2) {
	const f32x2 q2x = pk(qx, qx), q2y = pk(qy, qy), q2z = pk(qz, qz);
	const int c0 = (ci - HX) & ~1;
	const float kf = (float)k;
	f32x2 acc = pk(0.0f, 0.0f);
	float cnt = 0.0f;
	auto row = [&](int b) {
#pragma unroll
		for (int p = 0; p <= HX; p++) {
			const f32x2 X = *reinterpret_cast<const f32x2 *>(tx + b + 2 * p), Y = *reinterpret_cast<const f32x2 *>(ty + b + 2 * p), Z = *reinterpret_cast<const f32x2 *>(tz + b + 2 * p);
			const f32x2 d0 = sub2(q2x, X), d1 = sub2(q2y, Y), d2 = sub2(q2z, Z);
			const f32x2 s = fma2(fma2(mul2(d0, d0), one2, mul2(d1, d1)), one2, mul2(d2, d2));
			float lo, hi;
			upk(s, lo, hi);
			acc = add2(acc, pk(lo <= thr ? 1.0f : 0.0f, hi <= thr ? 1.0f : 0.0f));
		}
	};
#if LS3D_ORG_ROWS2
	// the centre row, then rows -r and +r together: one exit test per two rows (same rows as below, so the same decision)
	row(c0);
	{ float a0, a1; upk(acc, a0, a1); cnt = a0 + a1; }
#pragma unroll 1
	for (int r = 1; r <= rv && cnt < kf; r++) {
		row(c0 - r * kOrgSW);
		row(c0 + r * kOrgSW);
		float a0, a1;
		upk(acc, a0, a1);
		cnt = a0 + a1;
	}
#else
#pragma unroll 1
	for (int j = 0; j <= 2 * rv && cnt < kf; j++) {
		const int dy = (j & 1) ? -((j + 1) >> 1) : (j >> 1);
		row(c0 + dy * kOrgSW);
		float a0, a1;
		upk(acc, a0, a1);
		cnt = a0 + a1;
	}
#endif
	return (int)cnt;
}

__device__ __forceinline__ int org_count_dispatch2(int hx, const float *__restrict__ tx, const float *__restrict__ ty, const float *__restrict__ tz,
	int ci, int rv, float qx, float qy, float qz, int k, float thr, f32x2 one2) {
	switch (hx) {
	case 0: case 1: return org_count_rows2<1>(tx, ty, tz, ci, rv, qx, qy, qz, k, thr, one2);
	case 2: return org_count_rows2<2>(tx, ty, tz, ci, rv, qx, qy, qz, k, thr, one2);
	case 3: return org_count_rows2<3>(tx, ty, tz, ci, rv, qx, qy, qz, k, thr, one2);
	case 4: return org_count_rows2<4>(tx, ty, tz, ci, rv, qx, qy, qz, k, thr, one2);
	case 5: return org_count_rows2<5>(tx, ty, tz, ci, rv, qx, qy, qz, k, thr, one2);
	case 6: return org_count_rows2<6>(tx, ty, tz, ci, rv, qx, qy, qz, k, thr, one2);
	default: return org_count_rows2<7>(tx, ty, tz, ci, rv, qx, qy, qz, k, thr, one2);
	}
}

// the per-run control block cleared by a kernel instead of a memset node, so that the organized count can be chained behind it
// programmatically (its blocks start while this one runs; they wait for it before their first write)
__global__ void __launch_bounds__(256) k_zero_control(uint4 *p, int n16) {
	pdl_trigger();
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i < n16) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(kOrgTW * kOrgRows, LS3D_ORG_MINBLOCKS) k_organized_count(const uint8_t *__restrict__ depth, const SensorDesc *__restrict__ sd,
	const float *__restrict__ rays, int s_first, Bounds6 bnd, FrameCtl *ctl, int k, float thr, uint8_t *__restrict__ keep_px, unsigned *tile_count, float one)
{
#if LS3D_ORG_PACKED
#if LS3D_ORG_ONEARRAY
	// one array, three planes: a candidate's y and z are its x address plus a constant (immediate offsets instead of three address computations per row)
	__shared__ __align__(16) float tile_xyz[3 * kOrgSH * kOrgSW];
	float *const tile_x = tile_xyz, *const tile_y = tile_xyz + kOrgSH * kOrgSW, *const tile_z = tile_xyz + 2 * kOrgSH * kOrgSW;
#else
	__shared__ __align__(16) float tile_x[kOrgSH * kOrgSW], tile_y[kOrgSH * kOrgSW], tile_z[kOrgSH * kOrgSW];
#endif
	auto put = [&](int i, float a, float b, float c) { tile_x[i] = a; tile_y[i] = b; tile_z[i] = c; };
	constexpr int kReachX = kOrgHalo - 1;       // pairs start at even columns: one halo column is spent on the alignment
#else
	__shared__ float4 tile[kOrgSH * kOrgSW];
	auto put = [&](int i, float a, float b, float c) { tile[i] = make_float4(a, b, c, 0.0f); };
	constexpr int kReachX = kOrgHalo;
#endif
	__shared__ unsigned s_kept;
	__shared__ int s_hx, s_hy;
	// the map kernel behind us (launch_map, after_count) may move into the SMs as soon as every block of this grid has started: it
	// waits for this grid's completion (griddepcontrol.wait) before it touches anything written here
	pdl_trigger();
	const int s = s_first + blockIdx.z;
	const int w = sd[s].w, h = sd[s].h;
	const int tx0 = blockIdx.x * kOrgTW, ty0 = blockIdx.y * kOrgTH;
	if (tx0 >= w || ty0 >= h) return;
	const PixelXform m = load_xform(sd, s, bnd.v);
	const unsigned short *dimg = reinterpret_cast<const unsigned short *>(depth + sd[s].depth_off);
	const float *xray = rays + sd[s].ray_off, *yray = xray + w;
	const float *xreach = yray + h, *yreach = xreach + w;     // fx*sqrt(1+xn^2), fy*sqrt(1+yn^2): the window reach per unit of r'/(Z-r')
	const int tid = threadIdx.x;
#if LS3D_ORG_NOBAR1
	__shared__ int s_wx[kOrgRows], s_wy[kOrgRows];      // every warp's reach: no initialisation, so no barrier before phase 0
	if (tid == 0) s_kept = 0;                          // first used behind the barriers below
#else
	if (tid == 0) { s_kept = 0; s_hx = 0; s_hy = 0; }
	__syncthreads();
#endif

	// ---- phase 0: this thread's own pixels -> world positions into the tile centre, and how far their windows reach ----
	const int lx = tid & (kOrgTW - 1), ly = tid / kOrgTW;
	const int x = tx0 + lx;
	const float qnan = __int_as_float(0x7fc00000);
	const float rp = sd[s].org_rp;
	float qx[kOrgPPT], qy[kOrgPPT], qz[kOrgPPT];
	int ru[kOrgPPT], rv[kOrgPPT];
	bool has[kOrgPPT], in_halo[kOrgPPT];
	int mu = 0, mv = 0;
#if LS3D_ORG_PACKED && LS3D_ORG_OWNPAIR && (LS3D_ORG_PPT % 2 == 0)
	// the thread's pixels two at a time (same column, rows 8 apart) through the packed pair map: bit-identical positions
	unsigned dd[kOrgPPT];
	{
		const float xn = x < w ? __ldg(xray + x) : 0.0f;
		const f32x2 one2 = pk(one, one);
#pragma unroll
		for (int p = 0; p < kOrgPPT; p += 2) {
			const int ya = ty0 + ly + kOrgRows * p, yb = ya + kOrgRows;
			const bool ina = x < w && ya < h, inb = x < w && yb < h;
			dd[p] = ina ? (unsigned)__ldg(dimg + (size_t)ya * w + x) : 0u;
			dd[p + 1] = inb ? (unsigned)__ldg(dimg + (size_t)yb * w + x) : 0u;
			f32x2 ax, ay, az;
			const unsigned ok = map_pixel_pair_xy(m, pk(xn, xn), pk(ina ? __ldg(yray + ya) : 0.0f, inb ? __ldg(yray + yb) : 0.0f), dd[p] | (dd[p + 1] << 16), one2, ax, ay, az);
			float a0, a1, b0, b1, g0, g1;
			upk(ax, a0, a1); upk(ay, b0, b1); upk(az, g0, g1);
			has[p] = (ok & 1u) != 0; has[p + 1] = (ok & 2u) != 0;
			qx[p] = has[p] ? a0 : qnan; qy[p] = has[p] ? b0 : qnan; qz[p] = has[p] ? g0 : qnan;
			qx[p + 1] = has[p + 1] ? a1 : qnan; qy[p + 1] = has[p + 1] ? b1 : qnan; qz[p + 1] = has[p + 1] ? g1 : qnan;
		}
	}
#pragma unroll
	for (int p = 0; p < kOrgPPT; p++) {
		const int y = ty0 + ly + kOrgRows * p;
		ru[p] = 0; rv[p] = 0;
		if (has[p]) {
			const float den = (float)dd[p] * 0.001f - rp;
			ru[p] = 1 << 28; rv[p] = 1 << 28;
			if (den > 1e-6f) {
				const float g = __fdividef(rp, den) * 1.002f;
				const float fu = __ldg(xreach + x) * g + 1e-2f, fv = __ldg(yreach + y) * g + 1e-2f;
				if (fu < 1e8f) ru[p] = (int)ceilf(fu);
				if (fv < 1e8f) rv[p] = (int)ceilf(fv);
			}
		}
		put((ly + kOrgRows * p + kOrgHalo) * kOrgSW + lx + kOrgHalo, qx[p], qy[p], qz[p]);
		in_halo[p] = has[p] && ru[p] <= kReachX && rv[p] <= kOrgHalo;
		if (in_halo[p]) { mu = max(mu, ru[p]); mv = max(mv, rv[p]); }
	}
#else
#pragma unroll
	for (int p = 0; p < kOrgPPT; p++) {
		const int y = ty0 + ly + kOrgRows * p;
		qx[p] = qnan; qy[p] = qnan; qz[p] = qnan;
		ru[p] = 0; rv[p] = 0; has[p] = false;
		if (x < w && y < h) {
			const unsigned d = (unsigned)__ldg(dimg + (size_t)y * w + x);
			float ax, ay, az;
			if (map_pixel(m, __ldg(xray + x), __ldg(yray + y), d, ax, ay, az)) {
				has[p] = true;
				qx[p] = ax; qy[p] = ay; qz[p] = az;
				const float den = (float)d * 0.001f - rp;
				ru[p] = 1 << 28; rv[p] = 1 << 28;
				if (den > 1e-6f) {
					// window reach in pixels; the 0.2 % + 0.01 px margin covers the approximate reciprocal and the rounded tables
					const float g = __fdividef(rp, den) * 1.002f;
					const float fu = __ldg(xreach + x) * g + 1e-2f, fv = __ldg(yreach + y) * g + 1e-2f;
					if (fu < 1e8f) ru[p] = (int)ceilf(fu);
					if (fv < 1e8f) rv[p] = (int)ceilf(fv);
				}
			}
		}
		put((ly + kOrgRows * p + kOrgHalo) * kOrgSW + lx + kOrgHalo, qx[p], qy[p], qz[p]);
		in_halo[p] = has[p] && ru[p] <= kReachX && rv[p] <= kOrgHalo;
		if (in_halo[p]) { mu = max(mu, ru[p]); mv = max(mv, rv[p]); }
	}
#endif
	const int hxw = __reduce_max_sync(kFull, mu), hyw = __reduce_max_sync(kFull, mv);      // this warp's reach
#if LS3D_ORG_NOBAR1
	if ((tid & 31) == 0) { s_wx[tid >> 5] = hxw; s_wy[tid >> 5] = hyw; }
	__syncthreads();
	const int bhx = __reduce_max_sync(kFull, s_wx[tid & (kOrgRows - 1)]), bhy = __reduce_max_sync(kFull, s_wy[tid & (kOrgRows - 1)]);
#else
	if ((tid & 31) == 0 && (hxw | hyw)) { atomicMax(&s_hx, hxw); atomicMax(&s_hy, hyw); }
	__syncthreads();
	const int bhx = s_hx, bhy = s_hy;
#endif
	const int hy = bhy;                       // block-uniform reach of the shared-memory windows (0,0: nobody needs the halo)
	const int hx = (LS3D_ORG_PACKED && (bhx | bhy)) ? max(bhx, 1) + 1 : bhx;      // packed: dispatch floor 1, one column more for the pair alignment

	// ---- phase 1: stage only the halo ring that some window reaches: hy rows above/below, hx columns left/right ----
	auto stage = [&](int r, int c) {          // tile coordinates
		const int gx = tx0 - kOrgHalo + c, gy = ty0 - kOrgHalo + r;
		float wx = qnan, wy = qnan, wz = qnan;
		if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
			float ax, ay, az;
			if (map_pixel(m, __ldg(xray + gx), __ldg(yray + gy), (unsigned)__ldg(dimg + (size_t)gy * w + gx), ax, ay, az)) { wx = ax; wy = ay; wz = az; }
		}
		put(r * kOrgSW + c, wx, wy, wz);
	};
#if LS3D_ORG_PACKED && LS3D_ORG_PAIRSTAGE
	// The ring as a flat list of even-aligned pixel PAIRS (2 hy band rows of 16 + hxe pairs, then hxe pairs beside each of the 16 tile
	// rows; hxe = hx rounded up to even), dealt out evenly to the 256 threads and mapped two pixels at a time with packed fp32: one
	// 32-bit depth load, one 64-bit ray load and three 64-bit shared-memory stores per pair.  (Until round 2 every thread staged single
	// pixels in a row loop x column loop whose second trips kept a third of the lanes busy: 41 % of the kernel's instructions.)
	// Needs an even image width and a 4-byte aligned image (block-uniform test); other images take the scalar loops below.
	if ((hx | hy) && !(w & 1) && !(reinterpret_cast<uintptr_t>(dimg) & 3)) {
		const int hxe = (hx + 1) & ~1;                                   // 2..8
		const int pw = kOrgTW / 2 + hxe;                                 // pairs per band row: 18..24
		const int nband = 2 * hy * pw, ntot = nband + kOrgTH * hxe;
		const unsigned inv_pw = 65536u / (unsigned)pw + 1u, inv_hx = 65536u / (unsigned)hxe + 1u;      // (i * inv) >> 16 == i / d for i < 2048
		const f32x2 one2 = pk(one, one);
		for (int i = tid; i < ntot; i += kOrgTW * kOrgRows) {
			int r, c;                                                    // tile coordinates of the pair's left (even) column
			if (i < nband) {
				const int br = (int)(((unsigned)i * inv_pw) >> 16), bc = i - br * pw;
				r = br < hy ? kOrgHalo - hy + br : kOrgHalo + kOrgTH - hy + br;
				c = kOrgHalo - hxe + 2 * bc;
			} else {
				const int j = i - nband;
				const int sr = (int)(((unsigned)j * inv_hx) >> 16), sc = j - sr * hxe;      // hxe / 2 pairs left of the tile, hxe / 2 right of it
				r = kOrgHalo + sr;
				c = 2 * sc < hxe ? kOrgHalo - hxe + 2 * sc : kOrgHalo + kOrgTW - hxe + 2 * sc;
			}
			const int gx = tx0 - kOrgHalo + c, gy = ty0 - kOrgHalo + r;         // gx is even and w is even: both pixels are inside or both outside
			f32x2 sx = pk(qnan, qnan), sy = sx, sz = sx;
			if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
				const unsigned dp = __ldg(reinterpret_cast<const unsigned *>(dimg + (size_t)gy * w + gx));
				const float2 xn = __ldg(reinterpret_cast<const float2 *>(xray + gx));
				f32x2 ax;
				const unsigned ok = map_pixel_pair(m, pk(xn.x, xn.y), __ldg(yray + gy), dp, one2, ax, sy, sz);
				float a0, a1;
				upk(ax, a0, a1);
				sx = pk((ok & 1u) ? a0 : qnan, (ok & 2u) ? a1 : qnan);      // a NaN x is enough to fail every distance test
			}
			const int idx = r * kOrgSW + c;
			*reinterpret_cast<f32x2 *>(tile_x + idx) = sx;
			*reinterpret_cast<f32x2 *>(tile_y + idx) = sy;
			*reinterpret_cast<f32x2 *>(tile_z + idx) = sz;
		}
	} else
#endif
	if (hx | hy) {
		for (int r = ly; r < 2 * hy; r += kOrgRows) {
			const int row = r < hy ? kOrgHalo - hy + r : kOrgHalo + kOrgTH + (r - hy);
			for (int c = lx; c < kOrgTW + 2 * hx; c += kOrgTW) stage(row, kOrgHalo - hx + c);
		}
		if (lx < 2 * hx) {
			const int c = lx < hx ? kOrgHalo - hx + lx : kOrgHalo + kOrgTW + (lx - hx);
#pragma unroll
			for (int p = 0; p < kOrgPPT; p++) stage(kOrgHalo + ly + kOrgRows * p, c);
		}
	}
	__syncthreads();

	// ---- phase 2: count ----
	pdl_wait();          // chained behind k_zero_control: the tile counts and the control block are clear before anything is written
	unsigned kept_total = 0;
#pragma unroll
	for (int p = 0; p < kOrgPPT; p++) {
		const int y = ty0 + ly + kOrgRows * p;
		bool kept = false;
		if (has[p]) {
			int cnt = 0;
			if (in_halo[p]) {
#if LS3D_ORG_PACKED
				cnt = org_count_dispatch2(hxw, tile_x, tile_y, tile_z, (ly + kOrgRows * p + kOrgHalo) * kOrgSW + lx + kOrgHalo, rv[p], qx[p], qy[p], qz[p], k, thr, pk(one, one));
#else
				cnt = org_count_dispatch(hxw, tile, (ly + kOrgRows * p + kOrgHalo) * kOrgSW + lx + kOrgHalo, rv[p], qx[p], qy[p], qz[p], k, thr);
#endif
			} else {
				// window larger than the halo (very near depth, large radii): walk it in global memory, recomputing candidates
				const int xa = max(0, x - min(ru[p], w)), xb = min(w - 1, x + min(ru[p], w));
				const int ya = max(0, y - min(rv[p], h)), yb = min(h - 1, y + min(rv[p], h));
				for (int yy = ya; yy <= yb && cnt < k; yy++) {
					const float yny = __ldg(yray + yy);
					for (int xx = xa; xx <= xb; xx++) {
						float ax, ay, az;
						if (map_pixel(m, __ldg(xray + xx), yny, (unsigned)__ldg(dimg + (size_t)yy * w + xx), ax, ay, az))
							cnt += dist2_ref(qx[p], qy[p], qz[p], ax, ay, az) <= thr ? 1 : 0;
					}
				}
			}
			kept = cnt >= k;
		}
		if (x < w && y < h) keep_px[sd[s].pix_begin + (size_t)y * w + x] = (uint8_t)(kept ? 1 : 0);
		// survivors per compaction tile of the map kernel (a warp is one 32-pixel row segment: it touches at most two tiles)
		const unsigned km = __ballot_sync(kFull, kept);
		if (km) {
			const int pix = min(y, h - 1) * w + min(x, w - 1);
			const int t = sd[s].tile_begin + pix / kTile;
			const int tA = __shfl_sync(kFull, t, 0), tB = __shfl_sync(kFull, t, 31);
			const unsigned inA = __ballot_sync(kFull, kept && t == tA), inB = km & ~inA;
			if ((tid & 31) == 0) {
				if (inA) atomicAdd(&tile_count[tA], (unsigned)__popc(inA));
				if (inB) atomicAdd(&tile_count[tB], (unsigned)__popc(inB));
				kept_total += (unsigned)__popc(km);
			}
		}
	}
	if ((tid & 31) == 0 && kept_total) atomicAdd(&s_kept, kept_total);
	__syncthreads();
	if (tid == 0 && s_kept) atomicAdd(&ctl->n_kept, (int)s_kept);
}

// ------------------------------------------------------------------------------------------------------
// K2-K5: neighbour count on an unorganized cloud — a hash of x-runs of voxels
// ------------------------------------------------------------------------------------------------------
// Voxels have edge h >= maxDist, so a point's neighbours lie in its 27 surrounding voxels = 9 rows of 3 voxels along x.  The hash
// key is a RUN of 8 voxels along x (bucket): one 64-byte slot holds the key, the run's first position in the voxel-sorted array and
// the 8 voxel counts, so one look-up (two sectors, all loads independent) yields a row's candidates as ONE contiguous range; a row
// straddles two runs for a quarter of the voxels.  Queries are processed in input order, one lane each, 9-11 look-ups per lane
// issued three rows at a time, then a private cursor over the ranges.  (Rounds 1-2 hashed single voxels and processed the points in
// hash-slot order: 27 look-ups of two dependent loads per distinct voxel, 16 voxels per warp one after the other — 263 us for the
// bench frame's 739 k points, latency-bound; bricks of 4^3 voxels with shared-memory staging were slower still, see DESIGN.md.)
// The table is never cleared as a whole: k_voxel_cleanup zeroes exactly the slots the run's points touched.
constexpr int kRunBits = 3;
struct alignas(64) VoxBucket {
	unsigned long long word;       // (key + 1) << 32 | points in the run; 0 = free
	unsigned start;                // first position of the run's points in the sorted array (k_bucket_alloc)
	unsigned wide;                 // 1: more than 65535 points in the run, inc16 is not valid, read cnt
	unsigned short inc16[8];       // inclusive prefix sums of the voxel counts: with the header, the ONE sector a row look-up reads
	unsigned cnt[8];               // points per voxel of the run; k_bucket_alloc turns them into the inclusive prefix sums
};
struct alignas(32) V8 { unsigned v[8]; };
__device__ __forceinline__ V8 ldg256(const void *p) {          // one 32-byte request per lane (LDG.E.256, sm_100)
	V8 r;
	asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
	return r;
}
__device__ __forceinline__ unsigned bucket_key(int bx, int cy, int cz) {       // bx = cx >> kRunBits
	return ((unsigned)cz << (2 * kCellBits - kRunBits)) | ((unsigned)cy << (kCellBits - kRunBits)) | (unsigned)bx;
}

// K2: per point — find or claim its run's slot, add to the run total and to the voxel's count (the old value is the point's rank).
// Lanes of a warp that share a voxel are aggregated.  slot_of[g] = slot << 3 | voxel-in-run; rank_of[g] bit 31: this point's group
// was the first to add to its run, so it will allocate the run's range (k_bucket_alloc).
__global__ void __launch_bounds__(256) k_voxel_insert(const uint4 *__restrict__ cloud, const SensorDesc *__restrict__ sd,
	const int *__restrict__ culled_starts, int s_first, int s_end, FrameCtl *ctl,
	VoxBucket *table, unsigned *__restrict__ slot_of, unsigned *__restrict__ rank_of)
{
	pdl_enter();
	const int N = ctl->n_culled;
	const int lane = threadIdx.x & 31;
	for (int g0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); g0 < N; g0 += gridDim.x * blockDim.x) {
		const int g = g0 + lane;
		const bool active = g < N;
		unsigned key = 0, h0 = 0;
		unsigned long long tag = ~0ull - (unsigned long long)lane;
		int s = s_first, cx = 0;
		if (active) {
			while (s + 1 < s_end && culled_starts[s + 1] <= g) s++;
			const uint4 p = cloud[g];
			cx = cell_coord(__uint_as_float(p.y), sd[s].gox, sd[s].ginv_h);
			const int cy = cell_coord(__uint_as_float(p.z), sd[s].goy, sd[s].ginv_h);
			const int cz = cell_coord(__uint_as_float(p.w), sd[s].goz, sd[s].ginv_h);
			key = bucket_key(cx >> kRunBits, cy, cz);
			h0 = hash_mix(hash_x(cx >> kRunBits) ^ hash_y(cy) ^ hash_z(cz));
			tag = ((unsigned long long)key << 3) | (unsigned long long)(cx & 7) | ((unsigned long long)s << 40);
		}
		const unsigned grp = __match_any_sync(kFull, tag);
		const int leader = __ffs(grp) - 1;
		unsigned slot = 0, base_rank = 0;
		if (active && lane == leader) {
			const unsigned n = __popc(grp);
			const unsigned toff = sd[s].tbl_off, tmask = sd[s].tbl_mask;
			const unsigned long long want = (unsigned long long)key + 1;
			unsigned h = h0 & tmask;
			bool found = false;
			for (unsigned probe = 0; probe <= tmask; probe++) {
				const unsigned long long w = table[toff + h].word;
				const unsigned long long kk = w >> 32;
				if (kk == want) { found = true; break; }
				if (kk == 0) {
					const unsigned long long prev = atomicCAS(&table[toff + h].word, 0ull, want << 32);
					if (prev == 0 || (prev >> 32) == want) { found = true; break; }
				}
				h = (h + 1) & tmask;
			}
			if (!found) atomicOr(&ctl->err, kErrProbeLimit);
			else {
				const unsigned long long old = atomicAdd(&table[toff + h].word, (unsigned long long)n);
				if ((old & kCountMask) + n > kCountMask) atomicOr(&ctl->err, kErrCellOverflow);
				base_rank = atomicAdd(&table[toff + h].cnt[cx & 7], n);
				if ((old & kCountMask) == 0) base_rank |= 0x80000000u;
			}
			slot = toff + h;
		}
		__syncwarp();
		slot = __shfl_sync(kFull, slot, leader);
		base_rank = __shfl_sync(kFull, base_rank, leader);
		if (active) {
			slot_of[g] = (slot << 3) | (unsigned)(cx & 7);
			rank_of[g] = lane == leader ? base_rank : (base_rank & 0x7fffffffu) + __popc(grp & ((1u << lane) - 1u));
		}
	}
}

// K3: every run gets its range of the sorted array (any order will do): the elected points carry their run's total through a
// block-wide scan, one atomicAdd on the cursor per block
__global__ void __launch_bounds__(256) k_bucket_alloc(const unsigned *__restrict__ slot_of, const unsigned *__restrict__ rank_of, VoxBucket *table, FrameCtl *ctl) {
	pdl_enter();
	__shared__ unsigned s_w[8];
	__shared__ unsigned s_base;
	const int N = ctl->n_culled;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int g0 = blockIdx.x * 256; g0 < N; g0 += gridDim.x * 256) {
		const int g = g0 + threadIdx.x;
		unsigned tot = 0, slot = 0;
		if (g < N && (rank_of[g] & 0x80000000u)) {
			slot = slot_of[g] >> 3;
			tot = (unsigned)(table[slot].word & kCountMask);
		}
		const unsigned incl = warp_incl_scan(tot, lane);
		__syncthreads();                      // s_w / s_base of the previous iteration are no longer read
		if (lane == 31) s_w[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			const unsigned v = lane < 8 ? s_w[lane] : 0u;
			const unsigned sc = warp_incl_scan(v, lane);
			if (lane < 8) s_w[lane] = sc - v;
			if (lane == 31) s_base = sc ? atomicAdd(&ctl->cursor, sc) : 0u;
		}
		__syncthreads();
		if (tot) {
			table[slot].start = s_base + s_w[warp] + incl - tot;
			// voxel counts -> inclusive prefix sums, in place (every rank is out by now): a row's range is two of these values
			uint4 *cp = reinterpret_cast<uint4 *>(table[slot].cnt);
			uint4 a = cp[0], b = cp[1];
			a.y += a.x; a.z += a.y; a.w += a.z; b.x += a.w; b.y += b.x; b.z += b.y; b.w += b.z;
			cp[0] = a; cp[1] = b;
			*reinterpret_cast<uint4 *>(table[slot].inc16) = make_uint4((a.x & 0xffffu) | (a.y << 16), (a.z & 0xffffu) | (a.w << 16), (b.x & 0xffffu) | (b.y << 16), (b.z & 0xffffu) | (b.w << 16));
			table[slot].wide = tot > 0xffffu ? 1u : 0u;
		}
	}
}

// K4: scatter into run / voxel order
__global__ void __launch_bounds__(256) k_cell_scatter(const uint4 *__restrict__ cloud, const unsigned *__restrict__ slot_of,
	const unsigned *__restrict__ rank_of, const VoxBucket *__restrict__ table, const FrameCtl *ctl, float4 *__restrict__ sorted)
{
	pdl_enter();
	const int N = ctl->n_culled;
	for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < N; g += gridDim.x * blockDim.x) {
		const uint4 p = cloud[g];
		const unsigned sc = slot_of[g], c = sc & 7u;
		const VoxBucket *b = table + (sc >> 3);
		const unsigned pre = c ? __ldg(b->cnt + c - 1) : 0u;
		const unsigned pos = __ldg(&b->start) + pre + (rank_of[g] & 0x7fffffffu);
		sorted[pos] = make_float4(__uint_as_float(p.y), __uint_as_float(p.z), __uint_as_float(p.w), __int_as_float(g));
	}
}

// K5: neighbour count, one lane per query in input order; rows visited plane by plane (z, z-1, z+1), the query's own row first
constexpr int kLaneIters = 96;
constexpr int kMaxRanges = 18;

__global__ void __launch_bounds__(kCountWarps * 32) k_neighbour_count(const VoxBucket *__restrict__ table, const uint4 *__restrict__ cloud,
	const float4 *__restrict__ sorted, const SensorDesc *__restrict__ sd,
	const int *__restrict__ culled_starts, int s_first, int s_end, FrameCtl *ctl, int k, float thr, uint8_t *__restrict__ keep)
{
	pdl_enter();
	__shared__ uint2 s_rng[kCountWarps][kMaxRanges][32];      // [warp][range][lane] = (start, count): conflict-free per-lane lists
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int N = ctl->n_culled;
	const int nbatches = (N + 31) >> 5;

	for (;;) {
		int b = 0;
		if (lane == 0) b = (int)atomicAdd(&ctl->work_counter, 1u);
		b = __shfl_sync(kFull, b, 0);
		if (b >= nbatches) break;
		const int g = (b << 5) + lane;
		const bool valid = g < N;
		float qx = 0.f, qy = 0.f, qz = 0.f;
		int nr = 0;
		unsigned tot = 0;
		if (valid) {
			int s = s_first;
			while (s + 1 < s_end && culled_starts[s + 1] <= g) s++;
			const uint4 q = cloud[g];
			qx = __uint_as_float(q.y); qy = __uint_as_float(q.z); qz = __uint_as_float(q.w);
			const int cx = cell_coord(qx, sd[s].gox, sd[s].ginv_h), cy = cell_coord(qy, sd[s].goy, sd[s].ginv_h), cz = cell_coord(qz, sd[s].goz, sd[s].ginv_h);
			const unsigned toff = sd[s].tbl_off, tmask = sd[s].tbl_mask;
			const int xlo = max(cx - 1, 0), xhi = min(cx + 1, kCellMax);
			const int blo = xlo >> kRunBits, bhi = xhi >> kRunBits;          // the runs a row touches (one, or two neighbours)
			// ---- look-ups: the three rows of a z-plane together (three 32-byte loads in flight before the first is examined); the
			// second run of a straddling row in a pass of its own, so the common case keeps its registers.  Rows outside the grid (a
			// voxel on its boundary) are skipped.  Measured and dropped: trimming the voxels the ball cannot reach from the query's
			// position inside its voxel (27 % fewer candidates, but 108-114 us against 91: more registers, more divergent ranges).
			const unsigned hy[3] = {hash_y(cy - 1), hash_y(cy), hash_y(cy + 1)};
			const unsigned kyb = (unsigned)cy << (kCellBits - kRunBits);
			const unsigned ok_y = (cy > 0 ? 1u : 0u) | 2u | (cy < kCellMax ? 4u : 0u);
			auto plane = [&](int dz, int bx) {
				const int z = cz + dz;
				if (z < 0 || z > kCellMax) return;
				// voxels [a, e] of run bx belong to the rows: known before anything is loaded
				const int base = bx << kRunBits;
				const int a = max(xlo - base, 0), e = min(xhi - base, 7);
				const unsigned hxz = hash_x(bx) ^ hash_z(z);
				const unsigned kxz = ((unsigned)z << (2 * kCellBits - kRunBits)) | (unsigned)bx;
				V8 hd[3];
				unsigned hh[3];
				// (the zero fill also keeps ptxas from hoisting all eighteen loads of the six planes to the top: 182 registers without it)
				const V8 zero8 = {{0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}};
#pragma unroll
				for (int j = 0; j < 3; j++) {              // j: dy = 0, -1, +1 (the query's own row first)
					const int yi = j == 0 ? 1 : (j == 1 ? 0 : 2);
					hh[j] = hash_mix(hxz ^ hy[yi]) & tmask;
					hd[j] = zero8;
					if ((ok_y >> yi) & 1u) hd[j] = ldg256(table + toff + hh[j]);
				}
#pragma unroll
				for (int j = 0; j < 3; j++) {
					const int yi = j == 0 ? 1 : (j == 1 ? 0 : 2);
					const unsigned want = kxz + (kyb + ((unsigned)yi << (kCellBits - kRunBits)) - (1u << (kCellBits - kRunBits))) + 1u;      // key(y = cy + yi - 1) + 1
					const VoxBucket *bp = table + toff + hh[j];
					if (hd[j].v[1] != want && hd[j].v[1] != 0u) {
						// another run sits in the slot: follow the probe sequence (rare at this table's load)
						unsigned h = hh[j];
						for (unsigned probe = 0; probe <= tmask; probe++) {
							h = (h + 1) & tmask;
							bp = table + toff + h;
							hd[j] = ldg256(bp);
							if (hd[j].v[1] == want || hd[j].v[1] == 0u) break;
						}
					}
					if (hd[j].v[1] != want) continue;
					// inclusive prefix sums [a - 1] and [e] bound the row's range
					auto inc = [&](int i) -> unsigned {
						const unsigned wsel = (i & 4) ? ((i & 2) ? hd[j].v[7] : hd[j].v[6]) : ((i & 2) ? hd[j].v[5] : hd[j].v[4]);
						return (i & 1) ? (wsel >> 16) : (wsel & 0xffffu);
					};
					unsigned pe = inc(e), pa = a ? inc(a - 1) : 0u;
					if (hd[j].v[3]) { pe = __ldg(bp->cnt + e); pa = a ? __ldg(bp->cnt + a - 1) : 0u; }      // more than 65535 points in the run
					const unsigned n = pe - pa;
					if (n) { s_rng[warp][nr][lane] = make_uint2(hd[j].v[2] + pa, n); nr++; tot += n; }
				}
			};
			plane(0, blo); plane(-1, blo); plane(1, blo);
			if (bhi != blo) { plane(0, bhi); plane(-1, bhi); plane(1, bhi); }
		}
		__syncwarp();

		// ---- private cursors (a point with fewer than k candidates in its 27 voxels cannot reach k) ----
		int cnt = 0, c = 0;
		unsigned p = 0, e = 0;
		if (tot < (unsigned)k) nr = 0;
		if (nr) { const uint2 r = s_rng[warp][0][lane]; p = r.x; e = r.x + r.y; }
		for (int it = 0; it < kLaneIters; it++) {
			const bool act = cnt < k && c < nr;
			if (!__any_sync(kFull, act)) break;
			if (act) {
				// one aligned 32-byte request per turn = two candidates (the pair that holds position p); counting past k is harmless
				const unsigned q0 = p & ~1u;
				const V8 cc = ldg256(sorted + q0);
				const bool lo_ok = !(p & 1u), hi_ok = q0 + 1u < e;
				cnt += (lo_ok && dist2_ref(qx, qy, qz, __uint_as_float(cc.v[0]), __uint_as_float(cc.v[1]), __uint_as_float(cc.v[2])) <= thr) ? 1 : 0;
				cnt += (hi_ok && dist2_ref(qx, qy, qz, __uint_as_float(cc.v[4]), __uint_as_float(cc.v[5]), __uint_as_float(cc.v[6])) <= thr) ? 1 : 0;
				p = min(q0 + 2u, e);
				if (p == e) {
					if (++c < nr) { const uint2 r = s_rng[warp][c][lane]; p = r.x; e = r.x + r.y; }
				}
			}
		}

		// ---- cooperative tail for the lanes still undecided (dense surroundings, large radii): candidates in lanes ----
		unsigned rem = __ballot_sync(kFull, cnt < k && c < nr);
		while (rem) {
			const int l = __ffs(rem) - 1;
			rem &= rem - 1;
			const float lx = __shfl_sync(kFull, qx, l), ly = __shfl_sync(kFull, qy, l), lz = __shfl_sync(kFull, qz, l);
			const int lnr = __shfl_sync(kFull, nr, l);
			int lc = __shfl_sync(kFull, c, l), lcnt = __shfl_sync(kFull, cnt, l);
			unsigned lp = __shfl_sync(kFull, p, l), le = __shfl_sync(kFull, e, l);
			while (lcnt < k && lc < lnr) {
				for (unsigned base = lp; base < le && lcnt < k; base += 32) {
					const unsigned pp = base + lane;
					bool ok = false;
					if (pp < le) {
						const float4 cd = __ldg(sorted + pp);
						ok = dist2_ref(lx, ly, lz, cd.x, cd.y, cd.z) <= thr;
					}
					lcnt += __popc(__ballot_sync(kFull, ok));
				}
				if (++lc < lnr) { const uint2 r = s_rng[warp][lc][l]; lp = r.x; le = r.x + r.y; }
			}
			if (lane == l) cnt = lcnt;
		}

		const bool kept = valid && cnt >= k;
		if (valid) keep[g] = (uint8_t)(kept ? 1 : 0);
		const unsigned km = __ballot_sync(kFull, kept);
		if (lane == 0 && km) atomicAdd(&ctl->n_kept, __popc(km));
		__syncwarp();
	}
}

// after the count: the slots this run's points touched go back to zero (the table is cleared by its users, not by a memset of
// all of it: 4 M slots for 0.1 M occupied ones on the bench frame)
__global__ void __launch_bounds__(256) k_voxel_cleanup(const unsigned *__restrict__ slot_of, VoxBucket *table, const FrameCtl *ctl) {
	pdl_enter();
	const int N = ctl->n_culled;
	const uint4 z = make_uint4(0u, 0u, 0u, 0u);
	for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < N; g += gridDim.x * blockDim.x) {
		uint4 *b = reinterpret_cast<uint4 *>(table + (slot_of[g] >> 3));
		b[0] = z; b[1] = z; b[2] = z; b[3] = z;
	}
}

// ------------------------------------------------------------------------------------------------------
// K6: stable compaction of the filter survivors == the merged cloud
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads) k_filter_compact(const uint4 *__restrict__ cloud, const uint8_t *__restrict__ keep,
	const int *__restrict__ culled_starts, int s_first, int s_end, FrameCtl *ctl, unsigned long long *status, int *final_starts,
	uint4 *__restrict__ out, const int *__restrict__ d_out_offset, int *__restrict__ old_to_new, PeerDst peers)
{
	pdl_enter();
	__shared__ uint4 stage[kTile];
	__shared__ unsigned sm[16];
	__shared__ int s_tile;
	const int N = ctl->n_culled;
	const int ntiles = (N + kTile - 1) / kTile;
	const int out_off = d_out_offset ? *d_out_offset : 0;
	const int tid = threadIdx.x;

	for (;;) {
		if (tid == 0) s_tile = (int)atomicAdd(&ctl->tile_counter_b, 1u);
		__syncthreads();
		const int tile = s_tile;
		if (tile >= ntiles) break;
		const int g0 = tile * kTile + tid * 8;
		unsigned valid = 0;
		if (g0 + 8 <= N) {
			const uint2 f = *reinterpret_cast<const uint2 *>(keep + g0);
#pragma unroll
			for (int j = 0; j < 4; j++) {
				if ((f.x >> (8 * j)) & 0xff) valid |= 1u << j;
				if ((f.y >> (8 * j)) & 0xff) valid |= 1u << (4 + j);
			}
		} else {
#pragma unroll
			for (int j = 0; j < 8; j++)
				if (g0 + j < N && keep[g0 + j]) valid |= 1u << j;
		}
		const unsigned cnt = __popc(valid);
		unsigned total, base;
		const unsigned off = tile_scan(cnt, sm, status, tile, &ctl->err, &total, &base);
		{
			unsigned o = off;
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const bool kept = (valid >> j) & 1;
				if (kept) stage[o] = cloud[g0 + j];
				if (g0 + j < N) old_to_new[g0 + j] = kept ? (int)(base + o) : -1;
				if (kept) o++;
			}
		}
		// sensor boundaries of the merged cloud
		for (int s = s_first + 1; s < s_end; s++) {
			const int bnd = culled_starts[s];
			if (bnd >= g0 && bnd < g0 + 8 && bnd < N)
				final_starts[s] = (int)(base + off + __popc(valid & ((1u << (bnd - g0)) - 1u)));
		}
		if (tid == 0 && tile == ntiles - 1) {
			const int n_final = (int)(base + total);
			ctl->n_final = n_final;
			for (int s = s_first + 1; s <= s_end; s++)
				if (culled_starts[s] >= N) final_starts[s] = n_final;
		}
		if (tid == 0 && tile == 0) final_starts[s_first] = 0;
		__syncthreads();
		if (peers.n == 0) {
			uint4 *o4 = out + out_off + base;
			for (unsigned i = tid; i < total; i += kScanThreads) o4[i] = stage[i];
		} else {
			for (int p = 0; p < peers.n; p++) {
				uint4 *o4 = peers.ptr[p] + out_off + base;
				for (unsigned i = tid; i < total; i += kScanThreads) o4[i] = stage[i];
			}
		}
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------------
// standalone filter helpers: pack Point3f + RGB into 16-byte records (+ bounding box), grid parameters on
// the device, unpack survivors
// ------------------------------------------------------------------------------------------------------
struct FilterBox { unsigned mn[3], mx[3]; };   // order-preserving encodings

__global__ void __launch_bounds__(256) k_filter_pack(const float *__restrict__ verts, const unsigned *__restrict__ colors, int n,
	uint4 *__restrict__ cloud, FilterBox *box)
{
	float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const float x = verts[3 * (size_t)i], y = verts[3 * (size_t)i + 1], z = verts[3 * (size_t)i + 2];
		cloud[i] = make_uint4(colors[i], __float_as_uint(x), __float_as_uint(y), __float_as_uint(z));
		if (isfinite(x)) { mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x); }
		if (isfinite(y)) { mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y); }
		if (isfinite(z)) { mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z); }
	}
#pragma unroll
	for (int a = 0; a < 3; a++) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			mn[a] = fminf(mn[a], __shfl_xor_sync(kFull, mn[a], o));
			mx[a] = fmaxf(mx[a], __shfl_xor_sync(kFull, mx[a], o));
		}
	}
	if ((threadIdx.x & 31) == 0) {
#pragma unroll
		for (int a = 0; a < 3; a++) {
			atomicMin(&box->mn[a], f2ord(mn[a]));
			atomicMax(&box->mx[a], f2ord(mx[a]));
		}
	}
}

// Voxel-grid parameters from a bounding box.  Cell edge h = 1.01 r: two points with fp32 d2 <= fl(r^2) are
// closer than r(1+1e-6) on every axis, the fp32 voxel coordinate u = fl(fl(x-o)*inv_h) carries an absolute
// error below 2e-3 (u <= 8192), so their coordinates differ by less than 1 and their voxels by at most 1:
// the 27-voxel neighbourhood provably contains every point that can count.  When the box would need more than
// 2^13 voxels per axis the cells grow instead (still correct, only more candidates).
__host__ __device__ inline void voxel_grid_params(const double lo[3], const double hi[3], double r, float origin[3], float *inv_h) {
	double ext = 0;
	for (int a = 0; a < 3; a++) {
		double e = hi[a] - lo[a];
		if (!(e >= 0)) e = 0;
		if (e > ext) ext = e;
	}
	if (!(ext < 1e30)) ext = 1e30;
	double h = r * 1.01;
	const double hmin = ext / (double)(kCellMax - 3);
	if (h < hmin) h = hmin * 1.0001;
	if (!(h > 1e-30)) h = 1e-30;
	for (int a = 0; a < 3; a++) {
		double l = lo[a];
		if (!(l > -1e30)) l = -1e30;
		if (!(l < 1e30)) l = 1e30;
		origin[a] = (float)(l - h);
	}
	*inv_h = (float)(1.0 / h);
}

__global__ void k_filter_grid_params(const FilterBox *box, SensorDesc *sd, FrameCtl *ctl, int *culled_starts, int n, float max_dist) {
	double lo[3], hi[3];
	for (int a = 0; a < 3; a++) {
		lo[a] = (double)ord2f(box->mn[a]);
		hi[a] = (double)ord2f(box->mx[a]);
		if (!(lo[a] <= hi[a])) { lo[a] = 0; hi[a] = 0; }       // no finite coordinate on this axis
	}
	float o[3], inv_h;
	voxel_grid_params(lo, hi, (double)max_dist, o, &inv_h);
	sd[0].gox = o[0]; sd[0].goy = o[1]; sd[0].goz = o[2]; sd[0].ginv_h = inv_h;
	ctl->n_culled = n;
	ctl->n_final = n;
	culled_starts[0] = 0;
	culled_starts[1] = n;
}

__global__ void __launch_bounds__(256) k_filter_unpack(const uint4 *__restrict__ cloud, const FrameCtl *ctl, float *__restrict__ verts, unsigned *__restrict__ colors) {
	const int n = ctl->n_final;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint4 p = cloud[i];
		colors[i] = p.x;
		verts[3 * (size_t)i] = __uint_as_float(p.y);
		verts[3 * (size_t)i + 1] = __uint_as_float(p.z);
		verts[3 * (size_t)i + 2] = __uint_as_float(p.w);
	}
}

}  // namespace ls3d

using namespace ls3d;

// ======================================================================================================
// Frame context
// ======================================================================================================
// stage slots of the optional timing pass (ls3d_frame_stage_ms)
enum { kTsMap = 0, kTsHashClear, kTsInsert, kTsRanges, kTsCount, kTsCompact, kTsOrganized, kTsWhole, kTsTriangles, kTsN };
enum { kModeAuto = 0, kModeVoxelHash = 1, kModeOrganized = 2 };

constexpr int kMaxChunks = 16, kEvN = 5 * kMaxChunks + 5;
struct HostGraphKey { const void *depth, *colors, *out; int first, n_run, chunks, pull; unsigned long long params_version; };
struct Ls3dFrame {
	int device = 0;
	int S = 0;
	std::vector<int> w, h;
	long long total_px = 0;
	int total_tiles = 0;
	int max_w = 0, max_h = 0;
	size_t depth_bytes = 0, color_bytes = 0;
	unsigned total_slots = 0;
	bool table_dirty = true;           // the voxel table is not known to be all zero (see frame_filter_stages)
	std::vector<SensorDesc> h_sd;       // S+1 entries (sentinel last)
	SensorDesc *pin_sd = nullptr;       // pinned staging for the descriptor upload
	float *pin_rays = nullptr;          // pinned staging for the ray tables
	cudaEvent_t ev_staged = nullptr;    // recorded after every upload out of pin_sd / pin_rays: waited for before they are rewritten
	size_t n_rays = 0;
	float bounds[6] = {0, 0, 0, 0, 0, 0};
	int filter_k = 0;
	float filter_max_dist = 0, filter_thr = 0;
	bool filter_on = false;
	bool params_set = false;
	std::vector<float> params_key;      // bit pattern of the arguments of the last successful frame_set_params
	int filter_mode = kModeAuto;
	bool organized_ok = false;          // every sensor's pose/intrinsics admit the pixel-window bound
	bool organized_wide = false;        // ... but typical windows exceed the halo: auto mode prefers the voxel hash
	bool last_organized = false;        // what the last count stage ran (the merge stage has to match)
	const void *last_depth = nullptr, *last_colors = nullptr;
	int sm_count = 148;

	// device memory
	DevBuf sd, tile_sensor, rays, zero, cloud0, sorted, final_, slot_of, rank_of, keep, keep_px, map, d2v, table, in_depth, in_colors, box, tri;
	// carve-outs of `zero` (re-zeroed by one memset per run)
	FrameCtl *ctl = nullptr;
	unsigned long long *status_a = nullptr, *status_b = nullptr;
	int *culled_starts = nullptr, *final_starts = nullptr, *tri_starts = nullptr;
	size_t zero_bytes = 0;
	// pinned read-back block: FrameCtl + starts
	int *pin_out = nullptr;
	bool want_d2v = false;
	// host-buffer path only: the colour upload runs on a second stream; K1 (the only consumer of colours) waits for this event
	cudaEvent_t colors_ready = nullptr;
	cudaEvent_t ev_colors = nullptr, ev_count = nullptr;
	cudaEvent_t ev_up[kEvN] = {};      // host path: per chunk depth-landed / colours-landed / counted / merged, then fork + 3 joins
	cudaStream_t st_merge = nullptr;
	cudaEvent_t ev_tr[4 * kMaxChunks + 1] = {};    // LS3D_E2E_TRACE only: timed events (external records, so they also work inside the graph)
	cudaGraphExec_t hg_exec = nullptr;  // host path: the captured per-frame schedule, valid for hg_key
	HostGraphKey hg_key = {};
	int hg_launches = 0;
	unsigned long long params_version = 0;   // bumped whenever frame_set_params changes anything on the device
	cudaStream_t st_colors = nullptr, st_out = nullptr;     // host path: upload stream, read-back stream
	// host path, pageable caller buffers: page-locked (mapped) staging the caller's depth / colours are copied into by the host copy pool
	unsigned char *stage_depth = nullptr, *stage_colors = nullptr;
	cudaGraphExec_t hg_exec_b = nullptr;   // ... and the second half of the schedule (merge + read-back), launched once the colours are staged
	HostGraphKey hg_key_b = {};
	int hg_launches_b = 0;
	bool want_triangles = false;   // run the triangle stage after K1 (unfiltered runs only)
	int *tri_override = nullptr;   // host mesh path: where this run's triangles go (a per-chunk region of `tri`)
	DevBuf acc;                    // host mesh path: running totals + per-chunk records (k_mesh_chunk_done)
	int *pin_acc = nullptr;
	cudaGraphExec_t hm_exec = nullptr;   // host mesh path: the captured schedule, valid for hm_key
	HostGraphKey hm_key = {};
	const void *hm_tri_out = nullptr;
	int hm_launches = 0;
	// optional per-stage timing (bench.py's roofline pass): a (begin, end) event pair per stage on the run's stream
	bool timing = false;
	cudaEvent_t ev[kTsN][2] = {};
	int ev_recorded = 0;     // bit mask of stages with both events recorded
};

static void stage_begin(Ls3dFrame *f, int i, cudaStream_t st) {
	if (!f->timing) return;
	for (int j = 0; j < 2; j++)
		if (!f->ev[i][j] && cudaEventCreate(&f->ev[i][j]) != cudaSuccess) { f->ev[i][j] = nullptr; return; }
	cudaEventRecord(f->ev[i][0], st);
}
static void stage_end(Ls3dFrame *f, int i, cudaStream_t st) {
	if (!f->timing || !f->ev[i][1]) return;
	if (cudaEventRecord(f->ev[i][1], st) == cudaSuccess) f->ev_recorded |= 1 << i;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void frame_free(Ls3dFrame *f) {
	if (!f) return;
	DevBuf *bufs[] = {&f->sd, &f->tile_sensor, &f->rays, &f->zero, &f->cloud0, &f->sorted, &f->final_, &f->slot_of, &f->rank_of, &f->keep, &f->keep_px, &f->map, &f->d2v,
		&f->table, &f->in_depth, &f->in_colors, &f->box, &f->tri};
	for (DevBuf *b : bufs) b->release();
	if (f->pin_sd) cudaFreeHost(f->pin_sd);
	if (f->pin_rays) cudaFreeHost(f->pin_rays);
	if (f->ev_staged) cudaEventDestroy(f->ev_staged);
	if (f->pin_out) cudaFreeHost(f->pin_out);
	for (auto &e : f->ev) for (cudaEvent_t x : e) if (x) cudaEventDestroy(x);
	if (f->ev_colors) cudaEventDestroy(f->ev_colors);
	if (f->ev_count) cudaEventDestroy(f->ev_count);
	for (cudaEvent_t x : f->ev_up) if (x) cudaEventDestroy(x);
	if (f->st_colors) cudaStreamDestroy(f->st_colors);
	if (f->st_out) cudaStreamDestroy(f->st_out);
	if (f->st_merge) cudaStreamDestroy(f->st_merge);
	if (f->stage_depth) cudaFreeHost(f->stage_depth);
	if (f->stage_colors) cudaFreeHost(f->stage_colors);
	if (f->hg_exec_b) cudaGraphExecDestroy(f->hg_exec_b);
	for (cudaEvent_t x : f->ev_tr) if (x) cudaEventDestroy(x);
	if (f->hg_exec) cudaGraphExecDestroy(f->hg_exec);
	if (f->hm_exec) cudaGraphExecDestroy(f->hm_exec);
	if (f->pin_acc) cudaFreeHost(f->pin_acc);
	f->acc.release();
	delete f;
}

extern "C" Ls3dFrame *ls3d_frame_create(int n_maps, const int *widths, const int *heights) {
	clear_error();
	if (!ensure_device()) return nullptr;
	if (n_maps <= 0 || n_maps > 65535 || !widths || !heights) { set_error("ls3d_frame_create: n_maps must be in 1..65535"); return nullptr; }
	Ls3dFrame *f = new Ls3dFrame();
	cudaGetDevice(&f->device);
	cudaDeviceGetAttribute(&f->sm_count, cudaDevAttrMultiProcessorCount, f->device);
	f->S = n_maps;
	f->w.assign(widths, widths + n_maps);
	f->h.assign(heights, heights + n_maps);
	f->h_sd.resize(n_maps + 1);
	long long px_acc = 0, depth_off = 0, color_off = 0;
	int tile_acc = 0;
	unsigned long long slot_acc = 0;
	std::vector<unsigned short> tile_sensor;
	for (int i = 0; i <= n_maps; i++) {
		SensorDesc &d = f->h_sd[i];
		memset(&d, 0, sizeof(d));
		d.tile_begin = tile_acc;
		d.depth_off = depth_off;
		d.color_off = color_off;
		d.pix_begin = px_acc;
		d.tbl_off = (unsigned)slot_acc;
		if (i == n_maps) break;
		if (widths[i] <= 0 || heights[i] <= 0 || (long long)widths[i] * heights[i] >= (1ll << 24)) {
			set_error("ls3d_frame_create: map %d has unsupported size %dx%d (need 0 < w*h < 2^24)", i, widths[i], heights[i]);
			delete f;
			return nullptr;
		}
		const long long px = (long long)widths[i] * heights[i];
		d.w = widths[i]; d.h = heights[i]; d.px = (int)px;
		f->max_w = std::max(f->max_w, d.w);
		f->max_h = std::max(f->max_h, d.h);
		unsigned long long cap = 64;
		while (cap < (unsigned long long)px * 9 / 8 + 64) cap <<= 1;       // runs of 8 voxels: never more of them than points
		d.tbl_mask = (unsigned)(cap - 1);
		slot_acc += cap;
		px_acc += px;
		depth_off += px * 2;
		color_off += px * 3;
		d.ray_off = (int)f->n_rays;
		f->n_rays += 2 * ((size_t)widths[i] + (size_t)heights[i]);      // xn[w], yn[h], then the window-reach tables xreach[w], yreach[h]
		const int nt = (int)((px + kTile - 1) / kTile);
		tile_sensor.insert(tile_sensor.end(), nt, (unsigned short)i);
		tile_acc += nt;
	}
	if (px_acc >= (1ll << 31) - kTile || slot_acc >= (1ull << 29)) { set_error("ls3d_frame_create: frame too large"); delete f; return nullptr; }
	f->total_px = px_acc;
	f->total_tiles = tile_acc;
	f->depth_bytes = (size_t)depth_off;
	f->color_bytes = (size_t)color_off;
	f->total_slots = (unsigned)slot_acc;

	const size_t n = (size_t)f->total_px;
	const size_t ctl_b = align_up(sizeof(FrameCtl), 256);
	const size_t st_b = align_up(sizeof(unsigned long long) * (size_t)(f->total_tiles + 1), 256);
	const size_t starts_b = align_up(sizeof(int) * (size_t)(n_maps + 1), 256);
	f->zero_bytes = ctl_b + 2 * st_b + 3 * starts_b;
	bool ok = f->sd.reserve(sizeof(SensorDesc) * (n_maps + 1), "alloc descriptors") && f->zero.reserve(f->zero_bytes, "alloc control block") &&
		f->tile_sensor.reserve(sizeof(unsigned short) * (tile_sensor.size() + 1), "alloc tile table") && f->rays.reserve(sizeof(float) * (f->n_rays + 4), "alloc ray tables") &&
		f->cloud0.reserve(16 * n, "alloc culled cloud") && f->sorted.reserve(16 * n + 64, "alloc sorted cloud") && f->final_.reserve(16 * n, "alloc merged cloud") &&
		f->slot_of.reserve(4 * n, "alloc slots") && f->rank_of.reserve(4 * n, "alloc ranks") && f->keep.reserve(align_up(n, 16) + 16, "alloc keep flags") &&
		f->keep_px.reserve(align_up(n, 16) + 16, "alloc pixel keep flags") &&
		f->map.reserve(4 * n, "alloc index map") && f->d2v.reserve(4 * n, "alloc pixel map") &&
		f->table.reserve(sizeof(VoxBucket) * (size_t)f->total_slots, "alloc voxel hash") &&
		f->box.reserve(sizeof(FilterBox), "alloc bbox");
	ok = ok && cuda_ok(cudaHostAlloc((void **)&f->pin_sd, sizeof(SensorDesc) * (n_maps + 1), cudaHostAllocDefault), "alloc pinned descriptors");
	ok = ok && cuda_ok(cudaHostAlloc((void **)&f->pin_rays, sizeof(float) * (f->n_rays + 4), cudaHostAllocDefault), "alloc pinned ray tables");
	ok = ok && cuda_ok(cudaEventCreateWithFlags(&f->ev_staged, cudaEventDisableTiming), "create staging event");
	ok = ok && cuda_ok(cudaHostAlloc((void **)&f->pin_out, sizeof(int) * (size_t)(16 + 3 * (n_maps + 1)), cudaHostAllocDefault), "alloc pinned read-back");
	ok = ok && cuda_ok(cudaMemcpy(f->tile_sensor.p, tile_sensor.data(), sizeof(unsigned short) * tile_sensor.size(), cudaMemcpyHostToDevice), "upload tile table");
	if (!ok) { frame_free(f); return nullptr; }
	uint8_t *z = f->zero.as<uint8_t>();
	f->ctl = reinterpret_cast<FrameCtl *>(z);
	f->status_a = reinterpret_cast<unsigned long long *>(z + ctl_b);
	f->status_b = reinterpret_cast<unsigned long long *>(z + ctl_b + st_b);
	f->culled_starts = reinterpret_cast<int *>(z + ctl_b + 2 * st_b);
	f->final_starts = reinterpret_cast<int *>(z + ctl_b + 2 * st_b + starts_b);
	f->tri_starts = reinterpret_cast<int *>(z + ctl_b + 2 * st_b + 2 * starts_b);
	cudaMemset(f->zero.p, 0, f->zero_bytes);
	return f;
}

extern "C" void ls3d_frame_destroy(Ls3dFrame *f) { frame_free(f); }

// Axis-aligned bounds of everything sensor `d` can produce (any u16 depth), intersected with the cull box.
static void sensor_world_bounds(const SensorDesc &d, const float *b, double lo[3], double hi[3]) {
	const double zmax = 65.535;
	const double fx = fabs((double)d.fx), fy = fabs((double)d.fy);
	const double xr = fmax(fabs(0.0 - d.cx), fabs((double)(d.w - 1) - d.cx)) / fx * zmax;
	const double yr = fmax(fabs((double)d.cy - 0.0), fabs((double)d.cy - (double)(d.h - 1))) / fy * zmax;
	const double clo[3] = {-xr + d.t[0], -yr + d.t[1], 0.0 + d.t[2]};
	const double chi[3] = {xr + d.t[0], yr + d.t[1], zmax + d.t[2]};
	for (int i = 0; i < 3; i++) {
		double l = 0, h = 0;
		for (int j = 0; j < 3; j++) {
			const double r = d.R[3 * i + j];
			const double a = r * clo[j], c = r * chi[j];
			l += fmin(a, c);
			h += fmax(a, c);
		}
		const double pad = 1e-3 * (fabs(l) + fabs(h)) + 1e-3;
		l -= pad; h += pad;
		if (!(l == l) || !(h == h) || !std::isfinite(l) || !std::isfinite(h)) { l = -1e30; h = 1e30; }
		const double bl = (double)b[i], bh = (double)b[3 + i];
		if (bl == bl && bl > l) l = bl;
		if (bh == bh && bh < h) h = bh;
		if (h < l) h = l;
		lo[i] = l; hi[i] = h;
	}
}

// smallest singular value of a row-major 3x3 (Jacobi on R^T R, fp64)
static double sigma_min3(const float *R) {
	double A[3][3];
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) {
			double s = 0;
			for (int k = 0; k < 3; k++) s += (double)R[3 * k + i] * (double)R[3 * k + j];
			A[i][j] = s;
		}
	for (int sweep = 0; sweep < 40; sweep++) {
		const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
		if (!(off > 1e-300)) break;
		for (int p = 0; p < 2; p++)
			for (int q = p + 1; q < 3; q++) {
				if (fabs(A[p][q]) < 1e-300) continue;
				const double th = 0.5 * atan2(2 * A[p][q], A[q][q] - A[p][p]);
				const double c = cos(th), sn = sin(th);
				for (int k = 0; k < 3; k++) { const double a = A[k][p], b2 = A[k][q]; A[k][p] = c * a - sn * b2; A[k][q] = sn * a + c * b2; }
				for (int k = 0; k < 3; k++) { const double a = A[p][k], b2 = A[q][k]; A[p][k] = c * a - sn * b2; A[q][k] = sn * a + c * b2; }
			}
	}
	const double e = fmin(A[0][0], fmin(A[1][1], A[2][2]));
	return (e > 0 && e == e) ? sqrt(e) : 0.0;
}

// n_set: how many leading sensors intr_params / wtransform_params describe (<= f->S)
static int frame_set_params(Ls3dFrame *f, int n_set, const float *intr_params, const float *wtransform_params,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ, int filter_k, float filter_maxDist, void *stream)
{
	if (!f || !intr_params || !wtransform_params) { set_error("ls3d_frame_set_params: null argument"); return -1; }
	{
		// a live rig sends the same calibration frame after frame: skip the table rebuild and both uploads when nothing changed
		// (the device copies are only ever written here, in stream order)
		std::vector<float> key;
		key.reserve(10 + 19 * (size_t)n_set);
		const float head[9] = {minX, minY, minZ, maxX, maxY, maxZ, (float)filter_k, filter_maxDist, (float)n_set};
		key.insert(key.end(), head, head + 9);
		key.insert(key.end(), intr_params, intr_params + 7 * (size_t)n_set);
		key.insert(key.end(), wtransform_params, wtransform_params + 12 * (size_t)n_set);
		if (f->params_set && key.size() == f->params_key.size() && !memcmp(key.data(), f->params_key.data(), sizeof(float) * key.size())) return 0;
		f->params_key.swap(key);
		f->params_set = false;
		f->params_version++;
	}
	// the previous parameter set may still be on its way out of the pinned staging blocks (set_params(A); run; set_params(B) on
	// a busy stream): wait for that copy before the host rewrites them
	if (!cuda_ok(cudaEventSynchronize(f->ev_staged), "wait for the previous parameter upload")) return -1;
	f->bounds[0] = minX; f->bounds[1] = minY; f->bounds[2] = minZ; f->bounds[3] = maxX; f->bounds[4] = maxY; f->bounds[5] = maxZ;
	f->filter_on = filter_k > 0 && filter_maxDist > 0;        // filter.cpp:38-41
	f->filter_k = filter_k;
	f->filter_max_dist = filter_maxDist;
	f->filter_thr = (float)pow((double)filter_maxDist, 2.0);  // filter.cpp:52: float = pow(float, int)
	f->organized_ok = f->filter_on;
	f->organized_wide = false;
	for (int i = 0; i < n_set && i < f->S; i++) {
		SensorDesc &d = f->h_sd[i];
		const float *ip = intr_params + 7 * i;                 // IntrinsicCameraParameters(float*), depthprocessing.h:96-97
		d.cx = ip[0]; d.cy = ip[1]; d.fx = ip[2]; d.fy = ip[3];
		const float *tp = wtransform_params + 12 * i;          // WorldTranformation(float*), depthprocessing.h:56-63
		memcpy(d.t, tp, 3 * sizeof(float));
		memcpy(d.R, tp + 3, 9 * sizeof(float));
		// ray table: the reference's own two fp32 operations per coordinate (depthprocessing.cpp:151-152), once per column / row
		{
			float *xr = f->pin_rays + d.ray_off, *yr = xr + d.w;
			for (int x = 0; x < d.w; x++) { volatile float a = (float)x - d.cx; xr[x] = a / d.fx; }
			for (int y = 0; y < d.h; y++) { volatile float a = d.cy - (float)y; yr[y] = a / d.fy; }
			// organized neighbour count: pixels a camera-space radius r' spans per unit of r'/(Z - r') (rounded up a little)
			float *xw = yr + d.h, *yw = xw + d.w;
			for (int x = 0; x < d.w; x++) xw[x] = (float)(fabs((double)d.fx) * sqrt(1.0 + (double)xr[x] * (double)xr[x]) * (1.0 + 1e-6));
			for (int y = 0; y < d.h; y++) yw[y] = (float)(fabs((double)d.fy) * sqrt(1.0 + (double)yr[y] * (double)yr[y]) * (1.0 + 1e-6));
		}
		d.org_rp = 0.0f;
		if (f->filter_on) {
			double lo[3], hi[3];
			sensor_world_bounds(d, f->bounds, lo, hi);
			float o[3];
			voxel_grid_params(lo, hi, (double)filter_maxDist, o, &d.ginv_h);
			d.gox = o[0]; d.goy = o[1]; d.goz = o[2];
			// organized path: world fp32 d2 <= thr  ==>  camera-space distance <= r'
			//   true world distance <= sqrt(thr)(1+1e-6); fp32 world coordinates carry <= ~4e-6 * |coordinate| of rounding;
			//   |R(a-b)| >= sigma_min |a-b|
			double mag = 0;
			for (int a = 0; a < 3; a++) mag = fmax(mag, fmax(fabs(lo[a]), fabs(hi[a])));
			const double smin = sigma_min3(d.R);
			const double rp = (sqrt((double)f->filter_thr) * (1.0 + 1e-5) + 4e-6 * mag) / smin * (1.0 + 1e-4);
			const bool fin = std::isfinite(d.fx) && std::isfinite(d.fy) && std::isfinite(d.cx) && std::isfinite(d.cy) && d.fx != 0.0f && d.fy != 0.0f;
			if (smin > 1e-3 && mag < 1e6 && fin && std::isfinite(rp) && rp > 0) d.org_rp = (float)rp;
			else f->organized_ok = false;
			// the pixel window of a point 1 m away should still fit the shared-memory halo; beyond that most windows fall back to the
			// global-memory walk and the voxel hash is the faster enumeration (measured at 1920x1080, r = 1 cm: 1.41 vs 0.77 ms)
			if (f->organized_ok && fmax(fabs((double)d.fx), fabs((double)d.fy)) * rp / (1.0 - fmin(rp, 0.5)) > (double)kOrgHalo) f->organized_wide = true;
		}
	}
	memcpy(f->pin_sd, f->h_sd.data(), sizeof(SensorDesc) * (f->S + 1));
	if (!cuda_ok(cudaMemcpyAsync(f->sd.p, f->pin_sd, sizeof(SensorDesc) * (f->S + 1), cudaMemcpyHostToDevice, (cudaStream_t)stream), "upload descriptors")) return -1;
	if (!cuda_ok(cudaMemcpyAsync(f->rays.p, f->pin_rays, sizeof(float) * f->n_rays, cudaMemcpyHostToDevice, (cudaStream_t)stream), "upload ray tables")) return -1;
	if (!cuda_ok(cudaEventRecord(f->ev_staged, (cudaStream_t)stream), "record parameter upload")) return -1;
	f->params_set = true;
	return 0;
}

extern "C" int ls3d_frame_set_params(Ls3dFrame *f, const float *intr_params, const float *wtransform_params,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ, int filter_k, float filter_maxDist, void *stream)
{
	return frame_set_params(f, f ? f->S : 0, intr_params, wtransform_params, minX, minY, minZ, maxX, maxY, maxZ, filter_k, filter_maxDist, stream);
}

extern "C" int ls3d_frame_set_filter_mode(Ls3dFrame *f, int mode) {
	if (!f || mode < kModeAuto || mode > kModeOrganized) { set_error("ls3d_frame_set_filter_mode: mode must be 0 (auto), 1 (voxel hash) or 2 (organized)"); return -1; }
	f->filter_mode = mode;
	return 0;
}

enum { kStageCount = 1, kStageMerge = 2 };

__global__ void __launch_bounds__(256) k_keep_all(uint8_t *keep, FrameCtl *ctl) {
	const int N = ctl->n_culled;
	for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < N; g += gridDim.x * blockDim.x) keep[g] = 1;
	if (blockIdx.x == 0 && threadIdx.x == 0) ctl->n_kept = N;
}

// Host path: copies the records that the tiles [tile_lo, tile_hi) of a run produced from the device buffer into the mapped
// pinned output block.  Every warp store is 512 contiguous bytes, so the PCIe writes are full-size; the record range comes
// from the per-tile survivor counts (no host round trip to size the read-back).
__global__ void __launch_bounds__(256) k_copy_out(const uint4 *__restrict__ src, uint4 *__restrict__ dst, const unsigned *__restrict__ tile_count,
	int tile0, int tile_lo, int tile_hi)
{
	__shared__ unsigned s_part[2][8];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	unsigned lo = 0, mid = 0;
	for (int t = tid; t < tile_hi; t += 256) {
		const unsigned c = __ldg(tile_count + tile0 + t);
		if (t < tile_lo) lo += c; else mid += c;
	}
	lo = warp_sum(lo); mid = warp_sum(mid);
	if (lane == 0) { s_part[0][warp] = lo; s_part[1][warp] = mid; }
	__syncthreads();
	unsigned begin = 0, n = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) { begin += s_part[0][i]; n += s_part[1][i]; }
	for (unsigned i = blockIdx.x * 256 + tid; i < n; i += gridDim.x * 256) dst[begin + i] = src[begin + i];
}

// Host mesh path (unfiltered frame + triangles, sensors processed in chunks): after a chunk's K1 + triangle kernels, record where its
// vertices and triangles go in the concatenated result and advance the running totals the next chunk's K1 reads as its output offset.
//   acc[0] vertices so far, acc[1] triangles so far, acc[2] error flags; acc[4 + 4c ..] = {first vertex, vertices, first triangle,
//   triangles} of chunk c; acc[4 + 4 * kMaxMeshChunks + s] = vertices of sensor s
constexpr int kMaxMeshChunks = 16;
__global__ void k_mesh_chunk_done(const FrameCtl *ctl, const int *culled_starts, int s_first, int s_end, int chunk, int *acc) {
	for (int s = s_first + (int)threadIdx.x; s < s_end; s += (int)blockDim.x) acc[4 + 4 * kMaxMeshChunks + s] = culled_starts[s + 1] - culled_starts[s];
	if (threadIdx.x == 0) {
		const int nv = ctl->n_final, nt = ctl->n_triangles;
		int *rec = acc + 4 + 4 * chunk;
		rec[0] = acc[0]; rec[1] = nv; rec[2] = acc[1]; rec[3] = nt;
		acc[0] += nv; acc[1] += nt; acc[2] |= ctl->err;
	}
}
// ... and store the chunk into the page-locked Mesh blocks: its records as they are, its (chunk-relative) triangle indices rebased
// by the chunk's first vertex — formMesh's index offset (depthprocessing.cpp:1611-1626) applied on the way out.
__global__ void __launch_bounds__(256) k_copy_mesh_out(const uint4 *__restrict__ verts, uint4 *__restrict__ v_host, const int *__restrict__ tris, int *__restrict__ t_host,
	const int *__restrict__ rec)
{
	const int v0 = rec[0], nv = rec[1], t0 = rec[2], nt = rec[3];
	for (int i = blockIdx.x * 256 + threadIdx.x; i < nv; i += gridDim.x * 256) v_host[v0 + i] = verts[v0 + i];
	// indices: whole 16-byte pieces of the destination (the PCIe writes stay 512 bytes per warp), ragged ends one by one
	const long long lo = 3ll * t0, hi = lo + 3ll * nt;                 // destination range, in ints
	const long long a4 = (lo + 3) / 4, b4 = hi / 4;                    // aligned int4 pieces fully inside it
	if (a4 < b4) {
		for (long long j = a4 + blockIdx.x * 256 + threadIdx.x; j < b4; j += (long long)gridDim.x * 256) {
			const int *src = tris + (4 * j - lo);
			reinterpret_cast<int4 *>(t_host)[j] = make_int4(src[0] + v0, src[1] + v0, src[2] + v0, src[3] + v0);
		}
		if (blockIdx.x == 0) {
			for (long long i = lo + threadIdx.x; i < 4 * a4; i += 256) t_host[i] = tris[i - lo] + v0;
			for (long long i = 4 * b4 + threadIdx.x; i < hi; i += 256) t_host[i] = tris[i - lo] + v0;
		}
	} else if (blockIdx.x == 0) {
		for (long long i = lo + threadIdx.x; i < hi; i += 256) t_host[i] = tris[i - lo] + v0;
	}
}

// K1 launcher.  keep_px != nullptr: AND the organized neighbour-count mask into the validity test.  after_count: the organized count
// of the same sensors is the previous launch in this stream (nothing in between) — K1 is then launched programmatically: its blocks
// move in as the count's last blocks leave, fetch their depths, and wait for the count's results only then.
static int launch_map(Ls3dFrame *f, const void *d_depth, const void *d_colors, int s_first, int s_end, uint4 *out, const int *d_off,
	const uint8_t *keep_px, const PeerDst &peers, cudaStream_t st, int tile_lo = 0, int tile_hi = -1, bool after_count = false)
{
	const int ntiles = f->h_sd[s_end].tile_begin - f->h_sd[s_first].tile_begin;
	if (tile_hi < 0) tile_hi = ntiles;
	const int blocks = std::max(1, std::min(tile_hi - tile_lo, f->sm_count * 8));
	Bounds6 b;
	memcpy(b.v, f->bounds, sizeof(b.v));
	const uint8_t *dd = (const uint8_t *)d_depth, *dc = (const uint8_t *)d_colors;
	const SensorDesc *sd = f->sd.as<SensorDesc>();
	const unsigned short *ts = f->tile_sensor.as<unsigned short>();
	const float *rays = f->rays.as<float>();
	const unsigned *tc = reinterpret_cast<const unsigned *>(f->status_b);      // per-tile survivor counts left by the organized count
	if (f->colors_ready && !cuda_ok(cudaStreamWaitEvent(st, f->colors_ready, 0), "wait for the colour upload")) return -1;
	const size_t stage_bytes = peers.n > 0 ? (size_t)kTile * sizeof(uint4) : 0;      // 32 KB on top of up to 24 KB static: needs the opt-in
	if (stage_bytes) {
		static bool attr = false;
		if (!attr) {
			if (!cuda_ok(cudaFuncSetAttribute(k_map_cull_compact<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes), "merge staging") ||
				!cuda_ok(cudaFuncSetAttribute(k_map_cull_compact<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes), "merge staging") ||
				!cuda_ok(cudaFuncSetAttribute(k_map_cull_compact<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes), "merge staging") ||
				!cuda_ok(cudaFuncSetAttribute(k_map_cull_compact<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes), "merge staging")) return -1;
			attr = true;
		}
	}
	stage_begin(f, kTsMap, st);
	if (f->want_d2v || f->want_triangles) {
		if (keep_px) launch_chain(after_count, k_map_cull_compact<true, true>, dim3(blocks), kScanThreads, stage_bytes, st, dd, dc, sd, ts, rays, s_first, s_end, b, f->ctl, f->status_a, f->culled_starts, out, d_off, f->d2v.as<int>(), keep_px, tc, peers, tile_lo, tile_hi, 1.0f);
		else launch_chain(false, k_map_cull_compact<true, false>, dim3(blocks), kScanThreads, stage_bytes, st, dd, dc, sd, ts, rays, s_first, s_end, b, f->ctl, f->status_a, f->culled_starts, out, d_off, f->d2v.as<int>(), nullptr, nullptr, peers, tile_lo, tile_hi, 1.0f);
	} else {
		if (keep_px) launch_chain(after_count, k_map_cull_compact<false, true>, dim3(blocks), kScanThreads, stage_bytes, st, dd, dc, sd, ts, rays, s_first, s_end, b, f->ctl, f->status_a, f->culled_starts, out, d_off, nullptr, keep_px, tc, peers, tile_lo, tile_hi, 1.0f);
		else launch_chain(false, k_map_cull_compact<false, false>, dim3(blocks), kScanThreads, stage_bytes, st, dd, dc, sd, ts, rays, s_first, s_end, b, f->ctl, f->status_a, f->culled_starts, out, d_off, nullptr, nullptr, nullptr, peers, tile_lo, tile_hi, 1.0f);
	}
	stage_end(f, kTsMap, st);
	count_launch(1);
	return cuda_ok(cudaGetLastError(), "k_map_cull_compact") ? 1 : -1;
}

static int frame_merge_stage(Ls3dFrame *f, int s_first, int s_end, long long n_max, uint4 *dst, const int *d_dst_offset, const PeerDst &peers, cudaStream_t st) {
	const int tiles = (int)((n_max + kTile - 1) / kTile);
	stage_begin(f, kTsCompact, st);
	launch_chain(!f->timing, k_filter_compact, dim3((unsigned)(std::max(1, std::min(tiles, f->sm_count * 6)))), kScanThreads, 0, st, f->cloud0.as<uint4>(), f->keep.as<uint8_t>(),
		f->culled_starts, s_first, s_end, f->ctl, f->status_b, f->final_starts, dst, d_dst_offset, f->map.as<int>(), peers);
	stage_end(f, kTsCompact, st);
	count_launch(1);
	return cuda_ok(cudaGetLastError(), "k_filter_compact") ? 1 : -1;
}

// voxel-run neighbour count on the culled cloud in cloud0 (5 kernels), optionally followed by the compaction
static int frame_filter_stages(Ls3dFrame *f, int s_first, int s_end, long long n_max, uint4 *dst, const int *d_dst_offset, const PeerDst &peers, cudaStream_t st, int stages = kStageCount | kStageMerge) {
	const SensorDesc *sd = f->sd.as<SensorDesc>();
	VoxBucket *table = f->table.as<VoxBucket>();
	stage_begin(f, kTsHashClear, st);
	if (f->table_dirty) {
		// first use, or an earlier run was cut short before its clean-up kernel was enqueued
		if (!cuda_ok(cudaMemsetAsync(table, 0, sizeof(VoxBucket) * (size_t)f->total_slots, st), "clear voxel hash")) return -1;
	}
	f->table_dirty = true;
	stage_end(f, kTsHashClear, st);
	const int pt_blocks = (int)std::max<long long>(1, std::min<long long>((n_max + 255) / 256, (long long)f->sm_count * 8));
	stage_begin(f, kTsInsert, st);
	launch_chain(!f->timing, k_voxel_insert, dim3((unsigned)(pt_blocks)), 256, 0, st, f->cloud0.as<uint4>(), sd, f->culled_starts, s_first, s_end, f->ctl,
		table, f->slot_of.as<unsigned>(), f->rank_of.as<unsigned>());
	stage_end(f, kTsInsert, st);
	stage_begin(f, kTsRanges, st);
	launch_chain(!f->timing, k_bucket_alloc, dim3((unsigned)(pt_blocks)), 256, 0, st, f->slot_of.as<unsigned>(), f->rank_of.as<unsigned>(), table, f->ctl);
	launch_chain(!f->timing, k_cell_scatter, dim3((unsigned)(pt_blocks)), 256, 0, st, f->cloud0.as<uint4>(), f->slot_of.as<unsigned>(), f->rank_of.as<unsigned>(), table, f->ctl, f->sorted.as<float4>());
	stage_end(f, kTsRanges, st);
	stage_begin(f, kTsCount, st);
	launch_chain(!f->timing, k_neighbour_count, dim3((unsigned)(f->sm_count * 12)), kCountWarps * 32, 0, st, table, f->cloud0.as<uint4>(), f->sorted.as<float4>(), sd, f->culled_starts, s_first, s_end, f->ctl,
		f->filter_k, f->filter_thr, f->keep.as<uint8_t>());
	launch_chain(!f->timing, k_voxel_cleanup, dim3((unsigned)(pt_blocks)), 256, 0, st, f->slot_of.as<unsigned>(), table, f->ctl);
	stage_end(f, kTsCount, st);
	count_launch(5);
	if (!cuda_ok(cudaGetLastError(), "filter kernels")) return -1;
	f->table_dirty = false;                 // everything up to the clean-up is in the stream
	if (!(stages & kStageMerge)) return 5;
	return frame_merge_stage(f, s_first, s_end, n_max, dst, d_dst_offset, peers, st) < 0 ? -1 : 6;
}

// K1o: per-pixel survivor mask of sensors [s_first, s_end) straight from the depth images (the cloud is only materialised by the
// merge stage); adds to ctl->n_kept and the per-tile survivor counts, which the caller has zeroed
static int launch_organized_count(Ls3dFrame *f, const void *d_depth, int s_first, int s_end, cudaStream_t st, bool chained = false) {
	Bounds6 b;
	memcpy(b.v, f->bounds, sizeof(b.v));
	int mw = 0, mh = 0;
	for (int i = s_first; i < s_end; i++) { mw = std::max(mw, f->w[i]); mh = std::max(mh, f->h[i]); }
	const dim3 grid((mw + kOrgTW - 1) / kOrgTW, (mh + kOrgTH - 1) / kOrgTH, s_end - s_first);
	launch_chain(chained, k_organized_count, grid, kOrgTW * kOrgRows, 0, st, (const uint8_t *)d_depth, f->sd.as<SensorDesc>(), f->rays.as<float>(), s_first, b, f->ctl, f->filter_k, f->filter_thr,
		f->keep_px.as<uint8_t>(), reinterpret_cast<unsigned *>(f->status_b), 1.0f);
	count_launch(1);
	return cuda_ok(cudaGetLastError(), "k_organized_count") ? 0 : -1;
}

// stages: kStageCount = everything up to the survivor decision (n_kept known on the device); kStageMerge = the
// final compaction that places the survivors.  Both together is the normal single-GPU run; the multi-GPU merge
// runs them separately with the count exchange in between.
static int frame_run_impl(Ls3dFrame *f, const void *d_depth, const void *d_colors, int first_map, int n_run,
	uint4 *dst, const int *d_dst_offset, const PeerDst &peers, cudaStream_t st, int stages = kStageCount | kStageMerge)
{
	if (!f) { set_error("ls3d_frame_run: null frame"); return -1; }
	if (!f->params_set) { set_error("ls3d_frame_run: ls3d_frame_set_params has not been called"); return -1; }
	if (n_run <= 0) { first_map = 0; n_run = f->S; }
	if (first_map < 0 || first_map + n_run > f->S) { set_error("ls3d_frame_run: sensor range [%d,%d) outside 0..%d", first_map, first_map + n_run, f->S); return -1; }
	const int s_first = first_map, s_end = first_map + n_run;
	const long long n_max = f->h_sd[s_end].pix_begin - f->h_sd[s_first].pix_begin;
	const PeerDst none = PeerDst{0, {}};
	int launched = 0;

	if (stages & kStageCount) {
		if (!d_depth || !d_colors) { set_error("ls3d_frame_run: null argument"); return -1; }
		if (f->filter_mode == kModeOrganized && f->filter_on && !f->organized_ok) {
			set_error("ls3d_frame_run: organized filter mode requested but a sensor's pose/intrinsics do not admit the pixel-window bound");
			return -1;
		}
		const bool organized = f->filter_on && f->organized_ok && f->filter_mode != kModeVoxelHash && !(f->filter_mode == kModeAuto && f->organized_wide);
		f->last_organized = organized;
		f->last_depth = d_depth;
		f->last_colors = d_colors;
		f->ev_recorded = 0;
		stage_begin(f, kTsWhole, st);
		const bool chain0 = organized && !f->timing && pdl_enabled() && f->zero_bytes % 16 == 0;
		if (chain0) {
			const int n16 = (int)(f->zero_bytes / 16);
			k_zero_control<<<(n16 + 255) / 256, 256, 0, st>>>(f->zero.as<uint4>(), n16);
			count_launch(1);
		} else if (!cuda_ok(cudaMemsetAsync(f->zero.p, 0, f->zero_bytes, st), "clear control block")) return -1;
		// (Measured and dropped: merging the first sensors on a second stream beside the count of the last ones — 67.7-69.6 us per
		// frame against 60.5 serial; two more launches and two cross-stream waits cost more than the overlap returns.)
		if (organized) {
			stage_begin(f, kTsOrganized, st);
			if (launch_organized_count(f, d_depth, s_first, s_end, st, chain0) < 0) return -1;
			stage_end(f, kTsOrganized, st);
			launched += 1;
		} else if (f->filter_on) {
			if (launch_map(f, d_depth, d_colors, s_first, s_end, f->cloud0.as<uint4>(), nullptr, nullptr, none, st) < 0) return -1;
			const int r = frame_filter_stages(f, s_first, s_end, n_max, nullptr, nullptr, none, st, kStageCount);
			if (r < 0) return -1;
			launched += 1 + r;
		} else if (!(stages & kStageMerge)) {
			// no filter, split run: keep the culled cloud local and mark everything as kept
			if (launch_map(f, d_depth, d_colors, s_first, s_end, f->cloud0.as<uint4>(), nullptr, nullptr, none, st) < 0) return -1;
			k_keep_all<<<(int)std::min<long long>((n_max + 255) / 256, (long long)f->sm_count * 8), 256, 0, st>>>(f->keep.as<uint8_t>(), f->ctl);
			count_launch(1);
			if (!cuda_ok(cudaGetLastError(), "k_keep_all")) return -1;
			launched += 2;
		}
	}
	if (stages & kStageMerge) {
		if (!dst && peers.n == 0) { set_error("ls3d_frame_run: no destination buffer"); return -1; }
		int r;
		if (f->filter_on && f->last_organized) {
			if (!f->last_depth) { set_error("ls3d_frame_merge: no count stage has run"); return -1; }
			r = launch_map(f, f->last_depth, f->last_colors, s_first, s_end, dst, d_dst_offset, f->keep_px.as<uint8_t>(), peers, st, 0, -1,
				((stages & kStageCount) || peers.n > 0) && !f->timing && !f->colors_ready);      // straight behind the count kernel (or the count exchange) in this stream
		} else if (f->filter_on || !(stages & kStageCount)) {
			r = frame_merge_stage(f, s_first, s_end, n_max, dst, d_dst_offset, peers, st);
		} else {
			r = launch_map(f, d_depth, d_colors, s_first, s_end, dst, d_dst_offset, nullptr, peers, st);     // no filter: K1 writes the result directly
			if (r >= 0 && f->want_triangles) {
				// K1t: the pixel->vertex map K1 just wrote + the raw depth -> index triples (status_b / tile_counter_b are free: no filter ran)
				if (!f->tri.reserve(sizeof(int) * 6 * (size_t)f->total_px + 64, "alloc triangles")) return -1;
				const int ntiles = f->h_sd[s_end].tile_begin - f->h_sd[s_first].tile_begin;
				stage_begin(f, kTsTriangles, st);
				int mw = 0;
				for (int i = s_first; i < s_end; i++) mw = std::max(mw, f->w[i]);
				const size_t tri_smem = sizeof(unsigned short) * ((size_t)kTile + 3 * (size_t)mw + 4);
				if (tri_smem > 200 * 1024) { set_error("triangle stage: image width %d too large for the staged depth rows", mw); return -1; }
				if (tri_smem > 48 * 1024 && !cuda_ok(cudaFuncSetAttribute(k_triangles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_smem), "triangle stage shared memory")) return -1;
				launch_chain(!f->timing, k_triangles, dim3((unsigned)(std::max(1, std::min(ntiles, f->sm_count * 8)))), kScanThreads, tri_smem, st, (const uint8_t *)d_depth, f->sd.as<SensorDesc>(), f->tile_sensor.as<unsigned short>(),
					f->d2v.as<int>(), s_first, s_end, f->ctl, f->status_b, f->tri_starts, f->tri_override ? f->tri_override : f->tri.as<int>());
				stage_end(f, kTsTriangles, st);
				count_launch(1);
				if (!cuda_ok(cudaGetLastError(), "k_triangles")) return -1;
				r += 1;
			}
		}
		if (r < 0) return -1;
		launched += r;
		stage_end(f, kTsWhole, st);
	}
	return launched;
}

extern "C" int ls3d_frame_run(Ls3dFrame *f, const void *d_depth_maps, const void *d_depth_colors, int first_map, int n_run, void *stream) {
	PeerDst none; none.n = 0;
	if (!f) { set_error("ls3d_frame_run: null frame"); return -1; }
	return frame_run_impl(f, d_depth_maps, d_depth_colors, first_map, n_run, f->final_.as<uint4>(), nullptr, none, (cudaStream_t)stream);
}

extern "C" int ls3d_frame_run_to(Ls3dFrame *f, const void *d_depth_maps, const void *d_depth_colors, int first_map, int n_run,
	void *d_dst_vertices, const int *d_dst_offset, void *stream) {
	PeerDst none; none.n = 0;
	if (!f || !d_dst_vertices) { set_error("ls3d_frame_run_to: null argument"); return -1; }
	return frame_run_impl(f, d_depth_maps, d_depth_colors, first_map, n_run, (uint4 *)d_dst_vertices, d_dst_offset, none, (cudaStream_t)stream);
}

extern "C" int ls3d_frame_run_count(Ls3dFrame *f, const void *d_depth_maps, const void *d_depth_colors, int first_map, int n_run, void *stream) {
	PeerDst none; none.n = 0;
	return frame_run_impl(f, d_depth_maps, d_depth_colors, first_map, n_run, nullptr, nullptr, none, (cudaStream_t)stream, kStageCount);
}

extern "C" int ls3d_frame_merge_peers(Ls3dFrame *f, int first_map, int n_run, int n_peers, void *const *peer_dst_vertices, const int *d_dst_offset, void *stream) {
	if (!f || !peer_dst_vertices || n_peers <= 0 || n_peers > kMaxPeers) { set_error("ls3d_frame_merge_peers: need 1..%d destination buffers", kMaxPeers); return -1; }
	PeerDst peers;
	peers.n = n_peers;
	for (int i = 0; i < kMaxPeers; i++) peers.ptr[i] = i < n_peers ? (uint4 *)peer_dst_vertices[i] : nullptr;
	return frame_run_impl(f, nullptr, nullptr, first_map, n_run, (uint4 *)peer_dst_vertices[0], d_dst_offset, peers, (cudaStream_t)stream, kStageMerge);
}

extern "C" int ls3d_frame_run_peers(Ls3dFrame *f, const void *d_depth_maps, const void *d_depth_colors, int first_map, int n_run,
	int n_peers, void *const *peer_dst_vertices, const int *d_dst_offset, void *stream) {
	if (!f || !peer_dst_vertices || n_peers <= 0 || n_peers > kMaxPeers) { set_error("ls3d_frame_run_peers: need 1..%d destination buffers", kMaxPeers); return -1; }
	PeerDst peers;
	peers.n = n_peers;
	for (int i = 0; i < kMaxPeers; i++) peers.ptr[i] = i < n_peers ? (uint4 *)peer_dst_vertices[i] : nullptr;
	return frame_run_impl(f, d_depth_maps, d_depth_colors, first_map, n_run, (uint4 *)peer_dst_vertices[0], d_dst_offset, peers, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------
// multi-GPU merge without a collective library: survivor counts and completion flags travel as peer stores
// ------------------------------------------------------------------------------------------------------
// Every rank owns one Ls3dFrameSync block (include/ls3d.h; 256 bytes of device memory mapped by all ranks).  Per frame:
//   k_frame_publish_counts  (1 block, after the neighbour count): store this rank's survivor count into EVERY rank's block, tag it
//                           with the frame number, wait until all ranks' counts of this frame have arrived here, and leave the
//                           exclusive prefix (where this rank's records start in the merged cloud) in sync->offset — which is
//                           what the compaction kernel reads as its base (ls3d_frame_merge_peers' d_dst_offset);
//   k_frame_wait_peers      (1 block, after the compaction): tell every rank that this rank's records have been delivered and
//                           wait for the same word from all of them: after it, the merged cloud in this rank's buffer is complete.
// Spins are bounded (kErrScanSpin) so a broken peer cannot hang the GPU.
struct FrameSync {
	unsigned epoch;            // frames completed so far
	int offset;                // exclusive prefix of the survivor counts for this rank
	int total;                 // sum of all ranks' counts
	int err;
	unsigned cnt_tag[kMaxPeers];   // cnt_tag[r] == frame number: cnt_val[r] is rank r's survivor count of that frame
	int cnt_val[kMaxPeers];
	unsigned done_tag[kMaxPeers];  // done_tag[r] == frame number: rank r's records of that frame are in this rank's buffer
	unsigned pad[36];
};
static_assert(sizeof(FrameSync) == 256, "Ls3dFrameSync layout");
struct SyncPeers { int world, rank; FrameSync *p[kMaxPeers]; };

__global__ void k_frame_publish_counts(const FrameCtl *ctl, SyncPeers sp) {
	pdl_enter();
	if (threadIdx.x != 0) return;
	FrameSync *me = sp.p[sp.rank];
	const unsigned e = ld_volatile_u32(&me->epoch) + 1u;
	const int n = ctl ? ctl->n_kept : 0;
	for (int r = 0; r < sp.world; r++) *reinterpret_cast<volatile int *>(&sp.p[r]->cnt_val[sp.rank]) = n;
	__threadfence_system();
	for (int r = 0; r < sp.world; r++) st_volatile_u32(&sp.p[r]->cnt_tag[sp.rank], e);
	int off = 0, tot = 0;
	for (int r = 0; r < sp.world; r++) {
		unsigned spins = 0;
		while (ld_volatile_u32(&me->cnt_tag[r]) != e) if (++spins > (1u << 25)) { me->err |= kErrScanSpin; break; }
		__threadfence_system();
		const int v = *reinterpret_cast<volatile int *>(&me->cnt_val[r]);
		if (r < sp.rank) off += v;
		tot += v;
	}
	me->offset = off;
	me->total = tot;
}

__global__ void k_frame_wait_peers(SyncPeers sp) {
	pdl_enter();
	if (threadIdx.x != 0) return;
	FrameSync *me = sp.p[sp.rank];
	const unsigned e = ld_volatile_u32(&me->epoch) + 1u;
	__threadfence_system();                                           // the compaction kernel before us has completed: its peer stores are performed
	for (int r = 0; r < sp.world; r++) st_volatile_u32(&sp.p[r]->done_tag[sp.rank], e);
	for (int r = 0; r < sp.world; r++) {
		unsigned spins = 0;
		while (ld_volatile_u32(&me->done_tag[r]) != e) if (++spins > (1u << 25)) { me->err |= kErrScanSpin; break; }
	}
	__threadfence_system();
	me->epoch = e;
}

static bool sync_peers(SyncPeers &sp, int rank, int world, void *const *peer_sync, const char *who) {
	if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !peer_sync) { set_error("%s: world must be 1..%d, rank inside it, pointer table non-null", who, kMaxPeers); return false; }
	sp.world = world; sp.rank = rank;
	for (int r = 0; r < kMaxPeers; r++) sp.p[r] = r < world ? (FrameSync *)peer_sync[r] : nullptr;
	for (int r = 0; r < world; r++) if (!sp.p[r]) { set_error("%s: missing sync block of rank %d", who, r); return false; }
	return true;
}

extern "C" int ls3d_frame_publish_counts(Ls3dFrame *f, int rank, int world, void *const *peer_sync, void *stream) {
	SyncPeers sp;
	if (!sync_peers(sp, rank, world, peer_sync, "ls3d_frame_publish_counts")) return -1;
	launch_chain(true, k_frame_publish_counts, dim3(1), 32, 0, (cudaStream_t)stream, f ? f->ctl : (FrameCtl *)nullptr, sp);      // f == NULL: a rank without sensors publishes 0
	count_launch(1);
	return cuda_ok(cudaGetLastError(), "k_frame_publish_counts") ? 0 : -1;
}

extern "C" int ls3d_frame_wait_peers(int rank, int world, void *const *peer_sync, void *stream) {
	SyncPeers sp;
	if (!sync_peers(sp, rank, world, peer_sync, "ls3d_frame_wait_peers")) return -1;
	launch_chain(true, k_frame_wait_peers, dim3(1), 32, 0, (cudaStream_t)stream, sp);
	count_launch(1);
	return cuda_ok(cudaGetLastError(), "k_frame_wait_peers") ? 0 : -1;
}

extern "C" void ls3d_frame_enable_timing(Ls3dFrame *f, int on) { if (f) f->timing = on != 0; }

// Stage durations of the last ls3d_frame_run (milliseconds; waits for it to finish); stages that did not run are 0:
//   [0] map/cull/compact (K1)  [1] hash clear  [2] voxel insert  [3] cell ranges + scatter  [4] voxel-hash neighbour count
//   [5] survivor compaction    [6] organized neighbour count (K1o)   [7] whole run incl. the control-block clear
extern "C" int ls3d_frame_stage_ms(Ls3dFrame *f, float *out) {
	if (!f || !out) { set_error("ls3d_frame_stage_ms: null argument"); return -1; }
	for (int i = 0; i < kTsN; i++) out[i] = 0.0f;
	if (!(f->ev_recorded & (1 << kTsWhole))) { set_error("ls3d_frame_stage_ms: no timed run (call ls3d_frame_enable_timing first)"); return -1; }
	if (!cuda_ok(cudaEventSynchronize(f->ev[kTsWhole][1]), "stage timing")) return -1;
	for (int i = 0; i < kTsN; i++)
		if (f->ev_recorded & (1 << i)) cudaEventElapsedTime(&out[i], f->ev[i][0], f->ev[i][1]);
	return 0;
}

extern "C" const void *ls3d_frame_vertices(Ls3dFrame *f) { return f ? f->final_.p : nullptr; }
extern "C" const void *ls3d_frame_culled_vertices(Ls3dFrame *f) {
	if (!f) return nullptr;
	if (!f->filter_on) return f->final_.p;
	return f->last_organized ? nullptr : f->cloud0.p;      // the organized path never materialises the unfiltered cloud
}
extern "C" const int *ls3d_frame_count_ptr(Ls3dFrame *f) { return f ? &f->ctl->n_final : nullptr; }
extern "C" const int *ls3d_frame_sensor_starts(Ls3dFrame *f) { return f ? ((f->filter_on && !f->last_organized) ? f->final_starts : f->culled_starts) : nullptr; }
extern "C" const int *ls3d_frame_culled_starts(Ls3dFrame *f) { return f ? f->culled_starts : nullptr; }
extern "C" const int *ls3d_frame_old_to_new(Ls3dFrame *f) { return (f && !(f->filter_on && f->last_organized)) ? f->map.as<int>() : nullptr; }
extern "C" const unsigned char *ls3d_frame_keep_mask(Ls3dFrame *f) { return (f && f->filter_on && f->last_organized) ? f->keep_px.as<unsigned char>() : nullptr; }
extern "C" void ls3d_frame_enable_triangles(Ls3dFrame *f, int on) { if (f) f->want_triangles = on != 0; }
extern "C" const int *ls3d_frame_triangles(Ls3dFrame *f) {
	if (!f) return nullptr;
	if (!f->tri.reserve(sizeof(int) * 6 * (size_t)f->total_px + 64, "alloc triangles")) return nullptr;
	return f->tri.as<int>();
}
extern "C" const int *ls3d_frame_triangle_starts(Ls3dFrame *f) { return f ? f->tri_starts : nullptr; }
extern "C" const int *ls3d_frame_depth_to_vertex(Ls3dFrame *f) {
	if (!f) return nullptr;
	f->want_d2v = true;      // produced from the next run on
	return f->d2v.as<int>();
}

// ======================================================================================================
// Host-buffer (drop-in) entry points
// ======================================================================================================
static Ls3dFrame *g_frame = nullptr;     // cached context for the host-buffer API, keyed by the size list
static int g_default_filter_mode = kModeAuto;

extern "C" int ls3d_set_default_filter_mode(int mode) {
	if (mode < kModeAuto || mode > kModeOrganized) { set_error("ls3d_set_default_filter_mode: mode must be 0, 1 or 2"); return -1; }
	g_default_filter_mode = mode;
	return 0;
}

static Ls3dFrame *cached_frame(int n_maps, const int *widths, const int *heights) {
	// a context built for a longer size list serves any prefix of it (generateVerticesFromDepthMap is called once
	// per sensor with the same arrays, KinectServer.cs:527-554)
	if (g_frame && g_frame->S >= n_maps && !memcmp(g_frame->w.data(), widths, sizeof(int) * n_maps) && !memcmp(g_frame->h.data(), heights, sizeof(int) * n_maps))
		return g_frame;
	if (g_frame) { frame_free(g_frame); g_frame = nullptr; }
	g_frame = ls3d_frame_create(n_maps, widths, heights);
	if (g_frame) {
		if (!g_frame->in_depth.reserve(g_frame->depth_bytes, "alloc depth input") || !g_frame->in_colors.reserve(g_frame->color_bytes, "alloc colour input")) {
			frame_free(g_frame);
			g_frame = nullptr;
		}
	}
	return g_frame;
}

static void mesh_reset(Mesh *m) {
	m->nVertices = 0;
	m->vertices = nullptr;
	m->nTriangles = 0;
	m->triangles = (int *)malloc(sizeof(int));    // valid empty allocation, like the reference's new int[0]
}

static bool frame_will_use_organized(const Ls3dFrame *f) {
	return f->filter_on && f->organized_ok && f->filter_mode != kModeVoxelHash && !(f->filter_mode == kModeAuto && f->organized_wide);
}

// Shared body: upload sensors [first, first+n_run), run, read back.  Returns total vertices or -1.
// Overlap: depth goes up on the library stream and the neighbour count starts as soon as it has landed, while the colours (60 % of
// the input bytes, needed only by the final map/merge kernel) go up on a second stream; on the organized path the survivor count
// is known before that kernel runs, so the host sizes the read-back while it executes and the call ends with a single wait.
static int host_frame(int n_maps, unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	float *intr_params, float *wtransform_params, Mesh *out_mesh, const float bounds[6], int first, int n_run,
	int filter_k, float filter_maxDist, int *per_map_counts, bool with_triangles = false)
{
	clear_error();
	if (out_mesh) mesh_reset(out_mesh);
	if (!out_mesh || !depth_maps || !depth_colors || !widths || !heights || !intr_params || !wtransform_params) { set_error("null argument"); return -1; }
	if (n_maps <= 0 || first < 0 || n_run <= 0 || first + n_run > n_maps) { set_error("sensor range [%d,%d) outside 0..%d", first, first + n_run, n_maps); return -1; }
	if (!ensure_device()) return -1;
	std::lock_guard<std::mutex> lk(api_mutex());
	cudaStream_t st = api_stream();
	if (!st) return -1;
	Ls3dFrame *f = cached_frame(n_maps, widths, heights);
	if (!f) return -1;
	const SensorDesc &a = f->h_sd[first], &z = f->h_sd[first + n_run];
	if (!f->st_colors) {
		if (!cuda_ok(cudaStreamCreateWithFlags(&f->st_colors, cudaStreamNonBlocking), "create colour stream") ||
			!cuda_ok(cudaStreamCreateWithFlags(&f->st_out, cudaStreamNonBlocking), "create read-back stream") ||
			!cuda_ok(cudaEventCreateWithFlags(&f->ev_colors, cudaEventDisableTiming), "create event") ||
			!cuda_ok(cudaEventCreateWithFlags(&f->ev_count, cudaEventDisableTiming), "create event")) return -1;
	}
	struct ColorsGuard { Ls3dFrame *f; ~ColorsGuard() { f->colors_ready = nullptr; if (f->st_colors) cudaStreamSynchronize(f->st_colors); if (f->st_out) cudaStreamSynchronize(f->st_out); if (f->st_merge) cudaStreamSynchronize(f->st_merge); } } guard{f};
	f->filter_mode = g_default_filter_mode;
	f->want_triangles = with_triangles && !(filter_k > 0 && filter_maxDist > 0);
	if (frame_set_params(f, n_maps, intr_params, wtransform_params, bounds[0], bounds[1], bounds[2], bounds[3], bounds[4], bounds[5], filter_k, filter_maxDist, st) < 0) return -1;
	int *po = f->pin_out;
	const FrameCtl *hc = reinterpret_cast<const FrameCtl *>(po);
	uint8_t *const dd = f->in_depth.as<uint8_t>(), *const dc = f->in_colors.as<uint8_t>();
	cudaStream_t up = f->st_colors;
	// LS3D_E2E_MODE: 2 (default) pipelined, colours pulled by the merge kernel when the caller's buffer is page-locked; 1 pipelined,
	// colours uploaded; 0 legacy (upload everything, count -> sized copy)
	static const int env_mode = getenv("LS3D_E2E_MODE") ? atoi(getenv("LS3D_E2E_MODE")) : 2;
	static const int env_graph = getenv("LS3D_E2E_GRAPH") ? atoi(getenv("LS3D_E2E_GRAPH")) : 1;
	static const int env_chunks = getenv("LS3D_E2E_CHUNKS") ? atoi(getenv("LS3D_E2E_CHUNKS")) : 4;
	static const int env_cblocks = getenv("LS3D_E2E_COPY_BLOCKS") ? atoi(getenv("LS3D_E2E_COPY_BLOCKS")) : 16;
	static const int env_trace = getenv("LS3D_E2E_TRACE") ? atoi(getenv("LS3D_E2E_TRACE")) : 0;    // 1: timed events in the schedule, timeline printed to stderr
	static const int env_direct = getenv("LS3D_E2E_DIRECT") ? atoi(getenv("LS3D_E2E_DIRECT")) : 1;  // 1: the merge kernel stores its tiles into the host block itself
	if (env_mode != 0 && frame_will_use_organized(f)) {
		// Pipelined over chunks of sensors on four streams (upload, count, merge, read-back): a chunk's neighbour count starts when
		// its depth has landed; its map/merge kernel places the survivors at the base the per-tile survivor counts give it; a copy
		// kernel then stores that record range into the page-locked output block (mapped host memory, 512-byte warp stores).  The
		// read-back of chunk c therefore overlaps the upload of chunk c+1 (PCIe is full duplex) and needs no survivor count on the
		// host.  When the caller's colour buffer is page-locked too it is never uploaded: the merge kernel reads the 24-byte colour
		// groups of surviving pixels straight out of it (a fraction of the image).  With page-locked inputs the whole schedule is one
		// CUDA graph, re-captured only when a pointer or a parameter changes: one launch and one wait per frame.
		const int Cn = std::max(1, std::min(std::min(env_chunks, n_run), kMaxChunks));
		const long long px_run = z.pix_begin - a.pix_begin;
		void *v = host_block_alloc((size_t)std::max<long long>(px_run, 1) * sizeof(VertexC4ubV3f));
		void *v_dev = nullptr;
		if (v && host_block_is_pinned(v) && cudaHostGetDevicePointer(&v_dev, v, 0) == cudaSuccess && v_dev) {
			auto locked = [](const void *p, void **dev) {
				cudaPointerAttributes pa;
				if (cudaPointerGetAttributes(&pa, p) != cudaSuccess) { cudaGetLastError(); return false; }
				if (pa.type != cudaMemoryTypeHost) return false;
				if (dev && (cudaHostGetDevicePointer(dev, const_cast<void *>(p), 0) != cudaSuccess || !*dev)) { cudaGetLastError(); return false; }
				return true;
			};
			void *col_dev = nullptr;
			bool ok = true;
			// pageable caller buffers (a P/Invoke caller's GC-pinned byte[], KinectServer.cs:354-374): cudaMemcpyAsync from them would be
			// staged by the driver on this thread, chunk after chunk, and nothing would overlap.  Instead the host copy pool moves the
			// depth into page-locked staging, the schedule below starts from there (same pointers every frame: the captured graphs are
			// reused) and is cut in two launches: upload + neighbour count first; the colours are copied while that runs; then merge +
			// read-back.  (No kernel ever waits for the host: a profiler that serialises launches would deadlock it.)
			static const int env_stage = getenv("LS3D_E2E_STAGE") ? atoi(getenv("LS3D_E2E_STAGE")) : 1;
			const bool staged = env_stage != 0 && env_mode == 2 && !locked(depth_maps, nullptr);
			if (staged) {
				if (!f->stage_depth) {
					ok = cuda_ok(cudaHostAlloc((void **)&f->stage_depth, std::max<size_t>(f->depth_bytes, 16), cudaHostAllocDefault), "alloc depth staging") &&
						cuda_ok(cudaHostAlloc((void **)&f->stage_colors, std::max<size_t>(f->color_bytes, 16), cudaHostAllocMapped), "alloc colour staging");
					if (!ok) { if (f->stage_depth) cudaFreeHost(f->stage_depth); if (f->stage_colors) cudaFreeHost(f->stage_colors); f->stage_depth = f->stage_colors = nullptr; }
				}
				if (!ok) { host_block_free(v); return -1; }
				parallel_memcpy(f->stage_depth + a.depth_off, depth_maps + a.depth_off, (size_t)(z.depth_off - a.depth_off));
			}
			const unsigned char *src_depth = staged ? f->stage_depth : depth_maps, *src_colors = staged ? f->stage_colors : depth_colors;
			const bool pull = env_mode == 2 && locked(src_colors, &col_dev);
			const bool graph_ok = env_graph != 0 && locked(src_depth, nullptr) && (pull || locked(src_colors, nullptr));
			const uint8_t *col_src = pull ? (const uint8_t *)col_dev : dc;
			for (int i = 0; i < kEvN && ok; i++)
				if (!f->ev_up[i]) ok = cuda_ok(cudaEventCreateWithFlags(&f->ev_up[i], cudaEventDisableTiming), "create event");
			for (int i = 0; i < 4 * kMaxChunks + 1 && ok && env_trace; i++)
				if (!f->ev_tr[i]) ok = cuda_ok(cudaEventCreate(&f->ev_tr[i]), "create trace event");
			auto trace = [&](int kind, int c, cudaStream_t s_) { return !env_trace || cuda_ok(cudaEventRecordWithFlags(f->ev_tr[kind < 0 ? 4 * kMaxChunks : kind * kMaxChunks + c], s_, cudaEventRecordExternal), "trace"); };
			if (!f->st_merge) ok = ok && cuda_ok(cudaStreamCreateWithFlags(&f->st_merge, cudaStreamNonBlocking), "create merge stream");
			cudaEvent_t *ev_d = f->ev_up, *ev_c = ev_d + kMaxChunks, *ev_n = ev_c + kMaxChunks, *ev_m = ev_n + kMaxChunks, *ev_x = ev_m + 2 * kMaxChunks;   // ev_x: fork + 3 joins
			cudaStream_t sm = f->st_merge, so = f->st_out;
			// chunk c = sensors [bound(c), bound(c+1))   (measured and dropped: a smaller last chunk, 3 / 5 / 6 / 8 chunks, depth halves on two
			// copy engines, the colour fetch as a kernel of its own — all within 3 % or slower)
			auto bound = [&](int c) { return first + (int)((long long)n_run * c / Cn); };
			PeerDst none; none.n = 0;
			// the merge kernel stages a tile's records in shared memory and stores them into the page-locked output block by whole warps
			// (the multi-GPU merge's peer-store path with the host block as the only "peer"): no device copy of the result, no copy kernel
			PeerDst host_dst; host_dst.n = 1; host_dst.ptr[0] = (uint4 *)v_dev;
			// part 0: the whole schedule; part 1: fork, uploads and neighbour counts only; part 2: merges, copy-outs and the count read-back
			// only (staged mode launches 1, copies the colours, then launches 2: stream order on `st` is the dependency between them)
			auto enqueue = [&](int part) -> bool {
				bool k = trace(-1, 0, st) && cuda_ok(cudaEventRecord(ev_x[0], st), "fork") &&
					(part == 2 || cuda_ok(cudaStreamWaitEvent(up, ev_x[0], 0), "fork")) &&
					(part == 1 || (cuda_ok(cudaStreamWaitEvent(sm, ev_x[0], 0), "fork") && cuda_ok(cudaStreamWaitEvent(so, ev_x[0], 0), "fork"))) &&
					(part == 2 || cuda_ok(cudaMemsetAsync(f->zero.p, 0, f->zero_bytes, st), "clear control block"));
				for (int c = 0; c < Cn && k; c++) {
					const SensorDesc &ca = f->h_sd[bound(c)], &cz = f->h_sd[bound(c + 1)];
					const int tlo = ca.tile_begin - a.tile_begin, thi = cz.tile_begin - a.tile_begin;
					if (part != 2) {
						k = cuda_ok(cudaMemcpyAsync(dd + ca.depth_off, src_depth + ca.depth_off, (size_t)(cz.depth_off - ca.depth_off), cudaMemcpyHostToDevice, up), "upload depth") &&
							cuda_ok(cudaEventRecord(ev_d[c], up), "record depth upload") && trace(0, c, up);
						if (k && !pull)
							k = cuda_ok(cudaMemcpyAsync(dc + ca.color_off, src_colors + ca.color_off, (size_t)(cz.color_off - ca.color_off), cudaMemcpyHostToDevice, up), "upload colours") &&
								cuda_ok(cudaEventRecord(ev_c[c], up), "record colour upload");
						k = k && cuda_ok(cudaStreamWaitEvent(st, ev_d[c], 0), "wait for the depth upload") && launch_organized_count(f, dd, bound(c), bound(c + 1), st) == 0 && trace(1, c, st);
					}
					if (part == 0)
						k = k && cuda_ok(cudaEventRecord(ev_n[c], st), "record count") && cuda_ok(cudaStreamWaitEvent(sm, ev_n[c], 0), "wait for the count") &&
							(pull || cuda_ok(cudaStreamWaitEvent(sm, ev_c[c], 0), "wait for the colour upload"));
					if (part != 1) {
						k = k && launch_map(f, dd, col_src, first, first + n_run, f->final_.as<uint4>(), nullptr, f->keep_px.as<uint8_t>(), env_direct ? host_dst : none, sm, tlo, thi) >= 0 &&
							trace(2, c, sm) && cuda_ok(cudaEventRecord(ev_m[c], sm), "record merge") && cuda_ok(cudaStreamWaitEvent(so, ev_m[c], 0), "wait for the merge");
						if (!k) break;
						if (!env_direct) {
							k_copy_out<<<std::max(1, env_cblocks), 256, 0, so>>>(f->final_.as<uint4>(), (uint4 *)v_dev, reinterpret_cast<const unsigned *>(f->status_b), a.tile_begin, tlo, thi);
							count_launch(1);
							k = cuda_ok(cudaGetLastError(), "k_copy_out");
						}
						k = k && trace(3, c, so);
					}
				}
				if (part != 2) k = k && cuda_ok(cudaEventRecord(ev_x[1], up), "join") && cuda_ok(cudaStreamWaitEvent(st, ev_x[1], 0), "join");
				if (part != 1)
					k = k && cuda_ok(cudaEventRecord(ev_x[2], sm), "join") && cuda_ok(cudaEventRecord(ev_x[3], so), "join") &&
						cuda_ok(cudaStreamWaitEvent(st, ev_x[2], 0), "join") && cuda_ok(cudaStreamWaitEvent(st, ev_x[3], 0), "join") &&
						cuda_ok(cudaMemcpyAsync(po, f->ctl, sizeof(FrameCtl), cudaMemcpyDeviceToHost, st), "read counts") &&
						cuda_ok(cudaMemcpyAsync(po + 16, f->culled_starts, sizeof(int) * (n_maps + 1), cudaMemcpyDeviceToHost, st), "read sensor starts");
				return k;
			};
			// one part of the schedule as a CUDA graph, re-captured only when a pointer or a parameter changes
			auto run_graph = [&](int part, cudaGraphExec_t &exec, HostGraphKey &stored, int &n_launches) -> bool {
				const HostGraphKey key{src_depth, src_colors, v, first, n_run, Cn, (pull ? 1 : 0) | (staged ? 2 : 0) | (part << 2), f->params_version};
				bool k = true;
				if (!exec || memcmp(&key, &stored, sizeof(key))) {
					cudaGraph_t g = nullptr;
					const long long l0 = g_launches.load();
					k = cuda_ok(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed), "begin capture");
					if (k) {
						const bool q = enqueue(part);
						char keep_err[512];
						snprintf(keep_err, sizeof(keep_err), "%s", ls3d_last_error());
						const cudaError_t e = cudaStreamEndCapture(st, &g);
						if (!q) { set_error("%s", keep_err); k = false; }
						else k = cuda_ok(e, "end capture") && g;
					}
					n_launches = (int)(g_launches.load() - l0);
					g_launches.fetch_sub(n_launches);        // nothing ran yet: launches are counted per graph launch below
					if (k && exec) {
						cudaGraphExecUpdateResultInfo info;
						if (cudaGraphExecUpdate(exec, g, &info) != cudaSuccess) { cudaGetLastError(); cudaGraphExecDestroy(exec); exec = nullptr; }
					}
					if (k && !exec) k = cuda_ok(cudaGraphInstantiate(&exec, g, 0), "instantiate the frame graph");
					if (g) cudaGraphDestroy(g);
					if (k) stored = key; else if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; }
				}
				k = k && cuda_ok(cudaGraphLaunch(exec, st), "launch the frame graph");
				if (k) count_launch(n_launches);
				return k;
			};
			f->last_organized = true;
			f->last_depth = dd;
			f->last_colors = dc;
			f->ev_recorded = 0;
			const bool use_graph = graph_ok && !f->timing;
			if (ok && !staged) {
				ok = use_graph ? run_graph(0, f->hg_exec, f->hg_key, f->hg_launches) : enqueue(0);
			} else if (ok) {
				// the device starts on the depth while the host copies the colours; merges and read-back follow in stream order
				ok = use_graph ? run_graph(1, f->hg_exec, f->hg_key, f->hg_launches) : enqueue(1);
				if (ok) parallel_memcpy(f->stage_colors + a.color_off, depth_colors + a.color_off, (size_t)(z.color_off - a.color_off));
				ok = ok && (use_graph ? run_graph(2, f->hg_exec_b, f->hg_key_b, f->hg_launches_b) : enqueue(2));
			}
			ok = cuda_ok(cudaStreamSynchronize(st), "frame pipeline") && ok;
			if (ok && env_trace) {
				static int trace_calls = 0;
				if (++trace_calls == env_trace + 8) {
					auto t = [&](int kind, int c) { float ms = -1; cudaEventElapsedTime(&ms, f->ev_tr[4 * kMaxChunks], f->ev_tr[kind * kMaxChunks + c]); return ms * 1000.0f; };
					fprintf(stderr, "[ls3d trace] chunk: depth landed / counted / merged / copied out (us after the fork)\n");
					for (int c = 0; c < Cn; c++) fprintf(stderr, "[ls3d trace] %2d: %7.1f %7.1f %7.1f %7.1f\n", c, t(0, c), t(1, c), t(2, c), t(3, c));
				}
			}
			if (ok && hc->err) { set_error("device reported error flags 0x%x in the frame pipeline", hc->err); ok = false; }
			if (ok && hc->n_final != hc->n_kept) { set_error("internal: merged %d vertices but the neighbour count announced %d", hc->n_final, hc->n_kept); ok = false; }
			if (!ok) { host_block_free(v); return -1; }
			if (per_map_counts) {
				for (int i = 0; i < n_maps; i++) per_map_counts[i] = 0;
				for (int i = first; i < first + n_run; i++) per_map_counts[i] = po[16 + i + 1] - po[16 + i];
			}
			out_mesh->vertices = (VertexC4ubV3f *)v;
			out_mesh->nVertices = hc->n_final;
			return hc->n_final;
		}
		if (v) host_block_free(v);
	}
	if (env_mode != 0 && f->want_triangles && !f->filter_on) {
		// The unfiltered mesh (what the live view asks for every frame) through the same kind of schedule: sensors in chunks; a chunk is a
		// complete run of K1 + the triangle kernel on its sensors (vertices placed behind the previous chunks' through the device-side
		// running total, triangles chunk-relative in the chunk's own region of the scratch buffer); k_mesh_chunk_done records the chunk's
		// ranges; k_copy_mesh_out stores them into the two page-locked Mesh blocks, rebasing the indices, while the next chunk uploads
		// and runs.  The 28 MB read-back, which used to start after everything else, overlaps the upload.
		const int Cn = std::max(1, std::min(std::min(env_chunks, n_run), kMaxMeshChunks));
		const long long px_run = z.pix_begin - a.pix_begin;
		const size_t acc_ints = 4 + 4 * (size_t)kMaxMeshChunks + (size_t)f->S + 4;
		void *v = host_block_alloc((size_t)std::max<long long>(px_run, 1) * sizeof(VertexC4ubV3f));
		void *t = host_block_alloc((size_t)std::max<long long>(px_run, 1) * 6 * sizeof(int));
		void *v_dev = nullptr, *t_dev = nullptr;
		bool ok = v && t && host_block_is_pinned(v) && host_block_is_pinned(t) && cudaHostGetDevicePointer(&v_dev, v, 0) == cudaSuccess && cudaHostGetDevicePointer(&t_dev, t, 0) == cudaSuccess &&
			v_dev && t_dev && f->tri.reserve(sizeof(int) * 6 * (size_t)f->total_px + 64, "alloc triangles") && f->acc.reserve(sizeof(int) * acc_ints, "alloc mesh totals") &&
			(f->pin_acc || cudaHostAlloc((void **)&f->pin_acc, sizeof(int) * acc_ints, cudaHostAllocDefault) == cudaSuccess);
		if (ok) {
			auto locked = [](const void *p) {
				cudaPointerAttributes pa;
				if (cudaPointerGetAttributes(&pa, p) != cudaSuccess) { cudaGetLastError(); return false; }
				return pa.type == cudaMemoryTypeHost;
			};
			const bool graph_ok = env_graph != 0 && locked(depth_maps) && locked(depth_colors) && !f->timing;
			for (int i = 0; i < kEvN && ok; i++)
				if (!f->ev_up[i]) ok = cuda_ok(cudaEventCreateWithFlags(&f->ev_up[i], cudaEventDisableTiming), "create event");
			cudaEvent_t *ev_d = f->ev_up, *ev_n = ev_d + 2 * kMaxChunks, *ev_x = ev_d + 5 * kMaxChunks;
			cudaStream_t so = f->st_out;
			int *acc = f->acc.as<int>();
			auto bound = [&](int c) { return first + (int)((long long)n_run * c / Cn); };
			PeerDst none; none.n = 0;
			auto enqueue = [&]() -> bool {
				bool k = cuda_ok(cudaEventRecord(ev_x[0], st), "fork") && cuda_ok(cudaStreamWaitEvent(up, ev_x[0], 0), "fork") && cuda_ok(cudaStreamWaitEvent(so, ev_x[0], 0), "fork") &&
					cuda_ok(cudaMemsetAsync(acc, 0, sizeof(int) * acc_ints, st), "clear mesh totals");
				for (int c = 0; c < Cn && k; c++) {
					const int sa = bound(c), sb = bound(c + 1);
					const SensorDesc &ca = f->h_sd[sa], &cz = f->h_sd[sb];
					k = cuda_ok(cudaMemcpyAsync(dd + ca.depth_off, depth_maps + ca.depth_off, (size_t)(cz.depth_off - ca.depth_off), cudaMemcpyHostToDevice, up), "upload depth") &&
						cuda_ok(cudaMemcpyAsync(dc + ca.color_off, depth_colors + ca.color_off, (size_t)(cz.color_off - ca.color_off), cudaMemcpyHostToDevice, up), "upload colours") &&
						cuda_ok(cudaEventRecord(ev_d[c], up), "record upload") && cuda_ok(cudaStreamWaitEvent(st, ev_d[c], 0), "wait for the upload");
					if (!k) break;
					f->tri_override = f->tri.as<int>() + 6 * (ca.pix_begin - a.pix_begin);
					const int r = frame_run_impl(f, dd, dc, sa, sb - sa, f->final_.as<uint4>(), acc, none, st);
					f->tri_override = nullptr;
					if (r < 0) return false;
					k_mesh_chunk_done<<<1, 64, 0, st>>>(f->ctl, f->culled_starts, sa, sb, c, acc);
					k = cuda_ok(cudaGetLastError(), "k_mesh_chunk_done") && cuda_ok(cudaEventRecord(ev_n[c], st), "record chunk") && cuda_ok(cudaStreamWaitEvent(so, ev_n[c], 0), "wait for the chunk");
					if (!k) break;
					k_copy_mesh_out<<<std::max(1, 2 * env_cblocks), 256, 0, so>>>(f->final_.as<uint4>(), (uint4 *)v_dev, f->tri.as<int>() + 6 * (ca.pix_begin - a.pix_begin), (int *)t_dev, acc + 4 + 4 * c);
					count_launch(2);
					k = cuda_ok(cudaGetLastError(), "k_copy_mesh_out");
				}
				return k && cuda_ok(cudaEventRecord(ev_x[1], up), "join") && cuda_ok(cudaEventRecord(ev_x[3], so), "join") && cuda_ok(cudaStreamWaitEvent(st, ev_x[1], 0), "join") &&
					cuda_ok(cudaStreamWaitEvent(st, ev_x[3], 0), "join") && cuda_ok(cudaMemcpyAsync(f->pin_acc, acc, sizeof(int) * acc_ints, cudaMemcpyDeviceToHost, st), "read mesh totals");
			};
			if (ok && graph_ok) {
				const HostGraphKey key{depth_maps, depth_colors, v, first, n_run, Cn, 2, f->params_version};
				if (!f->hm_exec || memcmp(&key, &f->hm_key, sizeof(key)) || f->hm_tri_out != t) {
					cudaGraph_t g = nullptr;
					const long long l0 = g_launches.load();
					ok = cuda_ok(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed), "begin capture");
					if (ok) {
						const bool k = enqueue();
						char keep_err[512];
						snprintf(keep_err, sizeof(keep_err), "%s", ls3d_last_error());
						const cudaError_t e = cudaStreamEndCapture(st, &g);
						if (!k) { set_error("%s", keep_err); ok = false; }
						else ok = cuda_ok(e, "end capture") && g;
					}
					f->hm_launches = (int)(g_launches.load() - l0);
					g_launches.fetch_sub(f->hm_launches);
					if (ok && f->hm_exec) {
						cudaGraphExecUpdateResultInfo info;
						if (cudaGraphExecUpdate(f->hm_exec, g, &info) != cudaSuccess) { cudaGetLastError(); cudaGraphExecDestroy(f->hm_exec); f->hm_exec = nullptr; }
					}
					if (ok && !f->hm_exec) ok = cuda_ok(cudaGraphInstantiate(&f->hm_exec, g, 0), "instantiate the mesh graph");
					if (g) cudaGraphDestroy(g);
					if (ok) { f->hm_key = key; f->hm_tri_out = t; } else if (f->hm_exec) { cudaGraphExecDestroy(f->hm_exec); f->hm_exec = nullptr; }
				}
				ok = ok && cuda_ok(cudaGraphLaunch(f->hm_exec, st), "launch the mesh graph");
				if (ok) count_launch(f->hm_launches);
			} else if (ok) {
				ok = enqueue();
			}
			ok = cuda_ok(cudaStreamSynchronize(st), "mesh pipeline") && ok;
			const int *ha = f->pin_acc;
			if (ok && ha[2]) { set_error("device reported error flags 0x%x in the mesh pipeline", ha[2]); ok = false; }
			if (!ok) { host_block_free(v); host_block_free(t); return -1; }
			if (per_map_counts) {
				for (int i = 0; i < n_maps; i++) per_map_counts[i] = 0;
				for (int i = first; i < first + n_run; i++) per_map_counts[i] = ha[4 + 4 * kMaxMeshChunks + i];
			}
			out_mesh->vertices = (VertexC4ubV3f *)v;
			out_mesh->nVertices = ha[0];
			free(out_mesh->triangles);
			out_mesh->triangles = (int *)t;
			out_mesh->nTriangles = ha[1];
			return ha[0];
		}
		if (v) host_block_free(v);
		if (t) host_block_free(t);
	}
	if (!cuda_ok(cudaMemcpyAsync(dd + a.depth_off, depth_maps + a.depth_off, (size_t)(z.depth_off - a.depth_off), cudaMemcpyHostToDevice, st), "upload depth")) return -1;
	if (!cuda_ok(cudaMemcpyAsync(dc + a.color_off, depth_colors + a.color_off, (size_t)(z.color_off - a.color_off), cudaMemcpyHostToDevice, up), "upload colours") ||
		!cuda_ok(cudaEventRecord(f->ev_colors, up), "record colour upload")) return -1;
	f->colors_ready = f->ev_colors;
	if (frame_will_use_organized(f)) {
		// no pinned output block to store into: count stage -> survivor total to the host -> merge stage; the read-back is sized and
		// enqueued while the merge kernel runs
		PeerDst none; none.n = 0;
		if (frame_run_impl(f, f->in_depth.p, f->in_colors.p, first, n_run, nullptr, nullptr, none, st, kStageCount) < 0) return -1;
		if (!cuda_ok(cudaMemcpyAsync(po, f->ctl, sizeof(FrameCtl), cudaMemcpyDeviceToHost, st), "read survivor count") || !cuda_ok(cudaEventRecord(f->ev_count, st), "record count")) return -1;
		if (frame_run_impl(f, nullptr, nullptr, first, n_run, f->final_.as<uint4>(), nullptr, none, st, kStageMerge) < 0) return -1;
		if (!cuda_ok(cudaEventSynchronize(f->ev_count), "neighbour count")) return -1;
		const int n_early = hc->n_kept;
		void *v = host_block_alloc((size_t)std::max(n_early, 1) * sizeof(VertexC4ubV3f));
		if (!v) { set_error("out of host memory for %d vertices", n_early); return -1; }
		bool ok = (n_early == 0 || cuda_ok(cudaMemcpyAsync(v, f->final_.p, (size_t)n_early * sizeof(VertexC4ubV3f), cudaMemcpyDeviceToHost, st), "read vertices")) &&
			cuda_ok(cudaMemcpyAsync(po, f->ctl, sizeof(FrameCtl), cudaMemcpyDeviceToHost, st), "read counts") &&
			cuda_ok(cudaMemcpyAsync(po + 16, ls3d_frame_sensor_starts(f), sizeof(int) * (n_maps + 1), cudaMemcpyDeviceToHost, st), "read sensor starts") &&
			cuda_ok(cudaStreamSynchronize(st), "frame pipeline");
		if (ok && hc->err) { set_error("device reported error flags 0x%x in the frame pipeline", hc->err); ok = false; }
		if (ok && hc->n_final != n_early) { set_error("internal: merged %d vertices but the neighbour count announced %d", hc->n_final, n_early); ok = false; }
		if (!ok) { host_block_free(v); return -1; }
		if (per_map_counts) {
			for (int i = 0; i < n_maps; i++) per_map_counts[i] = 0;
			for (int i = first; i < first + n_run; i++) per_map_counts[i] = po[16 + i + 1] - po[16 + i];
		}
		out_mesh->vertices = (VertexC4ubV3f *)v;
		out_mesh->nVertices = n_early;
		return n_early;
	}
	if (ls3d_frame_run(f, f->in_depth.p, f->in_colors.p, first, n_run, st) < 0) return -1;
	// read back the counts, then exactly the bytes that exist
	const int *starts = ls3d_frame_sensor_starts(f);
	if (!cuda_ok(cudaMemcpyAsync(po, f->ctl, sizeof(FrameCtl), cudaMemcpyDeviceToHost, st), "read counts")) return -1;
	if (!cuda_ok(cudaMemcpyAsync(po + 16, starts, sizeof(int) * (n_maps + 1), cudaMemcpyDeviceToHost, st), "read sensor starts")) return -1;
	if (!cuda_ok(cudaStreamSynchronize(st), "frame pipeline")) return -1;
	if (hc->err) { set_error("device reported error flags 0x%x in the frame pipeline", hc->err); return -1; }
	const int n = hc->n_final;
	if (per_map_counts) {
		for (int i = 0; i < n_maps; i++) per_map_counts[i] = 0;
		for (int i = first; i < first + n_run; i++) per_map_counts[i] = po[16 + i + 1] - po[16 + i];
	}
	if (n > 0) {
		void *v = host_block_alloc((size_t)n * sizeof(VertexC4ubV3f));
		if (!v) { set_error("out of host memory for %d vertices", n); return -1; }
		if (!cuda_ok(cudaMemcpyAsync(v, f->final_.p, (size_t)n * sizeof(VertexC4ubV3f), cudaMemcpyDeviceToHost, st), "read vertices") ||
			!cuda_ok(cudaStreamSynchronize(st), "read vertices")) { host_block_free(v); return -1; }
		out_mesh->vertices = (VertexC4ubV3f *)v;
	} else {
		out_mesh->vertices = (VertexC4ubV3f *)host_block_alloc(sizeof(VertexC4ubV3f));   // valid empty allocation
	}
	out_mesh->nVertices = n;
	const int nt = f->want_triangles ? hc->n_triangles : 0;
	if (nt > 0) {
		void *t = host_block_alloc((size_t)nt * 3 * sizeof(int));
		if (!t) { set_error("out of host memory for %d triangles", nt); return -1; }
		if (!cuda_ok(cudaMemcpyAsync(t, f->tri.p, (size_t)nt * 3 * sizeof(int), cudaMemcpyDeviceToHost, st), "read triangles") ||
			!cuda_ok(cudaStreamSynchronize(st), "read triangles")) { host_block_free(t); return -1; }
		free(out_mesh->triangles);
		out_mesh->triangles = (int *)t;
		out_mesh->nTriangles = nt;
	}
	return n;
}

extern "C" void generateVerticesFromDepthMap(unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	float *intr_params, float *wtransform_params, Mesh *out_mesh,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ, int depth_map_index)
{
	const float b[6] = {minX, minY, minZ, maxX, maxY, maxZ};
	// the reference only ever reads entries 0..depth_map_index of widths/heights (depthprocessing.cpp:1646-1650)
	host_frame(depth_map_index + 1, depth_maps, depth_colors, widths, heights, intr_params, wtransform_params, out_mesh, b, depth_map_index, 1, 0, 0.0f, nullptr);
}

extern "C" void generateMeshFromDepthMaps(int n_maps, unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	float *intr_params, float *wtransform_params, Mesh *out_mesh, int bcolor_transfer,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ, int bgenerate_triangles)
{
	const float b[6] = {minX, minY, minZ, maxX, maxY, maxZ};
	// The reference runs generateTriangles unconditionally (depthprocessing.cpp:1786); its two flags only add the colour
	// correction and the multi-view vertex merge (:1757-1778), both outside this path, so they are read and ignored here:
	// the result is the reference's (false, false) branch — vertices in sensor order plus their depth-grid triangles.
	const int r = host_frame(n_maps, depth_maps, depth_colors, widths, heights, intr_params, wtransform_params, out_mesh, b, 0, n_maps, 0, 0.0f, nullptr, true);
	// Only the low byte of each flag is meaningful: C# marshals a 4-byte BOOL into a C++ bool parameter (KinectServer.cs:36-38
	// vs depthprocessing.h:108-110).  The call itself succeeded; say which branch the caller got (non-fatal, prefix "note:").
	const bool ct = (bcolor_transfer & 0xff) != 0, gt = (bgenerate_triangles & 0xff) != 0;
	if (r >= 0 && (ct || gt) && !ls3d_last_error()[0])
		set_error("note: generateMeshFromDepthMaps returned the reference's (bcolor_transfer=false, bgenerate_triangles=false) branch; %s%s%s "
			"(depthprocessing.cpp:1757-1778) %s outside this library's path and %s not applied", ct ? "colour transfer" : "",
			ct && gt ? " and " : "", gt ? "the multi-view vertex merge (generateVerticesConfidence + mergeVerticesForViews)" : "", ct && gt ? "are" : "is", ct && gt ? "were" : "was");
}

extern "C" int ls3d_frame_pipeline(int n_maps, unsigned char *depth_maps, unsigned char *depth_colors, int *widths, int *heights,
	float *intr_params, float *wtransform_params, Mesh *out_mesh,
	float minX, float minY, float minZ, float maxX, float maxY, float maxZ, int filter_k, float filter_maxDist, int *per_map_counts)
{
	const float b[6] = {minX, minY, minZ, maxX, maxY, maxZ};
	return host_frame(n_maps, depth_maps, depth_colors, widths, heights, intr_params, wtransform_params, out_mesh, b, 0, n_maps, filter_k, filter_maxDist, per_map_counts);
}

// ------------------------------------------------------------------------------------------------------
// ls3d_filter: the reference's filter() on flat arrays
// ------------------------------------------------------------------------------------------------------
static Ls3dFrame *g_filter_frame = nullptr;
static DevBuf g_fv, g_fc;

extern "C" int ls3d_filter(Point3f *verts, RGB *colors, int n, int k, float maxDist, int *old_to_new) {
	clear_error();
	if (n < 0 || (n > 0 && (!verts || !colors))) { set_error("ls3d_filter: bad arguments"); return -1; }
	if (k <= 0 || maxDist <= 0) {                      // filter.cpp:38-41: nothing happens
		if (old_to_new) for (int i = 0; i < n; i++) old_to_new[i] = -2;
		return n;
	}
	if (n == 0) return 0;
	if (n >= (1 << 24)) { set_error("ls3d_filter: at most 2^24-1 points per call"); return -1; }
	if (!ensure_device()) return -1;
	std::lock_guard<std::mutex> lk(api_mutex());
	cudaStream_t st = api_stream();
	if (!st) return -1;
	// a one-"sensor" context whose pixel budget covers n points
	if (!g_filter_frame || g_filter_frame->total_px < n) {
		if (g_filter_frame) frame_free(g_filter_frame);
		int cap = 1 << 16;
		while (cap < n) cap <<= 1;
		if (cap >= (1 << 24)) cap = (1 << 24) - 1;
		const int one = 1;
		g_filter_frame = ls3d_frame_create(1, &cap, &one);
		if (!g_filter_frame) return -1;
	}
	Ls3dFrame *f = g_filter_frame;
	if (!g_fv.reserve(12 * (size_t)n, "alloc filter verts") || !g_fc.reserve(4 * (size_t)n, "alloc filter colours")) return -1;
	f->filter_on = true;
	f->filter_k = k;
	f->filter_max_dist = maxDist;
	f->filter_thr = (float)pow((double)maxDist, 2.0);
	f->params_set = true;
	FilterBox hb;
	for (int a = 0; a < 3; a++) { hb.mn[a] = 0xffffffffu; hb.mx[a] = 0u; }
	if (!cuda_ok(cudaEventSynchronize(f->ev_staged), "wait for the previous descriptor upload")) return -1;
	memcpy(f->pin_sd, f->h_sd.data(), sizeof(SensorDesc) * 2);
	bool ok = cuda_ok(cudaMemcpyAsync(g_fv.p, verts, 12 * (size_t)n, cudaMemcpyHostToDevice, st), "upload verts") &&
		cuda_ok(cudaMemcpyAsync(g_fc.p, colors, 4 * (size_t)n, cudaMemcpyHostToDevice, st), "upload colours") &&
		cuda_ok(cudaMemcpyAsync(f->sd.p, f->pin_sd, sizeof(SensorDesc) * 2, cudaMemcpyHostToDevice, st), "upload descriptor") &&
		cuda_ok(cudaMemcpyAsync(f->box.p, &hb, sizeof(hb), cudaMemcpyHostToDevice, st), "init bbox") &&
		cuda_ok(cudaMemsetAsync(f->zero.p, 0, f->zero_bytes, st), "clear control block");
	if (!ok) return -1;
	const int blocks = std::min((n + 255) / 256, f->sm_count * 8);
	k_filter_pack<<<blocks, 256, 0, st>>>(g_fv.as<float>(), g_fc.as<unsigned>(), n, f->cloud0.as<uint4>(), f->box.as<FilterBox>());
	k_filter_grid_params<<<1, 1, 0, st>>>(f->box.as<FilterBox>(), f->sd.as<SensorDesc>(), f->ctl, f->culled_starts, n, maxDist);
	count_launch(2);
	PeerDst none; none.n = 0;
	if (frame_filter_stages(f, 0, 1, n, f->final_.as<uint4>(), nullptr, none, st) < 0) return -1;
	k_filter_unpack<<<blocks, 256, 0, st>>>(f->final_.as<uint4>(), f->ctl, g_fv.as<float>(), g_fc.as<unsigned>());
	count_launch(1);
	int *po = f->pin_out;
	ok = cuda_ok(cudaGetLastError(), "filter kernels") && cuda_ok(cudaMemcpyAsync(po, f->ctl, sizeof(FrameCtl), cudaMemcpyDeviceToHost, st), "read counts") &&
		cuda_ok(cudaStreamSynchronize(st), "filter");
	if (!ok) return -1;
	const FrameCtl *hc = reinterpret_cast<const FrameCtl *>(po);
	if (hc->err) { set_error("device reported error flags 0x%x in the filter", hc->err); return -1; }
	const int m = hc->n_final;
	ok = true;
	if (m > 0) {
		ok = cuda_ok(cudaMemcpyAsync(verts, g_fv.p, 12 * (size_t)m, cudaMemcpyDeviceToHost, st), "read verts") &&
			cuda_ok(cudaMemcpyAsync(colors, g_fc.p, 4 * (size_t)m, cudaMemcpyDeviceToHost, st), "read colours");
	}
	if (ok && old_to_new) ok = cuda_ok(cudaMemcpyAsync(old_to_new, f->map.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st), "read index map");
	ok = ok && cuda_ok(cudaStreamSynchronize(st), "filter read-back");
	return ok ? m : -1;
}

// Device self-checks of the arithmetic shortcuts whose correctness is an exhaustive fact rather than an identity.
// Returns the number of failures (0 = all good), or -1 when no device is usable.
extern "C" int ls3d_selftest(void) {
	clear_error();
	if (!ensure_device()) return -1;
	std::lock_guard<std::mutex> lk(api_mutex());
	cudaStream_t st = api_stream();
	if (!st) return -1;
	static DevBuf buf;
	if (!buf.reserve(256, "alloc selftest")) return -1;
	int bad = -1;
	bool ok = cuda_ok(cudaMemsetAsync(buf.p, 0, 4, st), "selftest");
	if (ok) { k_selftest_div1000<<<256, 256, 0, st>>>(buf.as<int>()); count_launch(1); }
	ok = ok && cuda_ok(cudaGetLastError(), "k_selftest_div1000") && cuda_ok(cudaMemcpyAsync(&bad, buf.p, 4, cudaMemcpyDeviceToHost, st), "selftest read") &&
		cuda_ok(cudaStreamSynchronize(st), "selftest");
	if (!ok) return -1;
	if (bad) set_error("ls3d_selftest: %d of 65536 depth values divide by 1000 differently from IEEE division", bad);
	return bad;
}
