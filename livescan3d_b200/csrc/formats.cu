// formats.cu — SURVEY §8f N4: the data formats either side of the path.
//
//   in  : the client's frame blob (LiveScanClient::SerializeFrame, src/LiveScanClient/liveScanClient.cpp:185-290; receiver
//         KinectSocket.ReceiveFrame, LiveScanServer/KinectSocket.cs:211-304) and NativeUtils' frames dump
//         (storeAllFramesInformation / loadAllFramesInformation, src/NativeUtils/depthprocessing.cpp:1316-1385)
//   out : binary PLY (Utils.saveToPly, LiveScanServer/Utils.cs:173-293) and the TransferServer mesh frame
//         (formVerticesChunks / formMeshChunks, LiveScanServer/TransferServer.cs:179-272; TransferSocket.SendFrame,
//         LiveScanServer/TransferSocket.cs:50-105)
//
// File and socket I/O, zstd and header text are host code; the byte re-packing of the cloud (16-byte records -> 15-byte PLY
// vertices / 13-byte PLY faces / split xyz + rgb arrays) and the chunk re-indexing of the triangle list run on the device, on
// the buffers the frame pipeline leaves there, so only the bytes that go to disk or socket cross PCIe.
#include "ls3d_common.cuh"
#include "ls3d_internal.h"
#include "../../include/ls3d.h"
#include <dlfcn.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

using namespace ls3d;

// ======================================================================================================
// zstd (the client compresses its payload with zstd 1.1.3, include/zstd.h:57-59): bound at run time to the system's
// libzstd — the simple API used here (ZSTD_compress / ZSTD_decompress) is frozen since 1.0
// ======================================================================================================
namespace {
struct Zstd {
	size_t (*compress)(void *, size_t, const void *, size_t, int) = nullptr;
	size_t (*decompress)(void *, size_t, const void *, size_t) = nullptr;
	size_t (*compressBound)(size_t) = nullptr;
	unsigned (*isError)(size_t) = nullptr;
	unsigned long long (*getFrameContentSize)(const void *, size_t) = nullptr;
	bool ok = false;
};
const Zstd &zstd() {
	static Zstd z = [] {
		Zstd r;
		void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
		if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
		if (!h) return r;
		r.compress = (decltype(r.compress))dlsym(h, "ZSTD_compress");
		r.decompress = (decltype(r.decompress))dlsym(h, "ZSTD_decompress");
		r.compressBound = (decltype(r.compressBound))dlsym(h, "ZSTD_compressBound");
		r.isError = (decltype(r.isError))dlsym(h, "ZSTD_isError");
		r.getFrameContentSize = (decltype(r.getFrameContentSize))dlsym(h, "ZSTD_getFrameContentSize");
		r.ok = r.compress && r.decompress && r.compressBound && r.isError && r.getFrameContentSize;
		return r;
	}();
	return z;
}
int rd_i32(const unsigned char *p) { int v; memcpy(&v, p, 4); return v; }
void wr_i32(unsigned char *p, int v) { memcpy(p, &v, 4); }

// walks the body records of a payload (liveScanClient.cpp:236-270): returns their byte length or -1 if truncated
long long bodies_length(const unsigned char *p, long long avail, int *n_bodies) {
	if (avail < 4) return -1;
	const int nb = rd_i32(p);
	if (nb < 0) return -1;
	long long pos = 4;
	for (int i = 0; i < nb; i++) {
		if (pos + 5 > avail) return -1;
		const int nj = rd_i32(p + pos + 1);      // bool bTracked (1 byte), then int nJoints
		if (nj < 0) return -1;
		pos += 5 + (long long)nj * 28;           // JointType, TrackingState, 3 floats position, 2 floats colour-space position
		if (pos > avail) return -1;
	}
	*n_bodies = nb;
	return pos;
}
}  // namespace

extern "C" int ls3d_client_frame_header(const unsigned char *blob, long long blob_bytes, Ls3dClientFrameInfo *info) {
	clear_error();
	if (!blob || !info) { set_error("ls3d_client_frame_header: null argument"); return -1; }
	if (blob_bytes < 16) { set_error("ls3d_client_frame_header: %lld bytes is shorter than the 16-byte header", blob_bytes); return -1; }
	info->payload_bytes = rd_i32(blob);
	info->compressed = rd_i32(blob + 4);
	info->width = rd_i32(blob + 8);
	info->height = rd_i32(blob + 12);
	info->n_bodies = -1;
	info->raw_bytes = -1;
	if (info->payload_bytes <= 0) { set_error("ls3d_client_frame_header: payload size %d (the client's 'no more frames' marker)", info->payload_bytes); return -1; }
	if (info->width < 0 || info->height < 0 || (long long)info->width * info->height > (1ll << 28)) { set_error("ls3d_client_frame_header: image size %dx%d", info->width, info->height); return -1; }
	if (16 + (long long)info->payload_bytes > blob_bytes) { set_error("ls3d_client_frame_header: payload of %d bytes but only %lld present", info->payload_bytes, blob_bytes - 16); return -1; }
	return 0;
}

extern "C" int ls3d_client_frame_unpack(const unsigned char *blob, long long blob_bytes, unsigned char *depth_out, unsigned char *colors_out,
	unsigned char *bodies_out, long long bodies_cap, Ls3dClientFrameInfo *info)
{
	Ls3dClientFrameInfo local;
	if (!info) info = &local;
	if (ls3d_client_frame_header(blob, blob_bytes, info) < 0) return -1;
	const long long px = (long long)info->width * info->height;
	const unsigned char *payload = blob + 16;
	long long raw = info->payload_bytes;
	std::vector<unsigned char> tmp;
	if (info->compressed == 1) {                       // KinectSocket.cs:245-246
		const Zstd &z = zstd();
		if (!z.ok) { set_error("ls3d_client_frame_unpack: the frame is zstd-compressed and libzstd could not be loaded"); return -1; }
		const unsigned long long want = z.getFrameContentSize(payload, (size_t)info->payload_bytes);
		if (want == (unsigned long long)-1 || want == (unsigned long long)-2 || want > (1ull << 32)) { set_error("ls3d_client_frame_unpack: not a zstd frame with a known content size"); return -1; }
		tmp.resize((size_t)std::max<unsigned long long>(want, 1));
		const size_t got = z.decompress(tmp.data(), tmp.size(), payload, (size_t)info->payload_bytes);
		if (z.isError(got) || got != want) { set_error("ls3d_client_frame_unpack: zstd decompression failed"); return -1; }
		payload = tmp.data();
		raw = (long long)got;
	}
	info->raw_bytes = raw;
	if (raw < 5 * px + 4) { set_error("ls3d_client_frame_unpack: payload of %lld bytes cannot hold a %dx%d depth + colour image and a body count", raw, info->width, info->height); return -1; }
	int nb = 0;
	const long long blen = bodies_length(payload + 5 * px, raw - 5 * px, &nb);
	if (blen < 0) { set_error("ls3d_client_frame_unpack: truncated body records"); return -1; }
	info->n_bodies = nb;
	if (depth_out) memcpy(depth_out, payload, (size_t)(2 * px));                 // KinectSocket.cs:254
	if (colors_out) memcpy(colors_out, payload + 2 * px, (size_t)(3 * px));      // KinectSocket.cs:255
	if (bodies_out) {
		if (bodies_cap < blen) { set_error("ls3d_client_frame_unpack: %lld bytes of body records, room for %lld", blen, bodies_cap); return -1; }
		memcpy(bodies_out, payload + 5 * px, (size_t)blen);
	}
	return (int)std::min<long long>(blen, 0x7fffffff);
}

extern "C" long long ls3d_client_frame_pack(const unsigned char *depth, const unsigned char *colors, int width, int height,
	const unsigned char *bodies, long long bodies_bytes, int compression_level, unsigned char *out, long long out_cap)
{
	clear_error();
	if (!depth || !colors || width < 0 || height < 0) { set_error("ls3d_client_frame_pack: bad argument"); return -1; }
	const long long px = (long long)width * height;
	if (px > (1ll << 28)) { set_error("ls3d_client_frame_pack: image size %dx%d", width, height); return -1; }
	const unsigned char zero_bodies[4] = {0, 0, 0, 0};
	if (!bodies) { bodies = zero_bodies; bodies_bytes = 4; }
	int nb = 0;
	if (bodies_length(bodies, bodies_bytes, &nb) != bodies_bytes) { set_error("ls3d_client_frame_pack: body records do not parse to exactly %lld bytes", bodies_bytes); return -1; }
	const long long raw = 5 * px + bodies_bytes;
	if (raw > 0x7fffffffll) { set_error("ls3d_client_frame_pack: payload exceeds the header's 32-bit size"); return -1; }
	const Zstd &z = zstd();
	if (compression_level > 0 && !z.ok) { set_error("ls3d_client_frame_pack: compression requested and libzstd could not be loaded"); return -1; }
	const long long worst = 16 + (compression_level > 0 ? (long long)z.compressBound((size_t)raw) : raw);
	if (!out) return worst;                            // size query
	std::vector<unsigned char> payload((size_t)raw);
	memcpy(payload.data(), depth, (size_t)(2 * px));
	memcpy(payload.data() + 2 * px, colors, (size_t)(3 * px));
	memcpy(payload.data() + 5 * px, bodies, (size_t)bodies_bytes);
	long long size = raw;
	if (compression_level > 0) {                       // liveScanClient.cpp:266-279
		if (out_cap < worst) { set_error("ls3d_client_frame_pack: need %lld bytes of output, got %lld", worst, out_cap); return -1; }
		const size_t c = z.compress(out + 16, (size_t)(out_cap - 16), payload.data(), (size_t)raw, compression_level);
		if (z.isError(c)) { set_error("ls3d_client_frame_pack: zstd compression failed"); return -1; }
		size = (long long)c;
	} else {
		if (out_cap < 16 + raw) { set_error("ls3d_client_frame_pack: need %lld bytes of output, got %lld", 16 + raw, out_cap); return -1; }
		memcpy(out + 16, payload.data(), (size_t)raw);
	}
	wr_i32(out, (int)size);                            // liveScanClient.cpp:283-288
	wr_i32(out + 4, compression_level > 0 ? 1 : 0);
	wr_i32(out + 8, width);
	wr_i32(out + 12, height);
	return 16 + size;
}

// ======================================================================================================
// frames dump (depthprocessing.cpp:1316-1385): int n_maps | int widths[n] | int heights[n] | per map: depth u16[w*h], RGB u8[3*w*h]
// | float intr[7n] | float wtransform[12n]
// ======================================================================================================
extern "C" int ls3d_frames_info_store(const char *filename, int n_maps, const unsigned char *depth_maps, const unsigned char *depth_colors,
	const int *widths, const int *heights, const float *intr_params, const float *wtransform_params)
{
	clear_error();
	if (!filename || n_maps < 0 || (n_maps > 0 && (!depth_maps || !depth_colors || !widths || !heights || !intr_params || !wtransform_params))) { set_error("ls3d_frames_info_store: bad argument"); return -1; }
	FILE *f = fopen(filename, "wb");
	if (!f) { set_error("ls3d_frames_info_store: cannot open %s for writing", filename); return -1; }
	bool ok = fwrite(&n_maps, 4, 1, f) == 1;
	if (n_maps > 0) ok = ok && fwrite(widths, 4, (size_t)n_maps, f) == (size_t)n_maps && fwrite(heights, 4, (size_t)n_maps, f) == (size_t)n_maps;
	size_t pd = 0, pc = 0;
	for (int i = 0; i < n_maps && ok; i++) {
		const size_t px = (size_t)widths[i] * heights[i];
		ok = (px == 0 || (fwrite(depth_maps + pd, 1, 2 * px, f) == 2 * px && fwrite(depth_colors + pc, 1, 3 * px, f) == 3 * px));
		pd += 2 * px; pc += 3 * px;
	}
	if (n_maps > 0) ok = ok && fwrite(intr_params, 4, 7 * (size_t)n_maps, f) == 7 * (size_t)n_maps && fwrite(wtransform_params, 4, 12 * (size_t)n_maps, f) == 12 * (size_t)n_maps;
	ok = (fclose(f) == 0) && ok;
	if (!ok) { set_error("ls3d_frames_info_store: short write to %s", filename); return -1; }
	return 0;
}

extern "C" void ls3d_frames_info_free(Ls3dFramesInfo *info) {
	if (!info) return;
	host_block_free(info->depth_maps);
	host_block_free(info->depth_colors);
	free(info->widths);
	free(info->heights);
	free(info->intr_params);
	free(info->wtransform_params);
	memset(info, 0, sizeof(*info));
}

extern "C" int ls3d_frames_info_load(const char *filename, Ls3dFramesInfo *out) {
	clear_error();
	if (!filename || !out) { set_error("ls3d_frames_info_load: null argument"); return -1; }
	memset(out, 0, sizeof(*out));
	FILE *f = fopen(filename, "rb");
	if (!f) { set_error("ls3d_frames_info_load: cannot open %s", filename); return -1; }
	auto fail = [&](const char *why) { set_error("ls3d_frames_info_load: %s (%s)", why, filename); fclose(f); ls3d_frames_info_free(out); return -1; };
	int n = 0;
	if (fread(&n, 4, 1, f) != 1) return fail("missing sensor count");
	if (n < 0 || n > 4096) return fail("implausible sensor count");
	out->n_maps = n;
	out->widths = (int *)malloc(sizeof(int) * std::max(n, 1));
	out->heights = (int *)malloc(sizeof(int) * std::max(n, 1));
	out->intr_params = (float *)malloc(sizeof(float) * 7 * std::max(n, 1));
	out->wtransform_params = (float *)malloc(sizeof(float) * 12 * std::max(n, 1));
	if (!out->widths || !out->heights || !out->intr_params || !out->wtransform_params) return fail("out of memory");
	if (n > 0 && (fread(out->widths, 4, (size_t)n, f) != (size_t)n || fread(out->heights, 4, (size_t)n, f) != (size_t)n)) return fail("truncated size lists");
	size_t pd = 0, pc = 0;
	for (int i = 0; i < n; i++) {
		if (out->widths[i] < 0 || out->heights[i] < 0 || (long long)out->widths[i] * out->heights[i] > (1ll << 28)) return fail("implausible image size");
		const size_t px = (size_t)out->widths[i] * out->heights[i];
		pd += 2 * px; pc += 3 * px;
	}
	// page-locked, so the loaded frame takes ls3d_frame_pipeline's graph schedule as it is
	out->depth_maps = (unsigned char *)host_block_alloc(std::max<size_t>(pd, 1));
	out->depth_colors = (unsigned char *)host_block_alloc(std::max<size_t>(pc, 1));
	if (!out->depth_maps || !out->depth_colors) return fail("out of memory");
	pd = pc = 0;
	for (int i = 0; i < n; i++) {
		const size_t px = (size_t)out->widths[i] * out->heights[i];
		if (px && (fread(out->depth_maps + pd, 1, 2 * px, f) != 2 * px || fread(out->depth_colors + pc, 1, 3 * px, f) != 3 * px)) return fail("truncated image data");
		pd += 2 * px; pc += 3 * px;
	}
	if (n > 0 && (fread(out->intr_params, 4, 7 * (size_t)n, f) != 7 * (size_t)n || fread(out->wtransform_params, 4, 12 * (size_t)n, f) != 12 * (size_t)n)) return fail("truncated parameters");
	fclose(f);
	return 0;
}

// ======================================================================================================
// device re-packing
// ======================================================================================================
// Body of the binary PLY: n vertices of 15 bytes (float x, y, z; uchar r, g, b — Utils.cs:257-266, or the vertex-only variant
// :200-212) then nt faces of 13 bytes (uchar 3; int a, b, c — :268-274).  One thread produces 16 consecutive output bytes
// (one STG.128); every byte is fetched from the record it comes from (the loads of a warp fall into ~550 contiguous input bytes,
// so they are L1 hits after the first touch of each line).
__device__ __forceinline__ unsigned ply_byte(const uint8_t *__restrict__ v, const uint8_t *__restrict__ t, long long vbytes, long long o) {
	if (o < vbytes) {
		const long long r = o / 15;
		const int k = (int)(o - 15 * r);
		return __ldg(v + 16 * r + (k < 12 ? 4 + k : k - 12));     // record = R,G,B,A,X,Y,Z
	}
	o -= vbytes;
	const long long r = o / 13;
	const int k = (int)(o - 13 * r);
	return k == 0 ? 3u : (unsigned)__ldg(t + 12 * r + (k - 1));
}

// 16 consecutive bytes of a byte string held in 8 little-endian words P[0..7], starting at byte k (0 <= k <= 15)
__device__ __forceinline__ uint4 bytes16_at(const unsigned (&P)[8], int k) {
	const unsigned sh = 8u * (unsigned)(k & 3);
	switch (k >> 2) {
	case 0: return make_uint4(__funnelshift_r(P[0], P[1], sh), __funnelshift_r(P[1], P[2], sh), __funnelshift_r(P[2], P[3], sh), __funnelshift_r(P[3], P[4], sh));
	case 1: return make_uint4(__funnelshift_r(P[1], P[2], sh), __funnelshift_r(P[2], P[3], sh), __funnelshift_r(P[3], P[4], sh), __funnelshift_r(P[4], P[5], sh));
	case 2: return make_uint4(__funnelshift_r(P[2], P[3], sh), __funnelshift_r(P[3], P[4], sh), __funnelshift_r(P[4], P[5], sh), __funnelshift_r(P[5], P[6], sh));
	default: return make_uint4(__funnelshift_r(P[3], P[4], sh), __funnelshift_r(P[4], P[5], sh), __funnelshift_r(P[5], P[6], sh), __funnelshift_r(P[6], P[7], sh));
	}
}

// One thread = one aligned 16-byte piece of the body.  Pieces that lie inside one section are assembled in registers from whole
// words: the two 16-byte records a vertex piece touches (2 x LDG.128) are laid out as their 30-byte PLY string by fixed funnel
// shifts, a face piece does the same with the seven indices it can touch; the piece is then cut out at its phase (piece offset mod
// record size).  Only the piece that straddles the vertex/face boundary and the tail go byte by byte.
__global__ void __launch_bounds__(256) k_pack_ply_body(const uint8_t *__restrict__ verts, long long n, const uint8_t *__restrict__ tris, long long nt, uint8_t *__restrict__ out) {
	const long long vbytes = 15 * n, total = vbytes + 13 * nt;
	const uint4 *__restrict__ v4 = reinterpret_cast<const uint4 *>(verts);
	const unsigned *__restrict__ t4 = reinterpret_cast<const unsigned *>(tris);
	for (long long o = 16ll * (blockIdx.x * 256ll + threadIdx.x); o < total; o += 16ll * 256 * gridDim.x) {
		if (o + 16 <= vbytes) {
			const long long r = o / 15;
			const int k = (int)(o - 15 * r);
			const uint4 a = __ldg(v4 + r), b = __ldg(v4 + r + 1);          // (rgba, x, y, z); a 16-byte piece always reaches into record r+1
			const unsigned P[8] = {a.y, a.z, a.w, (a.x & 0xffffffu) | (b.y << 24), __funnelshift_r(b.y, b.z, 8), __funnelshift_r(b.z, b.w, 8),
				(b.w >> 8) | (b.x << 24), (b.x >> 8) & 0xffffu};
			*reinterpret_cast<uint4 *>(out + o) = bytes16_at(P, k);
		} else if (o >= vbytes && o + 16 <= total) {
			const long long q = o - vbytes, r = q / 13;
			const int k = (int)(q - 13 * r);
			const unsigned *src = t4 + 3 * r;
			const unsigned a0 = __ldg(src), b0 = __ldg(src + 1), c0 = __ldg(src + 2), a1 = __ldg(src + 3), b1 = __ldg(src + 4), c1 = __ldg(src + 5);
			const unsigned a2 = r + 2 < nt ? __ldg(src + 6) : 0u;           // only its low bytes, and only when the piece starts at k >= 10
			const unsigned b2 = r + 2 < nt ? __ldg(src + 7) : 0u;
			const unsigned P[8] = {3u | (a0 << 8), __funnelshift_r(a0, b0, 24), __funnelshift_r(b0, c0, 24), (c0 >> 24) | (3u << 8) | (a1 << 16),
				__funnelshift_r(a1, b1, 16), __funnelshift_r(b1, c1, 16), (c1 >> 16) | (3u << 16) | (a2 << 24), __funnelshift_r(a2, b2, 8)};
			*reinterpret_cast<uint4 *>(out + o) = bytes16_at(P, k);
		} else {
			unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
			for (int b = 0; b < 16; b++)
				if (o + b < total) w[b >> 2] |= ply_byte(verts, tris, vbytes, o + b) << (8 * (b & 3));
			if (o + 16 <= total) *reinterpret_cast<uint4 *>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
			else for (int b = 0; o + b < total; b++) out[o + b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
		}
	}
}

// Body of a TransferSocket frame after its header and chunk-size lists (TransferSocket.cs:66-100): float xyz[3n] | uchar rgb[3n] |
// int triangles[3nt].  xyz pieces are four word loads, triangle pieces a shifted copy; the rgb section (a ninth of the vertex
// bytes) and the pieces on section boundaries are gathered byte by byte.
__global__ void __launch_bounds__(256) k_pack_transfer_body(const uint8_t *__restrict__ verts, long long n, const uint8_t *__restrict__ tris, long long nt, uint8_t *__restrict__ out) {
	const long long xb = 12 * n, cb = 3 * n, tb = xb + cb, total = tb + 12 * nt;
	const unsigned *__restrict__ v1 = reinterpret_cast<const unsigned *>(verts);
	const unsigned *__restrict__ t4 = reinterpret_cast<const unsigned *>(tris);
	for (long long o = 16ll * (blockIdx.x * 256ll + threadIdx.x); o < total; o += 16ll * 256 * gridDim.x) {
		if (o + 16 <= xb) {
			unsigned w[4];
#pragma unroll
			for (int i = 0; i < 4; i++) { const long long j = o / 4 + i, r = j / 3; w[i] = __ldg(v1 + 4 * r + 1 + (j - 3 * r)); }
			*reinterpret_cast<uint4 *>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
		} else if (o >= tb && o + 16 <= total) {
			const long long q = o - tb;                                     // source byte offset in the index list
			const unsigned *src = t4 + q / 4;
			const unsigned sh = 8u * (unsigned)(q & 3);
			const unsigned s0 = __ldg(src), s1 = __ldg(src + 1), s2 = __ldg(src + 2), s3 = __ldg(src + 3), s4 = (sh && q / 4 + 4 < 3 * nt) ? __ldg(src + 4) : 0u;
			*reinterpret_cast<uint4 *>(out + o) = make_uint4(__funnelshift_r(s0, s1, sh), __funnelshift_r(s1, s2, sh), __funnelshift_r(s2, s3, sh), __funnelshift_r(s3, s4, sh));
		} else {
			unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
			for (int b = 0; b < 16; b++) {
				const long long q = o + b;
				if (q >= total) break;
				unsigned v;
				if (q < xb) { const long long r = q / 12; v = __ldg(verts + 16 * r + 4 + (q - 12 * r)); }
				else if (q < tb) { const long long c = q - xb, r = c / 3; v = __ldg(verts + 16 * r + (c - 3 * r)); }
				else v = __ldg(tris + (q - tb));
				w[b >> 2] |= v << (8 * (b & 3));
			}
			if (o + 16 <= total) *reinterpret_cast<uint4 *>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
			else for (int b = 0; o + b < total; b++) out[o + b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
		}
	}
}

static int pack_grid(long long total_bytes) {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	return (int)std::max<long long>(1, std::min<long long>((total_bytes + 4095) / 4096, (long long)sms * 8));
}

extern "C" int ls3d_pack_ply_body_device(const void *d_vertices, int n_vertices, const int *d_triangles, int n_triangles, void *d_out, void *stream) {
	clear_error();
	if (n_vertices < 0 || n_triangles < 0 || !d_out || (n_vertices && !d_vertices) || (n_triangles && !d_triangles)) { set_error("ls3d_pack_ply_body_device: bad argument"); return -1; }
	if (((uintptr_t)d_out) & 15) { set_error("ls3d_pack_ply_body_device: the output buffer must be 16-byte aligned"); return -1; }
	if (!ensure_device()) return -1;
	const long long total = 15ll * n_vertices + 13ll * n_triangles;
	if (total == 0) return 0;
	k_pack_ply_body<<<pack_grid(total), 256, 0, (cudaStream_t)stream>>>((const uint8_t *)d_vertices, n_vertices, (const uint8_t *)d_triangles, n_triangles, (uint8_t *)d_out);
	count_launch(1);
	return cuda_ok(cudaGetLastError(), "k_pack_ply_body") ? 1 : -1;
}

extern "C" int ls3d_pack_transfer_body_device(const void *d_vertices, int n_vertices, const int *d_triangles, int n_triangles, void *d_out, void *stream) {
	clear_error();
	if (n_vertices < 0 || n_triangles < 0 || !d_out || (n_vertices && !d_vertices) || (n_triangles && !d_triangles)) { set_error("ls3d_pack_transfer_body_device: bad argument"); return -1; }
	if (((uintptr_t)d_out) & 15) { set_error("ls3d_pack_transfer_body_device: the output buffer must be 16-byte aligned"); return -1; }
	if (!ensure_device()) return -1;
	const long long total = 15ll * n_vertices + 12ll * n_triangles;
	if (total == 0) return 0;
	k_pack_transfer_body<<<pack_grid(total), 256, 0, (cudaStream_t)stream>>>((const uint8_t *)d_vertices, n_vertices, (const uint8_t *)d_triangles, n_triangles, (uint8_t *)d_out);
	count_launch(1);
	return cuda_ok(cudaGetLastError(), "k_pack_transfer_body") ? 1 : -1;
}

// ---- PLY header (Utils.cs:180-190, :231-241).  StreamWriter.WriteLine ends the first write with Environment.NewLine, "\r\n" on the
// Windows machines the server runs on; everything else is written with explicit "\n".
static std::string ply_header(int n_vertices, int n_triangles) {
	std::string h = "ply\nformat binary_little_endian 1.0\r\n";
	h += "element vertex " + std::to_string(n_vertices) + "\n";
	h += "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n";
	if (n_triangles >= 0) {
		h += "element face " + std::to_string(n_triangles) + "\n";
		h += "property list uchar int vertex_index\n";
	}
	h += "end_header\n";
	return h;
}

extern "C" long long ls3d_ply_binary_size(int n_vertices, int n_triangles) {
	if (n_vertices < 0) return -1;
	return (long long)ply_header(n_vertices, n_triangles).size() + 15ll * n_vertices + 13ll * std::max(n_triangles, 0);
}

// scratch shared by the host-buffer writers below (under the API lock)
static DevBuf g_fmt_in, g_fmt_tri, g_fmt_out, g_fmt_aux;

extern "C" long long ls3d_write_ply_binary(const VertexC4ubV3f *vertices, int n_vertices, const int *triangles, int n_triangles, unsigned char *out, long long out_cap) {
	clear_error();
	if (n_vertices < 0 || (n_vertices && !vertices) || (n_triangles > 0 && !triangles) || !out) { set_error("ls3d_write_ply_binary: bad argument"); return -1; }
	const std::string h = ply_header(n_vertices, n_triangles);
	const int nt = std::max(n_triangles, 0);
	const long long body = 15ll * n_vertices + 13ll * nt, total = (long long)h.size() + body;
	if (out_cap < total) { set_error("ls3d_write_ply_binary: need %lld bytes of output, got %lld", total, out_cap); return -1; }
	if (!ensure_device()) return -1;
	std::lock_guard<std::mutex> lk(api_mutex());
	cudaStream_t st = api_stream();
	if (!st) return -1;
	memcpy(out, h.data(), h.size());
	if (body == 0) return total;
	if (!g_fmt_in.reserve(16 * (size_t)std::max(n_vertices, 1), "alloc vertices") || !g_fmt_tri.reserve(12 * (size_t)std::max(nt, 1), "alloc triangles") ||
		!g_fmt_out.reserve((size_t)body + 16, "alloc PLY body")) return -1;
	bool ok = (n_vertices == 0 || cuda_ok(cudaMemcpyAsync(g_fmt_in.p, vertices, 16 * (size_t)n_vertices, cudaMemcpyHostToDevice, st), "upload vertices")) &&
		(nt == 0 || cuda_ok(cudaMemcpyAsync(g_fmt_tri.p, triangles, 12 * (size_t)nt, cudaMemcpyHostToDevice, st), "upload triangles"));
	ok = ok && ls3d_pack_ply_body_device(g_fmt_in.p, n_vertices, g_fmt_tri.as<int>(), nt, g_fmt_out.p, st) >= 0 &&
		cuda_ok(cudaMemcpyAsync(out + h.size(), g_fmt_out.p, (size_t)body, cudaMemcpyDeviceToHost, st), "read PLY body");
	ok = cuda_ok(cudaStreamSynchronize(st), "ls3d_write_ply_binary") && ok;
	return ok ? total : -1;
}

// ---- ASCII PLY (Utils.saveToPly with binary=false, Utils.cs:204-214 / :276-289): text, formatted on the host like the other
// host codecs of this file.  Quirks kept: the header's first WriteLine already ends in "\n" and gets "\r\n" on top; every vertex
// line ends in a blank before the line end; a face line is "3 " followed by the three indices with NOTHING between them (:284-287).
// Numbers are Single.ToString(CultureInfo.InvariantCulture) of the .NET Framework 4.5 the server targets (LiveScanServer.csproj):
// general format with 7 significant digits — fixed notation for decimal exponents -5 < e < 7, else d.dddE+XX (two exponent digits at
// least), trailing zeros dropped, "NaN" / "Infinity" / "-Infinity", and "0" for either zero.  The CLR rounds through the C runtime's
// _ecvt, i.e. half away from zero on the exact value; glibc rounds the same except on an exact tie at the 8th digit (possible only
// for values such as 1234566.5f), which is detected and rounded by hand.  No .NET runtime exists here: parity unpinned, the format
// is restated from the documented behaviour and checked against an independent decimal implementation (oracle/formats_oracle.py).
static int fmt_single_net45(float v, char *o) {
	if (v != v) { memcpy(o, "NaN", 3); return 3; }
	if (v == INFINITY) { memcpy(o, "Infinity", 8); return 8; }
	if (v == -INFINITY) { memcpy(o, "-Infinity", 9); return 9; }
	if (v == 0.0f) { o[0] = '0'; return 1; }
	char e[160];
	snprintf(e, sizeof(e), "%.9E", (double)v);                 // [-]d.dddddddddE±XX
	const char *m = e + (e[0] == '-' ? 1 : 0);
	char dig[8];
	int exp10;
	if (m[8] == '5' && m[9] == '0' && m[10] == '0') {           // significant digit k >= 2 is m[k]
		// perhaps an exact tie at the 8th digit: round the exact expansion (a float has at most 112 significant digits) by hand —
		// a 5 there rounds up whether or not anything follows it
		snprintf(e, sizeof(e), "%.120E", (double)v);
		m = e + (e[0] == '-' ? 1 : 0);
		exp10 = atoi(strchr(m, 'E') + 1);
		long long d7 = (m[0] - '0');
		for (int i = 2; i < 8; i++) d7 = d7 * 10 + (m[i] - '0');
		if (m[8] >= '5') d7++;
		if (d7 == 10000000) { d7 = 1000000; exp10++; }
		for (int i = 6; i >= 0; i--) { dig[i] = (char)('0' + d7 % 10); d7 /= 10; }
	} else {
		snprintf(e, sizeof(e), "%.6E", (double)v);               // 7 significant digits, correctly rounded
		m = e + (e[0] == '-' ? 1 : 0);
		dig[0] = m[0];
		for (int i = 0; i < 6; i++) dig[1 + i] = m[2 + i];
		exp10 = atoi(strchr(m, 'E') + 1);
	}
	int nd = 7;
	while (nd > 1 && dig[nd - 1] == '0') nd--;
	char *q = o;
	if (v < 0) *q++ = '-';
	if (exp10 > -5 && exp10 < 7) {
		if (exp10 >= 0) {
			for (int i = 0; i <= exp10; i++) *q++ = i < nd ? dig[i] : '0';
			if (nd > exp10 + 1) { *q++ = '.'; for (int i = exp10 + 1; i < nd; i++) *q++ = dig[i]; }
		} else {
			*q++ = '0'; *q++ = '.';
			for (int i = 0; i < -exp10 - 1; i++) *q++ = '0';
			for (int i = 0; i < nd; i++) *q++ = dig[i];
		}
	} else {
		*q++ = dig[0];
		if (nd > 1) { *q++ = '.'; for (int i = 1; i < nd; i++) *q++ = dig[i]; }
		*q++ = 'E';
		*q++ = exp10 < 0 ? '-' : '+';
		const int a = exp10 < 0 ? -exp10 : exp10;
		if (a >= 100) *q++ = (char)('0' + a / 100);
		*q++ = (char)('0' + (a / 10) % 10);
		*q++ = (char)('0' + a % 10);
	}
	return (int)(q - o);
}
static int fmt_uint(unsigned v, char *o) {
	char t[12];
	int n = 0;
	do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
	for (int i = 0; i < n; i++) o[i] = t[n - 1 - i];
	return n;
}
static int fmt_int(int v, char *o) {
	if (v >= 0) return fmt_uint((unsigned)v, o);
	o[0] = '-';
	return 1 + fmt_uint(0u - (unsigned)v, o + 1);
}

static std::string ply_ascii_header(int n_vertices, int n_triangles) {
	std::string h = "ply\nformat ascii 1.0\n\r\n";
	h += "element vertex " + std::to_string(n_vertices) + "\n";
	h += "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n";
	if (n_triangles >= 0) {
		h += "element face " + std::to_string(n_triangles) + "\n";
		h += "property list uchar int vertex_index\n";
	}
	h += "end_header\n";
	return h;
}
constexpr long long kAsciiVertexMax = 64, kAsciiFaceMax = 40;      // "-1.234568E-38 " x 3 + "255 " x 3 + CRLF = 59; "3 " + 3 x 11 + CRLF = 37

extern "C" long long ls3d_ply_ascii_bound(int n_vertices, int n_triangles) {
	if (n_vertices < 0) return -1;
	return (long long)ply_ascii_header(n_vertices, n_triangles).size() + kAsciiVertexMax * n_vertices + kAsciiFaceMax * std::max(n_triangles, 0);
}

extern "C" long long ls3d_write_ply_ascii(const VertexC4ubV3f *vertices, int n_vertices, const int *triangles, int n_triangles, unsigned char *out, long long out_cap) {
	clear_error();
	if (n_vertices < 0 || (n_vertices && !vertices) || (n_triangles > 0 && !triangles) || !out) { set_error("ls3d_write_ply_ascii: bad argument"); return -1; }
	const long long bound = ls3d_ply_ascii_bound(n_vertices, n_triangles);
	if (out_cap < bound) { set_error("ls3d_write_ply_ascii: need ls3d_ply_ascii_bound() = %lld bytes of output, got %lld", bound, out_cap); return -1; }
	const std::string h = ply_ascii_header(n_vertices, n_triangles);
	const int nt = std::max(n_triangles, 0);
	// lines are formatted by a few host threads, each into its own slice of a scratch buffer (fixed upper bound per line), then packed
	const int n_thr = (int)std::max(1u, std::min(16u, std::min(std::thread::hardware_concurrency(), (unsigned)(((long long)n_vertices + nt) / 20000 + 1))));
	const long long items = (long long)n_vertices + nt;
	std::vector<std::vector<char>> part(n_thr);
	auto work = [&](int t) {
		const long long a = items * t / n_thr, b = items * (t + 1) / n_thr;
		std::vector<char> &buf = part[t];
		buf.resize((size_t)std::max<long long>(1, (b - a) * kAsciiVertexMax));
		char *q = buf.data();
		for (long long i = a; i < b; i++) {
			if (i < n_vertices) {
				const VertexC4ubV3f &v = vertices[i];
				q += fmt_single_net45(v.X, q); *q++ = ' ';
				q += fmt_single_net45(v.Y, q); *q++ = ' ';
				q += fmt_single_net45(v.Z, q); *q++ = ' ';
				q += fmt_uint(v.R, q); *q++ = ' ';
				q += fmt_uint(v.G, q); *q++ = ' ';
				q += fmt_uint(v.B, q); *q++ = ' ';
			} else {
				const int *tr = triangles + 3 * (i - n_vertices);
				*q++ = '3'; *q++ = ' ';
				q += fmt_int(tr[0], q); q += fmt_int(tr[1], q); q += fmt_int(tr[2], q);
			}
			*q++ = '\r'; *q++ = '\n';
		}
		buf.resize((size_t)(q - buf.data()));
	};
	if (n_thr == 1) work(0);
	else {
		std::vector<std::thread> th;
		for (int t = 0; t < n_thr; t++) th.emplace_back(work, t);
		for (auto &x : th) x.join();
	}
	unsigned char *q = out;
	memcpy(q, h.data(), h.size()); q += h.size();
	for (int t = 0; t < n_thr; t++) { memcpy(q, part[t].data(), part[t].size()); q += part[t].size(); }
	return (long long)(q - out);
}

// ======================================================================================================
// TransferServer chunking
// ======================================================================================================
static int kChunkLimit = 65000 - 3;         // TransferServer.cs:181,205 (ls3d_set_transfer_chunk_limit lets tests shrink it)
extern "C" int ls3d_set_transfer_chunk_limit(int limit) { if (limit < 1) { set_error("ls3d_set_transfer_chunk_limit: limit must be positive"); return -1; } kChunkLimit = limit; return 0; }

// formMeshChunks (TransferServer.cs:203-271) walks the index list once: a vertex met for the first time SINCE THE CURRENT CHUNK BEGAN
// is appended to the new vertex list and gets the next local index; a chunk ends at the first triangle end where it holds
// >= 64 997 vertices.  Parallel form: prev[i] = the previous position holding the same vertex index (independent of chunking), so
// "first in the chunk that starts at s" is prev[i] < s; a chunk's end is found by a scan of those flags from s on, its local
// indices are the scan values (repeat occurrences follow prev[] back to the first one inside the chunk).  Chunks are resolved
// one after the other (each needs the previous end), every chunk with grid-wide kernels.
__global__ void __launch_bounds__(256) k_occ_count(const int *__restrict__ tri, long long m, int n_vertices, unsigned *__restrict__ cnt, int *err) {
	pdl_enter();
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < m; i += 256ll * gridDim.x) {
		const int v = __ldg(tri + i);
		if (v < 0 || v >= n_vertices) { atomicOr(err, 1); continue; }
		atomicAdd(cnt + v, 1u);
	}
}
// exclusive scan of cnt[0..n) into start[0..n] by a single block (n <= a few million: tens of microseconds)
__global__ void __launch_bounds__(1024) k_scan_single(const unsigned *__restrict__ cnt, unsigned *__restrict__ start, long long n) {
	pdl_enter();
	__shared__ unsigned s_w[32];
	__shared__ unsigned s_carry;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_carry = 0;
	__syncthreads();
	for (long long base = 0; base < n; base += 1024) {
		const long long i = base + tid;
		const unsigned v = i < n ? cnt[i] : 0u;
		const unsigned incl = warp_incl_scan(v, lane);
		if (lane == 31) s_w[warp] = incl;
		__syncthreads();
		if (warp == 0) { const unsigned x = s_w[lane]; const unsigned s = warp_incl_scan(x, lane); s_w[lane] = s - x; }
		__syncthreads();
		const unsigned carry = s_carry;
		if (i < n) start[i] = carry + s_w[warp] + incl - v;
		__syncthreads();
		if (tid == 1023) s_carry = carry + s_w[warp] + incl;
		__syncthreads();
	}
	if (tid == 0) start[n] = s_carry;
}
// the same exclusive scan, grid-wide (decoupled look-back over 2048-element tiles, tiles handed out by ticket): the single-block
// version above needs ~1 us per 1024 elements, i.e. 0.7 ms for the 739 k vertices of an 8-sensor mesh
__global__ void __launch_bounds__(kScanThreads) k_scan_excl(const unsigned *__restrict__ cnt, unsigned *__restrict__ start, long long n, unsigned long long *status, unsigned *ticket, int *err) {
	pdl_enter();
	__shared__ unsigned sm[16];
	__shared__ int s_tile;
	const int tid = threadIdx.x;
	const int ntiles = (int)((n + kTile - 1) / kTile);
	for (;;) {
		if (tid == 0) s_tile = (int)atomicAdd(ticket, 1u);
		__syncthreads();
		const int tile = s_tile;
		if (tile >= ntiles) break;
		const long long p0 = (long long)tile * kTile + tid * 8;
		unsigned v[8], sum = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) { v[j] = p0 + j < n ? cnt[p0 + j] : 0u; sum += v[j]; }
		unsigned total, base;
		unsigned run = tile_scan(sum, sm, status, tile, err, &total, &base) + base;
#pragma unroll
		for (int j = 0; j < 8; j++)
			if (p0 + j < n) { start[p0 + j] = run; run += v[j]; }
		if (tile == ntiles - 1 && tid == 0) start[n] = base + total;
		__syncthreads();
	}
}
__global__ void __launch_bounds__(256) k_occ_fill(const int *__restrict__ tri, long long m, int n_vertices, const unsigned *__restrict__ start, unsigned *__restrict__ cursor, unsigned *__restrict__ pos) {
	pdl_enter();
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < m; i += 256ll * gridDim.x) {
		const int v = __ldg(tri + i);
		if (v < 0 || v >= n_vertices) continue;
		pos[start[v] + atomicAdd(cursor + v, 1u)] = (unsigned)i;
	}
}
// per vertex: sort its (short) occurrence list, then link each occurrence to the one before it
__global__ void __launch_bounds__(256) k_occ_link(int n_vertices, const unsigned *__restrict__ start, unsigned *__restrict__ pos, int *__restrict__ prev) {
	pdl_enter();
	for (int v = blockIdx.x * 256 + threadIdx.x; v < n_vertices; v += 256 * gridDim.x) {
		const unsigned a = start[v], b = start[v + 1];
		for (unsigned i = a + 1; i < b; i++) {          // insertion sort: a grid mesh vertex has at most 6 occurrences
			const unsigned x = pos[i];
			unsigned j = i;
			while (j > a && pos[j - 1] > x) { pos[j] = pos[j - 1]; j--; }
			pos[j] = x;
		}
		for (unsigned i = a; i < b; i++) prev[pos[i]] = i == a ? -1 : (int)pos[i - 1];
	}
}

struct ChunkCtl {
	unsigned tile_counter;
	int err;
	long long end;          // first triangle-end position (inclusive) at which the chunk holds >= limit vertices, or m-1
	unsigned total_new;     // new vertices in [s, end]
	unsigned pad;
	// device-driven loop (k_chunk_begin / k_chunk_close): the host only looks at these once per batch of chunks
	long long s;            // start of the chunk being resolved
	long long emit_s, emit_e;       // the chunk just resolved, for k_chunk_emit
	long long tri_chunk_start;      // TransferServer.cs:250
	unsigned vbase, emit_vbase;     // new vertices emitted before this chunk
	int n_chunks, done, overflow, cap;
};

// One chunk: positions [s, s + span) in tiles of 2048; flag = prev < s; device-wide inclusive scan (decoupled look-back); the chunk's
// end = min position with (pos+1)%3==0 and scan >= limit (atomicMin); local index of every position written optimistically — positions
// beyond the end are simply rewritten by the next chunk.
__global__ void __launch_bounds__(kScanThreads) k_chunk_scan(const int *__restrict__ prev, long long m, long long s, long long span,
	unsigned long long *status, ChunkCtl *ctl, unsigned *__restrict__ incl_out, int limit)
{
	pdl_enter();
	__shared__ unsigned sm[16];
	__shared__ int s_tile;
	const int tid = threadIdx.x;
	if (s < 0) {                       // device-driven loop: the chunk start lives in ctl, span is the window limit
		if (ctl->done) return;
		s = ctl->s;
		span = min(span, m - s);
	}
	const int ntiles = (int)((span + kTile - 1) / kTile);
	for (;;) {
		if (tid == 0) s_tile = (int)atomicAdd(&ctl->tile_counter, 1u);
		__syncthreads();
		const int tile = s_tile;
		if (tile >= ntiles) break;
		const long long p0 = s + (long long)tile * kTile + tid * 8;
		unsigned f = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const long long p = p0 + j;
			if (p < s + span && p < m && (long long)__ldg(prev + p) < s) f |= 1u << j;
		}
		unsigned total, base;
		const unsigned off = tile_scan(__popc(f), sm, status, tile, &ctl->err, &total, &base);
		unsigned run = base + off;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const long long p = p0 + j;
			if (p >= s + span || p >= m) break;
			run += (f >> j) & 1u;
			incl_out[p - s] = run;
			if (run >= (unsigned)limit && (p + 1) % 3 == 0) { atomicMin((unsigned long long *)&ctl->end, (unsigned long long)p); break; }
		}
		__syncthreads();
	}
}

// Emit one resolved chunk [s, e]: new triangle indices (chunk-local) and the vertex copies in first-occurrence order.
__global__ void __launch_bounds__(256) k_chunk_emit(const int *__restrict__ tri, const int *__restrict__ prev, long long s, long long e,
	const unsigned *__restrict__ incl, const uint4 *__restrict__ verts, unsigned vbase, int *__restrict__ new_tri, uint4 *__restrict__ new_verts, const ChunkCtl *ctl)
{
	pdl_enter();
	if (ctl) {                         // device-driven loop: the range k_chunk_close resolved
		if (ctl->done) return;
		s = ctl->emit_s; e = ctl->emit_e; vbase = ctl->emit_vbase;
	}
	for (long long p = s + blockIdx.x * 256ll + threadIdx.x; p <= e; p += 256ll * gridDim.x) {
		long long q = p;
		int pv = __ldg(prev + q);
		const bool first = pv < s;
		while (pv >= s) { q = pv; pv = __ldg(prev + q); }      // first occurrence inside the chunk
		const unsigned local = incl[q - s] - 1u;
		new_tri[p] = (int)local;
		if (first) new_verts[vbase + local] = verts[__ldg(tri + p)];
	}
}

// The chunk loop without the host: k_chunk_begin arms the scan of the chunk that starts at ctl->s (or marks the loop done), k_chunk_close
// turns the scan's result into the chunk's sizes, the range k_chunk_emit writes, and the next start.  A chunk longer than the scan
// window raises `overflow` and the host loop below takes over (its window grows on demand).
__global__ void __launch_bounds__(256) k_chunk_begin(ChunkCtl *ctl, unsigned long long *status, long long m, long long span_max) {
	pdl_enter();
	if (ctl->done) return;
	const long long s = ctl->s;
	if (s >= m) { if (blockIdx.x == 0 && threadIdx.x == 0) ctl->done = 1; return; }
	const long long tiles = (min(span_max, m - s) + kTile - 1) / kTile;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < tiles; i += 256ll * gridDim.x) status[i] = 0ull;
	if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->tile_counter = 0; ctl->end = 0x7fffffffffffffffll; }
}
__global__ void k_chunk_close(ChunkCtl *ctl, const unsigned *__restrict__ incl, long long m, long long span_max, int *chunk_v, int *chunk_t) {
	pdl_enter();
	if (ctl->done) return;
	const long long s = ctl->s, span = min(span_max, m - s);
	const bool closed = ctl->end != 0x7fffffffffffffffll;
	if ((!closed && s + span < m) || ctl->n_chunks >= ctl->cap) { ctl->overflow = 1; ctl->done = 1; return; }
	const long long e = closed ? ctl->end : m - 1;
	const unsigned n_new = incl[e - s];
	const int c = ctl->n_chunks;
	chunk_v[c] = (int)n_new;
	chunk_t[c] = (int)(((closed ? e : m) - ctl->tri_chunk_start) / 3);       // TransferServer.cs:246-251,256-260 (see the host loop)
	if (closed) ctl->tri_chunk_start = e;
	ctl->emit_s = s; ctl->emit_e = e; ctl->emit_vbase = ctl->vbase;
	ctl->vbase += n_new;
	ctl->s = e + 1;
	ctl->n_chunks = c + 1;
}

extern "C" long long ls3d_transfer_frame_size(int n_vertices, int n_triangles, int n_chunks) {
	if (n_vertices < 0 || n_triangles < 0 || n_chunks < 0) return -1;
	return 12ll + 8ll * n_chunks + 15ll * n_vertices + 12ll * n_triangles;
}

// Chunk a device-resident mesh (formVerticesChunks when n_triangles == 0, formMeshChunks otherwise) and leave the frame body
// (xyz | rgb | triangles) in d_body.  Host outputs: sizes of every chunk, totals.  Returns the number of chunks or -1.
static int transfer_chunk_device(const uint4 *d_verts, int n_vertices, const int *d_tri, int n_triangles, std::vector<int> &v_sizes, std::vector<int> &t_sizes,
	int *out_vertices, DevBuf &body, cudaStream_t st)
{
	v_sizes.clear(); t_sizes.clear();
	if (n_triangles == 0) {
		for (int cur = 0; cur < n_vertices; cur += kChunkLimit) { v_sizes.push_back(std::min(kChunkLimit, n_vertices - cur)); t_sizes.push_back(0); }    // TransferServer.cs:179-201
		*out_vertices = n_vertices;
		if (!body.reserve(15 * (size_t)std::max(n_vertices, 1) + 16, "alloc frame body")) return -1;
		if (n_vertices && ls3d_pack_transfer_body_device(d_verts, n_vertices, nullptr, 0, body.p, st) < 0) return -1;
		return (int)v_sizes.size();
	}
	const long long m = 3ll * n_triangles;
	// scratch: cnt/cursor[n+1] start[n+1] pos[m] prev[m] incl[m] new_tri[m] new_verts[m] status ctl
	const size_t nv1 = (size_t)n_vertices + 1;
	const int max_tiles = (int)((std::max<long long>(m, (long long)nv1) + kTile - 1) / kTile) + 1;
	size_t off = 0;
	auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
	const size_t o_cnt = carve(4 * nv1), o_cur = carve(4 * nv1), o_start = carve(4 * nv1), o_pos = carve(4 * (size_t)m), o_prev = carve(4 * (size_t)m), o_incl = carve(4 * (size_t)m),
		o_ntri = carve(4 * (size_t)m), o_nv = carve(16 * (size_t)m), o_status = carve(8 * (size_t)max_tiles), o_ctl = carve(sizeof(ChunkCtl));
	if (!g_fmt_aux.reserve(off, "alloc chunking scratch")) return -1;
	uint8_t *b = g_fmt_aux.as<uint8_t>();
	unsigned *cnt = (unsigned *)(b + o_cnt), *cur = (unsigned *)(b + o_cur), *start = (unsigned *)(b + o_start), *pos = (unsigned *)(b + o_pos), *incl = (unsigned *)(b + o_incl);
	int *prev = (int *)(b + o_prev), *new_tri = (int *)(b + o_ntri);
	uint4 *new_verts = (uint4 *)(b + o_nv);
	unsigned long long *status = (unsigned long long *)(b + o_status);
	ChunkCtl *ctl = (ChunkCtl *)(b + o_ctl);
	const int grid = pack_grid(4 * m);
	if (!cuda_ok(cudaMemsetAsync(b + o_cnt, 0, o_start - o_cnt, st), "clear counters") || !cuda_ok(cudaMemsetAsync(ctl, 0, sizeof(ChunkCtl), st), "clear control")) return -1;
	k_occ_count<<<grid, 256, 0, st>>>(d_tri, m, n_vertices, cnt, &ctl->err);
	if (n_vertices <= 8192) {
		launch_chain(true, k_scan_single, dim3((unsigned)(1)), 1024, 0, st, cnt, start, n_vertices);
	} else {
		const int tiles = (int)(((long long)n_vertices + kTile - 1) / kTile);
		if (!cuda_ok(cudaMemsetAsync(status, 0, 8 * (size_t)tiles, st), "clear scan status")) return -1;
		launch_chain(true, k_scan_excl, dim3((unsigned)(std::max(1, std::min(tiles, 148 * 8)))), kScanThreads, 0, st, cnt, start, n_vertices, status, &ctl->tile_counter, &ctl->err);
	}
	launch_chain(true, k_occ_fill, dim3((unsigned)(grid)), 256, 0, st, d_tri, m, n_vertices, start, cur, pos);
	launch_chain(true, k_occ_link, dim3((unsigned)(pack_grid(16ll * n_vertices))), 256, 0, st, n_vertices, start, pos, prev);
	count_launch(4);
	if (!cuda_ok(cudaGetLastError(), "occurrence lists")) return -1;
	ChunkCtl h;
	if (!cuda_ok(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st), "read control") || !cuda_ok(cudaStreamSynchronize(st), "occurrence lists")) return -1;
	if (h.err) { set_error("triangle index outside 0..%d", n_vertices - 1); return -1; }
	long long s = 0, tri_chunk_start = 0;
	unsigned vbase = 0;
	static const int env_host_loop = getenv("LS3D_CHUNK_HOST_LOOP") ? atoi(getenv("LS3D_CHUNK_HOST_LOOP")) : 0;
	if (!env_host_loop) {
		// device-driven: batches of 16 chunks (begin, scan, close, emit each) per host wait; kernels of chunks past the end return at once
		const long long span_max = 6ll * kChunkLimit + 3;
		const int cap = (int)std::min<long long>(m / std::max(kChunkLimit, 1) + 2, 1 << 24);
		static DevBuf sizes;
		static int *pin = nullptr;
		static size_t pin_cap = 0;
		const size_t need = sizeof(ChunkCtl) + 8 * (size_t)cap;
		if (pin_cap < need) { if (pin) cudaFreeHost(pin); pin = nullptr; pin_cap = 0; if (cudaHostAlloc((void **)&pin, need, cudaHostAllocDefault) == cudaSuccess) pin_cap = need; else cudaGetLastError(); }
		if (pin && sizes.reserve(8 * (size_t)cap, "alloc chunk sizes")) {
			int *cv = sizes.as<int>(), *ct = cv + cap;
			ChunkCtl init; memset(&init, 0, sizeof(init)); init.end = 0x7fffffffffffffffll; init.cap = cap;
			memcpy(pin, &init, sizeof(init));
			bool ok = cuda_ok(cudaMemcpyAsync(ctl, pin, sizeof(ChunkCtl), cudaMemcpyHostToDevice, st), "reset control");
			const int scan_grid = std::max(1, std::min((int)((std::min(span_max, m) + kTile - 1) / kTile), 148 * 8));
			const ChunkCtl *hc = reinterpret_cast<const ChunkCtl *>(pin);
			for (int batch = 0; ok; batch++) {
				for (int i = 0; i < 16; i++) {
					launch_chain(true, k_chunk_begin, dim3((unsigned)(32)), 256, 0, st, ctl, status, m, span_max);
					launch_chain(true, k_chunk_scan, dim3((unsigned)(scan_grid)), kScanThreads, 0, st, prev, m, -1, span_max, status, ctl, incl, kChunkLimit);
					launch_chain(true, k_chunk_close, dim3((unsigned)(1)), 1, 0, st, ctl, incl, m, span_max, cv, ct);
					launch_chain(true, k_chunk_emit, dim3((unsigned)(pack_grid(4 * std::min(span_max, m)))), 256, 0, st, d_tri, prev, 0, -1, incl, d_verts, 0u, new_tri, new_verts, ctl);
				}
				count_launch(64);
				ok = cuda_ok(cudaGetLastError(), "chunk loop") && cuda_ok(cudaMemcpyAsync(pin, ctl, sizeof(ChunkCtl), cudaMemcpyDeviceToHost, st), "read control") &&
					cuda_ok(cudaStreamSynchronize(st), "chunk loop");
				if (!ok) return -1;
				if (hc->err) { set_error("device reported error flags 0x%x while chunking", hc->err); return -1; }
				if (hc->done || hc->s >= m) break;
			}
			if (!hc->overflow) {
				const int nc = hc->n_chunks;
				if (nc > 0 && (!cuda_ok(cudaMemcpyAsync(pin + sizeof(ChunkCtl) / 4, cv, 4 * (size_t)nc, cudaMemcpyDeviceToHost, st), "read chunk sizes") ||
					!cuda_ok(cudaMemcpyAsync(pin + sizeof(ChunkCtl) / 4 + cap, ct, 4 * (size_t)nc, cudaMemcpyDeviceToHost, st), "read chunk sizes") ||
					!cuda_ok(cudaStreamSynchronize(st), "chunk sizes"))) return -1;
				const int *hv = pin + sizeof(ChunkCtl) / 4, *ht = hv + cap;
				v_sizes.assign(hv, hv + nc);
				t_sizes.assign(ht, ht + nc);
				vbase = hc->vbase;
				s = m;                       // the host loop below has nothing left to do
			}
			// overflow: a chunk did not fit the window — start over with the host loop, whose window grows
		}
	}
	while (s < m) {
		// a chunk of L vertices spans at least L positions; grow the window until the end is inside it
		long long span = std::min<long long>(m - s, 6ll * kChunkLimit + 3);
		for (;;) {
			const int tiles = (int)((span + kTile - 1) / kTile);
			ChunkCtl init; memset(&init, 0, sizeof(init)); init.end = 0x7fffffffffffffffll;
			if (!cuda_ok(cudaMemcpyAsync(ctl, &init, sizeof(init), cudaMemcpyHostToDevice, st), "reset control") ||
				!cuda_ok(cudaMemsetAsync(status, 0, 8 * (size_t)tiles, st), "clear scan status")) return -1;
			launch_chain(true, k_chunk_scan, dim3((unsigned)(std::max(1, std::min(tiles, 148 * 8)))), kScanThreads, 0, st, prev, m, s, span, status, ctl, incl, kChunkLimit);
			count_launch(1);
			if (!cuda_ok(cudaGetLastError(), "k_chunk_scan") || !cuda_ok(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st), "read control") ||
				!cuda_ok(cudaStreamSynchronize(st), "chunk scan")) return -1;
			if (h.err) { set_error("device reported error flags 0x%x while chunking", h.err); return -1; }
			if (h.end != 0x7fffffffffffffffll || s + span >= m) break;
			span = std::min<long long>(m - s, span * 2);
		}
		const bool closed = h.end != 0x7fffffffffffffffll;
		const long long e = closed ? h.end : m - 1;
		unsigned n_new = 0;
		if (!cuda_ok(cudaMemcpyAsync(&n_new, incl + (e - s), 4, cudaMemcpyDeviceToHost, st), "read chunk size")) return -1;
		launch_chain(true, k_chunk_emit, dim3((unsigned)(pack_grid(4 * (e - s + 1)))), 256, 0, st, d_tri, prev, s, e, incl, d_verts, vbase, new_tri, new_verts, nullptr);
		count_launch(1);
		if (!cuda_ok(cudaGetLastError(), "k_chunk_emit") || !cuda_ok(cudaStreamSynchronize(st), "chunk emit")) return -1;
		v_sizes.push_back((int)n_new);
		// TransferServer.cs:246-251,256-260: the triangle count is measured from the previous chunk's LAST index position (not the one
		// after it), so the first of several chunks reports one triangle too few — kept, it is what the receiver is sent
		t_sizes.push_back((int)(((closed ? e : m) - tri_chunk_start) / 3));
		if (closed) tri_chunk_start = e;
		vbase += n_new;
		s = e + 1;
	}
	*out_vertices = (int)vbase;
	if (!body.reserve(15 * (size_t)std::max<unsigned>(vbase, 1) + 12 * (size_t)n_triangles + 16, "alloc frame body")) return -1;
	if (ls3d_pack_transfer_body_device(new_verts, (int)vbase, new_tri, n_triangles, body.p, st) < 0) return -1;
	return (int)v_sizes.size();
}

extern "C" int ls3d_transfer_chunks_device(const void *d_vertices, int n_vertices, const int *d_triangles, int n_triangles,
	int *chunk_vertices, int *chunk_triangles, int chunk_cap, int *n_vertices_out, const void **d_body, void *stream)
{
	clear_error();
	if (n_vertices < 0 || n_triangles < 0 || (n_vertices && !d_vertices) || (n_triangles && !d_triangles) || !n_vertices_out || !d_body) { set_error("ls3d_transfer_chunks_device: bad argument"); return -1; }
	if (n_triangles > 0 && n_vertices == 0) { set_error("ls3d_transfer_chunks_device: triangles without vertices"); return -1; }
	if (!ensure_device()) return -1;
	std::lock_guard<std::mutex> lk(api_mutex());
	std::vector<int> vs, ts;
	const int chunks = transfer_chunk_device((const uint4 *)d_vertices, n_vertices, d_triangles, n_triangles, vs, ts, n_vertices_out, g_fmt_out, (cudaStream_t)stream);
	if (chunks < 0) return -1;
	if (chunks > chunk_cap) { set_error("ls3d_transfer_chunks_device: %d chunks, room for %d", chunks, chunk_cap); return -1; }
	for (int i = 0; i < chunks; i++) { if (chunk_vertices) chunk_vertices[i] = vs[i]; if (chunk_triangles) chunk_triangles[i] = ts[i]; }
	*d_body = g_fmt_out.p;
	return chunks;
}

extern "C" long long ls3d_write_transfer_frame(const VertexC4ubV3f *vertices, int n_vertices, const int *triangles, int n_triangles, unsigned char *out, long long out_cap) {
	clear_error();
	if (n_vertices < 0 || n_triangles < 0 || (n_vertices && !vertices) || (n_triangles && !triangles)) { set_error("ls3d_write_transfer_frame: bad argument"); return -1; }
	if (n_triangles > 0 && n_vertices == 0) { set_error("ls3d_write_transfer_frame: triangles without vertices"); return -1; }
	if (!ensure_device()) return -1;
	std::lock_guard<std::mutex> lk(api_mutex());
	cudaStream_t st = api_stream();
	if (!st) return -1;
	if (!g_fmt_in.reserve(16 * (size_t)std::max(n_vertices, 1), "alloc vertices") || !g_fmt_tri.reserve(12 * (size_t)std::max(n_triangles, 1), "alloc triangles")) return -1;
	if ((n_vertices && !cuda_ok(cudaMemcpyAsync(g_fmt_in.p, vertices, 16 * (size_t)n_vertices, cudaMemcpyHostToDevice, st), "upload vertices")) ||
		(n_triangles && !cuda_ok(cudaMemcpyAsync(g_fmt_tri.p, triangles, 12 * (size_t)n_triangles, cudaMemcpyHostToDevice, st), "upload triangles"))) return -1;
	std::vector<int> vs, ts;
	int nv_out = 0;
	const int chunks = transfer_chunk_device(g_fmt_in.as<uint4>(), n_vertices, g_fmt_tri.as<int>(), n_triangles, vs, ts, &nv_out, g_fmt_out, st);
	if (chunks < 0) { cudaStreamSynchronize(st); return -1; }
	const long long total = ls3d_transfer_frame_size(nv_out, n_triangles, chunks);
	if (!out) { cudaStreamSynchronize(st); return total; }
	if (out_cap < total) { cudaStreamSynchronize(st); set_error("ls3d_write_transfer_frame: need %lld bytes of output, got %lld", total, out_cap); return -1; }
	wr_i32(out, nv_out);                               // TransferSocket.cs:92-96
	wr_i32(out + 4, n_triangles);
	wr_i32(out + 8, chunks);
	if (chunks) { memcpy(out + 12, vs.data(), 4 * (size_t)chunks); memcpy(out + 12 + 4 * (size_t)chunks, ts.data(), 4 * (size_t)chunks); }
	const long long body = 15ll * nv_out + 12ll * n_triangles;
	bool ok = body == 0 || cuda_ok(cudaMemcpyAsync(out + 12 + 8 * (size_t)chunks, g_fmt_out.p, (size_t)body, cudaMemcpyDeviceToHost, st), "read frame body");
	ok = cuda_ok(cudaStreamSynchronize(st), "ls3d_write_transfer_frame") && ok;
	return ok ? total : -1;
}
