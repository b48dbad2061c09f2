// pcie_probe.cu — what the host link gives the end-to-end frame path: copy-engine transfers against SM loads / stores on
// mapped page-locked memory, alone and together.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/pcie_probe scripts/pcie_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

// every `stride`-th 128-byte line of src (16 bytes per lane, 8 lanes per line) -> dst (device)
__global__ void k_pull_lines(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n_lines, int stride) {
	const size_t lane8 = threadIdx.x & 7;
	for (size_t l = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; l * stride < n_lines; l += ((size_t)gridDim.x * blockDim.x) >> 3) {
		const size_t i = l * stride * 8 + lane8;
		dst[i] = __ldg(src + i);
	}
}
// one 32-byte sector out of every `stride` sectors, 8 bytes per lane (4 lanes per sector)
__global__ void k_pull_sectors(const uint2 *__restrict__ src, uint2 *__restrict__ dst, size_t n_sectors, int stride) {
	const size_t lane4 = threadIdx.x & 3;
	for (size_t s = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2; s * stride < n_sectors; s += ((size_t)gridDim.x * blockDim.x) >> 2) {
		const size_t i = s * stride * 4 + lane4;
		dst[i] = __ldg(src + i);
	}
}
__global__ void k_push(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// device-only traffic: a stand-in for a kernel that should not care about the host link
__global__ void k_device_work(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n, int reps) {
	for (int r = 0; r < reps; r++)
		for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
			uint4 v = src[i];
			v.x += r;
			dst[i] = v;
		}
}

template <typename F> static float timed(cudaStream_t st, int reps, F f) {
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	std::vector<float> ts;
	for (int i = 0; i < reps + 2; i++) {
		CK(cudaEventRecord(a, st));
		f();
		CK(cudaEventRecord(b, st));
		CK(cudaEventSynchronize(b));
		float ms;
		CK(cudaEventElapsedTime(&ms, a, b));
		if (i >= 2) ts.push_back(ms);
	}
	std::sort(ts.begin(), ts.end());
	CK(cudaEventDestroy(a)); CK(cudaEventDestroy(b));
	return ts[ts.size() / 2];
}

// `pcie_probe sustain <mode> <seconds>`: mode h2d | d2h | both — frame-sized copies back to back for that long; prints this process's
// rates.  Run one per GPU at the same time to see what the host side gives N devices together.
static int sustain(const char *mode, double seconds) {
	const size_t up = 3473408 + 1244192, down = 5004752;
	uint8_t *h_in, *h_out, *d_a, *d_b;
	CK(cudaHostAlloc((void **)&h_in, 8 << 20, cudaHostAllocDefault)); CK(cudaHostAlloc((void **)&h_out, 8 << 20, cudaHostAllocDefault));
	memset(h_in, 1, 8 << 20); memset(h_out, 0, 8 << 20);
	CK(cudaMalloc((void **)&d_a, 8 << 20)); CK(cudaMalloc((void **)&d_b, 8 << 20));
	cudaStream_t s0, s1;
	CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
	const bool do_up = strcmp(mode, "d2h") != 0, do_down = strcmp(mode, "h2d") != 0;
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	size_t n = 0;
	float total_ms = 0;
	while (total_ms < seconds * 1000) {
		CK(cudaEventRecord(a, s0));
		CK(cudaStreamWaitEvent(s1, a, 0));
		for (int i = 0; i < 50; i++) {
			if (do_up) CK(cudaMemcpyAsync(d_a, h_in, up, cudaMemcpyHostToDevice, s0));
			if (do_down) CK(cudaMemcpyAsync(h_out, d_b, down, cudaMemcpyDeviceToHost, s1));
		}
		CK(cudaEventRecord(b, s1));
		CK(cudaStreamWaitEvent(s0, b, 0));
		CK(cudaEventRecord(b, s0));
		CK(cudaEventSynchronize(b));
		float ms;
		CK(cudaEventElapsedTime(&ms, a, b));
		total_ms += ms;
		n += 50;
	}
	int dev = 0;
	CK(cudaGetDevice(&dev));
	const char *vis = getenv("CUDA_VISIBLE_DEVICES");
	printf("sustain %s gpu %s: %zu frames in %.0f ms: H2D %.1f GB/s  D2H %.1f GB/s  (%.0f frame-equivalents/s)\n", mode, vis ? vis : "?", n, total_ms,
		do_up ? up * n / total_ms / 1e6 : 0.0, do_down ? down * n / total_ms / 1e6 : 0.0, n / total_ms * 1000);
	return 0;
}

int main(int argc, char **argv) {
	if (argc >= 4 && !strcmp(argv[1], "sustain")) return sustain(argv[2], atof(argv[3]));
	const size_t bytes = 5210112;          // the bench frame's colour images
	const size_t out_bytes = 5004752;      // its merged cloud
	uint8_t *h_in, *h_out;
	CK(cudaHostAlloc((void **)&h_in, 16 << 20, cudaHostAllocMapped));
	CK(cudaHostAlloc((void **)&h_out, 16 << 20, cudaHostAllocMapped));
	memset(h_in, 1, 16 << 20);
	memset(h_out, 0, 16 << 20);
	uint8_t *d_a, *d_b, *d_c, *d_d;
	CK(cudaMalloc((void **)&d_a, 16 << 20)); CK(cudaMalloc((void **)&d_b, 16 << 20));
	CK(cudaMalloc((void **)&d_c, 64 << 20)); CK(cudaMalloc((void **)&d_d, 64 << 20));
	void *hd_in, *hd_out;
	CK(cudaHostGetDevicePointer(&hd_in, h_in, 0)); CK(cudaHostGetDevicePointer(&hd_out, h_out, 0));
	cudaStream_t s0, s1, s2;
	CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
	const int R = 30;
	auto gbs = [](size_t b, float ms) { return b / ms / 1e6; };

	float t = timed(s0, R, [&] { CK(cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s0)); });
	printf("copy engine H2D %zu B: %.1f us  %.1f GB/s\n", bytes, t * 1000, gbs(bytes, t));
	t = timed(s0, R, [&] { CK(cudaMemcpyAsync(d_a, h_in, bytes / 8, cudaMemcpyHostToDevice, s0)); });
	printf("copy engine H2D %zu B: %.1f us  %.1f GB/s\n", bytes / 8, t * 1000, gbs(bytes / 8, t));
	t = timed(s0, R, [&] { CK(cudaMemcpyAsync(h_out, d_a, out_bytes, cudaMemcpyDeviceToHost, s0)); });
	printf("copy engine D2H %zu B: %.1f us  %.1f GB/s\n", out_bytes, t * 1000, gbs(out_bytes, t));
	t = timed(s0, R, [&] { CK(cudaMemcpyAsync(h_out, d_a, out_bytes / 4, cudaMemcpyDeviceToHost, s0)); });
	printf("copy engine D2H %zu B: %.1f us  %.1f GB/s\n", out_bytes / 4, t * 1000, gbs(out_bytes / 4, t));

	const size_t n_lines = bytes / 128, n_sect = bytes / 32;
	for (int blocks : {16, 74, 148, 592, 1184}) {
		for (int stride : {1, 4}) {
			t = timed(s0, R, [&] { k_pull_lines<<<blocks, 256, 0, s0>>>((const uint4 *)hd_in, (uint4 *)d_a, n_lines, stride); });
			printf("SM pull, 128-byte lines, every %d-th, %4d blocks: %.1f us  %.1f GB/s (%zu B)\n", stride, blocks, t * 1000, gbs(bytes / stride, t), bytes / stride);
		}
		t = timed(s0, R, [&] { k_pull_sectors<<<blocks, 256, 0, s0>>>((const uint2 *)hd_in, (uint2 *)d_a, n_sect, 4); });
		printf("SM pull, 32-byte sectors, every 4th, %4d blocks: %.1f us  %.1f GB/s (%zu B)\n", blocks, t * 1000, gbs(bytes / 4, t), bytes / 4);
	}
	for (int blocks : {8, 16, 64, 148, 592}) {
		t = timed(s0, R, [&] { k_push<<<blocks, 256, 0, s0>>>((const uint4 *)d_a, (uint4 *)hd_out, out_bytes / 16); });
		printf("SM push (512 B per warp store), %4d blocks: %.1f us  %.1f GB/s\n", blocks, t * 1000, gbs(out_bytes, t));
	}

	// together: H2D by the copy engine while SMs pull lines / push records
	cudaEvent_t e0, e1, e2, e3, go;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3)); CK(cudaEventCreate(&go));
	auto both = [&](const char *what, auto fa, auto fb) {
		std::vector<float> ta, tb;
		for (int i = 0; i < R + 2; i++) {
			CK(cudaEventRecord(go, s2));
			CK(cudaStreamWaitEvent(s0, go, 0)); CK(cudaStreamWaitEvent(s1, go, 0));
			CK(cudaEventRecord(e0, s0)); fa(); CK(cudaEventRecord(e1, s0));
			CK(cudaEventRecord(e2, s1)); fb(); CK(cudaEventRecord(e3, s1));
			CK(cudaDeviceSynchronize());
			float a, b;
			CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e2, e3));
			if (i >= 2) { ta.push_back(a); tb.push_back(b); }
		}
		std::sort(ta.begin(), ta.end()); std::sort(tb.begin(), tb.end());
		printf("%s: first %.1f us, second %.1f us\n", what, ta[R / 2] * 1000, tb[R / 2] * 1000);
	};
	const size_t depth = 3473408;
	both("CE H2D 3.47 MB  +  SM pull of every 4th line (1.3 MB), 592 blocks",
		[&] { CK(cudaMemcpyAsync(d_b, h_in + (8 << 20), depth, cudaMemcpyHostToDevice, s0)); },
		[&] { k_pull_lines<<<592, 256, 0, s1>>>((const uint4 *)hd_in, (uint4 *)d_a, n_lines, 4); });
	both("CE H2D 3.47 MB  +  SM pull of every 4th sector (1.3 MB), 592 blocks",
		[&] { CK(cudaMemcpyAsync(d_b, h_in + (8 << 20), depth, cudaMemcpyHostToDevice, s0)); },
		[&] { k_pull_sectors<<<592, 256, 0, s1>>>((const uint2 *)hd_in, (uint2 *)d_a, n_sect, 4); });
	both("CE H2D 3.47 MB  +  CE H2D 1.3 MB",
		[&] { CK(cudaMemcpyAsync(d_b, h_in + (8 << 20), depth, cudaMemcpyHostToDevice, s0)); },
		[&] { CK(cudaMemcpyAsync(d_a, h_in, bytes / 4, cudaMemcpyHostToDevice, s1)); });
	both("CE H2D 3.47 MB  +  SM push 5.0 MB, 16 blocks",
		[&] { CK(cudaMemcpyAsync(d_b, h_in + (8 << 20), depth, cudaMemcpyHostToDevice, s0)); },
		[&] { k_push<<<16, 256, 0, s1>>>((const uint4 *)d_a, (uint4 *)hd_out, out_bytes / 16); });
	both("CE H2D 3.47 MB  +  CE D2H 5.0 MB",
		[&] { CK(cudaMemcpyAsync(d_b, h_in + (8 << 20), depth, cudaMemcpyHostToDevice, s0)); },
		[&] { CK(cudaMemcpyAsync(h_out, d_a, out_bytes, cudaMemcpyDeviceToHost, s1)); });
	// does a device-only kernel slow down next to SM pulls?
	t = timed(s0, R, [&] { k_device_work<<<592, 256, 0, s0>>>((const uint4 *)d_c, (uint4 *)d_d, (32 << 20) / 16, 4); });
	printf("device-only kernel alone: %.1f us\n", t * 1000);
	both("device-only kernel  +  SM pull of every 4th sector, 592 blocks",
		[&] { k_device_work<<<592, 256, 0, s0>>>((const uint4 *)d_c, (uint4 *)d_d, (32 << 20) / 16, 4); },
		[&] { k_pull_sectors<<<592, 256, 0, s1>>>((const uint2 *)hd_in, (uint2 *)d_a, n_sect, 4); });
	both("device-only kernel  +  SM pull of every 4th line, 592 blocks",
		[&] { k_device_work<<<592, 256, 0, s0>>>((const uint4 *)d_c, (uint4 *)d_d, (32 << 20) / 16, 4); },
		[&] { k_pull_lines<<<592, 256, 0, s1>>>((const uint4 *)hd_in, (uint4 *)d_a, n_lines, 4); });
	both("device-only kernel  +  SM pull of every 4th line, 64 blocks",
		[&] { k_device_work<<<592, 256, 0, s0>>>((const uint4 *)d_c, (uint4 *)d_d, (32 << 20) / 16, 4); },
		[&] { k_pull_lines<<<64, 256, 0, s1>>>((const uint4 *)hd_in, (uint4 *)d_a, n_lines, 4); });
	both("device-only kernel  +  CE H2D 3.47 MB",
		[&] { k_device_work<<<592, 256, 0, s0>>>((const uint4 *)d_c, (uint4 *)d_d, (32 << 20) / 16, 4); },
		[&] { CK(cudaMemcpyAsync(d_b, h_in + (8 << 20), depth, cudaMemcpyHostToDevice, s1)); });
	return 0;
}
