// runtime.cu — error side channel, launch accounting, library stream, pinned host blocks, Mesh lifetime.
#include "ls3d_internal.h"
#include "../../include/ls3d.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <map>
#include <thread>
#include <vector>

namespace ls3d {

static thread_local char t_error[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(t_error, sizeof(t_error), fmt, ap);
	va_end(ap);
}
void clear_error() { t_error[0] = 0; }

bool cuda_ok(cudaError_t e, const char *what) {
	if (e == cudaSuccess) return true;
	set_error("%s: %s", what, cudaGetErrorString(e));
	return false;
}

std::mutex &api_mutex() {
	static std::mutex m;
	return m;
}

static std::mutex g_rt_mutex;
static cudaStream_t g_stream = nullptr;
static int g_device_state = 0;   // 0 unknown, 1 ok, -1 failed
static char g_version[256] = "";

bool ensure_device() {
	std::lock_guard<std::mutex> lk(g_rt_mutex);
	if (g_device_state == 1) return true;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		set_error("no CUDA device available (%s); libls3d_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
		cudaGetLastError();
		g_device_state = -1;
		return false;
	}
	int dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceProp prop;
	if (!cuda_ok(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties")) { g_device_state = -1; return false; }
	if (prop.major < 10) {
		set_error("device %s is compute capability %d.%d; this library carries sm_100a code only", prop.name, prop.major, prop.minor);
		g_device_state = -1;
		return false;
	}
	snprintf(g_version, sizeof(g_version), "ls3d-b200 0.1 sm_100a; device: %s (cc %d.%d, %d SMs)", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
	g_device_state = 1;
	return true;
}

cudaStream_t api_stream() {
	std::lock_guard<std::mutex> lk(g_rt_mutex);
	if (!g_stream) {
		if (!cuda_ok(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking), "cudaStreamCreate")) g_stream = nullptr;
	}
	return g_stream;
}

// ---- pinned host blocks -----------------------------------------------------------------------------
struct HostBlock { size_t cap; bool pinned; };
static std::mutex g_hb_mutex;
static std::map<void *, HostBlock> g_hb_live;                  // blocks currently owned by a caller
static std::multimap<size_t, std::pair<void *, bool>> g_hb_free;   // recycled blocks by capacity
static size_t g_hb_free_bytes = 0;

static size_t round_block(size_t b) {
	size_t c = 1 << 16;
	while (c < b) c <<= 1;
	return c;
}

void *host_block_alloc(size_t bytes) {
	const size_t cap = round_block(bytes ? bytes : 1);
	{
		std::lock_guard<std::mutex> lk(g_hb_mutex);
		auto it = g_hb_free.lower_bound(cap);
		if (it != g_hb_free.end() && it->first <= cap * 2) {
			void *p = it->second.first;
			g_hb_live[p] = HostBlock{it->first, it->second.second};
			g_hb_free_bytes -= it->first;
			g_hb_free.erase(it);
			return p;
		}
	}
	void *p = nullptr;
	bool pinned = true;
	if (cudaHostAlloc(&p, cap, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
		cudaGetLastError();
		p = malloc(cap);     // plain host memory is still valid output memory; the copy is just slower
		pinned = false;
		if (!p) return nullptr;
	}
	std::lock_guard<std::mutex> lk(g_hb_mutex);
	g_hb_live[p] = HostBlock{cap, pinned};
	return p;
}

bool host_block_is_pinned(void *p) {
	std::lock_guard<std::mutex> lk(g_hb_mutex);
	auto it = g_hb_live.find(p);
	return it != g_hb_live.end() && it->second.pinned;
}

void host_block_free(void *p) {
	if (!p) return;
	std::lock_guard<std::mutex> lk(g_hb_mutex);
	auto it = g_hb_live.find(p);
	if (it == g_hb_live.end()) { free(p); return; }    // not ours (e.g. a zero-length malloc): plain free
	HostBlock hb = it->second;
	g_hb_live.erase(it);
	if (g_hb_free_bytes + hb.cap <= (size_t)1 << 30) {   // keep at most 1 GiB parked
		g_hb_free.emplace(hb.cap, std::make_pair(p, hb.pinned));
		g_hb_free_bytes += hb.cap;
		return;
	}
	if (hb.pinned) cudaFreeHost(p); else free(p);
}

// ---- host copy pool -----------------------------------------------------------------------------------
// One job at a time (callers hold api_mutex()).  A job is cut into 256 KB slices handed out by an atomic counter; the workers
// and the caller pull slices until none is left.  After a job the workers keep polling for ~200 us before they go back to
// sleep on the condition variable, so back-to-back frames do not pay a wake-up each.
namespace {
struct CopyPool {
	std::vector<std::thread> workers;
	std::mutex m;
	std::condition_variable cv;
	std::atomic<unsigned long long> generation{0};
	std::atomic<bool> stop{false};
	// job description: atomics, because a worker that is late leaving the previous job may look at them while the next is set up
	// (it then simply takes part in the next job: `next` is published last, with release semantics)
	std::atomic<unsigned char *> dst{nullptr};
	std::atomic<const unsigned char *> src{nullptr};
	std::atomic<size_t> bytes{0}, n_slices{0};
	std::atomic<size_t> next{0}, done{0};
	static constexpr size_t kSlice = 256 * 1024;

	void work() {
		for (;;) {
			const size_t i = next.fetch_add(1, std::memory_order_acq_rel);
			if (i >= n_slices.load(std::memory_order_relaxed)) return;
			const size_t off = i * kSlice, n = std::min(kSlice, bytes.load(std::memory_order_relaxed) - off);
			memcpy(dst.load(std::memory_order_relaxed) + off, src.load(std::memory_order_relaxed) + off, n);
			done.fetch_add(1, std::memory_order_release);
		}
	}
	void worker_main() {
		unsigned long long seen = 0;
		for (;;) {
			// poll briefly, then sleep
			bool got = false;
			for (int spin = 0; spin < 20000 && !got; spin++) {
				if (stop.load(std::memory_order_relaxed)) return;
				got = generation.load(std::memory_order_acquire) != seen;
				if (!got) std::this_thread::yield();
			}
			if (!got) {
				std::unique_lock<std::mutex> lk(m);
				cv.wait(lk, [&] { return stop.load() || generation.load(std::memory_order_acquire) != seen; });
				if (stop.load()) return;
			}
			seen = generation.load(std::memory_order_acquire);
			work();
		}
	}
	explicit CopyPool(int n) {
		for (int i = 0; i < n; i++) workers.emplace_back([this] { worker_main(); });
	}
	~CopyPool() {
		{ std::lock_guard<std::mutex> lk(m); stop.store(true); }
		cv.notify_all();
		for (auto &t : workers) t.join();
	}
	void run(void *d, const void *s, size_t n) {
		const size_t ns = (n + kSlice - 1) / kSlice;
		next.store(~(size_t)0 / 2, std::memory_order_relaxed);          // park late workers while the job is being described
		dst.store((unsigned char *)d); src.store((const unsigned char *)s); bytes.store(n); n_slices.store(ns);
		done.store(0, std::memory_order_relaxed);
		next.store(0, std::memory_order_release);
		{ std::lock_guard<std::mutex> lk(m); generation.fetch_add(1, std::memory_order_release); }
		cv.notify_all();
		work();
		while (done.load(std::memory_order_acquire) < ns) std::this_thread::yield();
	}
};
CopyPool *g_copy_pool = nullptr;
std::mutex g_copy_mutex;
}  // namespace

void parallel_memcpy(void *dst, const void *src, size_t bytes) {
	if (bytes < 2 * CopyPool::kSlice) { memcpy(dst, src, bytes); return; }
	std::lock_guard<std::mutex> lk(g_copy_mutex);
	if (!g_copy_pool) {
		static const int env = getenv("LS3D_COPY_THREADS") ? atoi(getenv("LS3D_COPY_THREADS")) : -1;
		const unsigned hw = std::thread::hardware_concurrency();
		const int n = env >= 0 ? env : (int)std::max(1u, std::min(6u, hw > 2 ? hw / 2 - 1 : 0u));      // + the caller
		g_copy_pool = new CopyPool(std::max(0, std::min(n, 16)));      // never destroyed: worker threads must not be joined from a library destructor
	}
	g_copy_pool->run(dst, src, bytes);
}

bool DevBuf::reserve(size_t bytes, const char *what) {
	if (bytes <= cap && p) return true;
	release();
	size_t want = bytes < 256 ? 256 : bytes;
	if (!cuda_ok(cudaMalloc(&p, want), what)) { p = nullptr; cap = 0; return false; }
	cap = want;
	return true;
}
void DevBuf::release() {
	if (p) cudaFree(p);
	p = nullptr;
	cap = 0;
}

}  // namespace ls3d

using namespace ls3d;

extern "C" {

const char *ls3d_last_error(void) {
	return t_error;
}

const char *ls3d_version(void) {
	if (!ensure_device()) return t_error;
	return g_version;
}

long long ls3d_launch_count(void) { return g_launches.load(); }
void ls3d_reset_launch_count(void) { g_launches.store(0); }

void *ls3d_dev_alloc(unsigned long long bytes) {
	clear_error();
	if (!ensure_device()) return nullptr;
	void *p = nullptr;
	if (!cuda_ok(cudaMalloc(&p, bytes ? (size_t)bytes : 256), "ls3d_dev_alloc")) return nullptr;
	return p;
}
void ls3d_dev_free(void *p) { if (p) cudaFree(p); }

int ls3d_ipc_export(void *p, unsigned char handle[64]) {
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
	cudaIpcMemHandle_t h;
	if (!cuda_ok(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle")) return -1;
	memcpy(handle, &h, 64);
	return 0;
}
void *ls3d_ipc_open(const unsigned char handle[64]) {
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, 64);
	void *p = nullptr;
	if (!cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return nullptr;
	return p;
}
void ls3d_ipc_close(void *p) { if (p) cudaIpcCloseMemHandle(p); }

// createMesh / deleteMesh, depthprocessing.cpp:1818-1835
Mesh *createMesh(void) {
	Mesh *m = (Mesh *)malloc(sizeof(Mesh));
	if (!m) return nullptr;
	m->nVertices = 0;
	m->vertices = nullptr;
	m->nTriangles = 0;
	m->triangles = nullptr;
	return m;
}

void deleteMesh(Mesh *mesh) {
	if (!mesh) return;
	if (mesh->triangles) host_block_free(mesh->triangles);     // a recycled pinned block, or plain malloc memory (freed as such)
	if (mesh->vertices) host_block_free(mesh->vertices);
	// like the reference, neither pointer is cleared and the struct itself is not freed
}

}  // extern "C"
