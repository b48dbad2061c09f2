// ls3d_internal.h — host-side internals shared by the translation units of libls3d_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstddef>
#include <cstdlib>
#include <mutex>

namespace ls3d {

// thread-local error text behind ls3d_last_error()
void set_error(const char *fmt, ...);
void clear_error();
// returns true when e == cudaSuccess, otherwise records "<what>: <cuda error string>"
bool cuda_ok(cudaError_t e, const char *what);

extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// One lock for the host-buffer (drop-in) entry points: LiveScanServer calls them from two worker threads
// (MainWindowForm.cs:238,304) and they share cached device contexts.
std::mutex &api_mutex();

// Lazily created library stream (non-blocking) for the host-buffer entry points; nullptr on failure.
cudaStream_t api_stream();
// Makes sure a CUDA device is usable; false (with the error recorded) otherwise.  No CPU fallback exists.
bool ensure_device();

// Pinned host blocks handed out as Mesh::vertices so the device->host copy lands directly in the memory the
// caller reads (no staging copy); recycled by deleteMesh.
void *host_block_alloc(size_t bytes);
void host_block_free(void *p);
bool host_block_is_pinned(void *p);     // true: page-locked and mapped, a kernel may store into it

// memcpy spread over a small pool of host threads (the caller takes part): for moving the caller's pageable frame buffers into
// the library's page-locked staging at more than one core's copy bandwidth.  Small sizes fall through to plain memcpy.
void parallel_memcpy(void *dst, const void *src, size_t bytes);

// Kernel launch, optionally with the programmatic-stream-serialization attribute (see pdl_enter in ls3d_common.cuh): the kernel may
// become resident while its predecessor in the stream still runs and orders itself with griddepcontrol.wait.  LS3D_PDL=0 turns every
// such launch into a plain one (A/B).  Errors are picked up by the caller's cudaGetLastError.
inline bool pdl_enabled() {
	static const int v = getenv("LS3D_PDL") ? atoi(getenv("LS3D_PDL")) : 1;
	return v != 0;
}
template <typename... KArgs, typename... Args>
inline void launch_chain(bool programmatic, void (*kernel)(KArgs...), dim3 grid, unsigned block, size_t smem, cudaStream_t st, Args... args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = dim3(block);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = programmatic && pdl_enabled() ? 1 : 0;
	cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// device scratch with grow-only semantics
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	bool reserve(size_t bytes, const char *what);
	void release();
	template <typename T> T *as() const { return static_cast<T *>(p); }
};

}  // namespace ls3d
