// api.cu -- error plumbing and device helpers of libbgs_b200.so.
#include "bgs_common.cuh"

#include <cstring>
#include <mutex>

namespace bgs {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_error(cudaError_t e, const char* what) {
    cudaGetLastError();  // clear the sticky flag of non-fatal errors
    return set_error(BGS_ECUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return set_error(BGS_ENODEVICE,
                         "no CUDA device available (%s); libbgs_b200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
    }
    return BGS_OK;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

// ---------------------------------------------------------------------------------------------
// Per-device workspace.  Allocated once with cudaMalloc and never returned: stream-ordered
// allocations (cudaMallocAsync) hand their memory back to the driver at every synchronisation
// when the pool's release threshold is 0, which costs milliseconds on the next launch.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kMaxDevices = 64;
constexpr int kCounters = 1024;
struct DeviceWorkspace {
    unsigned int* counters = nullptr;
    unsigned next = 0;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaMemPool_t pool = nullptr;  // stream-ordered temporaries (kept across synchronisations)
};
DeviceWorkspace g_ws[kMaxDevices];
std::mutex g_ws_mutex;
}  // namespace

// A claim counter for one launch (round-robin over 1024 slots, so launches in flight on different
// streams never share one).  The caller zeroes it on its stream.
int next_counter(unsigned int** out) {
    int dev = 0;
    BGS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return set_error(BGS_EINVAL, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    DeviceWorkspace& w = g_ws[dev];
    if (!w.counters) BGS_CUDA_TRY(cudaMalloc((void**)&w.counters, kCounters * sizeof(unsigned int)));
    *out = w.counters + (w.next++ % kCounters);
    return BGS_OK;
}

// Write-only scratch of at least `bytes` bytes (contents are never read, so sharing it between
// concurrent launches is harmless).  A buffer that has been handed out is NEVER freed while the
// library is loaded: another host thread may still be launching kernels that write to it, so a
// larger request allocates a new buffer (at least twice the old size, which bounds the retired
// memory by the final size) and the old one stays allocated.
int scratch_buffer(size_t bytes, void** out) {
    int dev = 0;
    BGS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return set_error(BGS_EINVAL, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    DeviceWorkspace& w = g_ws[dev];
    if (w.scratch_bytes < bytes) {
        size_t want = bytes > 2 * w.scratch_bytes ? bytes : 2 * w.scratch_bytes;
        void* fresh = nullptr;
        cudaError_t e = cudaMalloc(&fresh, want);
        if (e != cudaSuccess && want > bytes) {  // no room for the doubled size: take exactly what is needed
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&fresh, want);
        }
        if (e != cudaSuccess) return cuda_error(e, "cudaMalloc(scratch)");
        w.scratch = fresh;  // the previous buffer, if any, is retired, not freed (see above)
        w.scratch_bytes = want;
    }
    *out = w.scratch;
    return BGS_OK;
}

// Stream-ordered temporary from the library's own per-device pool.  The pool's release threshold is
// unlimited, so freed blocks stay with the pool across synchronisations instead of going back to the
// driver (the default pool, threshold 0, costs milliseconds per re-allocation).
int temp_alloc(void** out, size_t bytes, cudaStream_t stream) {
    int dev = 0;
    BGS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return set_error(BGS_EINVAL, "device index %d out of range", dev);
    cudaMemPool_t pool = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_ws_mutex);
        DeviceWorkspace& w = g_ws[dev];
        if (!w.pool) {
            cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            BGS_CUDA_TRY(cudaMemPoolCreate(&w.pool, &props));
            unsigned long long keep = ~0ull;
            BGS_CUDA_TRY(cudaMemPoolSetAttribute(w.pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        pool = w.pool;
    }
    BGS_CUDA_TRY(cudaMallocFromPoolAsync(out, bytes ? bytes : 1, pool, stream));
    return BGS_OK;
}

void temp_free(void* ptr, cudaStream_t stream) {
    if (ptr) cudaFreeAsync(ptr, stream);
}

}  // namespace bgs

extern "C" int bgs_version(void) { return BGS_VERSION; }

extern "C" const char* bgs_last_error(void) { return bgs::g_err; }

extern "C" int bgs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
