// api.cu -- error plumbing and device helpers of libbgs_b200.so.
#include "bgs_common.cuh"

#include <cstring>

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
