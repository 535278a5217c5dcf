// Library-level entry points and error plumbing of libscipnp.
#include <stdarg.h>

#include "common.cuh"

namespace scipnp {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    // clear the sticky-free error so that the next call reports its own failure
    (void)cudaGetLastError();
    return (e == cudaErrorMemoryAllocation) ? SCIPNP_ENOMEM : SCIPNP_ECUDA;
}

int num_sms() {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return 148;   // B200
    }
    cached = n;
    return n;
}

}  // namespace scipnp

extern "C" {

int scipnp_version(void) { return 100; }   // 0.1.0

const char* scipnp_last_error(void) { return scipnp::g_err; }

int scipnp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

long long scipnp_launch_count(void) { return scipnp::g_launches.load(); }

}  // extern "C"
