// Shared helpers of libscipnp (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/scipnp.h"

namespace scipnp {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SCIPNP_CUDA(expr)                                             \
    do {                                                              \
        cudaError_t _e = (expr);                                      \
        if (_e != cudaSuccess) return ::scipnp::cuda_fail(_e, #expr); \
    } while (0)

#define SCIPNP_REQUIRE(cond, msg)                      \
    do {                                               \
        if (!(cond)) {                                 \
            ::scipnp::set_error("%s: %s", __func__, msg); \
            return SCIPNP_EINVAL;                      \
        }                                              \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return SCIPNP_OK;
}

// launch counter (per process); the solver reports it as `gpu_launches`
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms();

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace scipnp
