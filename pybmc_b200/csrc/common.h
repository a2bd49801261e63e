// Shared host-side helpers for the C ABI (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include "../../include/bmc_b200.h"

namespace bmc {

void set_error(const char* fmt, ...);

inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

// SMs of the current device (148 on B200): grids are sized from this, never from a literal.
inline int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        return 148;
    return n;
}

#define BMC_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            bmc::set_error(__VA_ARGS__);  \
            return BMC_ERR_ARG;           \
        }                                 \
    } while (0)

#define BMC_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t err__ = (expr);                                                      \
        if (err__ != cudaSuccess) {                                                      \
            bmc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),    \
                           __FILE__, __LINE__);                                          \
            return BMC_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define BMC_LAUNCH_CHECK() BMC_CUDA(cudaGetLastError())

}  // namespace bmc
