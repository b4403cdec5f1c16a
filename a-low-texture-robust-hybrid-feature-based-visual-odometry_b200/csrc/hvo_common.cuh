// Shared host/device helpers for libhvofront (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/hvo_capi.h"

namespace hvo {

void set_error(const char* fmt, ...);

#define HVO_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            ::hvo::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return HVO_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define HVO_CHECK_ARG(cond, msg)                            \
    do {                                                    \
        if (!(cond)) {                                      \
            ::hvo::set_error("bad argument: %s", msg);      \
            return HVO_ERR_ARG;                             \
        }                                                   \
    } while (0)

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ int reflect101(int p, int n) {
    // single reflection is enough for |overshoot| < n (halo radii here are <= 21 px)
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

}  // namespace hvo
