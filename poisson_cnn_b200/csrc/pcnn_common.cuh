// Shared helpers for the pcnn CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../include/pcnn.h"

namespace pcnn {

// thread-local last-error string, exported through pcnn_last_error()
void set_error(const char* fmt, ...);
void count_launch();

#define PCNN_CHECK_ARG(cond, ...)                    \
    do {                                             \
        if (!(cond)) {                               \
            ::pcnn::set_error(__VA_ARGS__);          \
            return PCNN_ERR_INVALID_ARGUMENT;        \
        }                                            \
    } while (0)

#define PCNN_CHECK_CUDA(expr)                                                             \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::pcnn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                              __FILE__, __LINE__);                                        \
            return PCNN_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

// every kernel launch of the library goes through this macro, which also feeds pcnn_launch_count()
#define PCNN_CHECK_LAUNCH()                      \
    do {                                         \
        ::pcnn::count_launch();                  \
        PCNN_CHECK_CUDA(cudaGetLastError());     \
    } while (0)

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == PCNN_ACT_LEAKY_RELU) return v >= 0.f ? v : v * 0.2f;   // tf.nn.leaky_relu alpha
    if (act == PCNN_ACT_TANH) return tanhf(v);
    return v;
}

// source index of tf.pad for an out-of-range coordinate; result clamped so that partial tiles
// never read out of bounds (their outputs are discarded).
__device__ __forceinline__ int pad_src_index(int i, int n, int mode) {
    if (i >= 0 && i < n) return i;
    if (mode == PCNN_PAD_SYMMETRIC) i = (i < 0) ? (-1 - i) : (2 * n - 1 - i);
    else if (mode == PCNN_PAD_REFLECT) i = (i < 0) ? (-i) : (2 * n - 2 - i);
    return min(max(i, 0), n - 1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace pcnn
