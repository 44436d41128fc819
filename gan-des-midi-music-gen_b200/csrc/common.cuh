// Shared helpers for the mmgan_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define MMG_OK 0
#define MMG_EINVAL (-1)       // bad pointer / size / alignment
#define MMG_EUNSUPPORTED (-2) // shape outside what the kernel was written for
#define MMG_EWORKSPACE (-3)   // workspace too small

#define MMG_NUM_SMS 148

void mmg_set_error(const char* fmt, ...);

#define MMG_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            mmg_set_error(__VA_ARGS__);         \
            return (code);                      \
        }                                       \
    } while (0)

#define MMG_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            mmg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                          __FILE__, __LINE__);                                      \
            return (int)_e;                                                         \
        }                                                                           \
    } while (0)

// every kernel launch of this library goes through MMG_LAUNCH_CHECK, which also counts it
extern unsigned long long g_mmg_launches;
#define MMG_LAUNCH_CHECK()                 \
    do {                                   \
        ++g_mmg_launches;                  \
        MMG_CUDA(cudaGetLastError());      \
    } while (0)

static inline int mmg_grid(long long work_items, int threads, int max_ctas_per_sm = 8) {
    long long g = (work_items + threads - 1) / threads;
    long long cap = (long long)MMG_NUM_SMS * max_ctas_per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// activation codes shared by the C-ABI (include/mmgan_b200.h)
enum { MMG_ACT_NONE = 0, MMG_ACT_LRELU = 1, MMG_ACT_RELU = 2, MMG_ACT_SIGMOID = 3 };

__device__ __forceinline__ float mmg_act(float z, int act) {
    switch (act) {
        case MMG_ACT_LRELU: return z > 0.f ? z : 0.2f * z;
        case MMG_ACT_RELU: return z > 0.f ? z : 0.f;
        case MMG_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
        default: return z;
    }
}
// derivative expressed through the OUTPUT y = act(z) (what the backward kernels keep)
__device__ __forceinline__ float mmg_act_grad(float y, int act) {
    switch (act) {
        case MMG_ACT_LRELU: return y > 0.f ? 1.f : 0.2f;
        case MMG_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case MMG_ACT_SIGMOID: return y * (1.f - y);
        default: return 1.f;
    }
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
