// Shared helpers for the b2s (B200 DiffSinger sampling) CUDA kernels.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/b2s.h"

namespace b2s {

// ---- error plumbing (C-ABI: every entry point returns 0 or a negative code; text via b2s_last_error) ----
void set_error(const char* fmt, ...);

#define B2S_CHECK_ARG(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            ::b2s::set_error(__VA_ARGS__);                         \
            return B2S_ERR_INVALID_ARGUMENT;                       \
        }                                                          \
    } while (0)

#define B2S_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::b2s::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                             __FILE__, __LINE__);                                         \
            return B2S_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define B2S_CHECK_LAUNCH() B2S_CHECK_CUDA(cudaGetLastError())

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- activations shared by the fp32 and tensor-core paths -------------------------------------------
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_MISH = 2, ACT_GELU = 3, ACT_SILU = 4, ACT_LRELU = 5 };
constexpr float LRELU_SLOPE = 0.1f;                        // nsf_hifigan/models.py:15

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(x, 0.0f);
        case ACT_MISH: {                                   // x * tanh(softplus(x)), torch threshold 20
            float sp = x > 20.0f ? x : log1pf(expf(x));
            return x * tanhf(sp);
        }
        case ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));   // exact erf GELU
        case ACT_SILU: return x * sigmoid_acc(x);
        case ACT_LRELU: return x > 0.0f ? x : __fmul_rn(x, LRELU_SLOPE);
        default: return x;
    }
}

// fast variants for the half-precision tensor-core epilogues (MUFU tanh)
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }

}  // namespace b2s
