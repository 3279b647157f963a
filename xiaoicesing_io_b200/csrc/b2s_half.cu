// 16-bit side of the elementwise helpers: fp32 -> bf16/fp16 casts that feed the TMA-staged
// tensor-core operands (HBM-bound, vectorised, grid-stride).
#include <stdlib.h>
#include "b2s_tc.cuh"

namespace b2s {

template <int BF16>
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, long long n,
                                                   int* __restrict__ zero, int nzero) {
    // programmatic dependent launch: this kernel may start while its predecessor drains; nothing is read before the wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // optional: reset the tile flags of the persistent denoiser kernel that consumes `out` (saves a memset launch)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nzero; i += gridDim.x * blockDim.x) zero[i] = 0;
    const long long n8 = n >> 3;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
        uint4 o;
        o.x = tc::Half16<BF16>::pack2(a.x, a.y);
        o.y = tc::Half16<BF16>::pack2(a.z, a.w);
        o.z = tc::Half16<BF16>::pack2(b.x, b.y);
        o.w = tc::Half16<BF16>::pack2(b.z, b.w);
        reinterpret_cast<uint4*>(out)[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        const long long i = (n8 << 3) + threadIdx.x;
        const uint32_t w = tc::Half16<BF16>::pack2(in[i], 0.f);
        out[i] = (uint16_t)(w & 0xFFFFu);
    }
}

}  // namespace b2s

using namespace b2s;

static int cast_impl(const float* in, void* out, int64_t n, int* zero, int nzero, int bf16, void* stream);

extern "C" int b2s_cast_f32_h(const float* in, void* out, int64_t n, int bf16, void* stream) {
    return cast_impl(in, out, n, nullptr, 0, bf16, stream);
}

extern "C" int b2s_cast_f32_h_reset(const float* in, void* out, int64_t n, int* flags, int n_flags, int bf16, void* stream) {
    B2S_CHECK_ARG(n_flags == 0 || flags, "b2s_cast_f32_h_reset: null flags");
    return cast_impl(in, out, n, flags, n_flags, bf16, stream);
}

static int cast_impl(const float* in, void* out, int64_t n, int* zero, int nzero, int bf16, void* stream) {
    B2S_CHECK_ARG(n >= 0 && (n == 0 || (in && out)), "b2s_cast_f32_h: null pointer");
    B2S_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "b2s_cast_f32_h: pointers must be 16B aligned");
    if (n == 0 && nzero == 0) return B2S_OK;
    long long blocks = ((n >> 3) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const bool ew_pdl = getenv("B2S_EW_PDL") != nullptr && atoi(getenv("B2S_EW_PDL")) != 0;
    cfg.numAttrs = ew_pdl ? 1 : 0;
    uint16_t* o = (uint16_t*)out;
    long long nn = n;
    if (bf16) B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cast_kernel<1>, in, o, nn, zero, nzero));
    else B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cast_kernel<0>, in, o, nn, zero, nzero));
    return B2S_OK;
}

// ====================================================================================================
// LYNXNet layer helpers of the 16-bit path (lynxnet.py:76-87, 52-62): the pointwise convs run on the tensor
// cores (b2s_tc_lynx_glu / b2s_tc_linear_residual); these two HBM-bound kernels feed them.
// ====================================================================================================
namespace b2s {

__device__ __forceinline__ float4 ld_h4(const uint16_t* p, int bf16) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    float2 a, b;
    if (bf16) { a = tc::Half16<1>::unpack2(u.x); b = tc::Half16<1>::unpack2(u.y); }
    else { a = tc::Half16<0>::unpack2(u.x); b = tc::Half16<0>::unpack2(u.y); }
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st_h4(uint16_t* p, float4 v, int bf16) {
    uint2 u;
    if (bf16) { u.x = tc::Half16<1>::pack2(v.x, v.y); u.y = tc::Half16<1>::pack2(v.z, v.w); }
    else { u.x = tc::Half16<0>::pack2(v.x, v.y); u.y = tc::Half16<0>::pack2(v.z, v.w); }
    *reinterpret_cast<uint2*>(p) = u;
}

// One warp per frame row, a lane owns 4 consecutive channels per 128-channel group (16-byte fp32 / 8-byte 16-bit
// accesses, fully coalesced).  PRE: u = x + cond + d, residual write-back x <- x + cond when strong_cond
// (front_cond_inject, lynxnet.py:77-84); h = LayerNorm_C(u) * gamma + beta (eps 1e-5) in 16 bits.
template <bool PRE>
__global__ void __launch_bounds__(256) layernorm_h_kernel(float* __restrict__ x, const uint16_t* __restrict__ cond, int ld_cond,
                                                          const float* __restrict__ dvec, int d_stride,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          uint16_t* __restrict__ h, int rows, int T, int C, int strong, int bf16) {
    extern __shared__ float srow[];                 // [warps][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= rows) return;
    float* u = srow + (long long)warp * C;
    const int b = PRE ? r / T : 0;
    float sum = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(x + (long long)r * C + c);
        if (PRE) {
            if (cond) {          // NULL: the cond add was folded into the previous layer's residual epilogue (b2s_tc_linear_residual_cond)
                const float4 cc = ld_h4(cond + (long long)r * ld_cond + c, bf16);
                v = make_float4(v.x + cc.x, v.y + cc.y, v.z + cc.z, v.w + cc.w);
                if (strong) *reinterpret_cast<float4*>(x + (long long)r * C + c) = v;
            }
            const float4 d = __ldg(reinterpret_cast<const float4*>(dvec + (long long)b * d_stride + c));
            v = make_float4(v.x + d.x, v.y + d.y, v.z + d.z, v.w + d.w);
        }
        *reinterpret_cast<float4*>(u + c) = v;
        sum += (v.x + v.y) + (v.z + v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float var = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(u + c);
        const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
        var += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)C + 1e-5f);
    for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(u + c);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c));
        st_h4(h + (long long)r * C + c,
              make_float4((v.x - mean) * rstd * g.x + bb.x, (v.y - mean) * rstd * g.y + bb.y, (v.z - mean) * rstd * g.z + bb.z,
                          (v.w - mean) * rstd * g.w + bb.w), bf16);
    }
}

// Depthwise conv along time + bias + activation, 16-bit in / out, fp32 math (lynxnet.py:57-58).
// Thread = 2 adjacent channels (4-byte accesses, a warp reads 128 contiguous bytes per frame), one block = 256 channels x
// a strip of DWH_STRIP frames of one utterance.  Per iteration a thread produces G = 8 consecutive output frames from a
// register window of K+7 input frames (8 x 2 independent FMA chains), then shifts the window by 8 (7.5 moves per output)
// and takes the 8 new frames that were requested one iteration earlier (memory-level parallelism).  The loop body is
// ~700 instructions: an earlier version unrolled K-fold (11.5 k instructions) and was instruction-cache bound.
// Weights are K-MAJOR [K][inner] so that they load coalesced.
// SiLU / ReLU variants go through ONE out-of-line call (PReLU, the default, is inline): keeps the unrolled body small
__device__ __noinline__ float2 act2_slow(float2 v, int act) { return make_float2(apply_act(v.x, act), apply_act(v.y, act)); }

constexpr int DWH_MAXK = 31;
constexpr int DWH_G = 8;
constexpr int DWH_STRIP = 128;

template <int K, int BF16>
__global__ void __launch_bounds__(128, 2) dwconv_h_kernel(const uint16_t* __restrict__ g, const float* __restrict__ WdwT,
                                                          const float* __restrict__ bias, const float* __restrict__ slope,
                                                          uint16_t* __restrict__ p, int T, int inner, int act) {
    const int ch = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (ch >= inner) return;
    constexpr int PAD = K / 2, G = DWH_G, W = K + G - 1;
    const int b = blockIdx.z;
    const int t_begin = blockIdx.y * DWH_STRIP;
    const int t_end = min(T, t_begin + DWH_STRIP);
    float2 w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = __ldg(reinterpret_cast<const float2*>(WdwT + (long long)k * inner + ch));
    const float2 bb = __ldg(reinterpret_cast<const float2*>(bias + ch));
    const float2 sl = slope ? __ldg(reinterpret_cast<const float2*>(slope + ch)) : make_float2(0.f, 0.f);
    const uint16_t* gb = g + (long long)b * T * inner + ch;
    uint16_t* pb = p + (long long)b * T * inner + ch;
    auto fetch = [&](int s) -> uint32_t {
        return (s >= 0 && s < T) ? __ldg(reinterpret_cast<const uint32_t*>(gb + (long long)s * inner)) : 0u;
    };
    // window: win[i] = input frame (t0 - PAD + i) for the current group of outputs t0 .. t0+G-1
    float2 win[W];
#pragma unroll
    for (int i = 0; i < W; ++i) win[i] = tc::Half16<BF16>::unpack2(fetch(t_begin - PAD + i));
    uint32_t nxt[G];
#pragma unroll
    for (int j = 0; j < G; ++j) nxt[j] = fetch(t_begin - PAD + W + j);
#pragma unroll 1
    for (int t0 = t_begin; t0 < t_end; t0 += G) {
        uint32_t cur[G];
#pragma unroll
        for (int j = 0; j < G; ++j) cur[j] = nxt[j];
        if (t0 + G < t_end) {
#pragma unroll
            for (int j = 0; j < G; ++j) nxt[j] = fetch(t0 + G - PAD + W + j);
        }
        float2 acc[G];
#pragma unroll
        for (int j = 0; j < G; ++j) acc[j] = bb;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int j = 0; j < G; ++j) {
                acc[j].x = fmaf(w[k].x, win[j + k].x, acc[j].x);
                acc[j].y = fmaf(w[k].y, win[j + k].y, acc[j].y);
            }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (t0 + j < t_end) {
                float o0, o1;
                if (act == 0) { o0 = acc[j].x >= 0.f ? acc[j].x : sl.x * acc[j].x; o1 = acc[j].y >= 0.f ? acc[j].y : sl.y * acc[j].y; }
                else { const float2 o = act2_slow(acc[j], act); o0 = o.x; o1 = o.y; }
                *reinterpret_cast<uint32_t*>(pb + (long long)(t0 + j) * inner) = tc::Half16<BF16>::pack2(o0, o1);
            }
        }
        // slide the window by G frames
#pragma unroll
        for (int i = 0; i < W - G; ++i) win[i] = win[i + G];
#pragma unroll
        for (int j = 0; j < G; ++j) win[W - G + j] = tc::Half16<BF16>::unpack2(cur[j]);
    }
}

template <int K>
static int launch_dwconv_h(const void* g_h, const float* WdwT, const float* bias, const float* slope, void* p_h, int B, int T,
                           int inner, int act, int bf16, cudaStream_t st) {
    dim3 grid(ceil_div(inner / 2, 128), ceil_div(T, DWH_STRIP), B);
    if (bf16) dwconv_h_kernel<K, 1><<<grid, 128, 0, st>>>((const uint16_t*)g_h, WdwT, bias, slope, (uint16_t*)p_h, T, inner, act);
    else dwconv_h_kernel<K, 0><<<grid, 128, 0, st>>>((const uint16_t*)g_h, WdwT, bias, slope, (uint16_t*)p_h, T, inner, act);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

}  // namespace b2s

extern "C" int b2s_lynx_prenorm_h(float* x, const void* cond_h, int ld_cond, const float* dvec, int d_stride,
                                  const float* gamma, const float* beta, void* h_h, int B, int T, int C, int strong_cond,
                                  int bf16, void* stream) {
    B2S_CHECK_ARG(x && dvec && gamma && beta && h_h, "b2s_lynx_prenorm_h: null pointer");
    B2S_CHECK_ARG(C > 0 && C <= 8192 && C % 4 == 0 && ld_cond % 4 == 0 && d_stride % 4 == 0, "b2s_lynx_prenorm_h: C, ld_cond, d_stride must be multiples of 4");
    const int rows = B * T;
    if (rows <= 0) return B2S_OK;
    const int warps = C <= 1024 ? 8 : (C <= 2048 ? 4 : 1);
    size_t smem = (size_t)warps * C * sizeof(float);
    layernorm_h_kernel<true><<<ceil_div(rows, warps), warps * 32, smem, (cudaStream_t)stream>>>(
        x, (const uint16_t*)cond_h, ld_cond, dvec, d_stride, gamma, beta, (uint16_t*)h_h, rows, T, C, strong_cond, bf16);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_layernorm_h(const float* x, const float* gamma, const float* beta, void* h_h, int rows, int C, int bf16,
                               void* stream) {
    B2S_CHECK_ARG(x && gamma && beta && h_h, "b2s_layernorm_h: null pointer");
    B2S_CHECK_ARG(C > 0 && C <= 8192 && C % 4 == 0, "b2s_layernorm_h: C must be a multiple of 4");
    if (rows <= 0) return B2S_OK;
    const int warps = C <= 1024 ? 8 : (C <= 2048 ? 4 : 1);
    size_t smem = (size_t)warps * C * sizeof(float);
    layernorm_h_kernel<false><<<ceil_div(rows, warps), warps * 32, smem, (cudaStream_t)stream>>>(
        const_cast<float*>(x), nullptr, 0, nullptr, 0, gamma, beta, (uint16_t*)h_h, rows, 1, C, 0, bf16);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_lynx_dwconv_h(const void* g_h, const float* WdwT, const float* bias, const float* slope, void* p_h, int B,
                                 int T, int inner, int ksize, int act, int bf16, void* stream) {
    B2S_CHECK_ARG(g_h && WdwT && bias && p_h, "b2s_lynx_dwconv_h: null pointer");
    B2S_CHECK_ARG(ksize >= 1 && ksize <= DWH_MAXK && (ksize & 1), "b2s_lynx_dwconv_h: kernel size must be odd and <= %d", DWH_MAXK);
    B2S_CHECK_ARG(inner % 2 == 0, "b2s_lynx_dwconv_h: inner must be even");
    B2S_CHECK_ARG(act != 0 || slope, "b2s_lynx_dwconv_h: PReLU needs slope");
    B2S_CHECK_ARG(B < 65536, "b2s_lynx_dwconv_h: B too large");
    if (B <= 0 || T <= 0) return B2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ksize) {          // the window ring is a compile-time structure: one instantiation per (odd) kernel size
#define B2S_DW_CASE(KK) case KK: return launch_dwconv_h<KK>(g_h, WdwT, bias, slope, p_h, B, T, inner, act, bf16, st);
        B2S_DW_CASE(1) B2S_DW_CASE(3) B2S_DW_CASE(5) B2S_DW_CASE(7) B2S_DW_CASE(9) B2S_DW_CASE(11) B2S_DW_CASE(13) B2S_DW_CASE(15)
        B2S_DW_CASE(17) B2S_DW_CASE(19) B2S_DW_CASE(21) B2S_DW_CASE(23) B2S_DW_CASE(25) B2S_DW_CASE(27) B2S_DW_CASE(29) B2S_DW_CASE(31)
#undef B2S_DW_CASE
    }
    set_error("b2s_lynx_dwconv_h: unsupported kernel size %d", ksize);
    return B2S_ERR_UNSUPPORTED;
}
