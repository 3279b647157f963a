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

// The same with the row held in REGISTERS (C a multiple of 128, at most 2048): no shared-memory round trip, every load of the row
// issued before the first reduction.  The staged variant above took 72 us at config 3's shape (276 MB: 36 us at the HBM roofline).
template <bool PRE, int NQ>
__global__ void __launch_bounds__(256) layernorm_h_reg_kernel(float* __restrict__ x, const uint16_t* __restrict__ cond, int ld_cond,
                                                              const float* __restrict__ dvec, int d_stride,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              uint16_t* __restrict__ h, int rows, int T, int strong, int bf16) {
    constexpr int C = NQ * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= rows) return;
    const int b = PRE ? r / T : 0;
    float4 v[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) v[i] = *reinterpret_cast<const float4*>(x + (long long)r * C + i * 128 + lane * 4);
    if (PRE) {
        if (cond) {
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const float4 cc = ld_h4(cond + (long long)r * ld_cond + i * 128 + lane * 4, bf16);
                v[i] = make_float4(v[i].x + cc.x, v[i].y + cc.y, v[i].z + cc.z, v[i].w + cc.w);
                if (strong) *reinterpret_cast<float4*>(x + (long long)r * C + i * 128 + lane * 4) = v[i];
            }
        }
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            const float4 d = __ldg(reinterpret_cast<const float4*>(dvec + (long long)b * d_stride + i * 128 + lane * 4));
            v[i] = make_float4(v[i].x + d.x, v[i].y + d.y, v[i].z + d.z, v[i].w + d.w);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NQ; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        var += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)C + 1e-5f);
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        const int c = i * 128 + lane * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c));
        st_h4(h + (long long)r * C + c,
              make_float4((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y, (v[i].z - mean) * rstd * g.z + bb.z,
                          (v[i].w - mean) * rstd * g.w + bb.w), bf16);
    }
}

template <bool PRE>
static bool launch_layernorm_reg(float* x, const uint16_t* cond, int ld_cond, const float* dvec, int d_stride, const float* gamma,
                                 const float* beta, uint16_t* h, int rows, int T, int C, int strong, int bf16, cudaStream_t st) {
    const int grid = ceil_div(rows, 8);
    switch (C) {
#define B2S_LN_CASE(NQ) case NQ * 128: layernorm_h_reg_kernel<PRE, NQ><<<grid, 256, 0, st>>>(x, cond, ld_cond, dvec, d_stride, gamma, beta, h, rows, T, strong, bf16); return true;
        B2S_LN_CASE(1) B2S_LN_CASE(2) B2S_LN_CASE(3) B2S_LN_CASE(4) B2S_LN_CASE(6) B2S_LN_CASE(8) B2S_LN_CASE(12) B2S_LN_CASE(16)
#undef B2S_LN_CASE
    }
    return false;
}

// LayerNorm over the channels of 16-bit rows -> 16-bit rows, fp32 statistics, eps as an argument (ConvNeXt block, convnext.py:44:
// eps 1e-6, input = the depthwise conv's output).  One warp per row, the row kept in registers (C <= 2048).
__global__ void __launch_bounds__(256) layernorm_hh_kernel(const uint16_t* __restrict__ in, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, uint16_t* __restrict__ out, int rows,
                                                           int C, float eps, int bf16) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= rows) return;
    float4 v[16];                                   // C / 128 quads per lane
    const int nq = C >> 7;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (i < nq) {
            v[i] = ld_h4(in + (long long)r * C + i * 128 + lane * 4, bf16);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (i < nq) {
            const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
            var += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)C + eps);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (i < nq) {
            const int c = i * 128 + lane * 4;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c));
            st_h4(out + (long long)r * C + c,
                  make_float4((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y,
                              (v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w), bf16);
        }
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

// ---------------------------------------------------------------------------------------------------------------------------
// Depthwise conv on the tensor cores.  The register-window kernel above is issue-bound (1.46 instructions per output: 198 us for
// the 2048-channel, 45 056-frame conv of config 3, 40 % of the fp32 pipe).  Per channel the conv is a product with a Toeplitz
// matrix: with t = 16 a + b,
//     out[16 a + b] = sum_{j = 0}^{47} in[16 a + j - 15] * Wt[j][b],      Wt[j][b] = w31[j - b]   (0 outside the 31 taps)
// i.e. a [T/16 x 48] x [48 x 16] GEMM whose A operand is the input read with overlapping rows (a Hankel view of the same shared
// memory) - 1.55 x the MACs, on mma.sync.m16n8k16 (fp32 accumulate) instead of FFMA: 0.22 instructions per output.
//   * block = 256 threads = one utterance x 64 channels, looping over 256-frame tiles with double-buffered cp.async staging
//     (frames outside the utterance zero-filled = the conv's padding); warp w owns channels 8 w .. 8 w + 7 (one 16-byte chunk)
//   * A fragments: a thread needs frame pairs (r, r + 1) of its channels: two LDS.128 (8 channels of one frame each) + 8 PRMT
//     give 8 channels' registers at once; the chunk index is XOR-swizzled with frame bits 1, 2, 4 -> conflict-free quarter-warps
//   * B fragments: Toeplitz, so all six (k-block, n-block) fragments of a channel are windows of SEVEN packed tap pairs
//     P[m] = (w31[d + 8 m], w31[d + 8 m + 1]), d = 2 (lane % 4) - lane / 4, m = -1 .. 5, kept in registers for the whole block
//   * taps are rounded to the activations' 16-bit type (like every other weight of the 16-bit path); kernel sizes below 31 are
//     centred in the 31-tap window
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool ok) {
    const int n = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int BF16>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int DWM_TILE = 256;                 // output frames per tile
constexpr int DWM_ROWS = DWM_TILE + 32;       // staged frames: row i = frame f0 - 15 + i, rows 0 .. 287 are read
constexpr int DWM_BUF = DWM_ROWS * 128;       // 64 channels x 2 B per row
constexpr int DWM_SMEM = 3 * DWM_BUF;            // + 31 x 65 fp32 taps behind it
constexpr int DWM_SMEM_ALL = DWM_SMEM + 31 * 65 * 4;

__device__ __forceinline__ int dwm_swz(int row) { return ((row >> 1) & 3) | (((row >> 4) & 1) << 2); }

template <int BF16, bool PRELU>
__global__ void __launch_bounds__(256, 1) dwconv_mma_kernel(const uint16_t* __restrict__ g, const float* __restrict__ WdwT,
                                                            const float* __restrict__ bias, const float* __restrict__ slope,
                                                            uint16_t* __restrict__ p, int B, int T, int inner, int ksize,
                                                            int act) {
    extern __shared__ __align__(128) uint8_t dw_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane & 3, gq = lane >> 2;
    const int ch0 = blockIdx.x * 64;                       // inner % 64 == 0 (host-checked)
    const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(dw_smem);
    const int n_tiles = (T + DWM_TILE - 1) / DWM_TILE;
    // work items of this block: utterances blockIdx.y, blockIdx.y + gridDim.y, ... x their tiles, in order; item j is staged in
    // buffer j % 3, two items ahead of the one being computed (a group is committed per item even when it is empty, so that
    // wait_group counts stay uniform)
    const int n_items = ((B - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y) * n_tiles;
    auto load_item = [&](int j) {
        if (j < n_items) {
            const int b = blockIdx.y + (j / n_tiles) * gridDim.y, it = j % n_tiles;
            const uint16_t* gb = g + (long long)b * T * inner + ch0 + (lane & 7) * 8;
            const int f0 = it * DWM_TILE - 15;
            const uint32_t dst = smem_s + (j % 3) * DWM_BUF;
            // 288 rows, 8 chunks of 16 B each: a warp instruction covers 4 rows, the block 32 rows
#pragma unroll
            for (int r0 = 0; r0 < DWM_ROWS; r0 += 32) {
                const int r = r0 + warp * 4 + (lane >> 3);
                const int s = f0 + r;
                const bool ok = s >= 0 && s < T;
                cp_async16_zfill(dst + r * 128 + (((lane & 7) ^ dwm_swz(r)) << 4), gb + (long long)(ok ? s : 0) * inner, ok);
            }
        }
        cp_async_commit();
    };
    load_item(0);
    load_item(1);

    // the block's 31 x 64 taps (fp32, centred in the 31-tap window) staged once, coalesced; pitch 65 keeps the per-lane reads of
    // different taps on different banks
    float* wsm = reinterpret_cast<float*>(dw_smem + DWM_SMEM);
    const int koff = 15 - ksize / 2;                       // w31[k] = w[k - koff]
    for (int i = threadIdx.x; i < 31 * 64; i += 256) {
        const int k = i >> 6, c = i & 63, kk = k - koff;
        wsm[k * 65 + c] = (kk >= 0 && kk < ksize) ? __ldg(WdwT + (long long)kk * inner + ch0 + c) : 0.f;
    }
    __syncthreads();
    // this warp's 8 channels: Toeplitz tap pairs, bias, PReLU slope
    const int chw = ch0 + warp * 8;
    const int d = 2 * q - gq;
    uint32_t P[8][7];
#pragma unroll
    for (int m = 0; m < 7; ++m) {
        const int i0 = d + 8 * (m - 1), i1 = i0 + 1;
        const bool ok0 = i0 >= 0 && i0 < 31, ok1 = i1 >= 0 && i1 < 31;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float w0 = ok0 ? wsm[i0 * 65 + warp * 8 + c] : 0.f;
            const float w1 = ok1 ? wsm[i1 * 65 + warp * 8 + c] : 0.f;
            P[c][m] = tc::Half16<BF16>::pack2(w0, w1);
        }
    }
    float bs[8], sl[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { bs[c] = __ldg(bias + chw + c); sl[c] = (PRELU ? __ldg(slope + chw + c) : 1.f) - 1.f; }
    for (int j = 0; j < n_items; ++j) {
        const int buf = j % 3;
        const int b = blockIdx.y + (j / n_tiles) * gridDim.y, it = j % n_tiles;
        load_item(j + 2);             // into the buffer item j - 1 used (released by the barrier that ended its iteration)
        cp_async_wait<2>();
        __syncthreads();
        const uint8_t* tile = dw_smem + buf * DWM_BUF;
        float acc[8][2][4];
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[c][nb][i] = bs[c];
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
            uint32_t A[8][4];
            // a0 = rows (g + kb) pair e = 0, a1 = rows (g + 8 + kb) e = 0, a2 = (g + kb) e = 1, a3 = (g + 8 + kb) e = 1
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const int r = 16 * (gq + kb + ((f & 1) ? 8 : 0)) + ((f & 2) ? 8 : 0) + 2 * q;
                const uint4 x0 = *reinterpret_cast<const uint4*>(tile + r * 128 + ((warp ^ dwm_swz(r)) << 4));
                const uint4 x1 = *reinterpret_cast<const uint4*>(tile + (r + 1) * 128 + ((warp ^ dwm_swz(r + 1)) << 4));
                A[0][f] = __byte_perm(x0.x, x1.x, 0x5410); A[1][f] = __byte_perm(x0.x, x1.x, 0x7632);
                A[2][f] = __byte_perm(x0.y, x1.y, 0x5410); A[3][f] = __byte_perm(x0.y, x1.y, 0x7632);
                A[4][f] = __byte_perm(x0.z, x1.z, 0x5410); A[5][f] = __byte_perm(x0.z, x1.z, 0x7632);
                A[6][f] = __byte_perm(x0.w, x1.w, 0x5410); A[7][f] = __byte_perm(x0.w, x1.w, 0x7632);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) mma_16816<BF16>(acc[c][nb], A[c], P[c][2 * kb - nb + 1], P[c][2 * kb - nb + 2]);
            }
        }
        // epilogue: activation, 16-bit; the tile is transposed through the (consumed) input buffer so that the global stores are
        // whole 128-byte rows (a direct store from the accumulator layout touches 32 rows per instruction).  A warp reads and
        // writes only ITS chunk column of the buffer, so a warp-level barrier orders its reads before its writes
        __syncwarp();
        uint8_t* otile = dw_smem + buf * DWM_BUF;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 16 * (gq + ((i & 2) ? 8 : 0)) + 8 * nb + 2 * q + (i & 1);
                float o[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float v = acc[c][nb][i];
                    // PReLU as v + (slope - 1) * min(v, 0): one FMNMX + one FFMA
                    o[c] = PRELU ? fmaf(sl[c], fminf(v, 0.f), v) : apply_act(v, act);
                }
                uint4 w;
                w.x = tc::Half16<BF16>::pack2(o[0], o[1]);
                w.y = tc::Half16<BF16>::pack2(o[2], o[3]);
                w.z = tc::Half16<BF16>::pack2(o[4], o[5]);
                w.w = tc::Half16<BF16>::pack2(o[6], o[7]);
                *reinterpret_cast<uint4*>(otile + r * 128 + ((warp ^ dwm_swz(r)) << 4)) = w;
            }
        }
        __syncthreads();
        {
            uint16_t* po = p + ((long long)b * T + (long long)it * DWM_TILE) * inner + ch0 + (lane & 7) * 8;
            const int rows = min(DWM_TILE, T - it * DWM_TILE);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = j * 32 + warp * 4 + (lane >> 3);
                if (r < rows)
                    *reinterpret_cast<uint4*>(po + (long long)r * inner) =
                        *reinterpret_cast<const uint4*>(otile + r * 128 + (((lane & 7) ^ dwm_swz(r)) << 4));
            }
        }
        __syncthreads();          // this buffer is refilled by the next iteration's load_item
    }
}

template <int BF16>
static int launch_dwconv_mma(const void* g_h, const float* WdwT, const float* bias, const float* slope, void* p_h, int B, int T,
                             int inner, int ksize, int act, cudaStream_t st) {
    static tc::PerDevice configured;
    if (configured.first()) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(dwconv_mma_kernel<BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DWM_SMEM_ALL));
        B2S_CHECK_CUDA(cudaFuncSetAttribute(dwconv_mma_kernel<BF16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DWM_SMEM_ALL));
    }
    // persistent over utterances: one block per SM (255 registers), channel tile fastest so that concurrently running blocks
    // read neighbouring 128-byte columns of the same frames
    const int n_ct = inner / 64;
    dim3 grid(n_ct, std::max(1, std::min(B, tc::num_sms() / n_ct)));
    if (act == 0)
        dwconv_mma_kernel<BF16, true><<<grid, 256, DWM_SMEM_ALL, st>>>((const uint16_t*)g_h, WdwT, bias, slope, (uint16_t*)p_h, B, T, inner, ksize, act);
    else
        dwconv_mma_kernel<BF16, false><<<grid, 256, DWM_SMEM_ALL, st>>>((const uint16_t*)g_h, WdwT, bias, slope, (uint16_t*)p_h, B, T, inner, ksize, act);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

static bool dwconv_use_mma() {
    static const bool on = [] { const char* e = getenv("B2S_DWCONV_MMA"); return !e || atoi(e) != 0; }();
    return on;
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
    if (launch_layernorm_reg<true>(x, (const uint16_t*)cond_h, ld_cond, dvec, d_stride, gamma, beta, (uint16_t*)h_h, rows, T, C,
                                   strong_cond, bf16, (cudaStream_t)stream)) {
        B2S_CHECK_LAUNCH();
        return B2S_OK;
    }
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
    if (launch_layernorm_reg<false>(const_cast<float*>(x), nullptr, 0, nullptr, 0, gamma, beta, (uint16_t*)h_h, rows, 1, C, 0, bf16,
                                    (cudaStream_t)stream)) {
        B2S_CHECK_LAUNCH();
        return B2S_OK;
    }
    const int warps = C <= 1024 ? 8 : (C <= 2048 ? 4 : 1);
    size_t smem = (size_t)warps * C * sizeof(float);
    layernorm_h_kernel<false><<<ceil_div(rows, warps), warps * 32, smem, (cudaStream_t)stream>>>(
        const_cast<float*>(x), nullptr, 0, nullptr, 0, gamma, beta, (uint16_t*)h_h, rows, 1, C, 0, bf16);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_layernorm_hh(const void* in_h, const float* gamma, const float* beta, void* out_h, int rows, int C, float eps,
                                int bf16, void* stream) {
    B2S_CHECK_ARG(in_h && gamma && beta && out_h, "b2s_layernorm_hh: null pointer");
    B2S_CHECK_ARG(C > 0 && C <= 2048 && C % 128 == 0, "b2s_layernorm_hh: C must be a multiple of 128, at most 2048 (C=%d)", C);
    B2S_CHECK_ARG(tc::al16(in_h) && tc::al16(out_h) && tc::al16(gamma) && tc::al16(beta), "b2s_layernorm_hh: misaligned pointer");
    if (rows <= 0) return B2S_OK;
    layernorm_hh_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)in_h, gamma, beta, (uint16_t*)out_h,
                                                                             rows, C, eps, bf16);
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
    // tensor-core formulation (Toeplitz GEMM on mma.sync): 64-channel tiles, 16-byte rows
    if (inner % 64 == 0 && tc::al16(g_h) && tc::al16(p_h) && dwconv_use_mma())
        return bf16 ? launch_dwconv_mma<1>(g_h, WdwT, bias, slope, p_h, B, T, inner, ksize, act, st)
                    : launch_dwconv_mma<0>(g_h, WdwT, bias, slope, p_h, B, T, inner, ksize, act, st);
    switch (ksize) {          // the window ring is a compile-time structure: one instantiation per (odd) kernel size
#define B2S_DW_CASE(KK) case KK: return launch_dwconv_h<KK>(g_h, WdwT, bias, slope, p_h, B, T, inner, act, bf16, st);
        B2S_DW_CASE(1) B2S_DW_CASE(3) B2S_DW_CASE(5) B2S_DW_CASE(7) B2S_DW_CASE(9) B2S_DW_CASE(11) B2S_DW_CASE(13) B2S_DW_CASE(15)
        B2S_DW_CASE(17) B2S_DW_CASE(19) B2S_DW_CASE(21) B2S_DW_CASE(23) B2S_DW_CASE(25) B2S_DW_CASE(27) B2S_DW_CASE(29) B2S_DW_CASE(31)
#undef B2S_DW_CASE
    }
    set_error("b2s_lynx_dwconv_h: unsupported kernel size %d", ksize);
    return B2S_ERR_UNSUPPORTED;
}
