// HBM-bound helper kernels of the sampling path: layout transposes, the sampler update
// (linear combination), the sinusoidal step embedding, LayerNorm(+cond/step add) and the LYNXNet
// depthwise convolution.  All fp32, vectorised, grid sized from the problem (grid-stride where useful).
#include <stdlib.h>
#include "b2s_common.cuh"

#include <string.h>

namespace b2s {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// [batch, rows, cols] -> [batch, cols, rows]; 32x32 tile through shared memory, both sides coalesced
// ------------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    const long long boff = (long long)blockIdx.z * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[boff + (long long)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[boff + (long long)c * rows + r] = tile[threadIdx.x][i];
    }
}

// ------------------------------------------------------------------------------------------------
// dst = sum_i coef[i] * src[i]
// ------------------------------------------------------------------------------------------------
struct LinCombArgs {
    const float* src[8];
    float* dst;
    const float* coef;
    int n_src;
    long long n4;     // number of float4 groups
    long long n;      // total elements
    uint16_t* out_h;  // optional 16-bit copy of dst (the denoiser's A operand), else null
    int bf16;
    int* zero;        // optional tile flags of the persistent denoiser kernel to re-arm
    int nzero;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b, int bf16) {
    if (bf16) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int NS>
__global__ void __launch_bounds__(256) lincomb_kernel(const LinCombArgs a) {
    // programmatic dependent launch: overlap this kernel's launch with the tail of its predecessor
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    float c[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) c[i] = __ldg(a.coef + i);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.nzero; i += (int)stride) a.zero[i] = 0;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < a.n4; idx += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(a.src[i] + idx * 4);
            acc.x = fmaf(c[i], v.x, acc.x);
            acc.y = fmaf(c[i], v.y, acc.y);
            acc.z = fmaf(c[i], v.z, acc.z);
            acc.w = fmaf(c[i], v.w, acc.w);
        }
        *reinterpret_cast<float4*>(a.dst + idx * 4) = acc;
        if (a.out_h) {
            uint2 h;
            h.x = pack_h2(acc.x, acc.y, a.bf16);
            h.y = pack_h2(acc.z, acc.w, a.bf16);
            *reinterpret_cast<uint2*>(a.out_h + idx * 4) = h;
        }
    }
    // scalar tail (n % 4)
    if (blockIdx.x == 0 && threadIdx.x < (a.n - a.n4 * 4)) {
        long long idx = a.n4 * 4 + threadIdx.x;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < NS; ++i) acc = fmaf(c[i], a.src[i][idx], acc);
        a.dst[idx] = acc;
        if (a.out_h) a.out_h[idx] = (uint16_t)(pack_h2(acc, 0.f, a.bf16) & 0xFFFFu);
    }
}

// ------------------------------------------------------------------------------------------------
// SinusoidalPosEmb (common_layers.py:266-278): freq_j = exp(-j * ln(1e4)/(half-1)) in fp32,
// out = [sin(t*freq), cos(t*freq)]
// ------------------------------------------------------------------------------------------------
__global__ void sinusoid_kernel(const float* __restrict__ t, float* __restrict__ out, int n, int dim) {
    const int half = dim / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * half) return;
    const int k = idx / half, j = idx - k * half;
    const float step = (float)(9.210340371976184 / (double)(half - 1));   // ln(10000)/(half-1), rounded as fp32 like torch
    const float freq = expf((float)j * -step);
    const float arg = t[k] * freq;
    out[(long long)k * dim + j] = sinf(arg);
    out[(long long)k * dim + half + j] = cosf(arg);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over channels, one warp per frame row; optional fused (x + cond + d) prologue with the
// residual write-back of LYNXNetResidualLayer (lynxnet.py:76-84).  eps = 1e-5 (nn.LayerNorm default).
// ------------------------------------------------------------------------------------------------
template <bool PRE>
__global__ void __launch_bounds__(256) layernorm_kernel(float* __restrict__ x, const float* __restrict__ cond, int ld_cond,
                                                        const float* __restrict__ dvec, int d_stride,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ h, int rows, int T, int C, int strong) {
    extern __shared__ float srow[];                 // [warps][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= rows) return;
    float* u = srow + (long long)warp * C;
    const int b = PRE ? r / T : 0;
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) {
        float v = x[(long long)r * C + c];
        if (PRE) {
            const float cc = cond[(long long)r * ld_cond + c];
            const float xc = v + cc;
            if (strong) { x[(long long)r * C + c] = xc; }       // res = x + cond  (front_cond_inject)
            v = xc + dvec[(long long)b * d_stride + c];
        }
        u[c] = v;
        sum += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float var = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float d = u[c] - mean;
        var += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)C + 1e-5f);
    for (int c = lane; c < C; c += 32) h[(long long)r * C + c] = (u[c] - mean) * rstd * gamma[c] + beta[c];
}

// ------------------------------------------------------------------------------------------------
// Depthwise conv along time + bias + activation (lynxnet.py:57-58).  Time-major: thread = channel
// (coalesced across channels), loops over a strip of frames keeping nothing but the weights in
// registers; zero padding at utterance edges.
// ------------------------------------------------------------------------------------------------
constexpr int DW_TSTRIP = 16;
constexpr int DW_MAXK = 31;

__global__ void __launch_bounds__(128) dwconv_kernel(const float* __restrict__ g, const float* __restrict__ Wdw,
                                                     const float* __restrict__ bias, const float* __restrict__ slope,
                                                     float* __restrict__ p, int T, int inner, int ksize, int act) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= inner) return;
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * DW_TSTRIP;
    const int pad = ksize / 2;
    float w[DW_MAXK];
#pragma unroll
    for (int k = 0; k < DW_MAXK; ++k) w[k] = k < ksize ? __ldg(Wdw + (long long)ch * ksize + k) : 0.f;
    const float bb = __ldg(bias + ch);
    const float sl = slope ? __ldg(slope + ch) : 0.f;
    const float* gb = g + (long long)b * T * inner + ch;
    float* pb = p + (long long)b * T * inner + ch;
    // sliding window over DW_TSTRIP + ksize - 1 input frames
    float win[DW_MAXK];
#pragma unroll
    for (int k = 0; k < DW_MAXK; ++k) {
        const int ts = t0 - pad + k;
        win[k] = (k < ksize && ts >= 0 && ts < T) ? gb[(long long)ts * inner] : 0.f;
    }
    for (int i = 0; i < DW_TSTRIP; ++i) {
        const int t = t0 + i;
        if (t >= T) break;
        float acc = bb;
#pragma unroll
        for (int k = 0; k < DW_MAXK; ++k) acc = fmaf(w[k], win[k], acc);
        float o;
        if (act == 0) o = acc >= 0.f ? acc : sl * acc;        // PReLU
        else o = apply_act(acc, act);
        pb[(long long)t * inner] = o;
        // slide
#pragma unroll
        for (int k = 0; k < DW_MAXK - 1; ++k) win[k] = win[k + 1];
        const int tn = t + 1 - pad + (ksize - 1);
        const float nv = (tn >= 0 && tn < T) ? gb[(long long)tn * inner] : 0.f;
        // place the incoming frame at index ksize-1 (the window is left-aligned)
#pragma unroll
        for (int k = 0; k < DW_MAXK; ++k)
            if (k == ksize - 1) win[k] = nv;
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_abi_version(void) { return B2S_ABI_VERSION; }
extern "C" const char* b2s_last_error(void) { return g_err; }

extern "C" int b2s_transpose_f32(const float* in, float* out, int batch, int rows, int cols, void* stream) {
    B2S_CHECK_ARG(batch >= 0 && rows >= 0 && cols >= 0 && batch < 65536, "b2s_transpose_f32: bad dims");
    if (batch == 0 || rows == 0 || cols == 0) return B2S_OK;       // empty input: nothing to do, pointers may be null
    B2S_CHECK_ARG(in && out, "b2s_transpose_f32: null pointer");
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), batch), block(32, 8);
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, rows, cols);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

static int lincomb_impl(const char* who, float* dst, const float* const* srcs_host, const float* coef, int n_src, int64_t n,
                        void* out_h, int bf16, int* flags, int n_flags, void* stream) {
    B2S_CHECK_ARG(n_src >= 1 && n_src <= 8, "%s: n_src must be in [1, 8] (got %d)", who, n_src);
    if (n <= 0) return B2S_OK;
    B2S_CHECK_ARG(dst && srcs_host && coef, "%s: null pointer", who);
    LinCombArgs a{};
    for (int i = 0; i < n_src; ++i) {
        B2S_CHECK_ARG(srcs_host[i] && (reinterpret_cast<uintptr_t>(srcs_host[i]) & 15) == 0,
                      "%s: src %d null or not 16B aligned", who, i);
        a.src[i] = srcs_host[i];
    }
    B2S_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "%s: dst not 16B aligned", who);
    B2S_CHECK_ARG((reinterpret_cast<uintptr_t>(out_h) & 7) == 0, "%s: out_h not 8B aligned", who);
    B2S_CHECK_ARG(n_flags == 0 || flags, "%s: null flags", who);
    a.dst = dst; a.coef = coef; a.n_src = n_src; a.n = n; a.n4 = n / 4;
    a.out_h = (uint16_t*)out_h; a.bf16 = bf16; a.zero = flags; a.nzero = n_flags;
    long long want = (a.n4 + 255) / 256;
    int blocks = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const bool ew_pdl = getenv("B2S_EW_PDL") != nullptr && atoi(getenv("B2S_EW_PDL")) != 0;
    cfg.numAttrs = ew_pdl ? 1 : 0;
    cudaError_t err;
    switch (n_src) {
        case 1: err = cudaLaunchKernelEx(&cfg, lincomb_kernel<1>, a); break;
        case 2: err = cudaLaunchKernelEx(&cfg, lincomb_kernel<2>, a); break;
        case 3: err = cudaLaunchKernelEx(&cfg, lincomb_kernel<3>, a); break;
        case 4: err = cudaLaunchKernelEx(&cfg, lincomb_kernel<4>, a); break;
        case 5: err = cudaLaunchKernelEx(&cfg, lincomb_kernel<5>, a); break;
        case 6: err = cudaLaunchKernelEx(&cfg, lincomb_kernel<6>, a); break;
        case 7: err = cudaLaunchKernelEx(&cfg, lincomb_kernel<7>, a); break;
        default: err = cudaLaunchKernelEx(&cfg, lincomb_kernel<8>, a); break;
    }
    B2S_CHECK_CUDA(err);
    return B2S_OK;
}

// ------------------------------------------------------------------------------------------------
// norm_spec / denorm_spec (ddpm.py:379-383, reflow.py:140-144) fused with the layout change between the caller's
// [B, T, M] / [B, F, T, M] and the sampler's time-major state [B*T, F*M] (the reference transposes to [B, F, M, T] at
// ddpm.py:370-373 and back at :350): one pass each, once per sampling call
// ------------------------------------------------------------------------------------------------
template <int INVERSE>
__global__ void spec_norm_kernel(const float* __restrict__ in, const float* __restrict__ lo, const float* __restrict__ hi,
                                 float* __restrict__ out, int B, int F, int T, int M) {
    const long long n = (long long)B * F * T * M;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // i indexes the time-major state [b, t, f, m]; the spec tensor is [b, f, t, m]
        const int m = (int)(i % M);
        long long r = i / M;
        const int f = (int)(r % F);
        r /= F;
        const int t = (int)(r % T);
        const long long b = r / T;
        const long long j = ((b * F + f) * T + t) * M + m;
        const float l = __ldg(lo + f * M + m), h = __ldg(hi + f * M + m);
        // the reference's expressions op by op (round-to-nearest intrinsics: no FMA contraction), so the result is bit-identical to
        // torch's (x + 1) / 2 * (max - min) + min  and  (x - min) / (max - min) * 2 - 1
        if (INVERSE) out[j] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(in[i], 1.0f), 2.0f), __fsub_rn(h, l)), l);
        else out[i] = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(in[j], l), __fsub_rn(h, l)), 2.0f), 1.0f);
    }
}

static int spec_norm_impl(const char* who, int inverse, const float* in, const float* lo, const float* hi, float* out, int B, int F,
                          int T, int M, void* stream) {
    B2S_CHECK_ARG(B >= 0 && F >= 1 && T >= 0 && M >= 1, "%s: bad dims B=%d F=%d T=%d M=%d", who, B, F, T, M);
    const long long n = (long long)B * F * T * M;
    if (n == 0) return B2S_OK;
    B2S_CHECK_ARG(in && lo && hi && out, "%s: null pointer", who);
    long long want = (n + 255) / 256;
    const int blocks = (int)(want > 148 * 16 ? 148 * 16 : want);
    if (inverse) spec_norm_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, lo, hi, out, B, F, T, M);
    else spec_norm_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, lo, hi, out, B, F, T, M);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_spec_norm_f32(const float* spec, const float* spec_min, const float* spec_max, float* state, int B, int F, int T,
                                 int M, void* stream) {
    return spec_norm_impl("b2s_spec_norm_f32", 0, spec, spec_min, spec_max, state, B, F, T, M, stream);
}
extern "C" int b2s_spec_denorm_f32(const float* state, const float* spec_min, const float* spec_max, float* spec, int B, int F, int T,
                                   int M, void* stream) {
    return spec_norm_impl("b2s_spec_denorm_f32", 1, state, spec_min, spec_max, spec, B, F, T, M, stream);
}

extern "C" int b2s_sampler_lincomb_f32(float* dst, const float* const* srcs_host, const float* coef, int n_src,
                                       int64_t n, void* stream) {
    return lincomb_impl("b2s_sampler_lincomb_f32", dst, srcs_host, coef, n_src, n, nullptr, 0, nullptr, 0, stream);
}

extern "C" int b2s_sampler_lincomb_f32_h(float* dst, const float* const* srcs_host, const float* coef, int n_src, int64_t n,
                                         void* out_h, int bf16, int* flags, int n_flags, void* stream) {
    B2S_CHECK_ARG(out_h || n <= 0, "b2s_sampler_lincomb_f32_h: null out_h");
    return lincomb_impl("b2s_sampler_lincomb_f32_h", dst, srcs_host, coef, n_src, n, out_h, bf16, flags, n_flags, stream);
}

extern "C" int b2s_sinusoid_f32(const float* t, float* out, int n, int dim, void* stream) {
    B2S_CHECK_ARG(t && out, "b2s_sinusoid_f32: null pointer");
    B2S_CHECK_ARG(dim >= 4 && dim % 2 == 0, "b2s_sinusoid_f32: dim must be even and >= 4");
    if (n <= 0) return B2S_OK;
    int total = n * (dim / 2);
    sinusoid_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(t, out, n, dim);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_lynx_prenorm_f32(float* x, const float* cond, int ld_cond, const float* dvec, int d_stride,
                                    const float* gamma, const float* beta, float* h, int B, int T, int C,
                                    int strong_cond, void* stream) {
    B2S_CHECK_ARG(x && cond && dvec && gamma && beta && h, "b2s_lynx_prenorm_f32: null pointer");
    B2S_CHECK_ARG(C > 0 && C <= 8192, "b2s_lynx_prenorm_f32: C out of range");
    const int rows = B * T;
    if (rows <= 0) return B2S_OK;
    const int warps = C <= 1024 ? 8 : (C <= 2048 ? 4 : 1);
    size_t smem = (size_t)warps * C * sizeof(float);
    layernorm_kernel<true><<<ceil_div(rows, warps), warps * 32, smem, (cudaStream_t)stream>>>(
        x, cond, ld_cond, dvec, d_stride, gamma, beta, h, rows, T, C, strong_cond);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_layernorm_f32(const float* x, const float* gamma, const float* beta, float* h, int rows, int C,
                                 void* stream) {
    B2S_CHECK_ARG(x && gamma && beta && h, "b2s_layernorm_f32: null pointer");
    B2S_CHECK_ARG(C > 0 && C <= 8192, "b2s_layernorm_f32: C out of range");
    if (rows <= 0) return B2S_OK;
    const int warps = C <= 1024 ? 8 : (C <= 2048 ? 4 : 1);
    size_t smem = (size_t)warps * C * sizeof(float);
    layernorm_kernel<false><<<ceil_div(rows, warps), warps * 32, smem, (cudaStream_t)stream>>>(
        const_cast<float*>(x), nullptr, 0, nullptr, 0, gamma, beta, h, rows, 1, C, 0);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_lynx_dwconv_f32(const float* g, const float* Wdw, const float* bias, const float* slope, float* p,
                                   int B, int T, int inner, int ksize, int act, void* stream) {
    B2S_CHECK_ARG(g && Wdw && bias && p, "b2s_lynx_dwconv_f32: null pointer");
    B2S_CHECK_ARG(ksize >= 1 && ksize <= DW_MAXK && (ksize & 1), "b2s_lynx_dwconv_f32: kernel size must be odd and <= %d", DW_MAXK);
    B2S_CHECK_ARG(act != 0 || slope, "b2s_lynx_dwconv_f32: PReLU needs slope");
    B2S_CHECK_ARG(B < 65536, "b2s_lynx_dwconv_f32: B too large");
    if (B <= 0 || T <= 0) return B2S_OK;
    dim3 grid(ceil_div(inner, 128), ceil_div(T, DW_TSTRIP), B);
    dwconv_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(g, Wdw, bias, slope, p, T, inner, ksize, act);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

// ---- training-branch losses (forward values: validation, SURVEY.md section 8 row f-4 forward half) ---------------------------------
// mean over [B, F, M, T] of w_b * loss(a * m, b * m), loss = |.| or (.)^2, m = non_padding[b, t, (m)] (modules/losses/diff_loss.py:17-37,
// reflow_loss.py:18-50; the log-normal time weights of reflow_loss.py:26-33 are computed here from t).  Two kernels: per-block partial
// sums (fp32 inside a thread, double across threads), then ONE block adds the partials in a fixed order - deterministic.
namespace b2s {
constexpr int LOSS_BLOCK = 256, LOSS_MAX_BLOCKS = 1024;

__global__ void __launch_bounds__(LOSS_BLOCK) masked_loss_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                         const float* __restrict__ mask, int mask_m,
                                                                         const float* __restrict__ tw, int B, int F, int M, int T, int l1,
                                                                         double* __restrict__ partial) {
    const long long n = (long long)B * F * M * T;
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)LOSS_BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * LOSS_BLOCK) {
        const int t = (int)(i % T);
        const long long r = i / T;
        const int m = (int)(r % M);
        const int bb = (int)(r / ((long long)M * F));
        float x = a[i], y = b[i];
        if (mask) {
            const float k = __ldg(mask + ((long long)bb * T + t) * mask_m + (mask_m > 1 ? m : 0));
            x = __fmul_rn(x, k);
            y = __fmul_rn(y, k);
        }
        const float d = __fsub_rn(x, y);
        float v = l1 ? fabsf(d) : __fmul_rn(d, d);
        if (tw) {                                                     // reflow_loss.py:26-33
            const float eps = 1e-7f;
            const float tt = fminf(fmaxf(__ldg(tw + bb), eps), 1.0f - eps);
            const float lg = logf(__fdiv_rn(tt, 1.0f - tt));
            const float w = __fadd_rn(__fmul_rn(__fdiv_rn(__fdiv_rn(0.398942f, tt), 1.0f - tt), expf(__fmul_rn(-0.5f, __fmul_rn(lg, lg)))), eps);
            v = __fmul_rn(w, v);
        }
        acc += (double)v;
    }
    __shared__ double sh[LOSS_BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < LOSS_BLOCK / 32; ++w) s += sh[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void masked_loss_final_kernel(const double* __restrict__ partial, int n_partial, double inv_n, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n_partial; ++i) s += partial[i];
        out[0] = (float)(s * inv_n);
    }
}
}  // namespace b2s

extern "C" int b2s_masked_loss_workspace_bytes(void) { return b2s::LOSS_MAX_BLOCKS * (int)sizeof(double); }

extern "C" int b2s_masked_loss_f32(const float* a, const float* b, const float* mask, int mask_m, const float* t_weights, int B, int F,
                                   int M, int T, int l1, void* workspace, float* out, void* stream) {
    using namespace b2s;
    B2S_CHECK_ARG(a && b && workspace && out, "b2s_masked_loss_f32: null pointer");
    B2S_CHECK_ARG(B > 0 && F > 0 && M > 0 && T > 0, "b2s_masked_loss_f32: empty tensor (the reference's mean of nothing is NaN)");
    B2S_CHECK_ARG(!mask || mask_m == 1 || mask_m == M, "b2s_masked_loss_f32: mask [B, T, %d] does not broadcast over %d bins", mask_m, M);
    const long long n = (long long)B * F * M * T;
    long long g = (n + LOSS_BLOCK - 1) / LOSS_BLOCK;
    const int grid = (int)(g > LOSS_MAX_BLOCKS ? LOSS_MAX_BLOCKS : g);
    masked_loss_partial_kernel<<<grid, LOSS_BLOCK, 0, (cudaStream_t)stream>>>(a, b, mask, mask_m, t_weights, B, F, M, T, l1,
                                                                            reinterpret_cast<double*>(workspace));
    B2S_CHECK_LAUNCH();
    masked_loss_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(workspace), grid, 1.0 / (double)n, out);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}
