// NSF-HiFiGAN vocoder (mel + f0 -> waveform, the step AFTER the sampling loop; reference modules/nsf_hifigan/models.py:206-289): the
// kernels that are NOT GEMMs.  The dense convolutions (conv_pre, the transposed convs as 3-tap convs over u * C columns, the dilated
// residual-block convs) run on the tcgen05 kernels of b2s_tc_gemm.cu (b2s_tc_conv1d_dil / b2s_tc_conv1d_residual); activations are
// time-major rows r = b * T_i + t, folded (f samples per GEMM row) or zero-padded to a multiple of 64 channels (vocoder.py).
//   voc_phase        per-frame phase accumulation of the sine source                   models.py:138-141 (SineGen), :253-259 (mini_nsf)
//   voc_source       1 + 8 harmonics, uv / noise mix, tanh(linear(.))                  models.py:142-147, :160-166, :197-200; :260-262
//   voc_source_add   x += noise_conv(source) (strided Conv1d with ONE input channel), 16-bit leaky_relu(x) copy   models.py:273-278, :62
//   voc_avg_act      leaky_relu((x_0 + x_1 + ...) / num_kernels, slope) -> 16-bit      models.py:279-285, :271
//   voc_post         tanh(conv_post(leaky_relu(mean of the blocks, 0.01)))             models.py:285-288
//   cast_scale       fp32 -> 16 bit with a scale (log10 -> ln mel, x 2.30259)          vocoders/nsf_hifigan.py:60-64
#include "b2s_common.cuh"
#include "b2s_tc.cuh"

namespace b2s {

constexpr int VOC_MAX_HARM = 16;
constexpr int VOC_MAX_BLOCKS = 4;

// One warp per utterance.  last[t] = (f0[t] / sr) * upp (+ the mini_nsf chirp term) is the phase advance of frame t in cycles;
// phase[t] = fmod(sum_{t' < t} wrap(last[t']), 1) with wrap(v) = fmod(v + 0.5, 1) - 0.5.  The prefix sum is a warp scan over chunks
// of 32 frames (fp32, like torch's CUDA cumsum; the order of the additions differs from a sequential sum by rounding only).
__global__ void voc_phase_kernel(const float* __restrict__ f0, float* __restrict__ phase, int T, float sr, int upp, int mini) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const float* f = f0 + (long long)b * T;
    float carry = 0.f;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        float v = 0.f;
        if (t < T) {
            const float s0 = __fdiv_rn(f[t], sr);
            float last = __fmul_rn(s0, (float)upp);
            if (mini) {                                                  // rad = s0 n + 0.5 ds0 n (n - 1) / upp at n = upp  (:256)
                const float s1 = t + 1 < T ? __fdiv_rn(f[t + 1], sr) : 0.f;
                const float ds0 = t + 1 < T ? __fsub_rn(s1, s0) : 0.f;
                const float c = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(0.5f, ds0), (float)upp), (float)(upp - 1)), (float)upp);
                last = __fadd_rn(last, c);
            }
            v = __fsub_rn(fmodf(__fadd_rn(last, 0.5f), 1.0f), 0.5f);
        }
        float s = v;                                                      // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float n = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += n;
        }
        const float incl = carry + s;
        if (t < T) phase[(long long)b * T + t] = fmodf(incl - v, 1.0f);   // exclusive: the sum over the frames BEFORE t
        carry = __shfl_sync(0xffffffffu, incl, 31);
    }
}

// One thread per output sample n = t * upp + i of utterance b.
//   harmonics (dim > 0): out = tanh(bias + sum_h w[h] * (sin(2 pi ((s0 (i+1) + phase[t]) (h+1) + ini[h])) * amp * uv + namp * noise[n, h]))
//   mini_nsf  (dim == 0): out = sin(2 pi (s0 (i+1) + 0.5 ds0 (i+1) i / upp + phase[t]))
__global__ void voc_source_kernel(const float* __restrict__ f0, const float* __restrict__ phase, const float* __restrict__ ini,
                                  const float* __restrict__ noise, const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ out, int B, int T, int upp, int dim, float sr, float amp, float noise_std,
                                  float thr) {
    const long long total = (long long)B * T * upp;
    const float two_pi = 6.283185307179586f;
    for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < total; n += (long long)gridDim.x * blockDim.x) {
        const long long ft = n / upp;                                    // b * T + t
        const int i = (int)(n - ft * upp);
        const int t = (int)(ft % T);
        const float f = __ldg(f0 + ft);
        const float s0 = __fdiv_rn(f, sr);
        const float ph = __ldg(phase + ft);
        if (dim == 0) {
            const float s1 = t + 1 < T ? __fdiv_rn(__ldg(f0 + ft + 1), sr) : 0.f;
            const float ds0 = t + 1 < T ? __fsub_rn(s1, s0) : 0.f;
            const float nn = (float)(i + 1);
            const float chirp = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(0.5f, ds0), nn), (float)i), (float)upp);
            const float rad = __fadd_rn(__fadd_rn(__fmul_rn(s0, nn), chirp), ph);
            out[n] = sinf(__fmul_rn(two_pi, rad));
            continue;
        }
        const float rad = __fadd_rn(__fmul_rn(s0, (float)(i + 1)), ph);
        const float uv = f > thr ? 1.f : 0.f;
        const float namp = f > thr ? noise_std : __fdiv_rn(amp, 3.0f);
        float acc = __ldg(bias);
        for (int h = 0; h < dim; ++h) {
            const float r = __fadd_rn(__fmul_rn(rad, (float)(h + 1)), h == 0 ? 0.f : __ldg(ini + h));
            const float sw = __fmul_rn(sinf(__fmul_rn(two_pi, r)), amp);
            const float v = __fadd_rn(__fmul_rn(sw, uv), __fmul_rn(namp, __ldg(noise + n * dim + h)));
            acc = fmaf(__ldg(w + h), v, acc);
        }
        out[n] = tanhf(acc);
    }
}

// x[r, c] += bias[c] + sum_j Wt[j, c] * src[b, t * stride - pad + j]   (zero outside [0, n_src)); lx_h[r, c] = leaky_relu(x[r, c]).
// A thread owns 4 consecutive channels of one row.  ksize == 0: no source term (mini_nsf stages without one), only the 16-bit copy.
template <int BF16>
__global__ void __launch_bounds__(256) voc_source_add_kernel(float* __restrict__ x, void* __restrict__ lx_h, const float* __restrict__ src,
                                                            const float* __restrict__ Wt, const float* __restrict__ bias, int B,
                                                            int T, int Cp, int ksize, int stride, int pad, int n_src, float slope) {
    const int q = Cp >> 2;
    const long long total = (long long)B * T * q;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / q;
        const int c = (int)(idx - r * q) * 4;
        const int b = (int)(r / T), t = (int)(r - (long long)b * T);
        float4 v = *reinterpret_cast<const float4*>(x + r * Cp + c);
        if (ksize > 0) {
            float4 a = __ldg(reinterpret_cast<const float4*>(bias + c));
            const float* s = src + (long long)b * n_src;
            const long long base = (long long)t * stride - pad;
            // taps outside [0, n_src) are the conv's zero padding: clip the tap range once, then a branch-free loop the compiler can
            // unroll (the loads of several taps in flight; with a `continue` per tap the 128-tap stage-0 conv was a latency chain)
            const int j0 = base < 0 ? (int)(-base) : 0;
            const int j1 = base + ksize > n_src ? (int)(n_src - base) : ksize;
#pragma unroll 8
            for (int j = j0; j < j1; ++j) {
                const float sv = __ldg(s + base + j);
                const float4 w = __ldg(reinterpret_cast<const float4*>(Wt + (long long)j * Cp + c));
                a.x = fmaf(w.x, sv, a.x); a.y = fmaf(w.y, sv, a.y); a.z = fmaf(w.z, sv, a.z); a.w = fmaf(w.w, sv, a.w);
            }
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            *reinterpret_cast<float4*>(x + r * Cp + c) = v;
        }
        uint2 o;
        o.x = tc::Half16<BF16>::pack2(v.x > 0.f ? v.x : v.x * slope, v.y > 0.f ? v.y : v.y * slope);
        o.y = tc::Half16<BF16>::pack2(v.z > 0.f ? v.z : v.z * slope, v.w > 0.f ? v.w : v.w * slope);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(lx_h) + r * Cp + c) = o;
    }
}

struct VocBlocks {
    const float* x[VOC_MAX_BLOCKS];
    int n;
};

__device__ __forceinline__ float voc_mean(const VocBlocks& xs, long long i) {
    float s = xs.x[0][i];
    for (int k = 1; k < xs.n; ++k) s = __fadd_rn(s, xs.x[k][i]);          // xs += resblock(x), in block order (:281-284)
    return __fdiv_rn(s, (float)xs.n);                                     // x = xs / num_kernels (:285)
}

template <int BF16>
__global__ void __launch_bounds__(256) voc_avg_act_kernel(VocBlocks xs, void* __restrict__ out_h, long long n4, float slope) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 s = *reinterpret_cast<const float4*>(xs.x[0] + 4 * i);
        for (int k = 1; k < xs.n; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(xs.x[k] + 4 * i);
            s.x = __fadd_rn(s.x, v.x); s.y = __fadd_rn(s.y, v.y); s.z = __fadd_rn(s.z, v.z); s.w = __fadd_rn(s.w, v.w);
        }
        const float d = (float)xs.n;
        s.x = __fdiv_rn(s.x, d); s.y = __fdiv_rn(s.y, d); s.z = __fdiv_rn(s.z, d); s.w = __fdiv_rn(s.w, d);
        uint2 o;
        o.x = tc::Half16<BF16>::pack2(s.x > 0.f ? s.x : s.x * slope, s.y > 0.f ? s.y : s.y * slope);
        o.y = tc::Half16<BF16>::pack2(s.z > 0.f ? s.z : s.z * slope, s.w > 0.f ? s.w : s.w * slope);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out_h) + 4 * i) = o;
    }
}

// wav[b, t] = tanh(b0 + sum_{j < ksize} sum_{c < C} W[j, c] * lrelu(mean_k x_k[b, t + j - ksize / 2, c], slope)), fp32 throughout.
// A block stages the (256 + ksize - 1) activated rows of its 256 samples in shared memory (row stride C + 1: conflict-free).
constexpr int VOC_POST_TILE = 256;
__global__ void __launch_bounds__(VOC_POST_TILE) voc_post_kernel(VocBlocks xs, const float* __restrict__ W, const float* __restrict__ b0,
                                                                 float* __restrict__ wav, int T, int C, int Cp, int ksize,
                                                                 float slope) {
    extern __shared__ float tile[];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * VOC_POST_TILE;
    const int half = ksize / 2, rows = VOC_POST_TILE + ksize - 1, ld = C + 1;
    float* wsm = tile + rows * ld;
    if ((C & 3) == 0 && (Cp & 3) == 0) {                                  // 16-byte loads (every shipped geometry: C = 16)
        const int q = C >> 2;
        for (int i = threadIdx.x; i < rows * q; i += VOC_POST_TILE) {
            const int rr = i / q, c = (i - rr * q) * 4;
            const int t = t0 + rr - half;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && t < T) {
                const long long off = ((long long)b * T + t) * Cp + c;
                v = *reinterpret_cast<const float4*>(xs.x[0] + off);
                for (int k = 1; k < xs.n; ++k) {
                    const float4 u = *reinterpret_cast<const float4*>(xs.x[k] + off);
                    v.x = __fadd_rn(v.x, u.x); v.y = __fadd_rn(v.y, u.y); v.z = __fadd_rn(v.z, u.z); v.w = __fadd_rn(v.w, u.w);
                }
                const float d = (float)xs.n;
                v.x = __fdiv_rn(v.x, d); v.y = __fdiv_rn(v.y, d); v.z = __fdiv_rn(v.z, d); v.w = __fdiv_rn(v.w, d);
                v.x = v.x > 0.f ? v.x : __fmul_rn(v.x, slope); v.y = v.y > 0.f ? v.y : __fmul_rn(v.y, slope);
                v.z = v.z > 0.f ? v.z : __fmul_rn(v.z, slope); v.w = v.w > 0.f ? v.w : __fmul_rn(v.w, slope);
            }
            float* dst = tile + rr * ld + c;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
    } else {
        for (int i = threadIdx.x; i < rows * C; i += VOC_POST_TILE) {
            const int rr = i / C, c = i - rr * C;
            const int t = t0 + rr - half;
            float v = 0.f;
            if (t >= 0 && t < T) {
                v = voc_mean(xs, ((long long)b * T + t) * Cp + c);
                v = v > 0.f ? v : __fmul_rn(v, slope);
            }
            tile[rr * ld + c] = v;
        }
    }
    for (int i = threadIdx.x; i < ksize * C; i += VOC_POST_TILE) wsm[i] = __ldg(W + i);
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= T) return;
    float acc = 0.f;
    for (int j = 0; j < ksize; ++j) {
        const float* row = tile + (threadIdx.x + j) * ld;
        const float* wj = wsm + j * C;
        for (int c = 0; c < C; ++c) acc = fmaf(wj[c], row[c], acc);
    }
    wav[(long long)b * T + t] = tanhf(__fadd_rn(acc, __ldg(b0)));
}

template <int BF16>
__global__ void cast_scale_kernel(const float* __restrict__ in, void* __restrict__ out, long long n, float scale) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = __fmul_rn(in[i], scale);
        if (BF16) reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
        else reinterpret_cast<__half*>(out)[i] = __float2half_rn(v);
    }
}

static int grid_for(long long n, int block) {
    long long g = (n + block - 1) / block;
    const long long cap = 148LL * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static bool al16v(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_voc_phase(const float* f0, float* phase, int B, int T, float sr, int upp, int mini_nsf, void* stream) {
    B2S_CHECK_ARG(f0 && phase && B >= 0 && T >= 0 && upp >= 1 && sr > 0.f, "b2s_voc_phase: bad arguments");
    if (B * T == 0) return B2S_OK;
    voc_phase_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(f0, phase, T, sr, upp, mini_nsf);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_voc_source(const float* f0, const float* phase, const float* rand_ini, const float* noise, const float* w,
                              const float* bias, float* out, int B, int T, int upp, int dim, float sr, float sine_amp, float noise_std,
                              float voiced_threshold, void* stream) {
    B2S_CHECK_ARG(f0 && phase && out, "b2s_voc_source: null pointer");
    B2S_CHECK_ARG(dim >= 0 && dim <= VOC_MAX_HARM && (dim == 0 || (rand_ini && noise && w && bias)),
                  "b2s_voc_source: dim=%d harmonics need rand_ini, noise, w and bias", dim);
    const long long n = (long long)B * T * upp;
    if (n == 0) return B2S_OK;
    voc_source_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(f0, phase, rand_ini, noise, w, bias, out, B, T, upp, dim, sr,
                                                                         sine_amp, noise_std, voiced_threshold);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_voc_source_add(float* x, void* lx_h, const float* src, const float* Wt, const float* bias, int B, int T, int Cp,
                                  int ksize, int stride, int pad, int n_src, float slope, int bf16, void* stream) {
    B2S_CHECK_ARG(x && lx_h && Cp > 0 && Cp % 4 == 0 && al16v(x) && al16v(lx_h), "b2s_voc_source_add: bad x / lx_h / Cp=%d", Cp);
    B2S_CHECK_ARG(ksize == 0 || (src && Wt && bias && stride >= 1 && al16v(Wt) && al16v(bias)), "b2s_voc_source_add: bad source conv");
    const long long n = (long long)B * T * (Cp / 4);
    if (n == 0) return B2S_OK;
    if (bf16) voc_source_add_kernel<1><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, lx_h, src, Wt, bias, B, T, Cp, ksize, stride, pad, n_src, slope);
    else voc_source_add_kernel<0><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, lx_h, src, Wt, bias, B, T, Cp, ksize, stride, pad, n_src, slope);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_voc_avg_act(const float* const* xs_host, int n_blocks, void* out_h, int64_t n, float slope, int bf16, void* stream) {
    B2S_CHECK_ARG(xs_host && out_h && n_blocks >= 1 && n_blocks <= VOC_MAX_BLOCKS && n % 4 == 0 && al16v(out_h),
                  "b2s_voc_avg_act: 1..%d blocks, n %% 4 == 0 (n_blocks=%d)", VOC_MAX_BLOCKS, n_blocks);
    VocBlocks xs{};
    xs.n = n_blocks;
    for (int k = 0; k < n_blocks; ++k) {
        B2S_CHECK_ARG(xs_host[k] && al16v(xs_host[k]), "b2s_voc_avg_act: null / misaligned block %d", k);
        xs.x[k] = xs_host[k];
    }
    if (n == 0) return B2S_OK;
    if (bf16) voc_avg_act_kernel<1><<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(xs, out_h, n / 4, slope);
    else voc_avg_act_kernel<0><<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(xs, out_h, n / 4, slope);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_voc_post(const float* const* xs_host, int n_blocks, const float* W, const float* b0, float* wav, int B, int T,
                            int C, int Cp, int ksize, float slope, void* stream) {
    B2S_CHECK_ARG(xs_host && W && b0 && wav && n_blocks >= 1 && n_blocks <= VOC_MAX_BLOCKS && C >= 1 && C <= Cp && (ksize & 1),
                  "b2s_voc_post: bad arguments (n_blocks=%d C=%d Cp=%d ksize=%d)", n_blocks, C, Cp, ksize);
    const size_t smem = ((size_t)(VOC_POST_TILE + ksize - 1) * (C + 1) + (size_t)ksize * C) * sizeof(float);
    B2S_CHECK_ARG(smem <= 48 * 1024, "b2s_voc_post: %d channels x %d taps need %zu bytes of shared memory (max 48 KB)", C, ksize, smem);
    VocBlocks xs{};
    xs.n = n_blocks;
    for (int k = 0; k < n_blocks; ++k) {
        B2S_CHECK_ARG(xs_host[k] != nullptr, "b2s_voc_post: null block %d", k);
        xs.x[k] = xs_host[k];
    }
    if (B * T == 0) return B2S_OK;
    B2S_CHECK_ARG(B <= 65535, "b2s_voc_post: at most 65535 utterances per call (B=%d)", B);
    dim3 grid((T + VOC_POST_TILE - 1) / VOC_POST_TILE, B);
    voc_post_kernel<<<grid, VOC_POST_TILE, smem, (cudaStream_t)stream>>>(xs, W, b0, wav, T, C, Cp, ksize, slope);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_cast_scale_f32_h(const float* in, void* out_h, int64_t n, float scale, int bf16, void* stream) {
    B2S_CHECK_ARG(in && out_h && n >= 0, "b2s_cast_scale_f32_h: bad arguments");
    if (n == 0) return B2S_OK;
    if (bf16) cast_scale_kernel<1><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out_h, n, scale);
    else cast_scale_kernel<0><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out_h, n, scale);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}
