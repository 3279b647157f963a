// fp32 (CUDA-core FFMA) GEMM family with fused prologue gathers and epilogues.
//
// This is the REFERENCE-PRECISION path (max-abs mel error <= 1e-3 vs the reference's fp32 PyTorch):
// SURVEY.md H3 shows single-pass TF32 cannot meet that bar, so operands stay true fp32.
// One templated kernel: out[M,N] = epi(A[M,K] . W[N,K]^T) with
//   * A gather:   plain rows, or the dilated k=3 conv expressed as an implicit GEMM over K = 3*C
//                 (tap-major), reading y at t-d, t, t+d with zero fill outside [0,T) per utterance
//                 (wavenet.py:22-28,38);
//   * epilogues:  LINEAR (bias/alpha/activation, optional y = out + step embedding), GATE
//                 (sigmoid*tanh, wavenet.py:41-42), RESSKIP (wavenet.py:47-48), SWIGLU
//                 (common_layers.py:116-117), RESIDUAL (lynxnet.py:86).
// Tile 128x128x16, 256 threads, 8x8 register micro-tile, register-staged double buffering.
#include "b2s_common.cuh"

namespace b2s {

constexpr int BM = 128, BN = 128, BK = 16, SPAD = 4, NTHREADS = 256;
constexpr int LDS_ = BM + SPAD;

enum Epi : int { EPI_LINEAR = 0, EPI_GATE = 1, EPI_RESSKIP = 2, EPI_SWIGLU = 3, EPI_RESIDUAL = 4 };

struct GemmP {
    const float* A; int lda;
    const float* W; int ldw;
    int M, N, K;
    int T, C, dil;                 // conv gather (template CONV): K == 3*C, rows are (b, t)
    const float* bias; float alpha; int act;
    float* out; int ldo;
    const float* add; int ldadd;   // GATE: cond table of this layer; RESIDUAL: residual input
    float* x; float* y; float* skip;
    const float* dvec; int d_stride; int first;
};

template <bool CONV>
__device__ __forceinline__ void load_tile_regs(const GemmP& p, bool is_a, int row_g, bool row_ok, int b, int t,
                                               int k0, float4& v0, float4& v1) {
    v0 = make_float4(0.f, 0.f, 0.f, 0.f);
    v1 = v0;
    if (!row_ok) return;
    const float* src;
    bool ok0, ok1;
    if (is_a) {
        if (CONV) {
            int tap = k0 / p.C;
            int c = k0 - tap * p.C;
            int ts = t + (tap - 1) * p.dil;
            if (tap > 2 || ts < 0 || ts >= p.T) return;
            src = p.A + ((long long)(b * p.T + ts)) * p.lda + c;
            ok0 = ok1 = true;                       // C % 16 == 0: an 8-wide slab never straddles taps
        } else {
            src = p.A + (long long)row_g * p.lda + k0;
            ok0 = k0 < p.K;
            ok1 = k0 + 4 < p.K;
        }
    } else {
        src = p.W + (long long)row_g * p.ldw + k0;
        ok0 = k0 < p.K;
        ok1 = k0 + 4 < p.K;
    }
    if (ok0) v0 = __ldg(reinterpret_cast<const float4*>(src));
    if (ok1) v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
}

__device__ __forceinline__ void store_tile_smem(float* S, int row, int kq, const float4& v0, const float4& v1) {
    float* d = S + (kq * 8) * LDS_ + row;
    d[0 * LDS_] = v0.x; d[1 * LDS_] = v0.y; d[2 * LDS_] = v0.z; d[3 * LDS_] = v0.w;
    d[4 * LDS_] = v1.x; d[5 * LDS_] = v1.y; d[6 * LDS_] = v1.z; d[7 * LDS_] = v1.w;
}

template <int EPI, bool CONV>
__global__ void __launch_bounds__(NTHREADS, 2) sgemm_fused_kernel(const GemmP p) {
    __shared__ __align__(16) float As[2][BK * LDS_];
    __shared__ __align__(16) float Bs[2][BK * LDS_];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // loader mapping: 128 rows x 2 k-halves
    const int lrow = tid & 127, lkq = tid >> 7;
    const int a_row = m0 + lrow, w_row = n0 + lrow;
    const bool a_ok = a_row < p.M, w_ok = w_row < p.N;
    int ab = 0, at = 0;
    if (CONV && a_ok) { ab = a_row / p.T; at = a_row - ab * p.T; }

    const int tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (p.K + BK - 1) / BK;
    float4 ra0, ra1, rb0, rb1;
    load_tile_regs<CONV>(p, true, a_row, a_ok, ab, at, lkq * 8, ra0, ra1);
    load_tile_regs<CONV>(p, false, w_row, w_ok, 0, 0, lkq * 8, rb0, rb1);
    store_tile_smem(As[0], lrow, lkq, ra0, ra1);
    store_tile_smem(Bs[0], lrow, lkq, rb0, rb1);
    __syncthreads();

    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_tile_regs<CONV>(p, true, a_row, a_ok, ab, at, (kt + 1) * BK + lkq * 8, ra0, ra1);
            load_tile_regs<CONV>(p, false, w_row, w_ok, 0, 0, (kt + 1) * BK + lkq * 8, rb0, rb1);
        }
        const float* as = As[cur];
        const float* bs = Bs[cur];
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(as + k * LDS_ + ty * 4);
            float4 a1 = *reinterpret_cast<const float4*>(as + k * LDS_ + 64 + ty * 4);
            float4 b0 = *reinterpret_cast<const float4*>(bs + k * LDS_ + tx * 4);
            float4 b1 = *reinterpret_cast<const float4*>(bs + k * LDS_ + 64 + tx * 4);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tile_smem(As[cur ^ 1], lrow, lkq, ra0, ra1);
            store_tile_smem(Bs[cur ^ 1], lrow, lkq, rb0, rb1);
        }
        __syncthreads();
    }

    // ---------------- epilogue ----------------
    const float inv_sqrt2 = 0.70710678118654752440f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
        if (r >= p.M) continue;
        const int b = (p.T > 0) ? r / p.T : 0;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int col = n0 + jh * 64 + tx * 4;
            if (col >= p.N) continue;
            float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
            if (EPI == EPI_LINEAR) {
                float4 o;
                float* po = &o.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float t = p.alpha * v[j] + (p.bias ? __ldg(p.bias + col + j) : 0.f);
                    po[j] = apply_act(t, p.act);
                }
                *reinterpret_cast<float4*>(p.out + (long long)r * p.ldo + col) = o;
                if (p.y) {
                    const float4 d = __ldg(reinterpret_cast<const float4*>(p.dvec + (long long)b * p.d_stride + col));
                    float4 yy = make_float4(o.x + d.x, o.y + d.y, o.z + d.z, o.w + d.w);
                    *reinterpret_cast<float4*>(p.y + (long long)r * p.ldo + col) = yy;
                }
            } else if (EPI == EPI_GATE) {
                const float4 c = __ldg(reinterpret_cast<const float4*>(p.add + (long long)r * p.ldadd + col));
                float2 z;
                z.x = sigmoid_acc(v[0] + c.x) * tanhf(v[1] + c.y);
                z.y = sigmoid_acc(v[2] + c.z) * tanhf(v[3] + c.w);
                *reinterpret_cast<float2*>(p.out + (long long)r * p.ldo + (col >> 1)) = z;
            } else if (EPI == EPI_SWIGLU) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                float2 z;
                float g0 = v[1] + bb.y, g1 = v[3] + bb.w;
                z.x = (v[0] + bb.x) * (g0 * sigmoid_acc(g0));
                z.y = (v[2] + bb.z) * (g1 * sigmoid_acc(g1));
                *reinterpret_cast<float2*>(p.out + (long long)r * p.ldo + (col >> 1)) = z;
            } else if (EPI == EPI_RESIDUAL) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                const float4 rr = *reinterpret_cast<const float4*>(p.add + (long long)r * p.ldadd + col);
                float4 o = make_float4(v[0] + bb.x + rr.x, v[1] + bb.y + rr.y, v[2] + bb.z + rr.z, v[3] + bb.w + rr.w);
                *reinterpret_cast<float4*>(p.out + (long long)r * p.ldo + col) = o;
            } else {  // EPI_RESSKIP
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                if (col < p.C) {
                    float* px = p.x + (long long)r * p.C + col;
                    const float4 xo = *reinterpret_cast<const float4*>(px);
                    float4 xn = make_float4((xo.x + (v[0] + bb.x)) * inv_sqrt2, (xo.y + (v[1] + bb.y)) * inv_sqrt2,
                                            (xo.z + (v[2] + bb.z)) * inv_sqrt2, (xo.w + (v[3] + bb.w)) * inv_sqrt2);
                    *reinterpret_cast<float4*>(px) = xn;
                    if (p.y) {
                        const float4 d = __ldg(reinterpret_cast<const float4*>(p.dvec + (long long)b * p.d_stride + col));
                        float4 yy = make_float4(xn.x + d.x, xn.y + d.y, xn.z + d.z, xn.w + d.w);
                        *reinterpret_cast<float4*>(p.y + (long long)r * p.C + col) = yy;
                    }
                } else {
                    float* ps = p.skip + (long long)r * p.C + (col - p.C);
                    float4 s = make_float4(v[0] + bb.x, v[1] + bb.y, v[2] + bb.z, v[3] + bb.w);
                    if (!p.first) {
                        const float4 so = *reinterpret_cast<const float4*>(ps);
                        s.x += so.x; s.y += so.y; s.z += so.z; s.w += so.w;
                    }
                    *reinterpret_cast<float4*>(ps) = s;
                }
            }
        }
    }
}

template <int EPI, bool CONV>
static int launch_gemm(const GemmP& p, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0) return B2S_OK;
    dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
    sgemm_fused_kernel<EPI, CONV><<<grid, NTHREADS, 0, st>>>(p);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_linear_f32(const float* A, int lda, const float* W, int ldw, const float* bias, float* out,
                              int ldo, int M, int N, int K, float alpha, int act, float* y, const float* dvec,
                              int d_stride, int T, void* stream) {
    B2S_CHECK_ARG(A && W && out, "b2s_linear_f32: null pointer");
    B2S_CHECK_ARG(K % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && ldo % 4 == 0,
                  "b2s_linear_f32: K, N and leading dimensions must be multiples of 4 (K=%d N=%d)", K, N);
    B2S_CHECK_ARG(aligned16(A) && aligned16(W) && aligned16(out), "b2s_linear_f32: pointers must be 16B aligned");
    B2S_CHECK_ARG(!y || (dvec && T > 0 && d_stride % 4 == 0), "b2s_linear_f32: y needs dvec, T > 0");
    GemmP p{};
    p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.M = M; p.N = N; p.K = K;
    p.T = T > 0 ? T : 0; p.bias = bias; p.alpha = alpha; p.act = act; p.out = out; p.ldo = ldo;
    p.y = y; p.dvec = dvec; p.d_stride = d_stride;
    return launch_gemm<EPI_LINEAR, false>(p, (cudaStream_t)stream);
}

extern "C" int b2s_wavenet_gate_f32(const float* y, const float* Wd, const float* cond, int ld_cond, float* z, int B,
                                    int T, int C, int dilation, void* stream) {
    B2S_CHECK_ARG(y && Wd && cond && z, "b2s_wavenet_gate_f32: null pointer");
    B2S_CHECK_ARG(C % 16 == 0, "b2s_wavenet_gate_f32: residual channels must be a multiple of 16 (C=%d)", C);
    B2S_CHECK_ARG(dilation >= 1 && ld_cond % 4 == 0, "b2s_wavenet_gate_f32: bad dilation / ld_cond");
    GemmP p{};
    p.A = y; p.lda = C; p.W = Wd; p.ldw = 3 * C; p.M = B * T; p.N = 2 * C; p.K = 3 * C;
    p.T = T; p.C = C; p.dil = dilation; p.out = z; p.ldo = C; p.add = cond; p.ldadd = ld_cond;
    return launch_gemm<EPI_GATE, true>(p, (cudaStream_t)stream);
}

extern "C" int b2s_wavenet_out_f32(const float* z, const float* Wo, const float* bo, float* x, float* y_next,
                                   float* skip, const float* dvec_next, int d_stride, int first_layer, int B, int T,
                                   int C, void* stream) {
    B2S_CHECK_ARG(z && Wo && bo && x && skip, "b2s_wavenet_out_f32: null pointer");
    B2S_CHECK_ARG(C % 4 == 0, "b2s_wavenet_out_f32: C %% 4 != 0");
    B2S_CHECK_ARG(!y_next || dvec_next, "b2s_wavenet_out_f32: y_next needs dvec_next");
    GemmP p{};
    p.A = z; p.lda = C; p.W = Wo; p.ldw = C; p.M = B * T; p.N = 2 * C; p.K = C;
    p.T = T; p.C = C; p.bias = bo; p.x = x; p.y = y_next; p.skip = skip; p.dvec = dvec_next; p.d_stride = d_stride;
    p.first = first_layer;
    return launch_gemm<EPI_RESSKIP, false>(p, (cudaStream_t)stream);
}

extern "C" int b2s_lynx_glu_f32(const float* h, const float* W, const float* bias, float* g, int rows, int C,
                                int inner, void* stream) {
    B2S_CHECK_ARG(h && W && bias && g, "b2s_lynx_glu_f32: null pointer");
    B2S_CHECK_ARG(C % 4 == 0 && inner % 2 == 0, "b2s_lynx_glu_f32: bad dims");
    GemmP p{};
    p.A = h; p.lda = C; p.W = W; p.ldw = C; p.M = rows; p.N = 2 * inner; p.K = C;
    p.bias = bias; p.out = g; p.ldo = inner;
    return launch_gemm<EPI_SWIGLU, false>(p, (cudaStream_t)stream);
}

extern "C" int b2s_linear_residual_f32(const float* pin, const float* W, const float* bias, float* x, int rows, int C,
                                       int inner, void* stream) {
    B2S_CHECK_ARG(pin && W && bias && x, "b2s_linear_residual_f32: null pointer");
    B2S_CHECK_ARG(C % 4 == 0 && inner % 4 == 0, "b2s_linear_residual_f32: bad dims");
    GemmP p{};
    p.A = pin; p.lda = inner; p.W = W; p.ldw = inner; p.M = rows; p.N = C; p.K = inner;
    p.bias = bias; p.out = x; p.ldo = C; p.add = x; p.ldadd = C;
    return launch_gemm<EPI_RESIDUAL, false>(p, (cudaStream_t)stream);
}
