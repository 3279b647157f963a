// FastSpeech2 acoustic encoder (the producer of the condition tensor, reference modules/fastspeech/acoustic_encoder.py:79-109): the
// kernels that are NOT GEMMs.  Token-rate work (a few hundred phoneme tokens per utterance) and one frame-rate gather; the GEMMs of
// the transformer layers (QKV / output projection / conv FFN / FFN linear) run on the tcgen05 kernels of b2s_tc_gemm.cu.
//   mel2ph_to_dur    frames per token                                   tts_modules.py:345-351
//   enc_embed        x = (sqrt(H) E[tok] + dur w + b) * keep             tts_modules.py:385-389, :415
//   rope             rotary position embedding of Q and K in place       rotary_embedding_torch.py:36-75, :174-188
//   attention        softmax(Q K^T / sqrt(d) + key padding mask) V       common_layers.py:191-207
//   mask_rows        x *= keep                                           common_layers.py:255, :262
//   layernorm_mask   final LayerNorm * keep -> padded table              tts_modules.py:423, acoustic_encoder.py:89
//   assemble         gather by mel2ph + speaker / pitch / variance / key-shift / speed embeddings   acoustic_encoder.py:89-107
#include "b2s_common.cuh"
#include "b2s_tc.cuh"

namespace b2s {

__global__ void mel2ph_to_dur_kernel(const long long* __restrict__ mel2ph, float* __restrict__ dur, int B, int T, int L) {
    const long long n = (long long)B * T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long ph = mel2ph[i];
        if (ph >= 1 && ph <= L) atomicAdd(dur + (i / T) * L + (ph - 1), 1.0f);      // counts: exact in fp32, order-independent
    }
}

// one warp per token row
__global__ void __launch_bounds__(256) enc_embed_kernel(const long long* __restrict__ tok, const float* __restrict__ dur,
                                                        const float* __restrict__ E, const float* __restrict__ wd,
                                                        const float* __restrict__ bd, float* __restrict__ x, float* __restrict__ keep,
                                                        int rows, int H, int vocab, float scale) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= rows) return;
    const long long t = tok[r];
    const bool ok = t != 0;
    if (lane == 0) keep[r] = ok ? 1.f : 0.f;
    const float d = dur[r];
    const long long tt = t < 0 ? 0 : (t >= vocab ? vocab - 1 : t);
    for (int c = lane; c < H; c += 32)
        x[(long long)r * H + c] = ok ? __fadd_rn(__fmul_rn(scale, __ldg(E + tt * H + c)), fmaf(d, __ldg(wd + c), __ldg(bd + c))) : 0.f;
}

// qkv [rows, 3H] fp32 (q | k | v, heads contiguous inside each): rotate the (2i, 2i+1) pairs of q and k by pos * freqs[i]
__global__ void rope_kernel(float* __restrict__ qkv, const float* __restrict__ freqs, int B, int L, int H, int hd) {
    const int half = hd >> 1, pairs_per_row = H;                 // H/2 pairs in q + H/2 pairs in k
    const long long n = (long long)B * L * pairs_per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int pr = (int)(i % pairs_per_row);
        const long long r = i / pairs_per_row;
        const int pos = (int)(r % L);
        const int which = pr / (H >> 1), p = pr - which * (H >> 1);          // 0: q, 1: k
        const int fi = p % half;
        float* v = qkv + r * 3LL * H + which * H + 2 * p;
        const float ang = __fmul_rn((float)pos, __ldg(freqs + fi));
        float s, c;
        sincosf(ang, &s, &c);
        const float x0 = v[0], x1 = v[1];
        v[0] = __fadd_rn(__fmul_rn(x0, c), __fmul_rn(-x1, s));
        v[1] = __fadd_rn(__fmul_rn(x1, c), __fmul_rn(x0, s));
    }
}

// One warp per query; lane owns dims [4 lane, 4 lane + 4) of the head (hd <= 128, multiple of 4).  Scores of the warp's query
// over all keys live in shared memory (two passes: max / exp-sum, then the weighted sum of V).
constexpr int ATT_WARPS = 8;
template <int BF16>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_kernel(const float* __restrict__ qkv, const float* __restrict__ keep,
                                                                   uint16_t* __restrict__ out, int L, int H, int nh, float scale) {
    extern __shared__ float att_s[];                              // [ATT_WARPS][L]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hd = H / nh;
    const int b = blockIdx.x / nh, h = blockIdx.x - b * nh;
    const int qi = blockIdx.y * ATT_WARPS + warp;
    if (qi >= L) return;
    float* sc = att_s + warp * L;
    const int d0 = lane * 4;
    const bool act = d0 < hd;
    const float* base = qkv + (long long)b * L * 3 * H;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) q = *reinterpret_cast<const float4*>(base + (long long)qi * 3 * H + h * hd + d0);
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) {
        float s = -INFINITY;
        if (__ldg(keep + b * L + j) != 0.f) {                    // key padding mask: common_layers.py:195-198
            float p = 0.f;
            if (act) {
                const float4 k = *reinterpret_cast<const float4*>(base + (long long)j * 3 * H + H + h * hd + d0);
                p = q.x * k.x + q.y * k.y + q.z * k.z + q.w * k.w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
            s = p * scale;
        }
        if (lane == 0) sc[j] = s;
        mx = fmaxf(mx, s);
    }
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
        const float e = sc[j] == -INFINITY ? 0.f : expf(sc[j] - mx);
        sc[j] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = sum > 0.f ? 1.f / sum : 0.f;          // (an utterance without a single token: zeros instead of the reference's NaN)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) {
        for (int j = 0; j < L; ++j) {
            const float w = sc[j];
            if (w != 0.f) {
                const float4 v = *reinterpret_cast<const float4*>(base + (long long)j * 3 * H + 2 * H + h * hd + d0);
                acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
            }
        }
        uint2 o2;
        o2.x = tc::Half16<BF16>::pack2(acc.x * inv, acc.y * inv);
        o2.y = tc::Half16<BF16>::pack2(acc.z * inv, acc.w * inv);
        *reinterpret_cast<uint2*>(out + ((long long)b * L + qi) * H + h * hd + d0) = o2;
    }
}

__global__ void mask_rows_kernel(float* __restrict__ x, const float* __restrict__ keep, long long rows, int H4) {
    const long long n = rows * H4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (__ldg(keep + i / H4) == 0.f) reinterpret_cast<float4*>(x)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// final LayerNorm * keep, written into the padded table enc [B, L + 1, H] (row 0 of every utterance = zeros: mel2ph 0 = padding
// frame, acoustic_encoder.py:89).  One warp per token row, fp32 in / out, two-pass variance.
__global__ void __launch_bounds__(256) layernorm_mask_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ keep,
                                                             float* __restrict__ enc, int B, int L, int H, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;       // over B * (L + 1) table rows
    if (r >= B * (L + 1)) return;
    const int b = r / (L + 1), l = r - b * (L + 1) - 1;
    float* dst = enc + (long long)r * H;
    if (l < 0 || __ldg(keep + b * L + l) == 0.f) {
        for (int c = lane; c < H; c += 32) dst[c] = 0.f;
        return;
    }
    const float* src = x + ((long long)b * L + l) * H;
    float sum = 0.f;
    for (int c = lane; c < H; c += 32) sum += src[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)H;
    float var = 0.f;
    for (int c = lane; c < H; c += 32) { const float a = src[c] - mean; var += a * a; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)H + eps);
    for (int c = lane; c < H; c += 32) dst[c] = (src[c] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
}

struct AssembleArgs {
    const float* val[8];        // per-frame scalars [B * T] (pitch first: f0, transformed to log(1 + f0 / 700) here)
    const float* w[8];          // [H]
    const float* bias[8];       // [H]
    int n;                      // number of scalar embeddings (pitch + variances + key shift + speed), in the reference's add order
    int n_var_first, n_var;     // [n_var_first, n_var_first + n_var): the variance embeddings, summed among themselves first (:62-67)
};

// cond[b, t, :] = enc[b, mel2ph[b, t], :] (+ spk[b, :]) + sum of scalar embeddings; one warp per frame
__global__ void __launch_bounds__(256) assemble_kernel(const float* __restrict__ enc, const long long* __restrict__ mel2ph,
                                                       const float* __restrict__ spk, int spk_per_frame, const AssembleArgs a,
                                                       float* __restrict__ cond, int B, int T, int L, int H) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= (long long)B * T) return;
    const int b = (int)(r / T);
    long long ph = mel2ph[r];
    if (ph < 0 || ph > L) ph = 0;
    const float* e = enc + ((long long)b * (L + 1) + ph) * H;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = i < a.n ? __ldg(a.val[i] + r) : 0.f;
    s[0] = logf(__fadd_rn(1.f, __fdiv_rn(s[0], 700.f)));                    // f0_mel = (1 + f0 / 700).log()   :101
    for (int c = lane; c < H; c += 32) {
        float v = e[c];
        if (spk) v = __fadd_rn(v, __ldg(spk + (spk_per_frame ? r : (long long)b) * H + c));   // :93-99 (spk_mix_embed [B, T, H] or one row per utterance)
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < a.n) {
                const float emb = fmaf(s[i], __ldg(a.w[i] + c), __ldg(a.bias[i] + c));
                if (i >= a.n_var_first && i < a.n_var_first + a.n_var) {
                    var = __fadd_rn(var, emb);
                    if (i == a.n_var_first + a.n_var - 1) v = __fadd_rn(v, var);
                } else {
                    v = __fadd_rn(v, emb);
                }
            }
        }
        cond[r * H + c] = v;
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_enc_mel2ph_to_dur(const int64_t* mel2ph, float* dur, int B, int T, int L, void* stream) {
    B2S_CHECK_ARG(B >= 0 && T >= 0 && L >= 0, "b2s_enc_mel2ph_to_dur: bad dims");
    if ((long long)B * L == 0) return B2S_OK;
    B2S_CHECK_ARG(mel2ph || (long long)B * T == 0, "b2s_enc_mel2ph_to_dur: null pointer");
    B2S_CHECK_ARG(dur, "b2s_enc_mel2ph_to_dur: null pointer");
    B2S_CHECK_CUDA(cudaMemsetAsync(dur, 0, sizeof(float) * (size_t)B * L, (cudaStream_t)stream));
    if ((long long)B * T == 0) return B2S_OK;
    const int blocks = ceil_div((long long)B * T, 256) > 1184 ? 1184 : ceil_div((long long)B * T, 256);
    mel2ph_to_dur_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)mel2ph, dur, B, T, L);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_enc_embed(const int64_t* tokens, const float* dur, const float* E, const float* w_dur, const float* b_dur, float* x,
                             float* keep, int rows, int H, int vocab, void* stream) {
    B2S_CHECK_ARG(rows >= 0 && H > 0 && vocab > 0, "b2s_enc_embed: bad dims");
    if (rows == 0) return B2S_OK;
    B2S_CHECK_ARG(tokens && dur && E && w_dur && b_dur && x && keep, "b2s_enc_embed: null pointer");
    enc_embed_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>((const long long*)tokens, dur, E, w_dur, b_dur, x, keep, rows, H,
                                                                         vocab, sqrtf((float)H));
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_enc_rope(float* qkv, const float* freqs, int B, int L, int H, int num_heads, void* stream) {
    B2S_CHECK_ARG(B >= 0 && L >= 0 && H > 0 && num_heads > 0 && H % num_heads == 0 && (H / num_heads) % 2 == 0, "b2s_enc_rope: bad dims");
    const long long n = (long long)B * L * H;
    if (n == 0) return B2S_OK;
    B2S_CHECK_ARG(qkv && freqs, "b2s_enc_rope: null pointer");
    const int blocks = ceil_div(n, 256) > 2368 ? 2368 : ceil_div(n, 256);
    rope_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(qkv, freqs, B, L, H, H / num_heads);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_enc_attention(const float* qkv, const float* keep, void* out_h, int B, int L, int H, int num_heads, int bf16,
                                 void* stream) {
    B2S_CHECK_ARG(B >= 0 && L >= 0 && H > 0 && num_heads > 0 && H % num_heads == 0, "b2s_enc_attention: bad dims");
    const int hd = num_heads > 0 ? H / num_heads : 0;
    B2S_CHECK_ARG(hd % 4 == 0 && hd <= 128 && H % 4 == 0, "b2s_enc_attention: head dim must be a multiple of 4, at most 128 (got %d)", hd);
    B2S_CHECK_ARG(L <= 1536, "b2s_enc_attention: at most 1536 tokens per utterance (got %d)", L);
    if ((long long)B * L == 0) return B2S_OK;
    B2S_CHECK_ARG(qkv && keep && out_h && tc::al16(qkv) && tc::al16(out_h), "b2s_enc_attention: null / misaligned pointer");
    const size_t smem = sizeof(float) * ATT_WARPS * (size_t)L;
    static tc::PerDevice configured;
    if (configured.first()) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
        B2S_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
    }
    dim3 grid(B * num_heads, ceil_div(L, ATT_WARPS));
    const float scale = 1.0f / sqrtf((float)hd);
    if (bf16) attention_kernel<1><<<grid, ATT_WARPS * 32, smem, (cudaStream_t)stream>>>(qkv, keep, (uint16_t*)out_h, L, H, num_heads, scale);
    else attention_kernel<0><<<grid, ATT_WARPS * 32, smem, (cudaStream_t)stream>>>(qkv, keep, (uint16_t*)out_h, L, H, num_heads, scale);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_enc_mask_rows(float* x, const float* keep, int rows, int H, void* stream) {
    B2S_CHECK_ARG(rows >= 0 && H > 0 && H % 4 == 0, "b2s_enc_mask_rows: H must be a multiple of 4");
    if (rows == 0) return B2S_OK;
    B2S_CHECK_ARG(x && keep && tc::al16(x), "b2s_enc_mask_rows: null / misaligned pointer");
    const long long n = (long long)rows * (H / 4);
    const int blocks = ceil_div(n, 256) > 2368 ? 2368 : ceil_div(n, 256);
    mask_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, keep, rows, H / 4);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_enc_layernorm_mask(const float* x, const float* gamma, const float* beta, const float* keep, float* enc, int B, int L,
                                      int H, float eps, void* stream) {
    B2S_CHECK_ARG(B >= 0 && L >= 0 && H > 0, "b2s_enc_layernorm_mask: bad dims");
    if (B == 0) return B2S_OK;
    B2S_CHECK_ARG(gamma && beta && enc && (L == 0 || (x && keep)), "b2s_enc_layernorm_mask: null pointer");
    layernorm_mask_kernel<<<ceil_div((long long)B * (L + 1), 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, keep, enc, B, L, H, eps);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_enc_assemble(const float* enc, const int64_t* mel2ph, const float* spk, int spk_per_frame, const float* const* vals_host,
                                const float* const* w_host, const float* const* bias_host, int n, int n_var_first, int n_var,
                                float* cond, int B, int T, int L, int H, void* stream) {
    B2S_CHECK_ARG(B >= 0 && T >= 0 && L >= 0 && H > 0, "b2s_enc_assemble: bad dims");
    B2S_CHECK_ARG(n >= 1 && n <= 8 && n_var >= 0 && n_var_first >= 0 && n_var_first + n_var <= n,
                  "b2s_enc_assemble: 1..8 scalar embeddings (the first is the pitch)");
    if ((long long)B * T == 0) return B2S_OK;
    B2S_CHECK_ARG(enc && mel2ph && cond && vals_host && w_host && bias_host, "b2s_enc_assemble: null pointer");
    AssembleArgs a{};
    for (int i = 0; i < n; ++i) {
        B2S_CHECK_ARG(vals_host[i] && w_host[i] && bias_host[i], "b2s_enc_assemble: null embedding %d", i);
        a.val[i] = vals_host[i]; a.w[i] = w_host[i]; a.bias[i] = bias_host[i];
    }
    a.n = n; a.n_var_first = n_var_first; a.n_var = n_var;
    assemble_kernel<<<ceil_div((long long)B * T, 8), 256, 0, (cudaStream_t)stream>>>(enc, (const long long*)mel2ph, spk, spk_per_frame, a, cond, B, T, L, H);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}
