// Tensor-core (tcgen05 / TMEM / TMA) GEMM family of the 16-bit path.  sm_100a only.
//
//   out[rows, N] = epilogue( A[rows, K] . W[N, K]^T )          A, W: bf16 or fp16, fp32 accumulate in TMEM
//
// One persistent, warp-specialised kernel (384 threads, 1 CTA / SM):
//   warp 0    TMA producer   - cp.async.bulk.tensor tiles (128B swizzle) into a 4-stage smem ring
//   warp 1    MMA issuer     - one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16)
//   warp 2    TMEM allocator - 512 columns = two 128x256 fp32 accumulators (double buffered)
//   warps 4-11 epilogue      - tcgen05.ld (32 lanes x 32 columns per warp per step), fused math, global stores;
//                              the epilogue of tile i overlaps the MMAs of tile i+1
// Tile 128 (frames) x 256 (output channels) x 64 (K per stage).
//
// The dilated k=3 convolution (wavenet.py:22-28,38) is an IMPLICIT GEMM over K = 3*C (tap-major): the
// A operand of k-block (tap, c0) is the activation tile at time offset (tap-1)*dilation, fetched with a
// 3-D tensor map [C, T, B]; TMA's out-of-bounds zero fill produces the per-utterance zero padding of
// y = x + step embedding for free, and tiles never straddle utterances.
//
// Epilogues (all fused, nothing but the named tensors touches HBM):
//   LINEAR   act(alpha*acc + bias) -> fp32 and/or 16-bit out, optional y = out + step-embedding row
//   GATE     [g|f] = acc + cond (hoisted conditioner projection, 16-bit, interleaved) ; z = sigmoid(g)*tanh(f)
//   RESSKIP  x <- (x + acc + b)/sqrt2 (fp32 residual stream), y_next <- x + d_next (16-bit), skip (+)= acc + b
//   SWIGLU   g = out * silu(gate)  (interleaved)             RESIDUAL   x <- x + acc + b
//   VRES     x <- x_src + acc + b, y_h <- leaky_relu(x) (the vocoder's residual blocks, nsf_hifigan/models.py:60-68)
// The shipped kernel is the cta_group::2 variant further down (tc_gemm_cg2_kernel): 256-row MMAs over a CTA pair, N tiles of 256 / 192 /
// 128 / 64 columns, and a last-wave split that cuts the tiles of a partial last wave along N (decode_vtile).
#include "b2s_tc.cuh"

#include <stdlib.h>

namespace b2s {
namespace tc {

constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 64, UMMA_K = 16, STAGES = 4;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;          // 16 KB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;          // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;          // 48 KB
constexpr int STG_LD = 36;                                // floats per staged row: 32 + 4 pad -> conflict-free 128-bit accesses
constexpr int EPI_WARPS = 8;                              // two per TMEM lane quarter, alternating 32-column chunks
constexpr int STG_BYTES = EPI_WARPS * 16 * STG_LD * 4;    // one 16x32 fp32 transpose buffer per epilogue warp (two passes per chunk)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;
constexpr int NTHREADS = 128 + EPI_WARPS * 32;

enum Epi : int { EPI_LINEAR = 0, EPI_GATE = 1, EPI_RESSKIP = 2, EPI_SWIGLU = 3, EPI_RESIDUAL = 4, EPI_VRES = 5 };

struct __align__(64) TcP {
    CUtensorMap mapA, mapW;
    int B, T;                  // utterance grid of the A tiles (flat GEMM: B = 1, T = rows)
    int T_utt;                 // frames per utterance, for the step-embedding row lookup (b = row / T_utt)
    int N, num_kb, kb_per_tap, dil;
    int tiles_m_per_b, tiles_n, num_tiles;
    const float* bias; float alpha; int act;
    float* out_f; int ldo;
    void* out_h; int ldoh;
    int oh_blk; long long oh_blk_stride;   // out_h column blocks: block k of oh_blk columns starts at k * oh_blk_stride
    int oh_tiled, oh_B, oh_tpb;            // out_h in the tile/chunk-major layout of the fused WaveNet kernels (see b2s.h)
    void* y_h; int ldy;
    const float* dvec; int d_stride;
    const void* cond; int ldc;
    float* x; float* skip; void* skip_h; int C; int first;
    int cg2;                   // 1: cta_group::2 kernel (tiles_m_per_b even, weight map box = 128 rows)
    int bn;                    // N tile of the cta_group::2 kernel: 256, or 192 when N is a multiple of 192 but not of 256 (C = 192 models)
    int k_layered;             // 1: K block group g = kb / kb_per_tap selects the THIRD coordinate of the A map (A = [L][rows][C], K = L*C)
    // EPI_GATE with dilation <= CONV3_HALO: a pipeline stage holds ONE 64-channel slab of the tile WITH its halo (mapA3: box of
    // BLOCK_M + 2 * CONV3_HALO rows) and the weight slabs of all three taps; the taps read it through row-shifted descriptors
    CUtensorMap mapA3;
    int conv3;
    int tap_c;                 // conv GEMMs: tap k reads rows t + (k - tap_c) * dil (1 for the 3-tap dilated conv, ksize / 2 in general)
    // EPI_VRES, the vocoder's residual blocks (x <- x_src + acc + bias): the stream that is READ (NULL: x itself) and a leaky ReLU on the 16-bit copy
    const float* x_src;
    int y_lrelu; float y_slope;
    // Last-wave split (cta_group::2 kernel): the tile pairs [0, vt_full) are whole tiles walked round robin by the CTA pairs; each of the
    // remaining num_pt - vt_full tile pairs (fewer than half of the CTA pairs) is cut along N into vt_split sub-tiles of bn / vt_split
    // columns, so that the last, partial wave of a launch (or a launch with fewer tiles than CTA pairs) spreads over the idle pairs.
    int vt_full, vt_split, vt_total;
    int relaxed_release;       // 1: the epilogue hands its TMEM accumulator back with a relaxed arrival (no wait for its global stores)
    int no_split;              // 1: never split (the vocoder's convs: three residual blocks run concurrently and fill each other's partial waves)
};

// ---- 16-bit helpers --------------------------------------------------------------------------------
template <int BF16>
__device__ __forceinline__ void store_h1(void* base, long long idx, float v) {
    if (BF16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
}
template <int BF16>
__device__ __forceinline__ void store_h4(void* base, long long idx, float4 v) {          // 4 values, 8 B
    uint2 o;
    o.x = Half16<BF16>::pack2(v.x, v.y);
    o.y = Half16<BF16>::pack2(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + idx) = o;
}
template <int BF16>
__device__ __forceinline__ void store_h2(void* base, long long idx, float a, float b) {  // 2 values, 4 B
    *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(base) + idx) = Half16<BF16>::pack2(a, b);
}
__device__ __forceinline__ uint2 ldg_nc_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__device__ __noinline__ float4 act4_slow(float4 v, int act) {
    return make_float4(apply_act(v.x, act), apply_act(v.y, act), apply_act(v.z, act), apply_act(v.w, act));
}

// Per-chunk constants that do not depend on the row (one float4 = the lane's 4 columns).
struct EpiConst {
    long long oh_off;  // offset of this column block of out_h (layer-major tables), 0 for plain row-major output
    float4 bias;
    float4 d;          // step-embedding row when it is shared by every utterance (d_stride == 0)
};

template <int EPI>
__device__ __forceinline__ EpiConst epilogue_consts(const TcP& p, int col) {
    EpiConst c;
    c.bias = make_float4(0.f, 0.f, 0.f, 0.f);
    c.d = c.bias;
    c.oh_off = (EPI == EPI_LINEAR && p.oh_blk) ? (long long)(col / p.oh_blk) * (p.oh_blk_stride - p.oh_blk) : 0;
    if (EPI == EPI_GATE) return c;
    if (p.bias) {
        if (EPI != EPI_LINEAR || col + 4 <= p.N) c.bias = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        else {
            float* bb = &c.bias.x;
            for (int i = 0; i < 4; ++i) bb[i] = col + i < p.N ? __ldg(p.bias + col + i) : 0.f;
        }
    }
    if ((EPI == EPI_LINEAR || EPI == EPI_RESSKIP) && p.y_h && p.d_stride == 0 && col + 4 <= p.N && (EPI != EPI_RESSKIP || col < p.C))
        c.d = __ldg(reinterpret_cast<const float4*>(p.dvec + col));
    if (EPI == EPI_RESIDUAL && p.dvec) c.d = __ldg(reinterpret_cast<const float4*>(p.dvec + col));     // layer scale (ConvNeXt gamma)
    return c;
}

// ---- epilogues in the COALESCED layout: a lane owns 4 consecutive columns [col, col+4) of one row; the 8
// lanes of a quarter-warp cover 128 contiguous bytes of an fp32 row.  Two phases per chunk so that the 8
// independent row loads of a lane are all in flight before the first dependent store is issued. ------------
template <int EPI, int BF16>
__device__ __forceinline__ float4 epilogue_load(const TcP& p, long long r, int col) {
    if (EPI == EPI_GATE) {
        // hoisted conditioner projection, same (gate j, filter j) interleave as the packed weight rows
        const uint2 c = ldg_nc_u2(reinterpret_cast<const uint16_t*>(p.cond) + r * p.ldc + col);
        const float2 c0 = Half16<BF16>::unpack2(c.x), c1 = Half16<BF16>::unpack2(c.y);
        return make_float4(c0.x, c0.y, c1.x, c1.y);
    } else if (EPI == EPI_RESIDUAL) {
        return *reinterpret_cast<const float4*>(p.x + r * p.C + col);
    } else if (EPI == EPI_VRES) {
        return *reinterpret_cast<const float4*>((p.x_src ? p.x_src : p.x) + r * p.C + col);
    } else if (EPI == EPI_RESSKIP) {
        if (col < p.C) return *reinterpret_cast<const float4*>(p.x + r * p.C + col);
        if (!p.first) return *reinterpret_cast<const float4*>(p.skip + r * p.C + (col - p.C));
    }
    return make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int EPI, int BF16>
__device__ __forceinline__ void epilogue_quad(const TcP& p, float4 acc, float4 in, const EpiConst& k, long long r, int b,
                                              int col, const uint2* pre_cond = nullptr) {
    if (EPI == EPI_LINEAR) {
        // ReLU / identity inline (the hot cases: stem, head, cond table); Mish / GELU / SiLU through ONE
        // out-of-line call so the unrolled epilogue stays small enough for the instruction cache
        const float lo = p.act == ACT_RELU ? 0.f : -INFINITY;
        float4 v = make_float4(fmaxf(fmaf(p.alpha, acc.x, k.bias.x), lo), fmaxf(fmaf(p.alpha, acc.y, k.bias.y), lo),
                               fmaxf(fmaf(p.alpha, acc.z, k.bias.z), lo), fmaxf(fmaf(p.alpha, acc.w, k.bias.w), lo));
        if (p.act == ACT_LRELU)      // inline like ReLU: every conv of the vocoder ends in it (an out-of-line call per quad made its
            v = make_float4(fmaxf(v.x, LRELU_SLOPE * v.x), fmaxf(v.y, LRELU_SLOPE * v.y),      // first convs epilogue-bound: 164 us
                            fmaxf(v.z, LRELU_SLOPE * v.z), fmaxf(v.w, LRELU_SLOPE * v.w));     // against 101 us for the residual conv)
        else if (p.act > ACT_RELU) v = act4_slow(v, p.act);
        if (col + 4 <= p.N) {
            if (p.out_f) *reinterpret_cast<float4*>(p.out_f + r * p.ldo + col) = v;
            if (p.out_h) {
                if (p.oh_tiled) {
                    // [l][b][tile][chunk j = (col % N2) / 32][16-byte piece k][row in tile][8 elements]
                    const int l = col / p.oh_blk, nn = col - l * p.oh_blk;
                    const int bb = (int)(r / p.T_utt), t = (int)(r - (long long)bb * p.T_utt);
                    const long long tile = ((long long)l * p.oh_B + bb) * p.oh_tpb + (t >> 7);
                    const long long idx = ((tile * (p.oh_blk >> 5) + (nn >> 5)) * 4 + ((nn & 31) >> 3)) * 1024 + (t & 127) * 8 + (nn & 7);
                    store_h4<BF16>(p.out_h, idx, v);
                } else {
                    store_h4<BF16>(p.out_h, k.oh_off + r * p.ldoh + col, v);
                }
            }
            if (p.y_h) {
                const float4 d = p.d_stride == 0 ? k.d : __ldg(reinterpret_cast<const float4*>(p.dvec + (long long)b * p.d_stride + col));
                store_h4<BF16>(p.y_h, r * p.ldy + col, add4(v, d));
            }
        } else {
            const float* vv = &v.x;
            for (int i = 0; i < 4; ++i) {
                if (col + i < p.N) {
                    if (p.out_f) p.out_f[r * p.ldo + col + i] = vv[i];
                    if (p.out_h) store_h1<BF16>(p.out_h, k.oh_off + r * p.ldoh + col + i, vv[i]);
                    if (p.y_h) store_h1<BF16>(p.y_h, r * p.ldy + col + i, vv[i] + __ldg(p.dvec + (long long)b * p.d_stride + col + i));
                }
            }
        }
    } else if (EPI == EPI_GATE) {
        const float z0 = sigmoid_fast(acc.x + in.x) * tanh_fast(acc.y + in.y);
        const float z1 = sigmoid_fast(acc.z + in.z) * tanh_fast(acc.w + in.w);
        store_h2<BF16>(p.out_h, r * p.ldoh + (col >> 1), z0, z1);
    } else if (EPI == EPI_SWIGLU) {
        const float g0 = acc.y + k.bias.y, g1 = acc.w + k.bias.w;
        store_h2<BF16>(p.out_h, r * p.ldoh + (col >> 1), (acc.x + k.bias.x) * (g0 * sigmoid_fast(g0)),
                       (acc.z + k.bias.z) * (g1 * sigmoid_fast(g1)));
    } else if (EPI == EPI_RESIDUAL) {
        float4 o = add4(acc, k.bias);
        if (p.dvec) o = make_float4(o.x * k.d.x, o.y * k.d.y, o.z * k.d.z, o.w * k.d.w);      // x + gamma * (acc + b), convnext.py:51-56
        float4 xn = add4(in, o);
        if (p.cond) {
            // strong_cond LYNXNet: the NEXT layer's front_cond_inject (x += cond_proj(cond), lynxnet.py:77-82) folded into this
            // layer's residual epilogue, so the next LayerNorm kernel neither reads the cond table nor writes x back
            const uint2 c = pre_cond ? *pre_cond : ldg_nc_u2(reinterpret_cast<const uint16_t*>(p.cond) + r * p.ldc + col);
            const float2 c0 = Half16<BF16>::unpack2(c.x), c1 = Half16<BF16>::unpack2(c.y);
            xn = add4(xn, make_float4(c0.x, c0.y, c1.x, c1.y));
        }
        *reinterpret_cast<float4*>(p.x + r * p.C + col) = xn;
        if (p.y_h) store_h4<BF16>(p.y_h, r * p.ldy + col, xn);                                 // 16-bit copy for the next depthwise conv
    } else if (EPI == EPI_VRES) {
        float4 xn = add4(in, add4(acc, k.bias));                                               // x = xt + x, nsf_hifigan/models.py:66
        *reinterpret_cast<float4*>(p.x + r * p.C + col) = xn;
        if (p.y_h) {                                                                           // leaky_relu(x) for the next conv (:62)
            if (p.y_lrelu)
                xn = make_float4(xn.x > 0.f ? xn.x : xn.x * p.y_slope, xn.y > 0.f ? xn.y : xn.y * p.y_slope,
                                 xn.z > 0.f ? xn.z : xn.z * p.y_slope, xn.w > 0.f ? xn.w : xn.w * p.y_slope);
            store_h4<BF16>(p.y_h, r * p.ldy + col, xn);
        }
    } else {   // EPI_RESSKIP: reference column order, [0, C) residual, [C, 2C) skip
        const float inv_sqrt2 = 0.70710678118654752440f;
        const float4 o = add4(acc, k.bias);
        if (col < p.C) {
            const float4 xn = make_float4((in.x + o.x) * inv_sqrt2, (in.y + o.y) * inv_sqrt2, (in.z + o.z) * inv_sqrt2,
                                          (in.w + o.w) * inv_sqrt2);
            *reinterpret_cast<float4*>(p.x + r * p.C + col) = xn;
            if (p.y_h) {
                const float4 d = p.d_stride == 0 ? k.d : __ldg(reinterpret_cast<const float4*>(p.dvec + (long long)b * p.d_stride + col));
                store_h4<BF16>(p.y_h, r * p.ldy + col, add4(xn, d));
            }
        } else {
            const float4 s = add4(o, in);                  // in = 0 on the first layer
            *reinterpret_cast<float4*>(p.skip + r * p.C + (col - p.C)) = s;
            if (p.skip_h) store_h4<BF16>(p.skip_h, r * p.C + (col - p.C), s);
        }
    }
}

#ifdef B2S_EXPERIMENTS     // the single-CTA kernel (B2S_GEMM_CG1=1): superseded by the cta_group::2 kernel below
// ---- the kernel ------------------------------------------------------------------------------------
template <int EPI, int BF16>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_kernel(const __grid_constant__ TcP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* stg_all = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STG_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.mapA);
        prefetch_tmap(&p.mapW);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], EPI_WARPS);   // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.tiles_n, m_tile = tile / p.tiles_n;
            const int b = m_tile / p.tiles_m_per_b, t0 = (m_tile - b * p.tiles_m_per_b) * BLOCK_M;
            const int n0 = n_tile * BLOCK_N;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    const int tap = kb / p.kb_per_tap;
                    const int c0 = (kb - tap * p.kb_per_tap) * BLOCK_K;
                    tma_load_3d(sa, &p.mapA, &full[stage], c0, t0 + (tap - p.tap_c) * p.dil, p.k_layered ? tap : b);
                    tma_load_2d(sa + A_BYTES, &p.mapW, &full[stage], kb * BLOCK_K, n0);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_f16(BLOCK_M, BLOCK_N, BF16);
        int stage = 0, as = 0;
        uint32_t phase = 0, aphase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[as], aphase ^ 1);          // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BLOCK_N;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        umma_ss(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UMMA_K * 2)),
                                make_sw128_kmajor_desc(b_addr + k * (UMMA_K * 2)), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[stage]);                       // frees the smem slot when the MMAs retire
                    if (kb == p.num_kb - 1) umma_commit(&tfull[as]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int e = warp - 4, q = e & 3, sub = e >> 2;   // TMEM lane quarter q; chunks j with (j & 1) == sub
        int as = 0;
        uint32_t aphase = 0;
        float* stg = stg_all + e * (16 * STG_LD);
        const int cl = (lane & 7) * 4, rsub = lane >> 3;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.tiles_n, m_tile = tile / p.tiles_n;
            const int bt = m_tile / p.tiles_m_per_b, t0 = (m_tile - bt * p.tiles_m_per_b) * BLOCK_M;
            const int n0 = n_tile * BLOCK_N;
            const int tq = t0 + q * 32 + rsub;             // row of iteration i: tq + 4*i
            // per-chunk inputs (cond / x / skip rows of the lane, bias and step-embedding quads) are requested ONE CHUNK AHEAD -
            // the first chunk's before the accumulator wait - so their L2 / HBM latency hides behind the TMEM drain of the
            // previous chunk (ncu: the GATE / RESSKIP epilogues were long-scoreboard bound, tensor pipe 46 % / 18 % active)
            EpiConst kcn;
            float4 inn[8];
            uint2 cnn[8];                                   // EPI_RESIDUAL with a folded cond add: the next layer's cond quads
            const bool fold = EPI == EPI_RESIDUAL && p.cond != nullptr;
            bool okn = false;
            auto load_chunk = [&](int j) {
                const int col = n0 + 32 * j + cl;
                okn = col < p.N;
                if (okn) {
                    kcn = epilogue_consts<EPI>(p, col);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        inn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        cnn[i] = make_uint2(0u, 0u);
                        if (tq + 4 * i < p.T) {
                            const long long r = (long long)bt * p.T + tq + 4 * i;
                            inn[i] = epilogue_load<EPI, BF16>(p, r, col);
                            if (fold) cnn[i] = ldg_nc_u2(reinterpret_cast<const uint16_t*>(p.cond) + r * p.ldc + col);
                        }
                    }
                }
            };
            if (n0 + 32 * sub < p.N) load_chunk(sub);
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
            for (int j = sub; j < BLOCK_N / 32; j += 2) {
                const int col0 = n0 + 32 * j;
                if (col0 >= p.N) break;
                float acc[32];
                tmem_ld32(taddr + j * 32, acc);
                const int col = col0 + cl;
                const bool colok = okn;
                const EpiConst kc = kcn;
                float4 in[8];
                uint2 cn[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { in[i] = inn[i]; cn[i] = cnn[i]; }
                if (j + 2 < BLOCK_N / 32 && col0 + 64 < p.N) load_chunk(j + 2);
                tmem_ld_wait();
                // transpose through the warp's private 16-row staging tile, two passes: thread = row -> lane = 4 columns
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if ((lane >> 4) == pass) {
                        float4* srow = reinterpret_cast<float4*>(stg + (lane & 15) * STG_LD);
#pragma unroll
                        for (int c = 0; c < 8; ++c) srow[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                    }
                    __syncwarp();
                    if (colok) {
#pragma unroll(EPI == EPI_LINEAR ? 1 : 4)
                        for (int i2 = 0; i2 < 4; ++i2) {
                            const int i = 4 * pass + i2;
                            const int t = tq + 4 * i;
                            if (t < p.T) {
                                const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i2 + rsub) * STG_LD + cl);
                                const int b = (p.d_stride != 0 && p.T_utt > 0) ? (bt * p.T + t) / p.T_utt : 0;
                                epilogue_quad<EPI, BF16>(p, v, in[i], kc, (long long)bt * p.T + t, b, col);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

#endif  // B2S_EXPERIMENTS

// ---- the same kernel with cta_group::2 MMAs (the default) ----------------------------------------------------
// Two CTAs of a cluster take two neighbouring M tiles of the same N tile and form ONE 256-row MMA: each CTA keeps its own
// A tile and HALF of the B tile (128 of its 256 rows), so a K slab costs 32 KB of shared memory per CTA instead of 48 KB
// (6 pipeline stages instead of 4, and half the L2 traffic for the weights: scripts/tc_probe.py showed the single-CTA
// mainloop sitting on the 6.3 KB/clk L2 slice cap).  Only the leader issues tcgen05.mma / commit; both CTAs' TMA loads
// complete on the leader's full barrier; both CTAs' epilogue warps arrive on the leader's tempty barrier.
// virtual tile v -> (tile pair pt, first column n0, width w) under the last-wave split (see TcP::vt_*)
struct VTile { int pt, n_off, w; };
__device__ __forceinline__ VTile decode_vtile(const TcP& p, int v) {
    VTile t;
    if (v < p.vt_full) { t.pt = v; t.n_off = 0; t.w = p.bn; }
    else {
        const int r = v - p.vt_full;
        t.pt = p.vt_full + r / p.vt_split;
        t.w = p.bn / p.vt_split;
        t.n_off = (r % p.vt_split) * t.w;
    }
    return t;
}

constexpr int STAGES2 = 6;
constexpr int BH_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;       // 16 KB
constexpr int STAGE2_BYTES = A_BYTES + BH_BYTES;            // 32 KB
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + STG_BYTES + 1024 + 256;
// Dilated conv (EPI_GATE) with the A slab resident across the three taps: the mainloop above sits on the L2 -> SM cap (32 KB per
// four MMAs and CTA = 11.7 TB/s chip-wide when MMA-bound, the measured cap); loading the slab once with its halo instead of once
// per tap moves 68 KB instead of 96 KB per three K blocks.  K order becomes (channel slab, tap) instead of (tap, channel slab).
constexpr int CONV3_HALO = 16;
constexpr int A3_ROWS = BLOCK_M + 2 * CONV3_HALO;            // 160
constexpr int A3_BYTES = A3_ROWS * BLOCK_K * 2;              // 20 KB
constexpr int STAGE3_BYTES = A3_BYTES + 3 * BH_BYTES;        // 68 KB
constexpr int STAGES3 = 3;
constexpr int SMEM3_BYTES = STAGES3 * STAGE3_BYTES + STG_BYTES + 1024 + 256;
static_assert(SMEM3_BYTES <= 232448, "conv3 ring does not fit");

template <int EPI, int BF16>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_cg2_kernel(const __grid_constant__ TcP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const bool conv3 = (EPI == EPI_GATE || EPI == EPI_LINEAR || EPI == EPI_VRES) && p.conv3 != 0;
    const int ring_bytes = conv3 ? STAGES3 * STAGE3_BYTES : STAGES2 * STAGE2_BYTES;
    float* stg_all = reinterpret_cast<float*>(smem + ring_bytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + ring_bytes + STG_BYTES);
    uint64_t* empty = full + STAGES2;
    uint64_t* tfull = empty + STAGES2;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;      // a cluster walks over (M-tile pair, N tile)
    const int num_pt = (p.B * p.tiles_m_per_b / 2) * p.tiles_n;      // tiles_m_per_b is even (host)

    if (warp == 0 && lane == 0) {
        prefetch_tmap(conv3 ? &p.mapA3 : &p.mapA);
        prefetch_tmap(&p.mapW);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES2; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 2 * EPI_WARPS);   // one arrival per epilogue warp of BOTH CTAs (on the leader's barrier)
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_cg2(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        // EPI_GATE: the hoisted conditioner-projection rows of a tile stream from HBM once per evaluation; ncu showed the epilogue
        // warps waiting on them (long scoreboard) although they are requested one chunk ahead.  The producer asks L2 for the rows
        // of the tile AFTER the current one (and of the first tile at kernel start), a whole mainloop before the epilogue needs them.
        auto prefetch_cond = [&](int pt2) {
            if (EPI != EPI_GATE || p.cond == nullptr || pt2 >= num_pt || lane != 0) return;
            const int m2 = 2 * (pt2 / p.tiles_n) + rank;
            const int b2 = m2 / p.tiles_m_per_b, t2 = (m2 - b2 * p.tiles_m_per_b) * BLOCK_M;
            const int nrows = min(BLOCK_M, p.T - t2);
            if (b2 >= p.B || nrows <= 0) return;
            const uint8_t* base = reinterpret_cast<const uint8_t*>(p.cond) + ((long long)b2 * p.T + t2) * p.ldc * 2;
            const int bytes = nrows * p.ldc * 2;                      // full rows: contiguous, a multiple of 16 bytes
            for (int off = 0; off < bytes; off += 16384) {
                const int n = min(16384, bytes - off);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(n) : "memory");
            }
        };
        // (the same prefetch for the fp32 residual-stream rows of the x-updating epilogues was measured: 1-2 % slower)
        if (pair < p.vt_total) prefetch_cond(decode_vtile(p, pair).pt);
        for (int v = pair; v < p.vt_total; v += npairs) {
            const VTile vt = decode_vtile(p, v);
            const int pt = vt.pt;
            const int n_tile = pt % p.tiles_n, m_tile = 2 * (pt / p.tiles_n) + rank;
            const int b = m_tile / p.tiles_m_per_b, t0 = (m_tile - b * p.tiles_m_per_b) * BLOCK_M;
            const int n0 = n_tile * p.bn + vt.n_off;
            const int wrow = n0 + rank * (vt.w / 2);             // this CTA's half of the (sub-)tile's weight rows; the box stays bn / 2 rows
            if (v + npairs < p.vt_full) prefetch_cond(v + npairs);
            if (conv3) {
                for (int cs = 0; cs < p.kb_per_tap; ++cs) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (lane == 0) {
                        uint8_t* sa = smem + stage * STAGE3_BYTES;
                        const uint32_t lbar = mapa_u32(&full[stage], 0);
                        if (rank == 0) mbar_expect_tx(&full[stage], 2 * (A3_BYTES + 3 * (p.bn / 2) * BLOCK_K * 2));
                        tma_load_3d_cg2(sa, &p.mapA3, lbar, cs * BLOCK_K, t0 - CONV3_HALO, b);
#pragma unroll
                        for (int tap = 0; tap < 3; ++tap)
                            tma_load_2d_cg2(sa + A3_BYTES + tap * BH_BYTES, &p.mapW, lbar, (tap * p.kb_per_tap + cs) * BLOCK_K, wrow);
                    }
                    __syncwarp();
                    if (++stage == STAGES3) { stage = 0; phase ^= 1; }
                }
                continue;
            }
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sa = smem + stage * STAGE2_BYTES;
                    const uint32_t lbar = mapa_u32(&full[stage], 0);
                    if (rank == 0) mbar_expect_tx(&full[stage], 2 * (A_BYTES + (p.bn / 2) * BLOCK_K * 2));
                    const int tap = kb / p.kb_per_tap;
                    const int c0 = (kb - tap * p.kb_per_tap) * BLOCK_K;
                    tma_load_3d_cg2(sa, &p.mapA, lbar, c0, t0 + (tap - p.tap_c) * p.dil, p.k_layered ? tap : b);
                    tma_load_2d_cg2(sa + A_BYTES, &p.mapW, lbar, kb * BLOCK_K, wrow);
                }
                __syncwarp();
                if (++stage == STAGES2) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (leader CTA): M = 256 across the pair =====================
        int stage = 0, as = 0;
        uint32_t phase = 0, aphase = 0;
        for (int v = pair; v < p.vt_total; v += npairs) {
            const uint32_t idesc = make_idesc_f16(2 * BLOCK_M, decode_vtile(p, v).w, BF16);      // sub-tiles of the last wave are narrower
            mbar_wait(&tempty[as], aphase ^ 1);          // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BLOCK_N;
            if (conv3) {
                for (int cs = 0; cs < p.kb_per_tap; ++cs) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t s_addr = smem_u32(smem + stage * STAGE3_BYTES);
#pragma unroll
                        for (int tap = 0; tap < 3; ++tap) {
                            // rows t0 + (tap - 1) * dil + i of the utterance = rows HALO + (tap - 1) * dil + i of the slab: the swizzle
                            // is a function of the absolute shared-memory address, so a row-shifted start address reads them in place
                            const uint32_t a_addr = s_addr + (CONV3_HALO + (tap - 1) * p.dil) * (BLOCK_K * 2);
                            const uint32_t b_addr = s_addr + A3_BYTES + tap * BH_BYTES;
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                umma_ss_cg2(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UMMA_K * 2)),
                                            make_sw128_kmajor_desc(b_addr + k * (UMMA_K * 2)), idesc, (cs | tap | k) != 0);
                            }
                        }
                        umma_commit_cg2_mcast(&empty[stage], 3);
                        if (cs == p.kb_per_tap - 1) umma_commit_cg2_mcast(&tfull[as], 3);
                    }
                    __syncwarp();
                    if (++stage == STAGES3) { stage = 0; phase ^= 1; }
                }
                as ^= 1;
                if (as == 0) aphase ^= 1;
                continue;
            }
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smem + stage * STAGE2_BYTES);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        umma_ss_cg2(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UMMA_K * 2)),
                                make_sw128_kmajor_desc(b_addr + k * (UMMA_K * 2)), idesc, (kb | k) != 0);
                    }
                    umma_commit_cg2_mcast(&empty[stage], 3);                       // frees the smem slot in BOTH CTAs
                    if (kb == p.num_kb - 1) umma_commit_cg2_mcast(&tfull[as], 3);  // accumulator complete, both CTAs
                }
                __syncwarp();
                if (++stage == STAGES2) { stage = 0; phase ^= 1; }
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int e = warp - 4, q = e & 3, sub = e >> 2;   // TMEM lane quarter q; chunks j with (j & 1) == sub
        int as = 0;
        uint32_t aphase = 0;
        float* stg = stg_all + e * (16 * STG_LD);
        const int cl = (lane & 7) * 4, rsub = lane >> 3;
        const bool relaxed_release = p.relaxed_release != 0;
        const bool direct16 = EPI == EPI_LINEAR && p.out_h && !p.out_f && !p.y_h && !p.oh_tiled && !p.oh_blk && (p.ldoh & 7) == 0 && (p.N & 31) == 0 &&
                              ((reinterpret_cast<uintptr_t>(p.out_h) & 15) == 0) && (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
        for (int v = pair; v < p.vt_total; v += npairs) {
            const VTile vt = decode_vtile(p, v);
            const int pt = vt.pt, tw = vt.w;                // tw: columns of this (sub-)tile
            const int n_tile = pt % p.tiles_n, m_tile = 2 * (pt / p.tiles_n) + rank;
            const int bt = m_tile / p.tiles_m_per_b, t0 = (m_tile - bt * p.tiles_m_per_b) * BLOCK_M;
            const int n0 = n_tile * p.bn + vt.n_off;
            const int tq = t0 + q * 32 + rsub;             // row of iteration i: tq + 4*i
            // per-chunk inputs (cond / x / skip rows of the lane, bias and step-embedding quads) are requested ONE CHUNK AHEAD -
            // the first chunk's before the accumulator wait - so their L2 / HBM latency hides behind the TMEM drain of the
            // previous chunk (ncu: the GATE / RESSKIP epilogues were long-scoreboard bound, tensor pipe 46 % / 18 % active)
            EpiConst kcn;
            float4 inn[8];
            uint2 cnn[8];                                   // EPI_RESIDUAL with a folded cond add: the next layer's cond quads
            const bool fold = EPI == EPI_RESIDUAL && p.cond != nullptr;
            bool okn = false;
            auto load_chunk = [&](int j) {
                const int col = n0 + 32 * j + cl;
                okn = col < p.N;
                if (okn) {
                    kcn = epilogue_consts<EPI>(p, col);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        inn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        cnn[i] = make_uint2(0u, 0u);
                        if (tq + 4 * i < p.T) {
                            const long long r = (long long)bt * p.T + tq + 4 * i;
                            inn[i] = epilogue_load<EPI, BF16>(p, r, col);
                            if (fold) cnn[i] = ldg_nc_u2(reinterpret_cast<const uint16_t*>(p.cond) + r * p.ldc + col);
                        }
                    }
                }
            };
            if (sub < tw / 32 && n0 + 32 * sub < p.N) load_chunk(sub);
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
            for (int j = sub; j < tw / 32; j += 2) {
                const int col0 = n0 + 32 * j;
                if (col0 >= p.N) break;
                float acc[32];
                tmem_ld32(taddr + j * 32, acc);
                if (EPI == EPI_LINEAR && direct16) {          // N % 32 == 0: every chunk of the tile takes this path
                    // 16-bit output only: after tcgen05.ld a thread holds 32 consecutive columns of ITS row - bias / activation / pack
                    // in registers and four 16-byte stores, no trip through the staging tile (that transpose exists to coalesce the
                    // fp32 read-modify-write epilogues).  The vocoder's first convs and conv_pre, the aux decoder's GELU GEMM.
                    tmem_ld_wait();
                    const int t = t0 + q * 32 + lane;
                    if (t < p.T) {
                        uint32_t pk[16];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 bq = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            float4 v = make_float4(fmaf(p.alpha, acc[4 * c], bq.x), fmaf(p.alpha, acc[4 * c + 1], bq.y),
                                                   fmaf(p.alpha, acc[4 * c + 2], bq.z), fmaf(p.alpha, acc[4 * c + 3], bq.w));
                            if (p.act == ACT_RELU) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
                            else if (p.act == ACT_LRELU)
                                v = make_float4(fmaxf(v.x, LRELU_SLOPE * v.x), fmaxf(v.y, LRELU_SLOPE * v.y), fmaxf(v.z, LRELU_SLOPE * v.z),
                                                fmaxf(v.w, LRELU_SLOPE * v.w));
                            else if (p.act > ACT_RELU) v = act4_slow(v, p.act);
                            pk[2 * c] = Half16<BF16>::pack2(v.x, v.y);
                            pk[2 * c + 1] = Half16<BF16>::pack2(v.z, v.w);
                        }
                        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out_h) + ((long long)bt * p.T + t) * p.ldoh + col0);
#pragma unroll
                        for (int c = 0; c < 4; ++c) dst[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                    }
                    continue;
                }
                const int col = col0 + cl;
                const bool colok = okn;
                const EpiConst kc = kcn;
                float4 in[8];
                uint2 cn[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { in[i] = inn[i]; cn[i] = cnn[i]; }
                if (j + 2 < tw / 32 && col0 + 64 < p.N) load_chunk(j + 2);
                tmem_ld_wait();
                // transpose through the warp's private 16-row staging tile, two passes: thread = row -> lane = 4 columns
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if ((lane >> 4) == pass) {
                        float4* srow = reinterpret_cast<float4*>(stg + (lane & 15) * STG_LD);
#pragma unroll
                        for (int c = 0; c < 8; ++c) srow[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                    }
                    __syncwarp();
                    if (colok) {
#pragma unroll(EPI == EPI_LINEAR ? 1 : 4)
                        for (int i2 = 0; i2 < 4; ++i2) {
                            const int i = 4 * pass + i2;
                            const int t = tq + 4 * i;
                            if (t < p.T) {
                                const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i2 + rsub) * STG_LD + cl);
                                const int b = (p.d_stride != 0 && p.T_utt > 0) ? (bt * p.T + t) / p.T_utt : 0;
                                epilogue_quad<EPI, BF16>(p, v, in[i], kc, (long long)bt * p.T + t, b, col, fold ? &cn[i] : nullptr);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (relaxed_release) mbar_arrive_cluster_relaxed(mapa_u32(&tempty[as], 0));
                else mbar_arrive_cluster(mapa_u32(&tempty[as], 0));
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, TMEM_COLS);
    }
}

// ---- host side (tensor-map builders live in b2s_tc.cuh) ------------------------------------------------
template <int EPI, int BF16>
static int launch_one(const TcP& p, cudaStream_t st) {
#ifdef B2S_EXPERIMENTS
    static PerDevice configured;
    if (configured.first()) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EPI, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
#endif
    if (p.cg2) {
        static PerDevice configured2;
        constexpr bool can3 = EPI == EPI_GATE || EPI == EPI_LINEAR || EPI == EPI_VRES;      // 3-tap convs with the A slab resident across the taps
        constexpr int smem_max = can3 ? (SMEM3_BYTES > SMEM2_BYTES ? SMEM3_BYTES : SMEM2_BYTES) : SMEM2_BYTES;
        if (configured2.first()) {
            B2S_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_cg2_kernel<EPI, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        }
        const int num_pt = (p.B * p.tiles_m_per_b / 2) * p.tiles_n;
        // last-wave split: when the tiles left for the last wave (or all tiles of a small launch) occupy at most half of the CTA pairs,
        // each is cut along N into 2 or 4 sub-tiles (>= 64 columns, whole 32-column epilogue chunks) - an MMA of width 64 / 128 costs
        // 0.45x / 0.63x of one of width 256, so the partial wave shrinks accordingly (config 5: 384 tile pairs on 74 CTA pairs = 5.2 waves)
        static const bool no_split = getenv("B2S_GEMM_NOSPLIT") != nullptr;
        const int maxp = num_sms() / 2;
        TcP q = p;
        static const bool relaxed = [] { const char* e = getenv("B2S_GEMM_RELAXED_RELEASE"); return !e || atoi(e) != 0; }();
        q.relaxed_release = relaxed ? 1 : 0;
        q.vt_full = num_pt >= maxp ? (num_pt / maxp) * maxp : 0;
        q.vt_split = 1;
        const int rem = num_pt - q.vt_full;
        if (!no_split && !p.no_split && rem > 0 && 2 * rem <= maxp) {
            q.vt_split = maxp / rem >= 4 ? 4 : 2;
            while (q.vt_split > 1 && ((p.bn / q.vt_split) % 32 != 0 || p.bn / q.vt_split < 64)) q.vt_split >>= 1;
        }
        if (q.vt_split == 1) q.vt_full = num_pt;
        q.vt_total = q.vt_full + (num_pt - q.vt_full) * q.vt_split;
        const int pairs = q.vt_total < maxp ? q.vt_total : maxp;
        cudaLaunchConfig_t cfg2{};
        cfg2.gridDim = dim3(2 * pairs);
        cfg2.blockDim = dim3(NTHREADS);
        cfg2.dynamicSmemBytes = (can3 && p.conv3) ? SMEM3_BYTES : SMEM2_BYTES;
        cfg2.stream = st;
        cudaLaunchAttribute attr2[2];
        attr2[0].id = cudaLaunchAttributeClusterDimension;
        attr2[0].val.clusterDim.x = 2;
        attr2[0].val.clusterDim.y = 1;
        attr2[0].val.clusterDim.z = 1;
        attr2[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr2[1].val.programmaticStreamSerializationAllowed = 1;
        cfg2.attrs = attr2;
        cfg2.numAttrs = 2;
        B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg2, tc_gemm_cg2_kernel<EPI, BF16>, q));
        return B2S_OK;
    }
#ifndef B2S_EXPERIMENTS
    set_error("the single-CTA GEMM kernel is an experiment: build with B2S_BUILD_EXPERIMENTS=1");
    return B2S_ERR_UNSUPPORTED;
#else
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<EPI, BF16>, p));
    return B2S_OK;
#endif
}
template <int EPI>
static int launch(const TcP& p, int bf16, cudaStream_t st) {
    if (p.num_tiles <= 0) return B2S_OK;
    return bf16 ? launch_one<EPI, 1>(p, st) : launch_one<EPI, 0>(p, st);
}

// common geometry: A is [B, T, Kcols]; per-utterance M tiles when conv, flat otherwise
static int setup(TcP& p, const void* A, int lda, int a_cols, int B, int T, bool per_utt, const void* W, int ldw, int N, int K,
                 int kb_per_tap, int dil, int bf16) {
    const int Bm = per_utt ? B : 1, Tm = per_utt ? T : B * T;
    int rc = make_map_act(&p.mapA, A, bf16, a_cols, lda, Tm, Bm, BLOCK_K, BLOCK_M);
    if (rc) return rc;
#ifdef B2S_EXPERIMENTS
    static const bool cg1 = getenv("B2S_GEMM_CG1") != nullptr;      // A/B switch: the single-CTA kernel
#else
    const bool cg1 = false;
#endif
    p.cg2 = cg1 ? 0 : 1;
    p.bn = (p.cg2 && N % 192 == 0 && N % 256 != 0) ? 192 : BLOCK_N;      // no half-empty second tile for the C = 192 models
    if (p.cg2 && N <= 128) p.bn = N <= 64 ? 64 : 128;                    // narrow outputs (the vocoder's late stages, mel heads): an MMA of
                                                                         // width 64 / 128 is 2.2x / 1.6x shorter than one of width 256
    rc = make_map_w(&p.mapW, W, bf16, K, N, ldw, BLOCK_K, p.cg2 ? p.bn / 2 : BLOCK_N);
    if (rc) return rc;
    p.B = Bm; p.T = Tm; p.T_utt = T;
    p.N = N;
    p.num_kb = ceil_div(K, BLOCK_K);
    p.kb_per_tap = kb_per_tap > 0 ? kb_per_tap : p.num_kb;
    p.dil = dil;
    p.tap_c = 1;
    p.tiles_m_per_b = ceil_div(Tm, BLOCK_M);
    if (p.cg2) p.tiles_m_per_b = (p.tiles_m_per_b + 1) & ~1;       // whole CTA pairs; a padding tile has no valid row
    p.tiles_n = ceil_div(N, p.bn);
    p.num_tiles = Bm * p.tiles_m_per_b * p.tiles_n;
    return B2S_OK;
}


}  // namespace tc
}  // namespace b2s

using namespace b2s;
using namespace b2s::tc;

extern "C" int b2s_tc_linear(const void* A, int lda, int rows, int T, const void* W, int ldw, const float* bias, int N,
                             int K, float alpha, int act, float* out_f32, int ldo, void* out_h, int ldoh, void* y_h,
                             int ldy, const float* dvec, int d_stride, int bf16, void* stream) {
    B2S_CHECK_ARG(A && W && (out_f32 || out_h || y_h), "b2s_tc_linear: null pointer");
    B2S_CHECK_ARG(rows >= 0 && N > 0 && K > 0, "b2s_tc_linear: bad shape rows=%d N=%d K=%d", rows, N, K);
    B2S_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && al16(A) && al16(W),
                  "b2s_tc_linear: 16-bit operands need 16B-aligned bases and leading dimensions that are multiples of 8");
    B2S_CHECK_ARG((!out_f32 || (ldo % 4 == 0 && al16(out_f32))) && (!out_h || (ldoh % 8 == 0 && al16(out_h))) &&
                      (!y_h || (ldy % 8 == 0 && al16(y_h))), "b2s_tc_linear: misaligned output");
    B2S_CHECK_ARG(!y_h || (dvec && d_stride % 4 == 0 && al16(dvec)), "b2s_tc_linear: y needs a 16B-aligned dvec");
    if (rows == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, A, lda, K, 1, rows, false, W, ldw, N, K, 0, 0, bf16);
    if (rc) return rc;
    p.T_utt = T > 0 ? T : 0;
    p.bias = bias; p.alpha = alpha; p.act = act;
    p.out_f = out_f32; p.ldo = ldo; p.out_h = out_h; p.ldoh = ldoh; p.y_h = y_h; p.ldy = ldy;
    p.dvec = dvec; p.d_stride = d_stride;
    return launch<EPI_LINEAR>(p, bf16, (cudaStream_t)stream);
}

extern "C" int b2s_tc_cond_table(const void* cond_h, int rows, const void* Wc_h, const float* bc, int L, int N2, int H,
                                 void* table_h, int bf16, void* stream) {
    B2S_CHECK_ARG(cond_h && Wc_h && table_h, "b2s_tc_cond_table: null pointer");
    B2S_CHECK_ARG(L > 0 && N2 > 0 && N2 % 32 == 0 && H % 8 == 0, "b2s_tc_cond_table: bad dims L=%d N2=%d H=%d", L, N2, H);
    B2S_CHECK_ARG(al16(cond_h) && al16(Wc_h) && al16(table_h), "b2s_tc_cond_table: misaligned pointer");
    if (rows == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, cond_h, H, H, 1, rows, false, Wc_h, H, L * N2, H, 0, 0, bf16);
    if (rc) return rc;
    p.bias = bc; p.alpha = 1.f; p.act = ACT_NONE;
    p.out_h = table_h; p.ldoh = N2; p.oh_blk = N2; p.oh_blk_stride = (long long)rows * N2;
    return launch<EPI_LINEAR>(p, bf16, (cudaStream_t)stream);
}

extern "C" int b2s_tc_cond_table_tiled(const void* cond_h, int B, int T, const void* Wc_h, const float* bc, int L, int N2,
                                       int H, void* table_h, int bf16, void* stream) {
    B2S_CHECK_ARG(cond_h && Wc_h && table_h, "b2s_tc_cond_table_tiled: null pointer");
    B2S_CHECK_ARG(L > 0 && N2 > 0 && N2 % 32 == 0 && H % 8 == 0, "b2s_tc_cond_table_tiled: bad dims L=%d N2=%d H=%d", L, N2, H);
    B2S_CHECK_ARG(al16(cond_h) && al16(Wc_h) && al16(table_h), "b2s_tc_cond_table_tiled: misaligned pointer");
    if (B * T == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, cond_h, H, H, B, T, false, Wc_h, H, L * N2, H, 0, 0, bf16);
    if (rc) return rc;
    p.bias = bc; p.alpha = 1.f; p.act = ACT_NONE;
    p.out_h = table_h; p.oh_blk = N2; p.oh_tiled = 1; p.oh_B = B; p.oh_tpb = (ceil_div(T, 128) + 1) & ~1;
    return launch<EPI_LINEAR>(p, bf16, (cudaStream_t)stream);
}

extern "C" int b2s_tc_wavenet_gate_ld(const void* y_h, const void* Wd_h, const void* cond_h, int ld_cond, void* z_h, int ld_z,
                                      int B, int T, int C, int dilation, int bf16, void* stream) {
    B2S_CHECK_ARG(y_h && Wd_h && cond_h && z_h, "b2s_tc_wavenet_gate: null pointer");
    B2S_CHECK_ARG(C % 64 == 0, "b2s_tc_wavenet_gate: the tensor-core path needs residual channels %% 64 == 0 (C=%d)", C);
    B2S_CHECK_ARG(dilation >= 1 && ld_cond % 8 == 0 && ld_z >= C && ld_z % 8 == 0 && al16(cond_h) && al16(y_h) && al16(z_h) && al16(Wd_h),
                  "b2s_tc_wavenet_gate: bad dilation / leading dimension / alignment");
    if (B * T == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, y_h, C, C, B, T, true, Wd_h, 3 * C, 2 * C, 3 * C, C / BLOCK_K, dilation, bf16);
    if (rc) return rc;
    p.cond = cond_h; p.ldc = ld_cond; p.out_h = z_h; p.ldoh = ld_z;
    static const bool conv3_on = [] { const char* e = getenv("B2S_GATE_CONV3"); return !e || atoi(e) != 0; }();
    if (p.cg2 && conv3_on && dilation <= CONV3_HALO) {
        rc = make_map_act(&p.mapA3, y_h, bf16, C, C, T, B, BLOCK_K, A3_ROWS);
        if (rc) return rc;
        p.conv3 = 1;
    }
    return launch<EPI_GATE>(p, bf16, (cudaStream_t)stream);
}

extern "C" int b2s_tc_wavenet_gate(const void* y_h, const void* Wd_h, const void* cond_h, int ld_cond, void* z_h, int B,
                                   int T, int C, int dilation, int bf16, void* stream) {
    return b2s_tc_wavenet_gate_ld(y_h, Wd_h, cond_h, ld_cond, z_h, C, B, T, C, dilation, bf16, stream);
}

extern "C" int b2s_tc_wavenet_res(const void* z_h, int ld_z, const void* Wres_h, const float* b_res, float* x, void* y_next_h,
                                  const float* dvec_next, int d_stride, int B, int T, int C, int bf16, void* stream) {
    B2S_CHECK_ARG(z_h && Wres_h && b_res && x, "b2s_tc_wavenet_res: null pointer");
    B2S_CHECK_ARG(C % 64 == 0 && ld_z >= C && ld_z % 8 == 0, "b2s_tc_wavenet_res: bad C=%d / ld_z=%d", C, ld_z);
    B2S_CHECK_ARG(!y_next_h || (dvec_next && d_stride % 4 == 0 && al16(dvec_next)), "b2s_tc_wavenet_res: y_next needs dvec_next");
    B2S_CHECK_ARG(al16(z_h) && al16(Wres_h) && al16(b_res) && al16(x), "b2s_tc_wavenet_res: misaligned pointer");
    if (B * T == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, z_h, ld_z, C, B, T, false, Wres_h, C, C, C, 0, 0, bf16);     // N = C: the residual rows only
    if (rc) return rc;
    p.bias = b_res; p.x = x; p.y_h = y_next_h; p.ldy = C; p.C = C;
    p.dvec = dvec_next; p.d_stride = d_stride;
    return launch<EPI_RESSKIP>(p, bf16, (cudaStream_t)stream);
}

extern "C" int b2s_tc_wavenet_out(const void* z_h, const void* Wo_h, const float* bo, float* x, void* y_next_h,
                                  float* skip, void* skip_h, const float* dvec_next, int d_stride, int first_layer, int B,
                                  int T, int C, int bf16, void* stream) {
    B2S_CHECK_ARG(z_h && Wo_h && bo && x && skip, "b2s_tc_wavenet_out: null pointer");
    B2S_CHECK_ARG(C % 64 == 0, "b2s_tc_wavenet_out: the tensor-core path needs residual channels %% 64 == 0 (C=%d)", C);
    B2S_CHECK_ARG(!y_next_h || (dvec_next && d_stride % 4 == 0 && al16(dvec_next)), "b2s_tc_wavenet_out: y_next needs dvec_next");
    B2S_CHECK_ARG(al16(z_h) && al16(Wo_h) && al16(bo) && al16(x) && al16(skip), "b2s_tc_wavenet_out: misaligned pointer");
    if (B * T == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, z_h, C, C, B, T, false, Wo_h, C, 2 * C, C, 0, 0, bf16);
    if (rc) return rc;
    p.bias = bo; p.x = x; p.y_h = y_next_h; p.ldy = C; p.skip = skip; p.skip_h = skip_h; p.C = C;
    p.dvec = dvec_next; p.d_stride = d_stride; p.first = first_layer;
    return launch<EPI_RESSKIP>(p, bf16, (cudaStream_t)stream);
}

extern "C" int b2s_tc_skip_sum(const void* z_all_h, const void* Wcat_h, const float* bias, void* out_h, int rows, int C, int L,
                               int bf16, void* stream) {
    B2S_CHECK_ARG(z_all_h && Wcat_h && bias && out_h, "b2s_tc_skip_sum: null pointer");
    B2S_CHECK_ARG(C % 64 == 0 && L >= 1, "b2s_tc_skip_sum: needs C %% 64 == 0 (C=%d) and L >= 1", C);
    B2S_CHECK_ARG(al16(z_all_h) && al16(Wcat_h) && al16(bias) && al16(out_h), "b2s_tc_skip_sum: misaligned pointer");
    if (rows == 0) return B2S_OK;
    TcP p{};
    // tiling of a flat [rows, L*C] x [L*C, C] GEMM; the A map is 3-D {C, rows, L} so that K block kb reads layer kb / (C/64)
    int rc = setup(p, z_all_h, C, C, 1, rows, false, Wcat_h, L * C, C, L * C, C / BLOCK_K, 0, bf16);
    if (rc) return rc;
    rc = make_map_act(&p.mapA, z_all_h, bf16, C, C, rows, L, BLOCK_K, BLOCK_M);
    if (rc) return rc;
    p.k_layered = 1;
    p.dil = 0;
    p.bias = bias; p.alpha = 1.0f; p.act = ACT_NONE; p.out_h = out_h; p.ldoh = C;
    return launch<EPI_LINEAR>(p, bf16, (cudaStream_t)stream);
}

extern "C" int b2s_tc_lynx_glu(const void* h_h, const void* W_h, const float* bias, void* g_h, int rows, int C, int inner,
                               int bf16, void* stream) {
    B2S_CHECK_ARG(h_h && W_h && bias && g_h, "b2s_tc_lynx_glu: null pointer");
    B2S_CHECK_ARG(C % 8 == 0 && inner % 16 == 0, "b2s_tc_lynx_glu: bad dims C=%d inner=%d", C, inner);
    B2S_CHECK_ARG(al16(h_h) && al16(W_h) && al16(bias) && al16(g_h), "b2s_tc_lynx_glu: misaligned pointer");
    if (rows == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, h_h, C, C, 1, rows, false, W_h, C, 2 * inner, C, 0, 0, bf16);
    if (rc) return rc;
    p.bias = bias; p.out_h = g_h; p.ldoh = inner;
    return launch<EPI_SWIGLU>(p, bf16, (cudaStream_t)stream);
}

static int linear_residual_impl(const void* p_h, const void* W_h, const float* bias, float* x, const void* cond_next_h, int ld_cond,
                                int rows, int C, int inner, int bf16, void* stream) {
    B2S_CHECK_ARG(p_h && W_h && bias && x, "b2s_tc_linear_residual: null pointer");
    B2S_CHECK_ARG(C % 32 == 0 && inner % 8 == 0, "b2s_tc_linear_residual: bad dims C=%d inner=%d", C, inner);
    B2S_CHECK_ARG(al16(p_h) && al16(W_h) && al16(bias) && al16(x), "b2s_tc_linear_residual: misaligned pointer");
    B2S_CHECK_ARG(!cond_next_h || (ld_cond % 4 == 0 && (reinterpret_cast<uintptr_t>(cond_next_h) & 7) == 0),
                  "b2s_tc_linear_residual_cond: cond table must be 8-byte aligned with ld_cond %% 4 == 0");
    if (rows == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, p_h, inner, inner, 1, rows, false, W_h, inner, C, inner, 0, 0, bf16);
    if (rc) return rc;
    p.bias = bias; p.x = x; p.C = C; p.cond = cond_next_h; p.ldc = ld_cond;
    return launch<EPI_RESIDUAL>(p, bf16, (cudaStream_t)stream);
}

// Dense Conv1d along time (stride 1, 'same' zero padding per utterance) as ONE GEMM over ksize taps: A = [B, T, Cin] 16-bit,
// W = [N, ksize * Cin] (column = tap * Cin + c), out = act(A * W + bias) as fp32 and / or 16-bit rows (aux_decoder/convnext.py:64-67, :73-76)
extern "C" int b2s_tc_conv1d(const void* a_h, const void* W_h, const float* bias, float* out_f32, int ldo, void* out_h, int ldoh,
                             int B, int T, int Cin, int N, int ksize, int act, int bf16, void* stream) {
    B2S_CHECK_ARG(a_h && W_h && (out_f32 || out_h), "b2s_tc_conv1d: null pointer");
    B2S_CHECK_ARG(Cin % 64 == 0 && N > 0 && ksize >= 1 && (ksize & 1) && ksize <= 63, "b2s_tc_conv1d: needs Cin %% 64 == 0 (Cin=%d) and an odd kernel size (%d)", Cin, ksize);
    B2S_CHECK_ARG(al16(a_h) && al16(W_h) && (!out_f32 || (ldo % 4 == 0 && al16(out_f32))) && (!out_h || (ldoh % 8 == 0 && al16(out_h))),
                  "b2s_tc_conv1d: misaligned pointer / leading dimension");
    if (B * T == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, a_h, Cin, Cin, B, T, true, W_h, ksize * Cin, N, ksize * Cin, Cin / BLOCK_K, 1, bf16);
    if (rc) return rc;
    p.tap_c = ksize / 2;
    p.bias = bias; p.alpha = 1.0f; p.act = act;
    p.out_f = out_f32; p.ldo = ldo; p.out_h = out_h; p.ldoh = ldoh;
    return launch<EPI_LINEAR>(p, bf16, (cudaStream_t)stream);
}

// Three taps with |shift| <= CONV3_HALO rows: ONE load of the activation slab with its halo serves all taps (row-shifted descriptors), as in
// the WaveNet gate GEMM - most of the vocoder's convs after row folding, and its transposed convs.  68 KB instead of 96 KB of L2 -> SM
// traffic per three K blocks; the wide (N = 256) folded convs sit on that cap.
static int conv3_if_three_taps(TcP& p, const void* a_h, int B, int T, int Cin, int ksize, int dil, int bf16) {
    static const bool on = [] { const char* e = getenv("B2S_VOC_CONV3"); return !e || atoi(e) != 0; }();
    if (!on || !p.cg2 || ksize != 3 || dil > CONV3_HALO) return B2S_OK;
    int rc = make_map_act(&p.mapA3, a_h, bf16, Cin, Cin, T, B, BLOCK_K, A3_ROWS);
    if (rc) return rc;
    p.conv3 = 1;
    return B2S_OK;
}

// Dilated dense Conv1d as one GEMM over ksize taps (the vocoder's residual-block convs, nsf_hifigan/models.py:39-58; its transposed
// convs are 3-tap convs over u * Cout output columns, see vocoder.py)
extern "C" int b2s_tc_conv1d_dil(const void* a_h, const void* W_h, const float* bias, float* out_f32, int ldo, void* out_h, int ldoh,
                                 int B, int T, int Cin, int N, int ksize, int dil, int act, int bf16, void* stream) {
    B2S_CHECK_ARG(a_h && W_h && (out_f32 || out_h), "b2s_tc_conv1d_dil: null pointer");
    B2S_CHECK_ARG(Cin % 64 == 0 && N > 0 && ksize >= 1 && (ksize & 1) && ksize <= 63 && dil >= 1,
                  "b2s_tc_conv1d_dil: needs Cin %% 64 == 0 (Cin=%d), an odd kernel size (%d) and dilation >= 1 (%d)", Cin, ksize, dil);
    B2S_CHECK_ARG(al16(a_h) && al16(W_h) && (!out_f32 || (ldo % 4 == 0 && al16(out_f32))) && (!out_h || (ldoh % 8 == 0 && al16(out_h))),
                  "b2s_tc_conv1d_dil: misaligned pointer / leading dimension");
    if (B * T == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, a_h, Cin, Cin, B, T, true, W_h, ksize * Cin, N, ksize * Cin, Cin / BLOCK_K, dil, bf16);
    if (rc) return rc;
    p.tap_c = ksize / 2;
    rc = conv3_if_three_taps(p, a_h, B, T, Cin, ksize, dil, bf16);
    if (rc) return rc;
    p.bias = bias; p.alpha = 1.0f; p.act = act;
    p.out_f = out_f32; p.ldo = ldo; p.out_h = out_h; p.ldoh = ldoh;
    p.no_split = 1;            // measured: 1.97 ms (no split) vs 2.03 ms per 8-s utterance with the blocks on three streams
    return launch<EPI_LINEAR>(p, bf16, (cudaStream_t)stream);
}

// x <- x_src + conv(a) + bias on the fp32 residual stream, y_h <- leaky_relu(x, y_slope) as 16-bit rows for the next conv
// (ResBlock1 / ResBlock2, nsf_hifigan/models.py:60-68, :90-95).  x_src NULL: x itself; y_h NULL: no 16-bit copy; y_slope 1: plain copy.
extern "C" int b2s_tc_conv1d_residual(const void* a_h, const void* W_h, const float* bias, const float* x_src, float* x, void* y_h,
                                      float y_slope, int B, int T, int Cin, int N, int ksize, int dil, int bf16, void* stream) {
    B2S_CHECK_ARG(a_h && W_h && x, "b2s_tc_conv1d_residual: null pointer");
    B2S_CHECK_ARG(Cin % 64 == 0 && N > 0 && N % 8 == 0 && ksize >= 1 && (ksize & 1) && ksize <= 63 && dil >= 1,
                  "b2s_tc_conv1d_residual: needs Cin %% 64 == 0 (Cin=%d), N %% 8 == 0 (N=%d), an odd kernel size (%d), dilation >= 1 (%d)",
                  Cin, N, ksize, dil);
    B2S_CHECK_ARG(al16(a_h) && al16(W_h) && al16(x) && (!x_src || al16(x_src)) && (!y_h || al16(y_h)) && (!bias || al16(bias)),
                  "b2s_tc_conv1d_residual: misaligned pointer");
    B2S_CHECK_ARG(y_h != a_h, "b2s_tc_conv1d_residual: y_h must not alias the conv input (neighbouring tiles read its halo rows)");
    if (B * T == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, a_h, Cin, Cin, B, T, true, W_h, ksize * Cin, N, ksize * Cin, Cin / BLOCK_K, dil, bf16);
    if (rc) return rc;
    p.tap_c = ksize / 2;
    rc = conv3_if_three_taps(p, a_h, B, T, Cin, ksize, dil, bf16);
    if (rc) return rc;
    p.bias = bias; p.x = x; p.x_src = x_src; p.C = N; p.y_h = y_h; p.ldy = N;
    p.y_lrelu = y_h != nullptr && y_slope != 1.0f; p.y_slope = y_slope;
    p.no_split = 1;
    return launch<EPI_VRES>(p, bf16, (cudaStream_t)stream);
}

// x <- x + gamma * (p * W^T + bias) (fp32 residual stream, per-channel layer scale), optionally also x as 16-bit rows
// (aux_decoder/convnext.py:49-57)
extern "C" int b2s_tc_linear_residual_scaled(const void* p_h, const void* W_h, const float* bias, const float* gamma, float* x,
                                             void* x_h, int rows, int C, int inner, int bf16, void* stream) {
    B2S_CHECK_ARG(p_h && W_h && bias && x, "b2s_tc_linear_residual_scaled: null pointer");
    B2S_CHECK_ARG(C % 32 == 0 && inner % 8 == 0, "b2s_tc_linear_residual_scaled: bad dims C=%d inner=%d", C, inner);
    B2S_CHECK_ARG(al16(p_h) && al16(W_h) && al16(bias) && al16(x) && (!gamma || al16(gamma)) && (!x_h || al16(x_h)),
                  "b2s_tc_linear_residual_scaled: misaligned pointer");
    if (rows == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, p_h, inner, inner, 1, rows, false, W_h, inner, C, inner, 0, 0, bf16);
    if (rc) return rc;
    p.bias = bias; p.x = x; p.C = C; p.dvec = gamma; p.y_h = x_h; p.ldy = C;
    return launch<EPI_RESIDUAL>(p, bf16, (cudaStream_t)stream);
}

extern "C" int b2s_tc_linear_residual(const void* p_h, const void* W_h, const float* bias, float* x, int rows, int C,
                                      int inner, int bf16, void* stream) {
    return linear_residual_impl(p_h, W_h, bias, x, nullptr, 0, rows, C, inner, bf16, stream);
}

extern "C" int b2s_tc_linear_residual_cond(const void* p_h, const void* W_h, const float* bias, float* x, const void* cond_next_h,
                                           int ld_cond, int rows, int C, int inner, int bf16, void* stream) {
    B2S_CHECK_ARG(cond_next_h, "b2s_tc_linear_residual_cond: null cond table");
    return linear_residual_impl(p_h, W_h, bias, x, cond_next_h, ld_cond, rows, C, inner, bf16, stream);
}
