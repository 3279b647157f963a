// Tensor-core (tcgen05 / TMEM / TMA) GEMM family of the 16-bit path.  sm_100a only.
//
//   out[rows, N] = epilogue( A[rows, K] . W[N, K]^T )          A, W: bf16 or fp16, fp32 accumulate in TMEM
//
// One persistent, warp-specialised kernel (256 threads, 1 CTA / SM):
//   warp 0    TMA producer   - cp.async.bulk.tensor tiles (128B swizzle) into a 4-stage smem ring
//   warp 1    MMA issuer     - one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16)
//   warp 2    TMEM allocator - 512 columns = two 128x256 fp32 accumulators (double buffered)
//   warps 4-7 epilogue       - tcgen05.ld (32 lanes x 32 columns per warp per step), fused math, global stores;
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
#include "b2s_tc.cuh"

namespace b2s {
namespace tc {

constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 64, UMMA_K = 16, STAGES = 4;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;          // 16 KB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;          // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;          // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;
constexpr int NTHREADS = 256;

enum Epi : int { EPI_LINEAR = 0, EPI_GATE = 1, EPI_RESSKIP = 2, EPI_SWIGLU = 3, EPI_RESIDUAL = 4 };

struct __align__(64) TcP {
    CUtensorMap mapA, mapW;
    int B, T;                  // utterance grid of the A tiles (flat GEMM: B = 1, T = rows)
    int T_utt;                 // frames per utterance, for the step-embedding row lookup (b = row / T_utt)
    int N, num_kb, kb_per_tap, dil;
    int tiles_m_per_b, tiles_n, num_tiles;
    const float* bias; float alpha; int act;
    float* out_f; int ldo;
    void* out_h; int ldoh;
    void* y_h; int ldy;
    const float* dvec; int d_stride;
    const void* cond; int ldc;
    float* x; float* skip; void* skip_h; int C; int first;
};

// ---- 16-bit helpers --------------------------------------------------------------------------------
template <int BF16>
__device__ __forceinline__ void store_h1(void* base, long long idx, float v) {
    if (BF16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
}
template <int BF16>
__device__ __forceinline__ void store_h32(void* base, long long idx, const float* v) {   // 32 values, 64 B
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + idx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 o;
        o.x = Half16<BF16>::pack2(v[8 * i + 0], v[8 * i + 1]);
        o.y = Half16<BF16>::pack2(v[8 * i + 2], v[8 * i + 3]);
        o.z = Half16<BF16>::pack2(v[8 * i + 4], v[8 * i + 5]);
        o.w = Half16<BF16>::pack2(v[8 * i + 6], v[8 * i + 7]);
        dst[i] = o;
    }
}
template <int BF16>
__device__ __forceinline__ void store_h16(void* base, long long idx, const float* v) {   // 16 values, 32 B
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + idx);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        uint4 o;
        o.x = Half16<BF16>::pack2(v[8 * i + 0], v[8 * i + 1]);
        o.y = Half16<BF16>::pack2(v[8 * i + 2], v[8 * i + 3]);
        o.z = Half16<BF16>::pack2(v[8 * i + 4], v[8 * i + 5]);
        o.w = Half16<BF16>::pack2(v[8 * i + 6], v[8 * i + 7]);
        dst[i] = o;
    }
}
__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---- epilogues: one thread = one output row, 32 consecutive columns [col0, col0+32) ------------------
template <int EPI, int BF16>
__device__ __forceinline__ void epilogue_chunk(const TcP& p, float* acc, long long r, int b, int col0) {
    if (EPI == EPI_LINEAR) {
        const int nvalid = min(32, p.N - col0);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float bias = (p.bias && i < nvalid) ? __ldg(p.bias + col0 + i) : 0.f;
            v[i] = apply_act(fmaf(p.alpha, acc[i], bias), p.act);
        }
        if (nvalid == 32) {
            if (p.out_f) {
                float4* dst = reinterpret_cast<float4*>(p.out_f + r * p.ldo + col0);
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (p.out_h) store_h32<BF16>(p.out_h, r * p.ldoh + col0, v);
            if (p.y_h) {
                const float4* d = reinterpret_cast<const float4*>(p.dvec + (long long)b * p.d_stride + col0);
                float yv[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 dd = __ldg(d + i);
                    yv[4 * i] = v[4 * i] + dd.x; yv[4 * i + 1] = v[4 * i + 1] + dd.y;
                    yv[4 * i + 2] = v[4 * i + 2] + dd.z; yv[4 * i + 3] = v[4 * i + 3] + dd.w;
                }
                store_h32<BF16>(p.y_h, r * p.ldy + col0, yv);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (i < nvalid) {
                    if (p.out_f) p.out_f[r * p.ldo + col0 + i] = v[i];
                    if (p.out_h) store_h1<BF16>(p.out_h, r * p.ldoh + col0 + i, v[i]);
                    if (p.y_h)
                        store_h1<BF16>(p.y_h, r * p.ldy + col0 + i, v[i] + __ldg(p.dvec + (long long)b * p.d_stride + col0 + i));
                }
            }
        }
    } else if (EPI == EPI_GATE || EPI == EPI_SWIGLU) {
        // packed columns: even = gate (GATE) / out (SWIGLU), odd = filter (GATE) / gate (SWIGLU)
        float z[16];
        if (EPI == EPI_GATE) {
            const uint16_t* cp = reinterpret_cast<const uint16_t*>(p.cond) + r * p.ldc + col0;
            uint4 c[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) c[i] = ldg_nc_u4(cp + 8 * i);
            const uint32_t* cw = reinterpret_cast<const uint32_t*>(c);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 cf = Half16<BF16>::unpack2(cw[i]);
                z[i] = sigmoid_fast(acc[2 * i] + cf.x) * tanh_fast(acc[2 * i + 1] + cf.y);
            }
        } else {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 bb = __ldg(bp + i);
                const float g0 = acc[4 * i + 1] + bb.y, g1 = acc[4 * i + 3] + bb.w;
                z[2 * i] = (acc[4 * i] + bb.x) * (g0 * sigmoid_fast(g0));
                z[2 * i + 1] = (acc[4 * i + 2] + bb.z) * (g1 * sigmoid_fast(g1));
            }
        }
        store_h16<BF16>(p.out_h, r * p.ldoh + (col0 >> 1), z);
    } else if (EPI == EPI_RESIDUAL) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
        float4* xp = reinterpret_cast<float4*>(p.x + r * p.C + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(bp + i);
            float4 xo = xp[i];
            xo.x += acc[4 * i] + bb.x; xo.y += acc[4 * i + 1] + bb.y;
            xo.z += acc[4 * i + 2] + bb.z; xo.w += acc[4 * i + 3] + bb.w;
            xp[i] = xo;
        }
    } else {   // EPI_RESSKIP: reference column order, [0, C) residual, [C, 2C) skip
        const float inv_sqrt2 = 0.70710678118654752440f;
        const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
        if (col0 < p.C) {
            float4* xp = reinterpret_cast<float4*>(p.x + r * p.C + col0);
            float xn[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 bb = __ldg(bp + i);
                const float4 xo = xp[i];
                xn[4 * i] = (xo.x + acc[4 * i] + bb.x) * inv_sqrt2;
                xn[4 * i + 1] = (xo.y + acc[4 * i + 1] + bb.y) * inv_sqrt2;
                xn[4 * i + 2] = (xo.z + acc[4 * i + 2] + bb.z) * inv_sqrt2;
                xn[4 * i + 3] = (xo.w + acc[4 * i + 3] + bb.w) * inv_sqrt2;
                xp[i] = make_float4(xn[4 * i], xn[4 * i + 1], xn[4 * i + 2], xn[4 * i + 3]);
            }
            if (p.y_h) {
                const float4* d = reinterpret_cast<const float4*>(p.dvec + (long long)b * p.d_stride + col0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 dd = __ldg(d + i);
                    xn[4 * i] += dd.x; xn[4 * i + 1] += dd.y; xn[4 * i + 2] += dd.z; xn[4 * i + 3] += dd.w;
                }
                store_h32<BF16>(p.y_h, r * p.ldy + col0, xn);
            }
        } else {
            float4* sp = reinterpret_cast<float4*>(p.skip + r * p.C + (col0 - p.C));
            float s[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 bb = __ldg(bp + i);
                s[4 * i] = acc[4 * i] + bb.x; s[4 * i + 1] = acc[4 * i + 1] + bb.y;
                s[4 * i + 2] = acc[4 * i + 2] + bb.z; s[4 * i + 3] = acc[4 * i + 3] + bb.w;
                if (!p.first) {
                    const float4 so = sp[i];
                    s[4 * i] += so.x; s[4 * i + 1] += so.y; s[4 * i + 2] += so.z; s[4 * i + 3] += so.w;
                }
                sp[i] = make_float4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
            }
            if (p.skip_h) store_h32<BF16>(p.skip_h, r * p.C + (col0 - p.C), s);
        }
    }
}

// ---- the kernel ------------------------------------------------------------------------------------
template <int EPI, int BF16>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_kernel(const __grid_constant__ TcP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
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
            mbar_init(&tempty[i], 4);           // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

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
                    tma_load_3d(sa, &p.mapA, &full[stage], c0, t0 + (tap - 1) * p.dil, b);
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
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.tiles_n, m_tile = tile / p.tiles_n;
            const int bt = m_tile / p.tiles_m_per_b, t0 = (m_tile - bt * p.tiles_m_per_b) * BLOCK_M;
            const int n0 = n_tile * BLOCK_N;
            const int t = t0 + q * 32 + lane;
            const bool valid = t < p.T;
            const long long r = (long long)bt * p.T + t;
            const int b = p.T_utt > 0 ? (int)(r / p.T_utt) : 0;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
            for (int j = 0; j < BLOCK_N / 32; ++j) {
                const int col0 = n0 + 32 * j;
                if (col0 >= p.N) break;
                float acc[32];
                tmem_ld32(taddr + j * 32, acc);
                tmem_ld_wait();
                if (valid) epilogue_chunk<EPI, BF16>(p, acc, r, b, col0);
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

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// activations [B, T, cols] (row stride ld elements) -> 3-D map {cols, T, B}, box {64, 128, 1}
static int make_map_act(CUtensorMap* m, const void* base, int bf16, int cols, int ld, int T, int B) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B2S_ERR_CUDA; }
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
    cuuint32_t box[3] = {BLOCK_K, BLOCK_M, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                     const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(activations cols=%d ld=%d T=%d B=%d) failed with CUresult %d", cols, ld, T, B, (int)rc);
        return B2S_ERR_CUDA;
    }
    return B2S_OK;
}
// weights [N, K] (row stride ldw) -> 2-D map {K, N}, box {64, 256}
static int make_map_w(CUtensorMap* m, const void* base, int bf16, int K, int N, int ldw) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B2S_ERR_CUDA; }
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)ldw * 2};
    cuuint32_t box[2] = {BLOCK_K, BLOCK_N};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights K=%d N=%d ldw=%d) failed with CUresult %d", K, N, ldw, (int)rc);
        return B2S_ERR_CUDA;
    }
    return B2S_OK;
}

static int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int EPI, int BF16>
static int launch_one(const TcP& p, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EPI, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    tc_gemm_kernel<EPI, BF16><<<grid, NTHREADS, SMEM_BYTES, st>>>(p);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
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
    int rc = make_map_act(&p.mapA, A, bf16, a_cols, lda, Tm, Bm);
    if (rc) return rc;
    rc = make_map_w(&p.mapW, W, bf16, K, N, ldw);
    if (rc) return rc;
    p.B = Bm; p.T = Tm; p.T_utt = T;
    p.N = N;
    p.num_kb = ceil_div(K, BLOCK_K);
    p.kb_per_tap = kb_per_tap > 0 ? kb_per_tap : p.num_kb;
    p.dil = dil;
    p.tiles_m_per_b = ceil_div(Tm, BLOCK_M);
    p.tiles_n = ceil_div(N, BLOCK_N);
    p.num_tiles = Bm * p.tiles_m_per_b * p.tiles_n;
    return B2S_OK;
}

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

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

extern "C" int b2s_tc_wavenet_gate(const void* y_h, const void* Wd_h, const void* cond_h, int ld_cond, void* z_h, int B,
                                   int T, int C, int dilation, int bf16, void* stream) {
    B2S_CHECK_ARG(y_h && Wd_h && cond_h && z_h, "b2s_tc_wavenet_gate: null pointer");
    B2S_CHECK_ARG(C % 64 == 0, "b2s_tc_wavenet_gate: the tensor-core path needs residual channels %% 64 == 0 (C=%d)", C);
    B2S_CHECK_ARG(dilation >= 1 && ld_cond % 8 == 0 && al16(cond_h) && al16(y_h) && al16(z_h) && al16(Wd_h),
                  "b2s_tc_wavenet_gate: bad dilation / alignment");
    if (B * T == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, y_h, C, C, B, T, true, Wd_h, 3 * C, 2 * C, 3 * C, C / BLOCK_K, dilation, bf16);
    if (rc) return rc;
    p.cond = cond_h; p.ldc = ld_cond; p.out_h = z_h; p.ldoh = C;
    return launch<EPI_GATE>(p, bf16, (cudaStream_t)stream);
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

extern "C" int b2s_tc_linear_residual(const void* p_h, const void* W_h, const float* bias, float* x, int rows, int C,
                                      int inner, int bf16, void* stream) {
    B2S_CHECK_ARG(p_h && W_h && bias && x, "b2s_tc_linear_residual: null pointer");
    B2S_CHECK_ARG(C % 32 == 0 && inner % 8 == 0, "b2s_tc_linear_residual: bad dims C=%d inner=%d", C, inner);
    B2S_CHECK_ARG(al16(p_h) && al16(W_h) && al16(bias) && al16(x), "b2s_tc_linear_residual: misaligned pointer");
    if (rows == 0) return B2S_OK;
    TcP p{};
    int rc = setup(p, p_h, inner, inner, 1, rows, false, W_h, inner, C, inner, 0, 0, bf16);
    if (rc) return rc;
    p.bias = bias; p.x = x; p.C = C;
    return launch<EPI_RESIDUAL>(p, bf16, (cudaStream_t)stream);
}
