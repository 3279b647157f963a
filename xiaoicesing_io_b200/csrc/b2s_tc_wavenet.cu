// One fused kernel per WaveNet residual layer (wavenet.py:33-48) on tcgen05 / TMEM / TMA.  sm_100a only.
//
// One CTA = one tile of 128 frames of one utterance; everything between the layer's input y = x + step embedding
// and its outputs (x', y_next, skip) stays on chip:
//
//   GEMM1  [128 x 768] . [768 x 512]   implicit-GEMM dilated k=3 conv: 3 TMA tiles per K slab at time offsets
//                                      -d, 0, +d (out-of-bounds zero fill = per-utterance zero padding, H1)
//          two N halves of 256 packed (gate j, filter j) columns -> all 512 TMEM columns
//   EPI1   + hoisted 16-bit conditioner projection, z = sigmoid(g) * tanh(f); z is written STRAIGHT into shared
//          memory in the 128B-swizzled K-major layout the tensor core reads (never touches HBM / L2)
//   GEMM2  [128 x 256] . [256 x 512]   output projection, A operand = the z tile in shared memory; the residual
//          half re-uses the TMEM columns of GEMM1's first half as soon as EPI1 has drained them
//   EPI2   x <- (x + r + b)/sqrt2 (fp32), y_next <- x + d_next (16-bit, other buffer: neighbours still read the
//          halo of y), skip (+)= s + b (fp32)          - coalesced through a per-warp smem transpose
//
// Warp roles: 0 = TMA producer (3-stage ring: 24 conv fills of A+B, 8 output-projection fills of B), 1 = MMA
// issuer, 2 = TMEM allocator, 4-11 = epilogue (two warps per TMEM lane quarter, alternating 32-column chunks).  EPI1 of half 0 overlaps the MMAs of half 1, GEMM2-residual K slabs
// 0-1 overlap EPI1 of half 1, EPI2-residual overlaps the skip MMAs.
// Programmatic dependent launch: the prologue (barrier init, TMEM alloc, descriptor prefetch) of layer l+1 overlaps
// the tail of layer l; dependent global data is only touched after griddepcontrol.wait.
#include "b2s_tc.cuh"

#include <math.h>
#include <stdlib.h>
#include <type_traits>

namespace b2s {
namespace tc {
namespace wl {

constexpr int C = 256;                                   // residual channels this kernel is specialised for
constexpr int BM = 128, BK = 64, BN = 256, UK = 16, STAGES = 3;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int Z_BYTES = BM * C * 2;                      // 4 K slabs of [128 x 64], 16 KB each
constexpr int G1_KB = 3 * C / BK, G2_KB = C / BK;        // 12, 4
constexpr int STG_LD = 36;
constexpr int STG_WARP_BYTES = 32 * STG_LD * 4;          // 4608
constexpr int SMEM_BYTES = Z_BYTES + STAGES * STAGE_BYTES + 256 + 1024;
constexpr int NTHREADS = 384;                           // 4 control warps + 8 epilogue warps (two per TMEM lane quarter)
constexpr int EPI_WARPS = 8;

struct __align__(64) LayerP {
    CUtensorMap mapY, mapWd, mapWo;
    int B, T, tiles_per_b, dil;
    const void* cond; int ldc;
    const float* bo;
    float* x; void* y_next; float* skip; void* skip_h;
    const float* dvec; int d_stride; int first;
};

__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_shared_u4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_shared_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef B2S_EXPERIMENTS     // version 1 (single CTA per tile): superseded by the 2-CTA cluster version below
template <int BF16>
__global__ void __launch_bounds__(NTHREADS, 1) wavenet_layer_kernel(const __grid_constant__ LayerP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* zs = smem;                                   // z tile, 4 swizzled K slabs
    uint8_t* stages = smem + Z_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* accb = empty + STAGES;                      // [4]: G1 half 0, G1 half 1, G2 residual, G2 skip
    uint64_t* zready = accb + 4;                          // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(zready + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.tiles_per_b, t0 = (blockIdx.x - b * p.tiles_per_b) * BM;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.mapY);
        prefetch_tmap(&p.mapWd);
        prefetch_tmap(&p.mapWo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&accb[i], 1);
        for (int i = 0; i < 2; ++i) mbar_init(&zready[i], EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    pdl_launch_dependents();       // the next layer may start its prologue on idle SMs
    pdl_wait();                    // everything below reads what the previous kernel wrote

    if (warp == 0) {
        // ===================== TMA producer: 24 conv fills, then 8 output-projection fills =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int f = 0; f < 2 * G1_KB + 2 * G2_KB; ++f) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0) {
                uint8_t* sa = stages + stage * STAGE_BYTES;
                if (f < 2 * G1_KB) {
                    const int h = f / G1_KB, kb = f - h * G1_KB;
                    const int tap = kb / (C / BK), c0 = (kb - tap * (C / BK)) * BK;
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_3d(sa, &p.mapY, &full[stage], c0, t0 + (tap - 1) * p.dil, b);
                    tma_load_2d(sa + A_BYTES, &p.mapWd, &full[stage], kb * BK, h * BN);
                } else {
                    const int g = (f - 2 * G1_KB) / G2_KB, kb = (f - 2 * G1_KB) - g * G2_KB;
                    mbar_expect_tx(&full[stage], B_BYTES);
                    tma_load_2d(sa + A_BYTES, &p.mapWo, &full[stage], kb * BK, g * BN);
                }
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_f16(BM, BN, BF16);
        int stage = 0;
        uint32_t phase = 0;
        // GEMM1: two halves of 256 packed columns
        for (int h = 0; h < 2; ++h) {
            const uint32_t d_tmem = tmem_base + h * BN;
            for (int kb = 0; kb < G1_KB; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(stages + stage * STAGE_BYTES), b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k)
                        umma_ss(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UK * 2)), make_sw128_kmajor_desc(b_addr + k * (UK * 2)),
                                idesc, (kb | k) != 0);
                    umma_commit(&empty[stage]);
                    if (kb == G1_KB - 1) umma_commit(&accb[h]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        // GEMM2: A = z tile in shared memory; g = 0 residual -> columns [0,256), g = 1 skip -> [256,512)
        for (int g = 0; g < 2; ++g) {
            const uint32_t d_tmem = tmem_base + g * BN;
            for (int kb = 0; kb < G2_KB; ++kb) {
                if (g == 0 && (kb == 0 || kb == 2)) {        // z K slabs 0-1 come from EPI1 half 0, 2-3 from half 1
                    mbar_wait(&zready[kb >> 1], 0);
                    tc_fence_after();
                }
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(zs + kb * A_BYTES);
                    const uint32_t b_addr = smem_u32(stages + stage * STAGE_BYTES) + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k)
                        umma_ss(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UK * 2)), make_sw128_kmajor_desc(b_addr + k * (UK * 2)),
                                idesc, (kb | k) != 0);
                    umma_commit(&empty[stage]);
                    if (kb == G2_KB - 1) umma_commit(&accb[2 + g]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 8 warps, warp e -> lane quarter e&3, chunks j with (j&1) == e>>2 =====================
        const int e = warp - 4, q = e & 3, sub = e >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        // ---- EPI1: thread = frame row; gate; z -> swizzled smem ----
        {
            const int row = q * 32 + lane, t = t0 + row;
            const bool valid = t < p.T;
            const uint16_t* crow = reinterpret_cast<const uint16_t*>(p.cond) + ((long long)b * p.T + t) * p.ldc;
            const uint32_t zrow = smem_u32(zs) + (row >> 3) * 1024 + (row & 7) * 128;
            const int sw = row & 7;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                // all of this warp's cond loads of the half are in flight BEFORE the wait on the accumulator
                uint4 c[4][4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        c[jj][i] = valid ? ldg_nc_u4(crow + h * BN + (2 * jj + sub) * 32 + 8 * i) : make_uint4(0, 0, 0, 0);
                mbar_wait(&accb[h], 0);
                tc_fence_after();
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 2 * jj + sub;
                    float acc[32];
                    tmem_ld32(taddr + h * BN + j * 32, acc);
                    tmem_ld_wait();
                    const uint32_t* cw = reinterpret_cast<const uint32_t*>(c[jj]);
                    uint32_t zp[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 ca = Half16<BF16>::unpack2(cw[2 * i]), cb = Half16<BF16>::unpack2(cw[2 * i + 1]);
                        const float z0 = sigmoid_fast(acc[4 * i] + ca.x) * tanh_fast(acc[4 * i + 1] + ca.y);
                        const float z1 = sigmoid_fast(acc[4 * i + 2] + cb.x) * tanh_fast(acc[4 * i + 3] + cb.y);
                        zp[i] = valid ? Half16<BF16>::pack2(z0, z1) : 0u;
                    }
                    // z channels [128h + 16j, +16) of this row: K slab 2h + j/4, 16-byte chunks 2(j%4), 2(j%4)+1
                    const uint32_t slab = zrow + (2 * h + (j >> 2)) * A_BYTES;
                    const int c16 = 2 * (j & 3);
                    st_shared_u4(slab + ((c16 ^ sw) << 4), make_uint4(zp[0], zp[1], zp[2], zp[3]));
                    st_shared_u4(slab + (((c16 + 1) ^ sw) << 4), make_uint4(zp[4], zp[5], zp[6], zp[7]));
                }
                fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core's async proxy
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&zready[h]);
            }
        }
        // ---- EPI2: coalesced layout through the warp's staging tile (aliases the A slots of the pipeline stages:
        //      the output-projection fills only write B slots, and every conv MMA has retired by accb[2]) ----
        float* stg = reinterpret_cast<float*>(stages + (e / 3) * STAGE_BYTES + (e % 3) * STG_WARP_BYTES);
        const int cl = (lane & 7) * 4, rsub = lane >> 3;
        const int tq = t0 + q * 32 + rsub;
        const float inv_sqrt2 = 0.70710678118654752440f;
        // software pipeline over the warp's 8 (g, j) chunks: the global loads of chunk n+1 are issued before chunk n
        // is processed
        float4 in[8], inn[8];
        auto load_inputs = [&](int g, int j, float4* dst) {
            const int col = j * 32 + cl;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int t = tq + 4 * i;
                if (t < p.T && (g == 0 || !p.first)) {
                    const float* src = (g == 0 ? p.x : p.skip) + ((long long)b * p.T + t) * C + col;
                    dst[i] = *reinterpret_cast<const float4*>(src);
                }
            }
        };
        load_inputs(0, sub, inn);
#pragma unroll 1
        for (int n = 0; n < 8; ++n) {
            const int g = n >> 2, j = 2 * (n & 3) + sub;
            if ((n & 3) == 0) {
                mbar_wait(&accb[2 + g], 0);
                tc_fence_after();
            }
            float acc[32];
            tmem_ld32(taddr + g * BN + j * 32, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) in[i] = inn[i];
            if (n + 1 < 8) load_inputs((n + 1) >> 2, 2 * ((n + 1) & 3) + sub, inn);
            const int col = j * 32 + cl;                                      // channel of x / skip
            const float4 bias = __ldg(reinterpret_cast<const float4*>(p.bo + g * C + col));
            float4 dsh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g == 0 && p.y_next && p.d_stride == 0) dsh = __ldg(reinterpret_cast<const float4*>(p.dvec + col));
            tmem_ld_wait();
            float4* srow = reinterpret_cast<float4*>(stg + lane * STG_LD);
#pragma unroll
            for (int c = 0; c < 8; ++c) srow[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int t = tq + 4 * i;
                if (t < p.T) {
                    const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i + rsub) * STG_LD + cl);
                    const long long r = (long long)b * p.T + t;
                    const float4 o = make_float4(v.x + bias.x, v.y + bias.y, v.z + bias.z, v.w + bias.w);
                    if (g == 0) {
                        const float4 xn = make_float4((in[i].x + o.x) * inv_sqrt2, (in[i].y + o.y) * inv_sqrt2,
                                                      (in[i].z + o.z) * inv_sqrt2, (in[i].w + o.w) * inv_sqrt2);
                        *reinterpret_cast<float4*>(p.x + r * C + col) = xn;
                        if (p.y_next) {
                            const float4 d = p.d_stride == 0
                                                 ? dsh
                                                 : __ldg(reinterpret_cast<const float4*>(p.dvec + (long long)b * p.d_stride + col));
                            uint2 yo;
                            yo.x = Half16<BF16>::pack2(xn.x + d.x, xn.y + d.y);
                            yo.y = Half16<BF16>::pack2(xn.z + d.z, xn.w + d.w);
                            *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.y_next) + r * C + col) = yo;
                        }
                    } else {
                        const float4 s = make_float4(o.x + in[i].x, o.y + in[i].y, o.z + in[i].z, o.w + in[i].w);
                        *reinterpret_cast<float4*>(p.skip + r * C + col) = s;
                        if (p.skip_h) {
                            uint2 so;
                            so.x = Half16<BF16>::pack2(s.x, s.y);
                            so.y = Half16<BF16>::pack2(s.z, s.w);
                            *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.skip_h) + r * C + col) = so;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int BF16>
static int launch_layer(const LayerP& p, int grid, cudaStream_t st) {
    static PerDevice configured;
    if (configured.first()) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(wavenet_layer_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
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
    B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg, wavenet_layer_kernel<BF16>, p));
    return B2S_OK;
}

#endif  // B2S_EXPERIMENTS

}  // namespace wl

// =====================================================================================================================
// Version 2: 2-CTA cluster, weight tiles multicast, A tiles shared by both N halves.
//
// Measured on B200 (profiles/): the L2 slice throughput cap (~6.3 KB / SM-clock chip-wide) bounds version 1, which pulls
// 1.4 MB of operands per 128-frame tile through L2.  Here
//   * one pipeline fill = one K slab for BOTH N halves: A [128 x 64] is fetched once (not twice), and
//   * the two CTAs of a cluster (two neighbouring time tiles) each fetch HALF of every weight tile (128 of its 256
//     rows) and TMA-multicast it into both CTAs' shared memory,
// so a CTA pulls 0.7 MB per tile.  Stage = A 16 KB + B(half 0) 32 KB + B(half 1) 32 KB, 2 stages (8 MMAs each).
// A stage is recycled only when BOTH CTAs' MMAs have consumed it: tcgen05.commit multicasts the arrive to the `empty`
// barrier (count 2) of both CTAs.
// =====================================================================================================================
namespace wl2 {

using wl::ldg_nc_u4;
using wl::pdl_launch_dependents;
using wl::pdl_wait;
using wl::st_shared_u4;
using wl::LayerP;

constexpr int C = 256;
constexpr int BM = 128, BK = 64, BN = 256, UK = 16, STAGES = 2, CLUSTER = 2;
constexpr int A_BYTES = BM * BK * 2;                     // 16 KB
constexpr int BH_BYTES = (BN / 2) * BK * 2;              // 16 KB: the 128 weight rows one CTA fetches
constexpr int B_BYTES = BN * BK * 2;                     // 32 KB
constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;       // 80 KB
constexpr int Z_BYTES = BM * C * 2;
constexpr int G1_KB = 3 * C / BK, G2_KB = C / BK;        // 12, 4
constexpr int STG_LD = 36;
constexpr int STG_WARP_BYTES = 32 * STG_LD * 4;
constexpr int SMEM_BYTES = Z_BYTES + STAGES * STAGE_BYTES + 256 + 1024;
constexpr int NTHREADS = 384;
constexpr int EPI_WARPS = 8;
constexpr uint16_t MASK = (1u << CLUSTER) - 1;

template <int BF16>
__global__ void __launch_bounds__(NTHREADS, 1) wavenet_layer_cl2_kernel(const __grid_constant__ LayerP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* zs = smem;
    uint8_t* stages = smem + Z_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* accb = empty + STAGES;                      // [3]: GEMM1 (both halves), GEMM2 residual, GEMM2 skip
    uint64_t* zready = accb + 3;                          // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(zready + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    // tiles_per_b is padded to an even count by the host; a padding tile has t0 >= T (all rows invalid, loads zero-filled)
    const int b = blockIdx.x / p.tiles_per_b, t0 = (blockIdx.x - b * p.tiles_per_b) * BM;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.mapY);
        prefetch_tmap(&p.mapWd);
        prefetch_tmap(&p.mapWo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], CLUSTER);                // one multicast commit from each CTA of the cluster
        }
        for (int i = 0; i < 3; ++i) mbar_init(&accb[i], 1);
        for (int i = 0; i < 2; ++i) mbar_init(&zready[i], EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // the peer's barriers exist before anything is multicast into this CTA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer: 12 conv fills (A + 2 x 2 weight half-tiles), 8 output-projection fills ====
        int stage = 0;
        uint32_t phase = 0;
        for (int f = 0; f < G1_KB + 2 * G2_KB; ++f) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0) {
                uint8_t* sa = stages + stage * STAGE_BYTES;
                uint8_t* sb = sa + A_BYTES;
                if (f < G1_KB) {
                    const int kb = f;
                    const int tap = kb / (C / BK), c0 = (kb - tap * (C / BK)) * BK;
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_3d(sa, &p.mapY, &full[stage], c0, t0 + (tap - 1) * p.dil, b);
                    // this CTA's 128 rows of both weight tiles, multicast to the whole cluster
                    tma_load_2d_mcast(sb + rank * BH_BYTES, &p.mapWd, &full[stage], kb * BK, rank * (BN / 2), MASK);
                    tma_load_2d_mcast(sb + B_BYTES + rank * BH_BYTES, &p.mapWd, &full[stage], kb * BK, BN + rank * (BN / 2), MASK);
                } else {
                    const int g = (f - G1_KB) / G2_KB, kb = (f - G1_KB) - g * G2_KB;
                    mbar_expect_tx(&full[stage], B_BYTES);
                    tma_load_2d_mcast(sb + rank * BH_BYTES, &p.mapWo, &full[stage], kb * BK, g * BN + rank * (BN / 2), MASK);
                }
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_f16(BM, BN, BF16);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < G1_KB; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(stages + stage * STAGE_BYTES), b_addr = a_addr + A_BYTES;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k)
                        umma_ss(tmem_base + h * BN, make_sw128_kmajor_desc(a_addr + k * (UK * 2)),
                                make_sw128_kmajor_desc(b_addr + h * B_BYTES + k * (UK * 2)), idesc, (kb | k) != 0);
                umma_commit_mcast(&empty[stage], MASK);
                if (kb == G1_KB - 1) umma_commit(&accb[0]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        for (int g = 0; g < 2; ++g) {
            const uint32_t d_tmem = tmem_base + g * BN;
            for (int kb = 0; kb < G2_KB; ++kb) {
                if (g == 0 && (kb == 0 || kb == 2)) {
                    mbar_wait(&zready[kb >> 1], 0);
                    tc_fence_after();
                }
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(zs + kb * A_BYTES);
                    const uint32_t b_addr = smem_u32(stages + stage * STAGE_BYTES) + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k)
                        umma_ss(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UK * 2)), make_sw128_kmajor_desc(b_addr + k * (UK * 2)),
                                idesc, (kb | k) != 0);
                    umma_commit_mcast(&empty[stage], MASK);
                    if (kb == G2_KB - 1) umma_commit(&accb[1 + g]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 8 warps, warp e -> lane quarter e&3, chunks j with (j&1) == e>>2 ===========
        const int e = warp - 4, q = e & 3, sub = e >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        {
            const int row = q * 32 + lane, t = t0 + row;
            const bool valid = t < p.T;
            const uint16_t* crow = reinterpret_cast<const uint16_t*>(p.cond) + ((long long)b * p.T + t) * p.ldc;
            const uint32_t zrow = smem_u32(zs) + (row >> 3) * 1024 + (row & 7) * 128;
            const int sw = row & 7;
            uint4 c[4][4];                                 // rolling cond prefetch: half 0 up front, half 1 as slots free up
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    c[jj][i] = valid ? ldg_nc_u4(crow + (2 * jj + sub) * 32 + 8 * i) : make_uint4(0, 0, 0, 0);
            mbar_wait(&accb[0], 0);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 2 * jj + sub;
                    float acc[32];
                    tmem_ld32(taddr + h * BN + j * 32, acc);
                    uint4 cc[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) cc[i] = c[jj][i];
                    if (h == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            c[jj][i] = valid ? ldg_nc_u4(crow + BN + j * 32 + 8 * i) : make_uint4(0, 0, 0, 0);
                    }
                    tmem_ld_wait();
                    const uint32_t* cw = reinterpret_cast<const uint32_t*>(cc);
                    uint32_t zp[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 ca = Half16<BF16>::unpack2(cw[2 * i]), cb = Half16<BF16>::unpack2(cw[2 * i + 1]);
                        const float z0 = sigmoid_fast(acc[4 * i] + ca.x) * tanh_fast(acc[4 * i + 1] + ca.y);
                        const float z1 = sigmoid_fast(acc[4 * i + 2] + cb.x) * tanh_fast(acc[4 * i + 3] + cb.y);
                        zp[i] = valid ? Half16<BF16>::pack2(z0, z1) : 0u;
                    }
                    const uint32_t slab = zrow + (2 * h + (j >> 2)) * A_BYTES;
                    const int c16 = 2 * (j & 3);
                    st_shared_u4(slab + ((c16 ^ sw) << 4), make_uint4(zp[0], zp[1], zp[2], zp[3]));
                    st_shared_u4(slab + (((c16 + 1) ^ sw) << 4), make_uint4(zp[4], zp[5], zp[6], zp[7]));
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&zready[h]);
            }
        }
        // ---- EPI2 (staging aliases pipeline memory the output-projection fills never write: the A slot of each stage
        //      and the second weight tile; every conv MMA has retired by accb[1]) ----
        const int si = e >> 1;                             // 0..3 within stage (e & 1)
        float* stg = reinterpret_cast<float*>(stages + (e & 1) * STAGE_BYTES +
                                              (si < 3 ? si * STG_WARP_BYTES : A_BYTES + B_BYTES));
        const int cl = (lane & 7) * 4, rsub = lane >> 3;
        const int tq = t0 + q * 32 + rsub;
        const float inv_sqrt2 = 0.70710678118654752440f;
        float4 in[8], inn[8];
        auto load_inputs = [&](int g, int j, float4* dst) {
            const int col = j * 32 + cl;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int t = tq + 4 * i;
                if (t < p.T && (g == 0 || !p.first)) {
                    const float* src = (g == 0 ? p.x : p.skip) + ((long long)b * p.T + t) * C + col;
                    dst[i] = *reinterpret_cast<const float4*>(src);
                }
            }
        };
        load_inputs(0, sub, inn);
#pragma unroll 1
        for (int n = 0; n < 8; ++n) {
            const int g = n >> 2, j = 2 * (n & 3) + sub;
            if ((n & 3) == 0) {
                mbar_wait(&accb[1 + g], 0);
                tc_fence_after();
            }
            float acc[32];
            tmem_ld32(taddr + g * BN + j * 32, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) in[i] = inn[i];
            if (n + 1 < 8) load_inputs((n + 1) >> 2, 2 * ((n + 1) & 3) + sub, inn);
            const int col = j * 32 + cl;
            const float4 bias = __ldg(reinterpret_cast<const float4*>(p.bo + g * C + col));
            float4 dsh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g == 0 && p.y_next && p.d_stride == 0) dsh = __ldg(reinterpret_cast<const float4*>(p.dvec + col));
            tmem_ld_wait();
            float4* srow = reinterpret_cast<float4*>(stg + lane * STG_LD);
#pragma unroll
            for (int c2 = 0; c2 < 8; ++c2) srow[c2] = make_float4(acc[4 * c2], acc[4 * c2 + 1], acc[4 * c2 + 2], acc[4 * c2 + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int t = tq + 4 * i;
                if (t < p.T) {
                    const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i + rsub) * STG_LD + cl);
                    const long long r = (long long)b * p.T + t;
                    const float4 o = make_float4(v.x + bias.x, v.y + bias.y, v.z + bias.z, v.w + bias.w);
                    if (g == 0) {
                        const float4 xn = make_float4((in[i].x + o.x) * inv_sqrt2, (in[i].y + o.y) * inv_sqrt2,
                                                      (in[i].z + o.z) * inv_sqrt2, (in[i].w + o.w) * inv_sqrt2);
                        *reinterpret_cast<float4*>(p.x + r * C + col) = xn;
                        if (p.y_next) {
                            const float4 d = p.d_stride == 0
                                                 ? dsh
                                                 : __ldg(reinterpret_cast<const float4*>(p.dvec + (long long)b * p.d_stride + col));
                            uint2 yo;
                            yo.x = Half16<BF16>::pack2(xn.x + d.x, xn.y + d.y);
                            yo.y = Half16<BF16>::pack2(xn.z + d.z, xn.w + d.w);
                            *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.y_next) + r * C + col) = yo;
                        }
                    } else {
                        const float4 s2 = make_float4(o.x + in[i].x, o.y + in[i].y, o.z + in[i].z, o.w + in[i].w);
                        *reinterpret_cast<float4*>(p.skip + r * C + col) = s2;
                        if (p.skip_h) {
                            uint2 so;
                            so.x = Half16<BF16>::pack2(s2.x, s2.y);
                            so.y = Half16<BF16>::pack2(s2.z, s2.w);
                            *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.skip_h) + r * C + col) = so;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // the peer may still multicast into / arrive on this CTA's shared memory until here
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int BF16>
static int launch_layer(const LayerP& p, int grid, cudaStream_t st) {
    static PerDevice configured;
    if (configured.first()) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(wavenet_layer_cl2_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg, wavenet_layer_cl2_kernel<BF16>, p));
    return B2S_OK;
}

}  // namespace wl2

// =====================================================================================================================
// Version 3: the WHOLE residual stack (all L layers of wavenet.py:92-94) as ONE persistent kernel.
//
// Grid = one CTA per 128-frame tile (all co-resident: grid <= number of SMs, the host splits larger batches by
// utterance), 2-CTA clusters with multicast weight tiles (as version 2), and the sequential-N-half structure of version 1
// so that phases overlap:
//     MMA       G1h0(l) | G1h1(l)    |            | G2res(l) | G2skip(l) G1h0(l+1) | G1h1(l+1) ...
//     epilogue          | EPI1h0(l)  | EPI1h1(l)  |          | EPI2res(l)          | EPI2skip(l) ...
// Layer l+1 of a tile needs y_{l+1} of its two neighbour tiles (dilation halo).  That hand-off goes through global
// memory with a release/acquire flag per tile (value = number of finished y buffers) instead of a kernel boundary:
//     writer  y stores -> bar.sync(epilogue warps) -> one thread: __threadfence, st.release.gpu flag[tile] = l+1
//     reader  (TMA producer) ld.acquire.gpu spin on flag[tile-1], flag[tile], flag[tile+1] -> fence.proxy.async -> TMA
// Weight tiles of layer l+1 are requested BEFORE the flag wait, so their latency hides behind the hand-off.
// TMEM columns are recycled through two more mbarriers (tfree[h]: EPI2 of layer l has drained half h).
// =====================================================================================================================
namespace ws {

using wl::ldg_nc_u4;
using wl::pdl_launch_dependents;
using wl::pdl_wait;
using wl::st_shared_u4;
using wl::st_shared_f4;
using wl::ld_shared_f4;

constexpr int C = 256, MAXL = 32;
constexpr int BM = 128, BK = 64, BN = 256, UK = 16, STAGES = 3, CLUSTER = 2;
constexpr int A_BYTES = BM * BK * 2, BH_BYTES = (BN / 2) * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int Z_BYTES = BM * C * 2;
constexpr int G1_KB = 3 * C / BK, G2_KB = C / BK;
constexpr int FILLS = 2 * G1_KB + 2 * G2_KB;             // 32 per layer
constexpr int STG_LD = 36, STG_ROWS = 16;
constexpr int STG_WARP_BYTES = STG_ROWS * STG_LD * 4;    // 2304: 16-row transpose buffer (two passes per 32-row chunk)
constexpr int NTHREADS = 384, EPI_WARPS = 8;
constexpr int SMEM_BYTES = Z_BYTES + STAGES * STAGE_BYTES + EPI_WARPS * STG_WARP_BYTES + 256;   // 231680
constexpr uint16_t MASK = (1u << CLUSTER) - 1;

struct __align__(64) StackP {
    CUtensorMap mapY[2], mapWd, mapWo;
    int B, T, tiles_per_b, L;
    int dil[MAXL];
    const void* cond; int ldc; long long cond_lstride;    // layer l's slab at cond + l * cond_lstride (elements)
    const float* bo;                                      // [L][2C]
    float* x; void* ybuf[2]; float* skip; void* skip_h;
    const float* dvec; int d_stride;                      // layer l's step embedding at dvec + b*d_stride + l*C
    int* flags;                                           // [B * tiles_per_b], zero before the launch
    int dbg;                                              // (profiling switches of earlier rounds; no longer read by this kernel)
    unsigned long long* tlog;                             // optional phase timestamps of CTA 2 (B2S_STACK_TLOG): [L][16] globaltimer ns
    // ---- whole denoiser in one launch (fuse = 1): stem (input projection, wavenet.py:86-88) before the stack and head
    //      (skip sum -> skip_projection -> ReLU -> output_projection, wavenet.py:96-99) after it
    int fuse, MF, kb_in;                                  // MF = in_dims * n_feats; kb_in = K slabs of the stem GEMM
    CUtensorMap mapXin, mapWin, mapWsp, mapWfin;
    const float* b_in; const float* b_sp; const float* b_fin;
    float alpha_head;                                     // 1 / sqrt(L)
    float* out;                                           // [rows, MF] fp32
    // ---- sampler update inside the head epilogue (upd_n > 0): x' = sum_i coef[i] * src_i, src_i = upd_src[i] or, where
    //      upd_src[i] is NULL, this evaluation's output; x' -> upd_x (fp32, may alias a source) and upd_xh (16-bit: the
    //      next evaluation's input).  `out` is not written then.
    int upd_n;
    const float* upd_src[3];
    const float* upd_coef;
    float* upd_x;
    void* upd_xh;
    int* flags_next;                                      // the NEXT launch's tile flags: every CTA zeroes its own entry
};

// phase timestamps of CTA 2 for scripts/stack_timeline.py; compiled in only with -DB2S_TLOG (B2S_BUILD_TLOG=1 python _build.py)
#ifdef B2S_TLOG
#define TLOG(slot) do { if (p.tlog && blockIdx.x == 2 && lane == 0) p.tlog[l * 16 + (slot)] = globaltimer_ns(); } while (0)
#else
#define TLOG(slot) do { } while (0)
#endif

__device__ __forceinline__ void wait_flag(const int* f, int want) {
    if (ld_acquire_gpu(f) >= want) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spin = 0;
    while (ld_acquire_gpu(f) < want) {
        if ((++spin & 255u) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
            printf("b2s: tile flag timeout (block %d want %d have %d)\n", blockIdx.x, want, ld_acquire_gpu(f));
            __trap();
        }
    }
}

template <int BF16>
__global__ void __launch_bounds__(NTHREADS, 1) wavenet_stack_kernel(const __grid_constant__ StackP p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* zs = smem;
    uint8_t* stages = smem + Z_BYTES;
    float* stg_all = reinterpret_cast<float*>(stages + STAGES * STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + STAGES * STAGE_BYTES + EPI_WARPS * STG_WARP_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* accb = empty + STAGES;                      // [4]: G1 half 0, G1 half 1, G2 residual, G2 skip
    uint64_t* zready = accb + 4;                          // [2]
    uint64_t* tfree = zready + 2;                         // [2]: EPI2 has drained TMEM half h
    uint64_t* stemb = tfree + 2;                          // stem accumulator complete
    uint64_t* stemfree = stemb + 1;                       // stem epilogue has drained TMEM columns [0,256)
    uint64_t* hz = stemfree + 1;                          // [2]: head operand tiles (skip sum, hidden) are in the z buffer
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(hz + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int b = blockIdx.x / p.tiles_per_b, ti = blockIdx.x - b * p.tiles_per_b, t0 = ti * BM;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("b2s: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.mapY[0]);
        prefetch_tmap(&p.mapY[1]);
        prefetch_tmap(&p.mapWd);
        prefetch_tmap(&p.mapWo);
        if (p.fuse) {
            prefetch_tmap(&p.mapXin);
            prefetch_tmap(&p.mapWin);
            prefetch_tmap(&p.mapWsp);
            prefetch_tmap(&p.mapWfin);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], CLUSTER);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&accb[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&zready[i], EPI_WARPS);
            mbar_init(&tfree[i], EPI_WARPS);
            mbar_init(&hz[i], EPI_WARPS);
        }
        mbar_init(stemb, 1);
        mbar_init(stemfree, EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // The dependent launch is released LATE (when the epilogue warps reach the head, see below): released here, the next
    // denoiser launch of a fused-update step would park its CTAs on the SMs this grid leaves idle and starve the side-stream
    // noise kernels that are supposed to run there meanwhile (measured: 2 % slower).
    pdl_wait();
    if (threadIdx.x == 64 && p.flags_next) p.flags_next[blockIdx.x] = 0;     // re-arm the next launch (it starts after this grid ends)

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        if (p.fuse) {                                       // stem: A = x_in tile, B = input_projection weights
            for (int kb = 0; kb < p.kb_in; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sa = stages + stage * STAGE_BYTES;
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d_mcast(sa + A_BYTES + rank * BH_BYTES, &p.mapWin, &full[stage], kb * BK, rank * (BN / 2), MASK);
                    tma_load_3d(sa, &p.mapXin, &full[stage], kb * BK, t0, b);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        for (int l = 0; l < p.L; ++l) {
            const int dil = p.dil[l];
            if (lane == 0 && l + 1 < p.L && t0 < p.T) {
                // the next layer's 128 KB slab of the hoisted conditioner projection -> L2 (it streams from HBM once per evaluation)
                const uint8_t* nxt = reinterpret_cast<const uint8_t*>(p.cond) + ((l + 1) * p.cond_lstride + (long long)blockIdx.x * (BM * 2 * C)) * 2;
                for (int i = 0; i < 8; ++i)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt + i * 16384), "r"(16384) : "memory");
            }
            for (int f = 0; f < FILLS; ++f) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sa = stages + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    if (f < 2 * G1_KB) {
                        const int h = f / G1_KB, kb = f - h * G1_KB;
                        const int tap = kb / (C / BK), c0 = (kb - tap * (C / BK)) * BK;
                        mbar_expect_tx(&full[stage], STAGE_BYTES);
                        tma_load_3d_mcast(sb + rank * BH_BYTES, &p.mapWd, &full[stage], kb * BK, h * BN + rank * (BN / 2), l, MASK);
                        if (f == 0 && l + p.fuse > 0 && t0 < p.T) {
                            // y_l of this tile and of its neighbours (dilation halo) must be complete
                            const int* fl = p.flags + b * p.tiles_per_b;
                            const int want = l + p.fuse;         // number of finished y buffers (the stem's y_0 counts when fused)
                            if (ti > 0) wait_flag(fl + ti - 1, want);
                            wait_flag(fl + ti, want);
                            if ((ti + 1) * BM < p.T) wait_flag(fl + ti + 1, want);
                            fence_proxy_async_all();
                        }
                        if (f == 0) TLOG(0);                               // neighbours' y ready, first A load issued
                        tma_load_3d(sa, &p.mapY[l & 1], &full[stage], c0, t0 + (tap - 1) * dil, b);
                    } else {
                        const int g = (f - 2 * G1_KB) / G2_KB, kb = (f - 2 * G1_KB) - g * G2_KB;
                        mbar_expect_tx(&full[stage], B_BYTES);
                        tma_load_3d_mcast(sb + rank * BH_BYTES, &p.mapWo, &full[stage], kb * BK, g * BN + rank * (BN / 2), l, MASK);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        if (p.fuse) {                                       // head: skip_projection then output_projection weights (B only)
            for (int i = 0; i < 2 * G2_KB; ++i) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sb = stages + stage * STAGE_BYTES + A_BYTES;
                    mbar_expect_tx(&full[stage], B_BYTES);
                    tma_load_2d_mcast(sb + rank * BH_BYTES, i < G2_KB ? &p.mapWsp : &p.mapWfin, &full[stage], (i % G2_KB) * BK,
                                      rank * (BN / 2), MASK);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_f16(BM, BN, BF16);
        int stage = 0;
        uint32_t phase = 0;
        auto mma_block = [&](uint32_t d_tmem, uint32_t a_addr, int kb, uint64_t* done) {
            // one K slab (4 MMAs of K = 16) with the B operand of the current stage; frees the stage, optionally signals `done`
            if (lane == 0) {
                const uint32_t b_addr = smem_u32(stages + stage * STAGE_BYTES) + A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UK; ++k)
                    umma_ss(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UK * 2)), make_sw128_kmajor_desc(b_addr + k * (UK * 2)), idesc,
                            (kb | k) != 0);
                umma_commit_mcast(&empty[stage], MASK);
                if (done) umma_commit(done);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };
        if (p.fuse) {                                       // stem GEMM -> TMEM columns [0,256)
            for (int kb = 0; kb < p.kb_in; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                mma_block(tmem_base, smem_u32(stages + stage * STAGE_BYTES), kb, kb == p.kb_in - 1 ? stemb : nullptr);
            }
        }
        for (int l = 0; l < p.L; ++l) {
            const uint32_t par = l & 1;
            for (int h = 0; h < 2; ++h) {
                if (l > 0) {                                   // EPI2 of the previous layer has drained these columns
                    mbar_wait(&tfree[h], par ^ 1);
                    tc_fence_after();
                } else if (p.fuse && h == 0) {                 // the stem epilogue has drained columns [0,256)
                    mbar_wait(stemfree, 0);
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + h * BN;
                for (int kb = 0; kb < G1_KB; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (kb == 0) TLOG(1 + h);                              // first K slab of G1 half h has landed
                    if (lane == 0) {
                        const uint32_t a_addr = smem_u32(stages + stage * STAGE_BYTES), b_addr = a_addr + A_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma_ss(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UK * 2)), make_sw128_kmajor_desc(b_addr + k * (UK * 2)),
                                    idesc, (kb | k) != 0);
                        umma_commit_mcast(&empty[stage], MASK);
                        if (kb == G1_KB - 1) umma_commit(&accb[h]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            for (int g = 0; g < 2; ++g) {
                const uint32_t d_tmem = tmem_base + g * BN;
                for (int kb = 0; kb < G2_KB; ++kb) {
                    if (g == 0 && (kb == 0 || kb == 2)) {
                        mbar_wait(&zready[kb >> 1], par);
                        tc_fence_after();
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_addr = smem_u32(zs + kb * A_BYTES);
                        const uint32_t b_addr = smem_u32(stages + stage * STAGE_BYTES) + A_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma_ss(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UK * 2)), make_sw128_kmajor_desc(b_addr + k * (UK * 2)),
                                    idesc, (kb | k) != 0);
                        umma_commit_mcast(&empty[stage], MASK);
                        if (kb == G2_KB - 1) umma_commit(&accb[2 + g]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        if (p.fuse) {
            // head GEMMs, A operand = the z buffer: skip sum (written by the last layer's EPI2-skip) then the hidden tile
            const uint32_t plast = (p.L - 1) & 1, pl = p.L & 1;
            (void)pl;
            for (int g = 0; g < 2; ++g) {
                mbar_wait(&tfree[g], plast);                // the last layer's EPI2 has drained this TMEM half
                mbar_wait(&hz[g], 0);                       // and the A tile is in shared memory
                tc_fence_after();
                for (int kb = 0; kb < G2_KB; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    mma_block(tmem_base + g * BN, smem_u32(zs + kb * A_BYTES), kb, kb == G2_KB - 1 ? &accb[g] : nullptr);
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 8 warps, warp e -> lane quarter e&3, chunks j with (j&1) == e>>2 ===========
        const int e = warp - 4, q = e & 3, sub = e >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int row = q * 32 + lane;
        const bool valid = t0 + row < p.T;
        const uint32_t zrow = smem_u32(zs) + (row >> 3) * 1024 + (row & 7) * 128;
        const int sw = row & 7;
        float* stg = stg_all + e * (STG_ROWS * STG_LD);
        const int cl = (lane & 7) * 4, rsub = lane >> 3;
        const int tq = t0 + q * 32 + rsub;
        const float inv_sqrt2 = 0.70710678118654752440f;
        // per-lane constants of the coalesced (EPI2) layout
        uint32_t vmask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) vmask |= (tq + 4 * i < p.T ? 1u : 0u) << i;
        const long long rowoff = ((long long)b * p.T + tq) * C + cl;          // element offset of (row 0 of the lane, its 4 columns)
        float* xrow = p.x + rowoff;
        float* srow_g = p.skip + rowoff;
        uint16_t* shrow = reinterpret_cast<uint16_t*>(p.skip_h) + rowoff;

        // z-buffer address of this lane's 4 channels (coalesced layout) for row 4i + rsub of the warp: used by the fused head
        auto z_addr_quad = [&](int i, int col) -> uint32_t {
            const int rt = q * 32 + 4 * i + rsub;
            return smem_u32(zs) + (col >> 6) * A_BYTES + (rt >> 3) * 1024 + (rt & 7) * 128 + ((((col & 63) >> 3) ^ (rt & 7)) << 4) + (col & 4) * 2;
        };
        if (p.fuse) {
            // ---- stem epilogue: x = relu(acc + b_in) (fp32), y_0 = x + d_0 (16-bit), then the tile flag ----
            const float* d0 = p.dvec + (long long)b * p.d_stride;
            uint16_t* y0 = reinterpret_cast<uint16_t*>(p.ybuf[0]) + rowoff;
            mbar_wait(stemb, 0);
            tc_fence_after();
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {
                const int j = 2 * jj + sub, col = j * 32 + cl;
                float acc[32];
                tmem_ld32(taddr + j * 32, acc);
                const float4 bias = __ldg(reinterpret_cast<const float4*>(p.b_in + col));
                const float4 d = __ldg(reinterpret_cast<const float4*>(d0 + col));
                tmem_ld_wait();
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if ((lane >> 4) == pass) {
                        float4* srow = reinterpret_cast<float4*>(stg + (lane & 15) * STG_LD);
#pragma unroll
                        for (int c2 = 0; c2 < 8; ++c2)
                            srow[c2] = make_float4(acc[4 * c2], acc[4 * c2 + 1], acc[4 * c2 + 2], acc[4 * c2 + 3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2) {
                        const int i = 4 * pass + i2;
                        if (vmask >> i & 1) {
                            const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i2 + rsub) * STG_LD + cl);
                            const float4 xv = make_float4(fmaxf(v.x + bias.x, 0.f), fmaxf(v.y + bias.y, 0.f), fmaxf(v.z + bias.z, 0.f),
                                                          fmaxf(v.w + bias.w, 0.f));
                            *reinterpret_cast<float4*>(xrow + j * 32 + i * 4 * C) = xv;
                            uint2 yv;
                            yv.x = Half16<BF16>::pack2(xv.x + d.x, xv.y + d.y);
                            yv.y = Half16<BF16>::pack2(xv.z + d.z, xv.w + d.w);
                            *reinterpret_cast<uint2*>(y0 + j * 32 + i * 4 * C) = yv;
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(stemfree);
            named_bar_sync(1, EPI_WARPS * 32);
            if (e == 0 && lane == 0) {
                __threadfence();
                st_release_gpu(p.flags + blockIdx.x, 1);
            }
        }

#pragma unroll 1
        for (int l = 0; l < p.L; ++l) {
            const uint32_t par = l & 1;
            const bool first = l == 0, last = l == p.L - 1;
            // ---- EPI1: thread = frame row; + cond; gate; z -> swizzled smem ----
            // tile/chunk-major table: (chunk j, 16-byte piece i) of this tile = 128 rows x 16 B contiguous
            const uint16_t* ctile = reinterpret_cast<const uint16_t*>(p.cond) + l * p.cond_lstride + (long long)blockIdx.x * (BM * 2 * C) + row * 8;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint4 c[4][4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        c[jj][i] = valid ? ldg_nc_u4(ctile + ((8 * h + 2 * jj + sub) * 4 + i) * (BM * 8)) : make_uint4(0, 0, 0, 0);
                mbar_wait(&accb[h], par);
                tc_fence_after();
                if (e == 0) TLOG(3 + 2 * h);                              // G1 half h complete
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 2 * jj + sub;
                    float acc[32];
                    tmem_ld32(taddr + h * BN + j * 32, acc);
                    tmem_ld_wait();
                    const uint32_t* cw = reinterpret_cast<const uint32_t*>(c[jj]);
                    uint32_t zp[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 ca = Half16<BF16>::unpack2(cw[2 * i]), cb = Half16<BF16>::unpack2(cw[2 * i + 1]);
                        const float z0 = sigmoid_fast(acc[4 * i] + ca.x) * tanh_fast(acc[4 * i + 1] + ca.y);
                        const float z1 = sigmoid_fast(acc[4 * i + 2] + cb.x) * tanh_fast(acc[4 * i + 3] + cb.y);
                        zp[i] = valid ? Half16<BF16>::pack2(z0, z1) : 0u;
                    }
                    const uint32_t slab = zrow + (2 * h + (j >> 2)) * A_BYTES;
                    const int c16 = 2 * (j & 3);
                    st_shared_u4(slab + ((c16 ^ sw) << 4), make_uint4(zp[0], zp[1], zp[2], zp[3]));
                    st_shared_u4(slab + (((c16 + 1) ^ sw) << 4), make_uint4(zp[4], zp[5], zp[6], zp[7]));
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&zready[h]);
                if (e == 0) TLOG(4 + 2 * h);                              // EPI1 half h done (this warp)
            }
            // ---- EPI2: coalesced layout through the warp's 16-row staging tile (two passes per 32-row chunk).
            //      Row i of the lane is frame tq + 4i: all addresses are (lane base) + (compile-time i) * 4C + (chunk) * 32,
            //      validity is a precomputed bit mask. ----
            const float* bo = p.bo + (long long)l * 2 * C;
            const float* dnext = p.dvec + (long long)(l + 1) * C + (long long)b * p.d_stride;
            uint16_t* ynext = last ? nullptr : reinterpret_cast<uint16_t*>(p.ybuf[(l + 1) & 1]) + rowoff;
            // Eight chunk bodies (4 residual, 4 skip), each specialised at compile time on the half it drains and fully
            // unrolled: the loop used to carry both halves, the profiling switches and a register copy of the prefetched inputs
            // in ONE body of ~800 SASS instructions per 32x32 chunk, and two warps per scheduler made EPI2 instruction-issue bound.
            // Inputs of chunk n+1 (x or skip rows, bias and step-embedding quads: L1 is ~0 KB here, every load is an L2 round
            // trip) are requested while chunk n's TMEM load is in flight, alternating between two register buffers.
            float4 bufA[8], bufB[8];
            float4 cbA, cdA, cbB, cdB;
            // chunk j = 2 * JJ + sub: `sub` is folded into the base pointers so that every address below is base + immediate
            const float* xs_ = xrow + sub * 32;
            const float* ss_ = srow_g + sub * 32;
            uint16_t* ys_ = ynext ? ynext + sub * 32 : nullptr;
            uint16_t* shs_ = shrow + sub * 32;
            const float* bos_ = bo + sub * 32 + cl;
            const float* dns_ = dnext + sub * 32 + cl;
            const uint32_t tsub = taddr + sub * 32;
            const uint32_t stg_w = smem_u32(stg) + (lane & 15) * (STG_LD * 4);          // this lane's staging row (its pass)
            const uint32_t stg_r = smem_u32(stg) + (rsub * STG_LD + cl) * 4;             // coalesced read-back: + i2 * 4 rows
            auto load_inputs = [&](auto gtag, auto jtag, float4* dst, float4& cb, float4& cd) {
                constexpr int g = decltype(gtag)::value, JJ = decltype(jtag)::value;
                const float* src = (g == 0 ? xs_ : ss_) + JJ * 64;
                const bool rd = (g == 0 || !first);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = (rd && (vmask >> i & 1)) ? *reinterpret_cast<const float4*>(src + i * 4 * C) : make_float4(0.f, 0.f, 0.f, 0.f);
                cb = __ldg(reinterpret_cast<const float4*>(bos_ + g * C + JJ * 64));
                cd = (g == 0 && ynext) ? __ldg(reinterpret_cast<const float4*>(dns_ + JJ * 64)) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto chunk = [&](auto gtag, auto jtag, const float4* in, const float4 bias, const float4 d, auto&& prefetch_next) {
                constexpr int g = decltype(gtag)::value, JJ = decltype(jtag)::value;
                float acc[32];
                tmem_ld32(tsub + g * BN + JJ * 64, acc);
                prefetch_next();
                float* xo = const_cast<float*>(xs_) + JJ * 64;
                float* so = const_cast<float*>(ss_) + JJ * 64;
                uint16_t* yo = ys_ + JJ * 64;
                uint16_t* sho = shs_ + JJ * 64;
                tmem_ld_wait();
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if ((lane >> 4) == pass) {
#pragma unroll
                        for (int c2 = 0; c2 < 8; ++c2)
                            st_shared_f4(stg_w + c2 * 16, acc[4 * c2], acc[4 * c2 + 1], acc[4 * c2 + 2], acc[4 * c2 + 3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2) {
                        const int i = 4 * pass + i2;                        // row 4*i + rsub of the warp's 32
                        if (vmask >> i & 1) {
                            const float4 v = ld_shared_f4(stg_r + i2 * (4 * STG_LD * 4));
                            const float4 o = make_float4(v.x + bias.x, v.y + bias.y, v.z + bias.z, v.w + bias.w);
                            if (g == 0) {
                                const float4 xn = make_float4((in[i].x + o.x) * inv_sqrt2, (in[i].y + o.y) * inv_sqrt2,
                                                              (in[i].z + o.z) * inv_sqrt2, (in[i].w + o.w) * inv_sqrt2);
                                *reinterpret_cast<float4*>(xo + i * 4 * C) = xn;
                                if (ynext) {
                                    uint2 yv;
                                    yv.x = Half16<BF16>::pack2(xn.x + d.x, xn.y + d.y);
                                    yv.y = Half16<BF16>::pack2(xn.z + d.z, xn.w + d.w);
                                    *reinterpret_cast<uint2*>(yo + i * 4 * C) = yv;
                                }
                            } else {
                                const float4 s2 = make_float4(o.x + in[i].x, o.y + in[i].y, o.z + in[i].z, o.w + in[i].w);
                                *reinterpret_cast<float4*>(so + i * 4 * C) = s2;
                                if (last && (p.skip_h || p.fuse)) {
                                    uint2 sv;
                                    sv.x = Half16<BF16>::pack2(s2.x, s2.y);
                                    sv.y = Half16<BF16>::pack2(s2.z, s2.w);
                                    if (p.fuse) {       // the head GEMM's A operand: straight into the (now idle) z buffer
                                        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(z_addr_quad(i, (2 * JJ + sub) * 32 + cl)), "r"(sv.x), "r"(sv.y) : "memory");
                                    } else {
                                        *reinterpret_cast<uint2*>(sho + i * 4 * C) = sv;
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            };
            using G0 = std::integral_constant<int, 0>;
            using G1 = std::integral_constant<int, 1>;
            using J0 = std::integral_constant<int, 0>;
            using J1 = std::integral_constant<int, 1>;
            using J2 = std::integral_constant<int, 2>;
            using J3 = std::integral_constant<int, 3>;
            load_inputs(G0{}, J0{}, bufA, cbA, cdA);
            // ---- residual half ----
            mbar_wait(&accb[2], par);
            tc_fence_after();
            if (e == 0) TLOG(7);                                        // G2 residual half complete
            chunk(G0{}, J0{}, bufA, cbA, cdA, [&] { load_inputs(G0{}, J1{}, bufB, cbB, cdB); });
            if (e == 0) TLOG(11);
            chunk(G0{}, J1{}, bufB, cbB, cdB, [&] { load_inputs(G0{}, J2{}, bufA, cbA, cdA); });
            if (e == 0) TLOG(12);
            chunk(G0{}, J2{}, bufA, cbA, cdA, [&] { load_inputs(G0{}, J3{}, bufB, cbB, cdB); });
            if (e == 0) TLOG(13);
            chunk(G0{}, J3{}, bufB, cbB, cdB, [&] { load_inputs(G1{}, J0{}, bufA, cbA, cdA); });
            if (e == 0) TLOG(14);
            if (e == 0) TLOG(15);                                       // this warp's last residual chunk stored
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tfree[0]);
            // hand-off: every epilogue thread has issued its y_next stores -> CTA barrier -> ONE thread publishes the tile flag.
            // st.release.gpu is itself a GPU-scope release fence, cumulative over the barrier: no separate __threadfence().
            named_bar_sync(1, EPI_WARPS * 32);
            if (e == 0 && lane == 0) st_release_gpu(p.flags + blockIdx.x, l + 1 + p.fuse);
            if (e == 0) TLOG(8);                                        // EPI2 residual half done, flag released
            // ---- skip half (overlaps GEMM1 of the next layer) ----
            mbar_wait(&accb[3], par);
            tc_fence_after();
            if (e == 0) TLOG(9);                                        // G2 skip half complete
            chunk(G1{}, J0{}, bufA, cbA, cdA, [&] { load_inputs(G1{}, J1{}, bufB, cbB, cdB); });
            chunk(G1{}, J1{}, bufB, cbB, cdB, [&] { load_inputs(G1{}, J2{}, bufA, cbA, cdA); });
            chunk(G1{}, J2{}, bufA, cbA, cdA, [&] { load_inputs(G1{}, J3{}, bufB, cbB, cdB); });
            chunk(G1{}, J3{}, bufB, cbB, cdB, [] {});
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tfree[1]);
            if (last && p.fuse) {
                fence_proxy_async_smem();                   // skip-sum tile (generic-proxy writes) -> tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&hz[0]);
            }
            if (e == 0) TLOG(10);                                       // EPI2 skip half done
        }
        pdl_launch_dependents();
        if (p.fuse) {
            const uint32_t pl = p.L & 1;
            // ---- EPI3 (thread = frame row): hidden = relu(alpha * acc + b_sp) -> 16-bit, back into the z buffer ----
            mbar_wait(&accb[0], pl);
            tc_fence_after();
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {
                const int j = 2 * jj + sub;
                float acc[32];
                tmem_ld32(taddr + j * 32, acc);
                float4 bs[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) bs[i] = __ldg(reinterpret_cast<const float4*>(p.b_sp + j * 32 + 4 * i));
                tmem_ld_wait();
                uint32_t hp[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float h0 = fmaxf(fmaf(p.alpha_head, acc[4 * i], bs[i].x), 0.f), h1 = fmaxf(fmaf(p.alpha_head, acc[4 * i + 1], bs[i].y), 0.f);
                    const float h2 = fmaxf(fmaf(p.alpha_head, acc[4 * i + 2], bs[i].z), 0.f), h3 = fmaxf(fmaf(p.alpha_head, acc[4 * i + 3], bs[i].w), 0.f);
                    hp[2 * i] = valid ? Half16<BF16>::pack2(h0, h1) : 0u;
                    hp[2 * i + 1] = valid ? Half16<BF16>::pack2(h2, h3) : 0u;
                }
                // 32 hidden channels [32j, +32) of this row: K slab j/2, 16-byte chunks 4(j%2) .. 4(j%2)+3
                const uint32_t slab = zrow + (j >> 1) * A_BYTES;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4)
                    st_shared_u4(slab + (((4 * (j & 1) + c4) ^ sw) << 4), make_uint4(hp[4 * c4], hp[4 * c4 + 1], hp[4 * c4 + 2], hp[4 * c4 + 3]));
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hz[1]);
            // ---- EPI4 (coalesced layout): out = acc + b_fin -> fp32 [rows, MF] ----
            mbar_wait(&accb[1], pl);
            tc_fence_after();
            const long long orow_off = ((long long)b * p.T + tq) * p.MF + cl;
            float uc[3] = {0.f, 0.f, 0.f};
            for (int n = 0; n < p.upd_n; ++n) uc[n] = __ldg(p.upd_coef + n);
            // the (at most two) buffer operands of the update: all 8 rows of a chunk are requested BEFORE the first store - the
            // destination aliases a source, so loads issued after a store would be serialised behind it
            const float* upA = nullptr;
            const float* upB = nullptr;
            for (int n = 0; n < p.upd_n; ++n)
                if (p.upd_src[n]) { if (!upA) upA = p.upd_src[n]; else upB = p.upd_src[n]; }
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {
                const int j = 2 * jj + sub, col = j * 32 + cl;
                if (j * 32 >= p.MF) break;
                float acc[32];
                tmem_ld32(taddr + BN + j * 32, acc);
                const float4 bias = col < p.MF ? __ldg(reinterpret_cast<const float4*>(p.b_fin + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 preA[8], preB[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long long off = orow_off + j * 32 + (long long)i * 4 * p.MF;
                    const bool okr = (vmask >> i & 1) && col < p.MF;
                    preA[i] = (upA && okr) ? *reinterpret_cast<const float4*>(upA + off) : make_float4(0.f, 0.f, 0.f, 0.f);
                    preB[i] = (upB && okr) ? *reinterpret_cast<const float4*>(upB + off) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                tmem_ld_wait();
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if ((lane >> 4) == pass) {
                        float4* srow = reinterpret_cast<float4*>(stg + (lane & 15) * STG_LD);
#pragma unroll
                        for (int c2 = 0; c2 < 8; ++c2)
                            srow[c2] = make_float4(acc[4 * c2], acc[4 * c2 + 1], acc[4 * c2 + 2], acc[4 * c2 + 3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2) {
                        const int i = 4 * pass + i2;
                        if ((vmask >> i & 1) && col < p.MF) {
                            const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i2 + rsub) * STG_LD + cl);
                            const float4 o = make_float4(v.x + bias.x, v.y + bias.y, v.z + bias.z, v.w + bias.w);
                            const long long off = orow_off + j * 32 + (long long)i * 4 * p.MF;
                            if (p.upd_n == 0) {
                                *reinterpret_cast<float4*>(p.out + off) = o;
                            } else {
                                // same operation order as lincomb_kernel: acc = 0; acc = fma(c_i, src_i, acc) in term order
                                float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
                                bool usedA = false;
                                for (int n = 0; n < p.upd_n; ++n) {
                                    const float c = uc[n];
                                    float4 sv = o;
                                    if (p.upd_src[n]) { sv = usedA ? preB[i] : preA[i]; usedA = true; }
                                    u.x = fmaf(c, sv.x, u.x);
                                    u.y = fmaf(c, sv.y, u.y);
                                    u.z = fmaf(c, sv.z, u.z);
                                    u.w = fmaf(c, sv.w, u.w);
                                }
                                *reinterpret_cast<float4*>(p.upd_x + off) = u;
                                uint2 h;
                                h.x = Half16<BF16>::pack2(u.x, u.y);
                                h.y = Half16<BF16>::pack2(u.z, u.w);
                                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.upd_xh) + off) = h;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Every CTA of this kernel must be resident at once (tiles spin on their neighbours' flags): the capacity is what the driver
// reports for THIS launch configuration on THIS device as it is now (2-CTA clusters, 231 KB of shared memory, 384 threads; MPS SM
// limits, MIG slices and green contexts included), not the SM count.
template <int BF16>
static int stack_capacity() {
    static int cap[MAX_DEVICES] = {};
    const int dev = current_device();
    if (!cap[dev]) {
        cudaFuncSetAttribute(wavenet_stack_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(CLUSTER);
        cfg.blockDim = dim3(NTHREADS);
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CLUSTER;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int q = 0;
        if (cudaOccupancyMaxActiveClusters(&q, wavenet_stack_kernel<BF16>, &cfg) != cudaSuccess) { cudaGetLastError(); q = 0; }
        cap[dev] = q > 0 ? q * CLUSTER : -1;
    }
    return cap[dev] > 0 ? cap[dev] : 0;
}

template <int BF16>
static int launch_stack(const StackP& p, int grid, cudaStream_t st) {
    static PerDevice configured;
    if (configured.first()) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(wavenet_stack_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
    if (grid > stack_capacity<BF16>()) {
        set_error("b2s_tc_wavenet_stack: %d tiles do not fit the device at once (the driver reports room for %d co-resident CTAs of "
                  "this kernel); split the batch by utterance", grid, stack_capacity<BF16>());
        return B2S_ERR_UNSUPPORTED;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg, wavenet_stack_kernel<BF16>, p));
    return B2S_OK;
}

}  // namespace ws

#ifdef B2S_EXPERIMENTS
// =====================================================================================================================
// Version 4: the whole-stack kernel with cta_group::2 MMAs.  The two CTAs of a cluster (neighbouring time tiles) form ONE
// 256-row MMA: each keeps its own A tile and HALF of every weight tile in shared memory (32 KB per K slab instead of
// 48 KB: SM ingress, not L2, is what bounds the per-layer GEMMs), only the leader issues tcgen05.mma / commit, both CTAs'
// TMA loads complete on the LEADER's full barrier, and the leader waits for BOTH CTAs' epilogues (remote mbarrier
// arrives) before it overwrites TMEM halves or reads the z tiles.
// =====================================================================================================================
namespace ws2 {

using ws::StackP;
using ws::wait_flag;

using wl::ldg_nc_u4;
using wl::pdl_launch_dependents;
using wl::pdl_wait;
using wl::st_shared_u4;

constexpr int C = 256, MAXL = 32;
constexpr int BM = 128, BK = 64, BN = 256, UK = 16, STAGES = 4, CLUSTER = 2;
constexpr int A_BYTES = BM * BK * 2, BH_BYTES = (BN / 2) * BK * 2, STAGE_BYTES = A_BYTES + BH_BYTES;   // each CTA keeps HALF of B
constexpr int Z_BYTES = BM * C * 2;
constexpr int G1_KB = 3 * C / BK, G2_KB = C / BK;
constexpr int FILLS = 2 * G1_KB + 2 * G2_KB;             // 32 per layer
constexpr int STG_LD = 36, STG_ROWS = 16;
constexpr int STG_WARP_BYTES = STG_ROWS * STG_LD * 4;    // 2304: 16-row transpose buffer (two passes per 32-row chunk)
constexpr int NTHREADS = 384, EPI_WARPS = 8;
constexpr int SMEM_BYTES = Z_BYTES + STAGES * STAGE_BYTES + EPI_WARPS * STG_WARP_BYTES + 256;   // 231680
constexpr uint16_t MASK = (1u << CLUSTER) - 1;

template <int BF16>
__global__ void __launch_bounds__(NTHREADS, 1) wavenet_stack_cg2_kernel(const __grid_constant__ StackP p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* zs = smem;
    uint8_t* stages = smem + Z_BYTES;
    float* stg_all = reinterpret_cast<float*>(stages + STAGES * STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + STAGES * STAGE_BYTES + EPI_WARPS * STG_WARP_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* accb = empty + STAGES;                      // [4]: G1 half 0, G1 half 1, G2 residual, G2 skip
    uint64_t* zready = accb + 4;                          // [2]
    uint64_t* tfree = zready + 2;                         // [2]: EPI2 has drained TMEM half h
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tfree + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int b = blockIdx.x / p.tiles_per_b, ti = blockIdx.x - b * p.tiles_per_b, t0 = ti * BM;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("b2s: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.mapY[0]);
        prefetch_tmap(&p.mapY[1]);
        prefetch_tmap(&p.mapWd);
        prefetch_tmap(&p.mapWo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);                       // one multicast commit from the leader's MMA thread
        }
        for (int i = 0; i < 4; ++i) mbar_init(&accb[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&zready[i], EPI_WARPS * CLUSTER);    // the leader's MMA thread waits for BOTH CTAs' epilogues
            mbar_init(&tfree[i], EPI_WARPS * CLUSTER);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_cg2(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int l = 0; l < p.L; ++l) {
            const int dil = p.dil[l];
            if (lane == 0 && l + 1 < p.L && t0 < p.T) {
                // the next layer's 128 KB slab of the hoisted conditioner projection -> L2 (it streams from HBM once per evaluation)
                const uint8_t* nxt = reinterpret_cast<const uint8_t*>(p.cond) + ((l + 1) * p.cond_lstride + (long long)blockIdx.x * (BM * 2 * C)) * 2;
                for (int i = 0; i < 8; ++i)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt + i * 16384), "r"(16384) : "memory");
            }
            for (int f = 0; f < FILLS; ++f) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sa = stages + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    const uint32_t lbar = mapa_u32(&full[stage], 0);          // the LEADER's full barrier collects both CTAs' bytes
                    if (f < 2 * G1_KB) {
                        const int h = f / G1_KB, kb = f - h * G1_KB;
                        const int tap = kb / (C / BK), c0 = (kb - tap * (C / BK)) * BK;
                        if (rank == 0) mbar_expect_tx(&full[stage], CLUSTER * STAGE_BYTES);
                        tma_load_3d_cg2(sb, &p.mapWd, lbar, kb * BK, h * BN + rank * (BN / 2), l);
                        if (f == 0 && l > 0 && t0 < p.T) {
                            const int* fl = p.flags + b * p.tiles_per_b;
                            if (ti > 0) wait_flag(fl + ti - 1, l);
                            wait_flag(fl + ti, l);
                            if ((ti + 1) * BM < p.T) wait_flag(fl + ti + 1, l);
                            fence_proxy_async_all();
                        }
                        if (f == 0) TLOG(0);
                        tma_load_3d_cg2(sa, &p.mapY[l & 1], lbar, c0, t0 + (tap - 1) * dil, b);
                    } else {
                        const int g = (f - 2 * G1_KB) / G2_KB, kb = (f - 2 * G1_KB) - g * G2_KB;
                        if (rank == 0) mbar_expect_tx(&full[stage], CLUSTER * BH_BYTES);
                        tma_load_3d_cg2(sb, &p.mapWo, lbar, kb * BK, g * BN + rank * (BN / 2), l);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (leader CTA only): M = 256 across the pair =====================
        const uint32_t idesc = make_idesc_f16(2 * BM, BN, BF16);
        int stage = 0;
        uint32_t phase = 0;
        for (int l = 0; l < p.L; ++l) {
            const uint32_t par = l & 1;
            for (int h = 0; h < 2; ++h) {
                if (l > 0) {                                   // EPI2 of the previous layer has drained these columns
                    mbar_wait(&tfree[h], par ^ 1);
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + h * BN;
                for (int kb = 0; kb < G1_KB; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (kb == 0) TLOG(1 + h);
                    if (lane == 0) {
                        const uint32_t a_addr = smem_u32(stages + stage * STAGE_BYTES), b_addr = a_addr + A_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma_ss_cg2(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UK * 2)), make_sw128_kmajor_desc(b_addr + k * (UK * 2)),
                                    idesc, (kb | k) != 0);
                        umma_commit_cg2_mcast(&empty[stage], MASK);
                        if (kb == G1_KB - 1) umma_commit_cg2_mcast(&accb[h], MASK);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            for (int g = 0; g < 2; ++g) {
                const uint32_t d_tmem = tmem_base + g * BN;
                for (int kb = 0; kb < G2_KB; ++kb) {
                    if (g == 0 && (kb == 0 || kb == 2)) {
                        mbar_wait(&zready[kb >> 1], par);
                        tc_fence_after();
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_addr = smem_u32(zs + kb * A_BYTES);
                        const uint32_t b_addr = smem_u32(stages + stage * STAGE_BYTES) + A_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma_ss_cg2(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UK * 2)), make_sw128_kmajor_desc(b_addr + k * (UK * 2)),
                                    idesc, (kb | k) != 0);
                        umma_commit_cg2_mcast(&empty[stage], MASK);
                        if (kb == G2_KB - 1) umma_commit_cg2_mcast(&accb[2 + g], MASK);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 8 warps, warp e -> lane quarter e&3, chunks j with (j&1) == e>>2 ===========
        const int e = warp - 4, q = e & 3, sub = e >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int row = q * 32 + lane;
        const bool valid = t0 + row < p.T;
        const uint32_t zrow = smem_u32(zs) + (row >> 3) * 1024 + (row & 7) * 128;
        const int sw = row & 7;
        float* stg = stg_all + e * (STG_ROWS * STG_LD);
        const int cl = (lane & 7) * 4, rsub = lane >> 3;
        const int tq = t0 + q * 32 + rsub;
        const float inv_sqrt2 = 0.70710678118654752440f;
        // per-lane constants of the coalesced (EPI2) layout
        uint32_t vmask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) vmask |= (tq + 4 * i < p.T ? 1u : 0u) << i;
        const long long rowoff = ((long long)b * p.T + tq) * C + cl;          // element offset of (row 0 of the lane, its 4 columns)
        float* xrow = p.x + rowoff;
        float* srow_g = p.skip + rowoff;
        uint16_t* shrow = reinterpret_cast<uint16_t*>(p.skip_h) + rowoff;

#pragma unroll 1
        for (int l = 0; l < p.L; ++l) {
            const uint32_t par = l & 1;
            const bool first = l == 0, last = l == p.L - 1;
            // ---- EPI1: thread = frame row; + cond; gate; z -> swizzled smem ----
            // tile/chunk-major table: (chunk j, 16-byte piece i) of this tile = 128 rows x 16 B contiguous
            const uint16_t* ctile = reinterpret_cast<const uint16_t*>(p.cond) + l * p.cond_lstride + (long long)blockIdx.x * (BM * 2 * C) + row * 8;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint4 c[4][4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        c[jj][i] = (valid && !(p.dbg & 1)) ? ldg_nc_u4(ctile + ((8 * h + 2 * jj + sub) * 4 + i) * (BM * 8)) : make_uint4(0, 0, 0, 0);
                mbar_wait(&accb[h], par);
                tc_fence_after();
                if (e == 0) TLOG(3 + 2 * h);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 2 * jj + sub;
                    float acc[32];
                    tmem_ld32(taddr + h * BN + j * 32, acc);
                    tmem_ld_wait();
                    const uint32_t* cw = reinterpret_cast<const uint32_t*>(c[jj]);
                    uint32_t zp[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 ca = Half16<BF16>::unpack2(cw[2 * i]), cb = Half16<BF16>::unpack2(cw[2 * i + 1]);
                        float z0, z1;
                        if (p.dbg & 8) {
                            z0 = (acc[4 * i] + ca.x) * (acc[4 * i + 1] + ca.y);
                            z1 = (acc[4 * i + 2] + cb.x) * (acc[4 * i + 3] + cb.y);
                        } else {
                            z0 = sigmoid_fast(acc[4 * i] + ca.x) * tanh_fast(acc[4 * i + 1] + ca.y);
                            z1 = sigmoid_fast(acc[4 * i + 2] + cb.x) * tanh_fast(acc[4 * i + 3] + cb.y);
                        }
                        zp[i] = valid ? Half16<BF16>::pack2(z0, z1) : 0u;
                    }
                    const uint32_t slab = zrow + (2 * h + (j >> 2)) * A_BYTES;
                    const int c16 = 2 * (j & 3);
                    st_shared_u4(slab + ((c16 ^ sw) << 4), make_uint4(zp[0], zp[1], zp[2], zp[3]));
                    st_shared_u4(slab + (((c16 + 1) ^ sw) << 4), make_uint4(zp[4], zp[5], zp[6], zp[7]));
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_u32(&zready[h], 0));
                if (e == 0) TLOG(4 + 2 * h);
            }
            // ---- EPI2: coalesced layout through the warp's 16-row staging tile (two passes per 32-row chunk).
            //      Row i of the lane is frame tq + 4i: all addresses are (lane base) + (compile-time i) * 4C + (chunk) * 32,
            //      validity is a precomputed bit mask. ----
            const float* bo = p.bo + (long long)l * 2 * C;
            const float* dnext = p.dvec + (long long)(l + 1) * C + (long long)b * p.d_stride;
            uint16_t* ynext = last ? nullptr : reinterpret_cast<uint16_t*>(p.ybuf[(l + 1) & 1]) + rowoff;
            float4 in[8], inn[8];
            auto load_inputs = [&](int g, int j, float4* dst) {
                const float* src = (g == 0 ? xrow : srow_g) + j * 32;
                const bool rd = (g == 0 || !first) && !(p.dbg & 2);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = (rd && (vmask >> i & 1)) ? *reinterpret_cast<const float4*>(src + i * 4 * C) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            load_inputs(0, sub, inn);
#pragma unroll 1
            for (int n = 0; n < 8; ++n) {
                const int g = n >> 2, j = 2 * (n & 3) + sub;
                if ((n & 3) == 0) {
                    mbar_wait(&accb[2 + g], par);
                    tc_fence_after();
                    if (e == 0) TLOG(7 + 2 * g);
                }
                float acc[32];
                tmem_ld32(taddr + g * BN + j * 32, acc);
#pragma unroll
                for (int i = 0; i < 8; ++i) in[i] = inn[i];
                if (n + 1 < 8) load_inputs((n + 1) >> 2, 2 * ((n + 1) & 3) + sub, inn);
                const int col = j * 32 + cl;
                const float4 bias = __ldg(reinterpret_cast<const float4*>(bo + g * C + col));
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g == 0 && ynext) d = __ldg(reinterpret_cast<const float4*>(dnext + col));
                float* xo = xrow + j * 32;
                float* so = srow_g + j * 32;
                uint16_t* yo = ynext + j * 32;
                uint16_t* sho = shrow + j * 32;
                tmem_ld_wait();
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if ((lane >> 4) == pass) {
                        float4* srow = reinterpret_cast<float4*>(stg + (lane & 15) * STG_LD);
#pragma unroll
                        for (int c2 = 0; c2 < 8; ++c2)
                            srow[c2] = make_float4(acc[4 * c2], acc[4 * c2 + 1], acc[4 * c2 + 2], acc[4 * c2 + 3]);
                    }
                    __syncwarp();
                    if (!(p.dbg & 4)) {
#pragma unroll
                        for (int i2 = 0; i2 < 4; ++i2) {
                            const int i = 4 * pass + i2;                        // row 4*i + rsub of the warp's 32
                            if (vmask >> i & 1) {
                                const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i2 + rsub) * STG_LD + cl);
                                const float4 o = make_float4(v.x + bias.x, v.y + bias.y, v.z + bias.z, v.w + bias.w);
                                if (g == 0) {
                                    const float4 xn = make_float4((in[i].x + o.x) * inv_sqrt2, (in[i].y + o.y) * inv_sqrt2,
                                                                  (in[i].z + o.z) * inv_sqrt2, (in[i].w + o.w) * inv_sqrt2);
                                    *reinterpret_cast<float4*>(xo + i * 4 * C) = xn;
                                    if (ynext) {
                                        uint2 yv;
                                        yv.x = Half16<BF16>::pack2(xn.x + d.x, xn.y + d.y);
                                        yv.y = Half16<BF16>::pack2(xn.z + d.z, xn.w + d.w);
                                        *reinterpret_cast<uint2*>(yo + i * 4 * C) = yv;
                                    }
                                } else {
                                    const float4 s2 = make_float4(o.x + in[i].x, o.y + in[i].y, o.z + in[i].z, o.w + in[i].w);
                                    *reinterpret_cast<float4*>(so + i * 4 * C) = s2;
                                    if (last && p.skip_h) {
                                        uint2 sv;
                                        sv.x = Half16<BF16>::pack2(s2.x, s2.y);
                                        sv.y = Half16<BF16>::pack2(s2.z, s2.w);
                                        *reinterpret_cast<uint2*>(sho + i * 4 * C) = sv;
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
                if ((n & 3) == 3) {
                    // this warp has finished TMEM half g of the layer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_u32(&tfree[g], 0));
                    if (g == 0) {
                        // hand-off: every epilogue thread has issued its y_next stores -> CTA barrier -> ONE thread makes them
                        // visible GPU-wide (the fence is cumulative over the barrier) and releases the tile flag
                        named_bar_sync(1, EPI_WARPS * 32);
                        if (e == 0 && lane == 0) {
                            __threadfence();
                            st_release_gpu(p.flags + blockIdx.x, l + 1);
                        }
                    }
                    if (e == 0) TLOG(8 + 2 * g);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, 512);
    }
}

template <int BF16>
static int launch_stack_cg2(const StackP& p, int grid, cudaStream_t st) {
    static PerDevice configured;
    if (configured.first()) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(wavenet_stack_cg2_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg, wavenet_stack_cg2_kernel<BF16>, p));
    return B2S_OK;
}

}  // namespace ws2
#endif  // B2S_EXPERIMENTS
}  // namespace tc
}  // namespace b2s

using namespace b2s;
using namespace b2s::tc;

extern "C" int b2s_tc_wavenet_layer(const void* y_h, const void* Wd_h, const void* cond_h, int ld_cond, const void* Wo_h,
                                    const float* bo, float* x, void* y_next_h, float* skip, void* skip_h,
                                    const float* dvec_next, int d_stride, int first_layer, int B, int T, int C, int dilation,
                                    int bf16, void* stream) {
    B2S_CHECK_ARG(y_h && Wd_h && cond_h && Wo_h && bo && x && skip, "b2s_tc_wavenet_layer: null pointer");
    if (C != wl::C) {
        set_error("b2s_tc_wavenet_layer: the fused layer kernel is specialised for %d residual channels (got %d); use "
                  "b2s_tc_wavenet_gate + b2s_tc_wavenet_out", wl::C, C);
        return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_ARG(y_next_h != y_h, "b2s_tc_wavenet_layer: y_next must not alias y (neighbouring tiles read its halo)");
    B2S_CHECK_ARG(!y_next_h || (dvec_next && d_stride % 4 == 0 && al16(dvec_next)), "b2s_tc_wavenet_layer: y_next needs dvec_next");
    B2S_CHECK_ARG(dilation >= 1 && ld_cond % 8 == 0 && al16(cond_h) && al16(y_h) && al16(Wd_h) && al16(Wo_h) && al16(bo) &&
                      al16(x) && al16(skip) && (!y_next_h || al16(y_next_h)) && (!skip_h || al16(skip_h)),
                  "b2s_tc_wavenet_layer: bad dilation / alignment");
    if (B * T == 0) return B2S_OK;
    // variant 2 (2-CTA cluster, multicast weights); experiment builds: variant 1 when B2S_LAYER_V1 is set (A/B comparison)
#ifdef B2S_EXPERIMENTS
    static const bool v1 = getenv("B2S_LAYER_V1") != nullptr;
#else
    const bool v1 = false;
#endif
    wl::LayerP p{};
    int rc = make_map_act(&p.mapY, y_h, bf16, C, C, T, B, wl::BK, wl::BM);
    if (rc) return rc;
    rc = make_map_w(&p.mapWd, Wd_h, bf16, 3 * C, 2 * C, 3 * C, wl::BK, v1 ? wl::BN : wl::BN / 2);
    if (rc) return rc;
    rc = make_map_w(&p.mapWo, Wo_h, bf16, C, 2 * C, C, wl::BK, v1 ? wl::BN : wl::BN / 2);
    if (rc) return rc;
    p.B = B; p.T = T; p.tiles_per_b = ceil_div(T, wl::BM); p.dil = dilation;
    if (!v1) p.tiles_per_b = (p.tiles_per_b + 1) & ~1;          // whole clusters per utterance
    p.cond = cond_h; p.ldc = ld_cond; p.bo = bo; p.x = x; p.y_next = y_next_h; p.skip = skip; p.skip_h = skip_h;
    p.dvec = dvec_next; p.d_stride = d_stride; p.first = first_layer;
    const int grid = B * p.tiles_per_b;
#ifdef B2S_EXPERIMENTS
    if (v1) return bf16 ? wl::launch_layer<1>(p, grid, (cudaStream_t)stream) : wl::launch_layer<0>(p, grid, (cudaStream_t)stream);
#endif
    return bf16 ? wl2::launch_layer<1>(p, grid, (cudaStream_t)stream) : wl2::launch_layer<0>(p, grid, (cudaStream_t)stream);
}

unsigned long long* g_tlog = nullptr;      // shared with b2s_tc_wavenet_t.cu
/* profiling hook (not part of the product API): device buffer [L][16] of globaltimer ns, or NULL to switch off */
extern "C" void b2s_debug_set_stack_tlog(void* buf) { g_tlog = (unsigned long long*)buf; }

/* 1 when the library was built with B2S_BUILD_EXPERIMENTS=1 (the measured-and-rejected variants of DESIGN.md section 3.3) */
extern "C" int b2s_has_experiments(void) {
#ifdef B2S_EXPERIMENTS
    return 1;
#else
    return 0;
#endif
}

extern "C" int b2s_tc_wavenet_stack_max_tiles(void) {
    const int a = ws::stack_capacity<1>(), b = ws::stack_capacity<0>();      // cudaOccupancyMaxActiveClusters x 2, per device
    return a < b ? a : b;
}

struct DenoiserIO {            // stem / head operands of b2s_tc_wavenet_denoiser (all NULL for the plain stack)
    const void* xin_h; int MF; const void* Win_h; int ld_win; const float* b_in;
    const void* Wsp_h; const float* b_sp; const void* Wfin_h; const float* b_fin; float* out;
    int upd_n; const float* upd_src[3]; const float* upd_coef; float* upd_x; void* upd_xh; int* flags_next;
};

static int stack_impl(void* y0_h, void* y1_h, const void* Wd_h, const void* cond_h, int ld_cond, int64_t cond_layer_stride,
                      const void* Wo_h, const float* bo, float* x, float* skip, void* skip_h, const float* dvec, int d_stride,
                      const int* dilations_host, int L, int B, int T, int C, int* flags, int bf16, void* stream,
                      const DenoiserIO* io) {
    B2S_CHECK_ARG(y0_h && y1_h && Wd_h && cond_h && Wo_h && bo && x && skip && dvec && dilations_host && flags,
                  "b2s_tc_wavenet_stack: null pointer");
    if (C != ws::C) {
        set_error("b2s_tc_wavenet_stack: specialised for %d residual channels (got %d)", ws::C, C);
        return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_ARG(L >= 1 && L <= ws::MAXL, "b2s_tc_wavenet_stack: 1 <= L <= %d (got %d)", ws::MAXL, L);
    B2S_CHECK_ARG(y0_h != y1_h && d_stride % 4 == 0 && cond_layer_stride % 8 == 0,
                  "b2s_tc_wavenet_stack: y buffers must differ; strides must keep 16B alignment");
    B2S_CHECK_ARG(al16(y0_h) && al16(y1_h) && al16(Wd_h) && al16(Wo_h) && al16(cond_h) && al16(bo) && al16(x) && al16(skip) &&
                      al16(dvec) && (!skip_h || al16(skip_h)), "b2s_tc_wavenet_stack: misaligned pointer");
    if (B * T == 0) return B2S_OK;
    ws::StackP p{};
    p.tiles_per_b = (ceil_div(T, ws::BM) + 1) & ~1;
    const int grid = B * p.tiles_per_b;
    if (grid > b2s_tc_wavenet_stack_max_tiles()) {
        set_error("b2s_tc_wavenet_stack: %d tiles do not fit the device at once (room for %d co-resident CTAs; every tile must be "
                  "resident); split the batch by utterance", grid, b2s_tc_wavenet_stack_max_tiles());
        return B2S_ERR_UNSUPPORTED;
    }
    int rc = make_map_act(&p.mapY[0], y0_h, bf16, C, C, T, B, ws::BK, ws::BM);
    if (rc) return rc;
    rc = make_map_act(&p.mapY[1], y1_h, bf16, C, C, T, B, ws::BK, ws::BM);
    if (rc) return rc;
    rc = make_map_w3(&p.mapWd, Wd_h, bf16, 3 * C, 2 * C, L, ws::BK, ws::BN / 2);
    if (rc) return rc;
    rc = make_map_w3(&p.mapWo, Wo_h, bf16, C, 2 * C, L, ws::BK, ws::BN / 2);
    if (rc) return rc;
    p.B = B; p.T = T; p.L = L;
    for (int l = 0; l < L; ++l) {
        B2S_CHECK_ARG(dilations_host[l] >= 1 && dilations_host[l] < ws::BM, "b2s_tc_wavenet_stack: bad dilation %d", dilations_host[l]);
        p.dil[l] = dilations_host[l];
    }
    p.cond = cond_h; p.ldc = ld_cond; p.cond_lstride = cond_layer_stride; p.bo = bo;
    p.x = x; p.ybuf[0] = y0_h; p.ybuf[1] = y1_h; p.skip = skip; p.skip_h = skip_h;
    p.dvec = dvec; p.d_stride = d_stride; p.flags = flags;
    if (io) {
        B2S_CHECK_ARG(io->xin_h && io->Win_h && io->b_in && io->Wsp_h && io->b_sp && io->Wfin_h && io->b_fin && (io->out || io->upd_n > 0),
                      "b2s_tc_wavenet_denoiser: null pointer");
        B2S_CHECK_ARG(io->upd_n >= 0 && io->upd_n <= 3, "b2s_tc_wavenet_denoiser_update: 1 <= n_terms <= 3 (got %d)", io->upd_n);
        if (io->upd_n > 0) {
            B2S_CHECK_ARG(io->upd_coef && io->upd_x && io->upd_xh && al16(io->upd_x) && al16(io->upd_xh),
                          "b2s_tc_wavenet_denoiser_update: null or misaligned update operand");
            for (int i = 0; i < io->upd_n; ++i)
                B2S_CHECK_ARG(al16(io->upd_src[i]), "b2s_tc_wavenet_denoiser_update: misaligned source %d", i);
        }
        B2S_CHECK_ARG(io->MF > 0 && io->MF <= ws::BN && io->MF % 8 == 0 && io->ld_win % 8 == 0,
                      "b2s_tc_wavenet_denoiser: in_dims*n_feats must be a multiple of 8 and <= %d (got %d)", ws::BN, io->MF);
        B2S_CHECK_ARG(al16(io->xin_h) && al16(io->Win_h) && al16(io->b_in) && al16(io->Wsp_h) && al16(io->b_sp) && al16(io->Wfin_h) &&
                          al16(io->b_fin) && (!io->out || al16(io->out)), "b2s_tc_wavenet_denoiser: misaligned pointer");
        p.fuse = 1; p.MF = io->MF; p.kb_in = ceil_div(io->MF, ws::BK);
        rc = make_map_act(&p.mapXin, io->xin_h, bf16, io->MF, io->MF, T, B, ws::BK, ws::BM);
        if (rc) return rc;
        rc = make_map_w(&p.mapWin, io->Win_h, bf16, io->MF, C, io->ld_win, ws::BK, ws::BN / 2);
        if (rc) return rc;
        rc = make_map_w(&p.mapWsp, io->Wsp_h, bf16, C, C, C, ws::BK, ws::BN / 2);
        if (rc) return rc;
        rc = make_map_w(&p.mapWfin, io->Wfin_h, bf16, C, io->MF, C, ws::BK, ws::BN / 2);
        if (rc) return rc;
        p.b_in = io->b_in; p.b_sp = io->b_sp; p.b_fin = io->b_fin; p.out = io->out;
        p.upd_n = io->upd_n; p.upd_coef = io->upd_coef; p.upd_x = io->upd_x; p.upd_xh = io->upd_xh; p.flags_next = io->flags_next;
        for (int i = 0; i < 3; ++i) p.upd_src[i] = i < io->upd_n ? io->upd_src[i] : nullptr;
        p.alpha_head = 1.0f / sqrtf((float)L);
    }
    static const int dbg = getenv("B2S_STACK_DBG") ? atoi(getenv("B2S_STACK_DBG")) : 0;
    p.dbg = dbg;
    p.tlog = g_tlog;
#ifdef B2S_EXPERIMENTS
    static const bool cg2 = getenv("B2S_STACK_CG2") != nullptr && atoi(getenv("B2S_STACK_CG2")) != 0;
    if (cg2 && !io) return bf16 ? ws2::launch_stack_cg2<1>(p, grid, (cudaStream_t)stream) : ws2::launch_stack_cg2<0>(p, grid, (cudaStream_t)stream);
#endif
    return bf16 ? ws::launch_stack<1>(p, grid, (cudaStream_t)stream) : ws::launch_stack<0>(p, grid, (cudaStream_t)stream);
}

extern "C" int b2s_tc_wavenet_stack(void* y0_h, void* y1_h, const void* Wd_h, const void* cond_h, int ld_cond,
                                    int64_t cond_layer_stride, const void* Wo_h, const float* bo, float* x, float* skip,
                                    void* skip_h, const float* dvec, int d_stride, const int* dilations_host, int L, int B,
                                    int T, int C, int* flags, int bf16, void* stream) {
    return stack_impl(y0_h, y1_h, Wd_h, cond_h, ld_cond, cond_layer_stride, Wo_h, bo, x, skip, skip_h, dvec, d_stride,
                      dilations_host, L, B, T, C, flags, bf16, stream, nullptr);
}

extern "C" int b2s_tc_wavenet_denoiser(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, void* y0_h,
                                       void* y1_h, const void* Wd_h, const void* cond_h, int64_t cond_layer_stride,
                                       const void* Wo_h, const float* bo, float* x, float* skip, const float* dvec, int d_stride,
                                       const int* dilations_host, int L, const void* Wsp_h, const float* b_sp, const void* Wfin_h,
                                       const float* b_fin, float* out, int B, int T, int C, int* flags, int bf16, void* stream) {
    DenoiserIO io{xin_h, MF, Win_h, ld_win, b_in, Wsp_h, b_sp, Wfin_h, b_fin, out, 0, {nullptr, nullptr, nullptr}, nullptr, nullptr, nullptr, nullptr};
    return stack_impl(y0_h, y1_h, Wd_h, cond_h, 2 * C, cond_layer_stride, Wo_h, bo, x, skip, nullptr, dvec, d_stride, dilations_host,
                      L, B, T, C, flags, bf16, stream, &io);
}

#ifdef B2S_EXPERIMENTS
extern "C" int b2s_tc_wavenet_denoiser_update(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, void* y0_h,
                                              void* y1_h, const void* Wd_h, const void* cond_h, int64_t cond_layer_stride,
                                              const void* Wo_h, const float* bo, float* x, float* skip, const float* dvec, int d_stride,
                                              const int* dilations_host, int L, const void* Wsp_h, const float* b_sp, const void* Wfin_h,
                                              const float* b_fin, int B, int T, int C, int* flags, int* flags_next, int n_terms,
                                              const float* const* srcs_host, const float* coef, float* x_out, void* x_out_h, int bf16,
                                              void* stream) {
    B2S_CHECK_ARG(n_terms >= 1 && n_terms <= 3 && srcs_host, "b2s_tc_wavenet_denoiser_update: 1 <= n_terms <= 3 (got %d)", n_terms);
    DenoiserIO io{xin_h, MF, Win_h, ld_win, b_in, Wsp_h, b_sp, Wfin_h, b_fin, nullptr, n_terms, {nullptr, nullptr, nullptr}, coef, x_out, x_out_h,
                  flags_next};
    for (int i = 0; i < n_terms; ++i) io.upd_src[i] = srcs_host[i];
    return stack_impl(y0_h, y1_h, Wd_h, cond_h, 2 * C, cond_layer_stride, Wo_h, bo, x, skip, nullptr, dvec, d_stride, dilations_host,
                      L, B, T, C, flags, bf16, stream, &io);
}
#endif  // B2S_EXPERIMENTS
