// WaveNet residual stack (wavenet.py:33-48, 92-94), version 5: the whole stack as ONE persistent kernel with the GEMMs
// TRANSPOSED - the tensor-core M dimension is the CHANNEL axis and N is the TIME axis:
//
//     G1   D1[ch, t] = sum_k Wd[ch, k] * Y[t (+ tap shift), k]      M = 128 channels per block (4 blocks: gate/filter x 2 halves)
//     G2   D2[ch, t] = sum_k Wo[ch, k] * Z[t, k]                    N = NT frames (a multiple of 16, <= 80), K = 768 / 256
//
// Why: tcgen05 charges max(M,128) x N per instruction, so a frame tile can only shrink along N.  With N = 80 the 16 x 690
// frames of BASELINE config 2 become 144 tiles - one per SM on 144 of the 148 SMs - instead of the 96 tiles of the 128-row
// formulation (version 3, b2s_tc_wavenet.cu).  Everything else follows from putting channels on the TMEM lanes:
//   * an epilogue thread owns ONE channel (lane) and a range of frames (columns): biases / step embeddings are per-thread
//     scalars, global accesses are coalesced across the warp without any shared-memory transposition;
//   * the fp32 residual stream x never leaves the SM: each thread keeps its NT/2 x 2 values in REGISTERS for all L layers;
//   * the skip sum never leaves the SM either: the G2 skip blocks ACCUMULATE over all layers in 2 x NT spare TMEM columns
//     (6 x NT <= 512) and are drained once, after the last layer;
//   * the activation tile y_l is loaded ONCE per layer as 4 K slabs of (NT + 2 * Dmax) rows; the three dilation taps are
//     row-shifted windows of the same slabs (matrix descriptors that start r rows into a SWIZZLE_128B slab);
//   * what streams per layer is only the weights: 64 tiles of [128 x 64] (16 KB) through an 8-deep ring, every tile fetched
//     half by each CTA of a 2-CTA cluster and multicast to both.
// Hand-off between layers (dilation halo of the neighbouring tiles) is the release / acquire tile flag of version 3.
#include <stdlib.h>

#include "b2s_common.cuh"
// EXPERIMENT (measured slower than the 128-row kernels, DESIGN.md section 3.3 'instruction floor'): compiled only with
// B2S_BUILD_EXPERIMENTS=1 (-DB2S_EXPERIMENTS).
#ifdef B2S_EXPERIMENTS
#include "b2s_tc.cuh"

namespace b2s {
namespace tc {
namespace wt {

constexpr int C = 256, MAXL = 32;
constexpr int BM = 128, BK = 64, UK = 16, NSTG = 8, CLUSTER = 2, PRE = 6;
constexpr int WT_BYTES = BM * BK * 2, WH_BYTES = WT_BYTES / 2;          // weight tile 16 KB, multicast half 8 KB
constexpr int G1_KB = 3 * C / BK, G2_KB = C / BK;                      // 12, 4
constexpr int G1_FILLS = 4 * G1_KB, FILLS = G1_FILLS + 4 * G2_KB;       // 48 + 16 weight tiles per layer
constexpr int NTHREADS = 320, EPI_WARPS = 8;                            // warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 epilogue
constexpr uint16_t MASK = (1u << CLUSTER) - 1;

struct __align__(64) StackTP {
    CUtensorMap mapY[2], mapWd, mapWo;
    int B, T, tiles_per_b, L, dmax, R;                   // R = NT + 2 * dmax rows per activation slab
    int dil[MAXL];
    const void* cond; long long cond_lstride;             // retiled table (b2s_tc_cond_retile), elements per layer
    const float* bo;                                      // [L][2C]  (residual | skip)
    const float* x;                                       // [rows, C] fp32: the stem's output (read once)
    void* ybuf[2];                                        // y_0 = x + d_0 in ybuf[0] on entry
    void* skip_h;                                         // out: sum of the L skip outputs, 16-bit [rows, C]
    const float* dvec; int d_stride;
    int* flags;
    int dbg;
    unsigned long long* tlog;                             // optional phase stamps of CTA 2 (-DB2S_TLOG builds): [L][16] globaltimer ns
};

#ifdef B2S_TLOG
#define TLOGT(slot) do { if (p.tlog && blockIdx.x == 2 && lane == 0) p.tlog[l * 16 + (slot)] = globaltimer_ns(); } while (0)
#else
#define TLOGT(slot) do { } while (0)
#endif

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
template <int BF16>
__device__ __forceinline__ uint16_t half1(float v) {
    if (BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
    return __half_as_ushort(__float2half_rn(v));
}

__device__ __forceinline__ void wait_flag(const int* f, int want) {
    if (ld_acquire_gpu(f) >= want) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spin = 0;
    while (ld_acquire_gpu(f) < want) {
        if ((++spin & 255u) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
            printf("b2s: tile flag timeout (transposed stack, block %d want %d have %d)\n", blockIdx.x, want, ld_acquire_gpu(f));
            __trap();
        }
    }
}

template <int NT, int BF16>
__global__ void __launch_bounds__(NTHREADS, 1) wavenet_stack_t_kernel(const __grid_constant__ StackTP p) {
    static_assert(NT % 16 == 0 && NT >= 16 && 6 * NT <= 512, "frame tile: multiple of 16, six accumulator blocks in 512 TMEM columns");
    constexpr int ZS_BYTES = NT * 128;                    // one K slab of the z tile: NT rows x 64 channels
    constexpr int NCH = NT / 16;                          // 8-column chunks per epilogue warp (it owns NT/2 columns)
    extern __shared__ __align__(1024) uint8_t smem[];
    const int YS_BYTES = p.R * 128;                       // one K slab of the activation tile
    uint8_t* ys = smem;
    uint8_t* zs = ys + 4 * YS_BYTES;
    uint8_t* wst = zs + 4 * ZS_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(wst + NSTG * WT_BYTES);
    uint64_t* empty = full + NSTG;
    uint64_t* yfull = empty + NSTG;
    uint64_t* accb = yfull + 1;                           // [3]: G1 blocks 0-1, G1 blocks 2-3, G2 residual blocks
    uint64_t* skipb = accb + 3;                           // skip accumulator complete (after the last layer)
    uint64_t* zready = skipb + 1;                         // [2]
    uint64_t* tfree = zready + 2;                         // residual accumulator drained (TMEM columns [0, 2NT) reusable)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tfree + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int b = blockIdx.x / p.tiles_per_b, ti = blockIdx.x - b * p.tiles_per_b, t0 = ti * NT;
    const bool tile_ok = b < p.B;                         // the grid is padded to an even size: the last CTA may be a dummy

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("b2s: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.mapY[0]);
        prefetch_tmap(&p.mapY[1]);
        prefetch_tmap(&p.mapWd);
        prefetch_tmap(&p.mapWo);
        for (int i = 0; i < NSTG; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], CLUSTER);
        }
        mbar_init(yfull, 1);
        for (int i = 0; i < 3; ++i) mbar_init(&accb[i], 1);
        mbar_init(skipb, 1);
        mbar_init(&zready[0], EPI_WARPS);
        mbar_init(&zready[1], EPI_WARPS);
        mbar_init(tfree, EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
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
        auto fill = [&](int l, int f) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0 && (p.dbg & 2)) {
                mbar_arrive(&full[stage]);                     // profiling experiment: no weight traffic at all (results are WRONG)
            } else if (lane == 0) {
                uint8_t* sw = wst + stage * WT_BYTES + rank * WH_BYTES;
                mbar_expect_tx(&full[stage], WT_BYTES);
                if (f < G1_FILLS) {
                    const int mb = f / G1_KB, kb = f - mb * G1_KB;
                    tma_load_3d_mcast(sw, &p.mapWd, &full[stage], kb * BK, mb * BM + rank * (BM / 2), l, MASK);
                } else {
                    const int g = f - G1_FILLS, mb = g / G2_KB, kb = g - mb * G2_KB;
                    tma_load_3d_mcast(sw, &p.mapWo, &full[stage], kb * BK, mb * BM + rank * (BM / 2), l, MASK);
                }
            }
            __syncwarp();
            if (++stage == NSTG) { stage = 0; phase ^= 1; }
        };
        for (int l = 0; l < p.L; ++l) {
            if (lane == 0 && l + 1 < p.L && tile_ok) {
                // the next layer's slab of the hoisted conditioner projection -> L2
                const uint8_t* nxt = reinterpret_cast<const uint8_t*>(p.cond) + ((l + 1) * p.cond_lstride + (long long)blockIdx.x * (NT * 2 * C)) * 2;
                for (int i = 0; i < NT / 16; ++i)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt + i * 16384), "r"(16384) : "memory");
            }
            for (int f = 0; f < PRE; ++f) fill(l, f);          // weights first: their latency hides behind the hand-off
            TLOGT(0);                                          // first PRE weight tiles requested
            if (l > 0) mbar_wait(tfree, (l - 1) & 1);          // this tile's layer l-1 is through: the y slabs may be overwritten
            TLOGT(1);                                          // own residual epilogue of layer l-1 done
            if (lane == 0) {
                if (l > 0 && tile_ok && t0 < p.T) {
                    const int* fl = p.flags + blockIdx.x;
                    if (ti > 0) wait_flag(fl - 1, l);
                    wait_flag(fl, l);
                    if ((ti + 1) * NT < p.T) wait_flag(fl + 1, l);
                    fence_proxy_async_all();
                }
                mbar_expect_tx(yfull, 4 * YS_BYTES);
                for (int s = 0; s < 4; ++s) tma_load_3d(ys + s * YS_BYTES, &p.mapY[l & 1], yfull, s * BK, t0 - p.dmax, b);
            }
            __syncwarp();
            TLOGT(2);                                          // neighbour flags acquired, y slabs requested
            for (int f = PRE; f < FILLS; ++f) fill(l, f);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_f16(BM, NT, BF16);
        int stage = 0;
        uint32_t phase = 0;
        for (int l = 0; l < p.L; ++l) {
            const uint32_t par = l & 1;
            const int dil = p.dil[l];
            mbar_wait(yfull, par);
            if (l > 0) mbar_wait(tfree, par ^ 1);
            tc_fence_after();
            TLOGT(3);                                          // y slabs landed: G1 may start
            for (int mb = 0; mb < 4; ++mb) {
                for (int kb = 0; kb < G1_KB; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const int tap = kb >> 2, s = kb & 3;
                        const uint32_t a_addr = smem_u32(wst + stage * WT_BYTES);
                        const uint32_t b_addr = smem_u32(ys + s * YS_BYTES) + (uint32_t)(p.dmax + (tap - 1) * dil) * 128u;
                        const uint32_t bo = (p.dbg & 1) ? (b_addr >> 7) : 0u;
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma_ss(tmem_base + mb * NT, make_sw128_kmajor_desc(a_addr + k * (UK * 2)),
                                    make_sw128_kmajor_desc_bo(b_addr + k * (UK * 2), bo), idesc, (kb | k) != 0);
                        umma_commit_mcast(&empty[stage], MASK);
                        if (kb == G1_KB - 1 && (mb & 1)) umma_commit(&accb[mb >> 1]);
                    }
                    __syncwarp();
                    if (kb == G1_KB - 1) TLOGT(4 + mb);              // G1 block mb issued (4..7)
                    if (++stage == NSTG) { stage = 0; phase ^= 1; }
                }
            }
            for (int mb = 0; mb < 4; ++mb) {
                // residual blocks reuse the columns of G1 blocks 0-1 (drained before zready[0]); skip blocks accumulate over layers
                const uint32_t d_tmem = tmem_base + (mb < 2 ? mb * NT : 4 * NT + (mb - 2) * NT);
                for (int kb = 0; kb < G2_KB; ++kb) {
                    if (mb == 0 && (kb == 0 || kb == 2)) {
                        mbar_wait(&zready[kb >> 1], par);
                        tc_fence_after();
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_addr = smem_u32(wst + stage * WT_BYTES);
                        const uint32_t b_addr = smem_u32(zs + kb * ZS_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma_ss(d_tmem, make_sw128_kmajor_desc(a_addr + k * (UK * 2)), make_sw128_kmajor_desc(b_addr + k * (UK * 2)), idesc,
                                    (mb < 2 ? 0 : l) | kb | k);
                        umma_commit_mcast(&empty[stage], MASK);
                        if (kb == G2_KB - 1 && mb == 1) umma_commit(&accb[2]);
                        if (kb == G2_KB - 1 && mb == 3 && l == p.L - 1) umma_commit(skipb);
                    }
                    __syncwarp();
                    if (kb == G2_KB - 1 && (mb & 1)) TLOGT(8 + (mb >> 1));   // G2 residual / skip blocks issued (8, 9)
                    if (++stage == NSTG) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue: 8 warps; warp -> TMEM lane quarter q = warp % 4 (channels), column half `sub` ==========
        const int e = warp - 2, q = warp & 3, sub = e >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int cl = q * 32 + lane;                           // channel within a 128-channel block
        const int col0 = sub * (NT / 2);                        // first frame (column) of this warp
        const float inv_sqrt2 = 0.70710678118654752440f;
        const long long row0 = (long long)b * p.T + t0 + col0;  // global frame row of column col0
        uint32_t vmask[(NT / 2 + 31) / 32];                     // valid frames of this warp's columns
#pragma unroll
        for (int w = 0; w < (NT / 2 + 31) / 32; ++w) vmask[w] = 0;
#pragma unroll
        for (int i = 0; i < NT / 2; ++i)
            if (tile_ok && t0 + col0 + i < p.T) vmask[i >> 5] |= 1u << (i & 31);
        auto ok = [&](int i) -> bool { return (vmask[i >> 5] >> (i & 31)) & 1u; };

        // the residual stream of this thread: channels cl and 128 + cl, frames col0 .. col0 + NT/2
        float xr[2][NT / 2];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int i = 0; i < NT / 2; ++i) xr[mb][i] = ok(i) ? __ldg(p.x + (row0 + i) * C + mb * BM + cl) : 0.f;

        // z-tile address pieces of this thread's channel (K slab of 64 channels, 16-byte chunk, byte within the chunk)
        const int cw = (q & 1) * 32 + lane;
        const uint32_t zbase = smem_u32(zs) + (q >> 1) * ZS_BYTES + (cw & 7) * 2;
        const int zchunk = cw >> 3;

#pragma unroll 1
        for (int l = 0; l < p.L; ++l) {
            const uint32_t par = l & 1;
            const bool last = l == p.L - 1;
            // per-channel scalars of the layer (L2 round trips: issued now, consumed in EPI2)
            const float* bo = p.bo + (long long)l * 2 * C;
            const float bres0 = __ldg(bo + cl), bres1 = __ldg(bo + BM + cl);
            const float* dnext = p.dvec + (long long)(l + 1) * C + (long long)(tile_ok ? b : 0) * p.d_stride;
            const float dn0 = last ? 0.f : __ldg(dnext + cl), dn1 = last ? 0.f : __ldg(dnext + BM + cl);
            // ---- EPI1: z = sigmoid(gate + cond) * tanh(filter + cond) -> K-major swizzled z tile ----
            const uint16_t* ctile = reinterpret_cast<const uint16_t*>(p.cond) + l * p.cond_lstride + (long long)blockIdx.x * (NT * 2 * C);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint4 cg[NCH], cf[NCH];
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    const uint16_t* cp = ctile + ((long long)(sub * NCH + ch) * (2 * C) + h * C + cl) * 8;
                    if (tile_ok) {
                        cg[ch] = ldg_nc_u4(cp);
                        cf[ch] = ldg_nc_u4(cp + BM * 8);
                    } else {
                        cg[ch] = make_uint4(0, 0, 0, 0);
                        cf[ch] = make_uint4(0, 0, 0, 0);
                    }
                }
                mbar_wait(&accb[h], par);
                tc_fence_after();
                if (e == 0) TLOGT(10 + h);                          // G1 half h complete (10, 11)
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    float g[8], f[8];
                    tmem_ld8(taddr + (2 * h) * NT + col0 + 8 * ch, g);
                    tmem_ld8(taddr + (2 * h + 1) * NT + col0 + 8 * ch, f);
                    tmem_ld_wait();
                    const uint32_t* gw = reinterpret_cast<const uint32_t*>(&cg[ch]);
                    const uint32_t* fw = reinterpret_cast<const uint32_t*>(&cf[ch]);
                    const uint32_t zrow = zbase + h * 2 * ZS_BYTES + (sub * NCH + ch) * 1024;
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2) {
                        const float2 cga = Half16<BF16>::unpack2(gw[i2]), cfa = Half16<BF16>::unpack2(fw[i2]);
                        const float z0 = sigmoid_fast(g[2 * i2] + cga.x) * tanh_fast(f[2 * i2] + cfa.x);
                        const float z1 = sigmoid_fast(g[2 * i2 + 1] + cga.y) * tanh_fast(f[2 * i2 + 1] + cfa.y);
                        const int i0 = 2 * i2, i1 = 2 * i2 + 1;         // frame within the 8-row swizzle atom
                        st_shared_u16(zrow + i0 * 128 + ((zchunk ^ i0) << 4), ok(8 * ch + i0) ? half1<BF16>(z0) : (uint16_t)0);
                        st_shared_u16(zrow + i1 * 128 + ((zchunk ^ i1) << 4), ok(8 * ch + i1) ? half1<BF16>(z1) : (uint16_t)0);
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&zready[h]);
                if (e == 0) TLOGT(12 + h);                          // EPI1 half h done (12, 13)
            }
            // ---- EPI2: x <- (x + residual + bias) / sqrt(2) in registers; y_{l+1} = x + d_{l+1} -> global (16-bit) ----
            uint16_t* ynext = last ? nullptr : reinterpret_cast<uint16_t*>(p.ybuf[(l + 1) & 1]) + row0 * C + cl;
            mbar_wait(&accb[2], par);
            tc_fence_after();
            if (e == 0) TLOGT(14);                                  // G2 residual complete
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const float bias = mb ? bres1 : bres0, dn = mb ? dn1 : dn0;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    float acc[8];
                    tmem_ld8(taddr + mb * NT + col0 + 8 * ch, acc);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float xn = (xr[mb][8 * ch + i] + acc[i] + bias) * inv_sqrt2;
                        xr[mb][8 * ch + i] = xn;
                        if (ynext && ok(8 * ch + i)) ynext[(long long)(8 * ch + i) * C + mb * BM] = half1<BF16>(xn + dn);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tfree);
            // hand-off: every epilogue thread has issued its y stores -> CTA barrier -> one thread publishes the tile flag
            named_bar_sync(1, EPI_WARPS * 32);
            if (e == 0 && lane == 0 && tile_ok) {
                __threadfence();
                st_release_gpu(p.flags + blockIdx.x, l + 1);
            }
            if (e == 0) TLOGT(15);                                  // EPI2 done, flag released
        }
        // ---- final: skip sum (accumulated over all layers in TMEM) + the summed skip biases -> 16-bit [rows, C] ----
        float bsum0 = 0.f, bsum1 = 0.f;
        for (int l = 0; l < p.L; ++l) {
            bsum0 += __ldg(p.bo + (long long)l * 2 * C + C + cl);
            bsum1 += __ldg(p.bo + (long long)l * 2 * C + C + BM + cl);
        }
        uint16_t* sk = reinterpret_cast<uint16_t*>(p.skip_h) + row0 * C + cl;
        mbar_wait(skipb, 0);
        tc_fence_after();
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                float acc[8];
                tmem_ld8(taddr + 4 * NT + mb * NT + col0 + 8 * ch, acc);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (ok(8 * ch + i)) sk[(long long)(8 * ch + i) * C + mb * BM] = half1<BF16>(acc[i] + (mb ? bsum1 : bsum0));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int NT, int BF16>
static int launch_t(const StackTP& p, int grid, int smem_bytes, cudaStream_t st) {
    static PerDevice configured;
    if (configured.first()) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(wavenet_stack_t_kernel<NT, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem_bytes;
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
    B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg, wavenet_stack_t_kernel<NT, BF16>, p));
    return B2S_OK;
}

// [L][rows][n2] (layer-major, frame rows) -> [L][tile][NT/8][n2][8 frames]: what an epilogue thread (= one channel) of the
// transposed kernel reads with one 16-byte load per 8 frames, the 32 lanes of a warp contiguous.
__global__ void __launch_bounds__(256) cond_retile_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int B, int T,
                                                          int n2, int NT, int tiles_per_b, int n_tiles) {
    const int l = blockIdx.y;
    const int unit = blockIdx.x;                          // (tile, 8-frame chunk)
    const int cpt = NT / 8;
    const int tile = unit / cpt, c8 = unit - tile * cpt;
    const int b = tile / tiles_per_b, ti = tile - b * tiles_per_b;
    const int tb = ti * NT + c8 * 8;
    const long long rows = (long long)B * T;
    for (int n = threadIdx.x * 2; n < n2; n += blockDim.x * 2) {
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = tb + i;
            v[i] = (b < B && t < T) ? __ldg(reinterpret_cast<const uint32_t*>(src + ((long long)l * rows + (long long)b * T + t) * n2 + n)) : 0u;
        }
        uint4 lo, hi;                                     // channel n: low halves; channel n + 1: high halves
        lo.x = (v[0] & 0xFFFFu) | (v[1] << 16);
        lo.y = (v[2] & 0xFFFFu) | (v[3] << 16);
        lo.z = (v[4] & 0xFFFFu) | (v[5] << 16);
        lo.w = (v[6] & 0xFFFFu) | (v[7] << 16);
        hi.x = (v[0] >> 16) | (v[1] & 0xFFFF0000u);
        hi.y = (v[2] >> 16) | (v[3] & 0xFFFF0000u);
        hi.z = (v[4] >> 16) | (v[5] & 0xFFFF0000u);
        hi.w = (v[6] >> 16) | (v[7] & 0xFFFF0000u);
        uint4* o = reinterpret_cast<uint4*>(dst + ((((long long)l * n_tiles + tile) * cpt + c8) * n2 + n) * 8);
        o[0] = lo;
        o[1] = hi;
    }
}

}  // namespace wt
}  // namespace tc
}  // namespace b2s

using namespace b2s;
using namespace b2s::tc;

extern unsigned long long* g_tlog;      // b2s_tc_wavenet.cu (b2s_debug_set_stack_tlog)

static bool nt_supported(int NT) { return NT == 32 || NT == 48 || NT == 64 || NT == 80; }

extern "C" int b2s_tc_wavenet_stack_t_tiles(int B, int T, int NT) {
    if (B <= 0 || T <= 0 || NT <= 0) return 0;
    const int n = B * ceil_div(T, NT);
    return (n + 1) & ~1;
}

extern "C" int b2s_tc_cond_retile(const void* table_h, int L, int B, int T, int n2, int NT, void* out_h, void* stream) {
    B2S_CHECK_ARG(table_h && out_h, "b2s_tc_cond_retile: null pointer");
    B2S_CHECK_ARG(nt_supported(NT) && n2 > 0 && n2 % 2 == 0 && L >= 1, "b2s_tc_cond_retile: unsupported frame tile %d / width %d", NT, n2);
    B2S_CHECK_ARG(al16(table_h) && al16(out_h), "b2s_tc_cond_retile: misaligned pointer");
    if (B * T == 0) return B2S_OK;
    const int tpb = ceil_div(T, NT), n_tiles = b2s_tc_wavenet_stack_t_tiles(B, T, NT);
    dim3 grid(n_tiles * (NT / 8), L);
    wt::cond_retile_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)table_h, (uint16_t*)out_h, B, T, n2, NT, tpb, n_tiles);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

extern "C" int b2s_tc_wavenet_stack_t(void* y0_h, void* y1_h, const void* Wd_h, const void* cond_t, const void* Wo_h, const float* bo,
                                      const float* x, void* skip_h, const float* dvec, int d_stride, const int* dilations_host, int L,
                                      int B, int T, int C, int NT, int* flags, int bf16, void* stream) {
    B2S_CHECK_ARG(y0_h && y1_h && Wd_h && cond_t && Wo_h && bo && x && skip_h && dvec && dilations_host && flags,
                  "b2s_tc_wavenet_stack_t: null pointer");
    if (C != wt::C) {
        set_error("b2s_tc_wavenet_stack_t: specialised for %d residual channels (got %d)", wt::C, C);
        return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_ARG(L >= 1 && L <= wt::MAXL, "b2s_tc_wavenet_stack_t: 1 <= L <= %d (got %d)", wt::MAXL, L);
    B2S_CHECK_ARG(nt_supported(NT), "b2s_tc_wavenet_stack_t: frame tile must be 32, 48, 64 or 80 (got %d)", NT);
    B2S_CHECK_ARG(y0_h != y1_h && d_stride % 4 == 0, "b2s_tc_wavenet_stack_t: y buffers must differ; d_stride must keep 16B alignment");
    B2S_CHECK_ARG(al16(y0_h) && al16(y1_h) && al16(Wd_h) && al16(Wo_h) && al16(cond_t) && al16(bo) && al16(x) && al16(skip_h) && al16(dvec),
                  "b2s_tc_wavenet_stack_t: misaligned pointer");
    if (B * T == 0) return B2S_OK;
    wt::StackTP p{};
    p.tiles_per_b = ceil_div(T, NT);
    const int grid = b2s_tc_wavenet_stack_t_tiles(B, T, NT);
    if (grid > num_sms()) {
        set_error("b2s_tc_wavenet_stack_t: %d tiles do not fit the %d SMs at once (every tile must be resident)", grid, num_sms());
        return B2S_ERR_UNSUPPORTED;
    }
    int dmax = 8;
    for (int l = 0; l < L; ++l) {
        B2S_CHECK_ARG(dilations_host[l] >= 1 && dilations_host[l] <= 16 && dilations_host[l] <= NT,
                      "b2s_tc_wavenet_stack_t: dilation %d not in [1, min(16, frame tile)]", dilations_host[l]);
        p.dil[l] = dilations_host[l];
        if (dilations_host[l] > dmax) dmax = 16;
    }
    p.dmax = dmax;
    p.R = NT + 2 * dmax;
    int rc = make_map_act(&p.mapY[0], y0_h, bf16, C, C, T, B, wt::BK, p.R);
    if (rc) return rc;
    rc = make_map_act(&p.mapY[1], y1_h, bf16, C, C, T, B, wt::BK, p.R);
    if (rc) return rc;
    rc = make_map_w3(&p.mapWd, Wd_h, bf16, 3 * C, 2 * C, L, wt::BK, wt::BM / 2);
    if (rc) return rc;
    rc = make_map_w3(&p.mapWo, Wo_h, bf16, C, 2 * C, L, wt::BK, wt::BM / 2);
    if (rc) return rc;
    p.B = B; p.T = T; p.L = L;
    p.cond = cond_t; p.cond_lstride = (long long)grid * NT * 2 * C; p.bo = bo;
    p.x = x; p.ybuf[0] = y0_h; p.ybuf[1] = y1_h; p.skip_h = skip_h;
    p.dvec = dvec; p.d_stride = d_stride; p.flags = flags;
    static const int dbg = getenv("B2S_STACKT_DBG") ? atoi(getenv("B2S_STACKT_DBG")) : 0;
    p.dbg = dbg;
    p.tlog = g_tlog;
    const int smem = 4 * p.R * 128 + 4 * NT * 128 + wt::NSTG * wt::WT_BYTES + 256;
    cudaStream_t st = (cudaStream_t)stream;
#define B2S_LAUNCH_T(N)                                                                                  \
    case N: return bf16 ? wt::launch_t<N, 1>(p, grid, smem, st) : wt::launch_t<N, 0>(p, grid, smem, st);
    switch (NT) {
        B2S_LAUNCH_T(32)
        B2S_LAUNCH_T(48)
        B2S_LAUNCH_T(64)
        B2S_LAUNCH_T(80)
    }
#undef B2S_LAUNCH_T
    return B2S_ERR_UNSUPPORTED;
}

#endif  // B2S_EXPERIMENTS
