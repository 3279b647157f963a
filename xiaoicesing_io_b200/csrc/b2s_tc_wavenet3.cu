// Whole WaveNet residual stack (wavenet.py:86-96: stem + all L residual layers) as ONE persistent kernel, third design.
// sm_100a only.  What changed against b2s_tc_wavenet.cu (namespace ws), and why:
//
//   * cta_group::2: a cluster = two neighbouring 128-frame tiles = ONE 256-row MMA.  Each CTA streams only HALF of every
//     weight tile (SM ingress, ~64 B/clk, was the limiter of GEMM1: 48 KB per 512 MMA cycles).
//   * the layer input y = x + step embedding is RESIDENT in shared memory: 4 K slabs of (16 + 128 + 16) rows.  The three
//     conv taps are row-shifted SWIZZLE_128B descriptors into the same slabs (the swizzle is a function of the absolute
//     shared-memory address, so a window may start at any row); y is fetched 0 times instead of 6 times per layer.  Only the
//     16 edge rows per side travel through global memory (TMA store -> release flag -> neighbour's TMA load, zero fill
//     outside [0, T) = the conv's zero padding).
//   * the fp32 residual stream never leaves TMEM: X_l = 2^(l/2) x_l - (bias terms) is the accumulator of the residual half of
//     the output projection, whose weights carry the factor 2^(l/2) (x_{l+1} = (x_l + r_l)/sqrt2  <=>  X_{l+1} = X_l + 2^(l/2) r_l).
//     The epilogue that produces the next layer's input is  y_{l+1} = 2^(-(l+1)/2) X_{l+1} + e_{l+1}[c]  (one FMA per element,
//     16-bit, straight into the swizzled shared-memory tile) - no global fp32 read-modify-write, no shared-memory transpose.
//   * the skip half of the output projection is deferred: the gated tile z_l (already in shared memory in the layout the tensor
//     core reads) is TMA-stored to z_all[l]; sum_l z_l Wskip_l^T is ONE K = L*C GEMM afterwards (b2s_tc_skip_sum), accumulated
//     in TMEM over all layers (north_star: "skip outputs accumulate on-chip across layers").
//   * GEMM1 runs in four N = 128 quarters through two 128-column TMEM buffers, so the gate epilogue of quarter q overlaps the
//     MMAs of quarter q+1; the residual GEMM's K slabs are interleaved as soon as their z slab exists.
//
// Warp roles (384 threads): 0 weight-tile TMA producer, 1 MMA issuer (leader CTA only), 2 TMEM allocator + step-embedding
// table stager, 3 IO (halo loads, edge / z TMA stores, tile flags), 4-11 epilogue (two warps per TMEM lane quarter).
#include "b2s_tc.cuh"

#include <math.h>
#include <stdlib.h>

namespace b2s {
namespace tc {
namespace ws3 {

constexpr int C = 256, MAXL = 32;
constexpr int BM = 128, BK = 64, UK = 16, CLUSTER = 2;
constexpr int HALO = 16, YR = BM + 2 * HALO;            // 160 rows per resident y slab
constexpr int YSLAB = YR * 128;                          // 20480 B: one K slab (64 channels) of the y tile incl. halo
constexpr int Y_BYTES = 4 * YSLAB;                       // 81920
constexpr int ZSLAB = BM * 128;                          // 16384
constexpr int Z_BYTES = 2 * ZSLAB;                       // 32768: a RING of two z K slabs (slab s lives in buffer s & 1)
constexpr int STAGES = 7, STAGE_BYTES = 16384;           // weight ring: [128 rows x 64 k] per CTA = half of a 256-column tile
constexpr int E_BYTES = 2 * C * 4;                       // two step-embedding rows e_m[c]
constexpr int NBARS = 38, BAR_BYTES = 320;
constexpr int SMEM_BYTES = Y_BYTES + Z_BYTES + STAGES * STAGE_BYTES + E_BYTES + BAR_BYTES;   // 231744
constexpr int NTHREADS = 384, EPI_WARPS = 8;
static_assert(NBARS * 8 + 4 <= BAR_BYTES, "barrier block");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

struct __align__(64) Stack3P {
    CUtensorMap mapXin, mapWin, mapWd, mapWres, mapYe[2], mapZ;
    int B, T, tiles_per_b, L, MF, kb_in;
    int dil[MAXL];
    const void* cond; long long cond_lstride;             // tile/chunk-major table of b2s_tc_cond_table_tiled; elements per layer
    const void* wd; const void* wres;                     // the packed weights again as plain pointers (L2 prefetch of the next layer)
    const float* b_in;                                    // [C]
    const float* bsum;                                    // [L][C]: sum_{k<m} 2^(k/2) b_res,k[c]
    const float* dvec; int d_stride;                      // step embedding of layer m at dvec + b*d_stride + m*C
    int* flags;                                           // [B * tiles_per_b], zero before the launch
    const int* lens;                                      // optional [B]: valid frames of utterance b (<= T); rows beyond are the conv's zero padding
    unsigned long long* tlog;                             // optional phase timestamps (B2S_TLOG builds)
    int dbg;                                              // B2S_TLOG builds only: 1 = no weight loads, 2 = no gate epilogue math (timing experiments, wrong results)
    // ---- fused skip sum + head (wavenet.py:96-99) on the CTAs behind the layer tiles (fuse_head = 1): blocks [n_layer_ctas, grid)
    //      accumulate S = sum_l z_l Wskip_l^T in TMEM as the z tiles are published (zflags[tile] = completed half-tile stores), then
    //      out = W_fin relu(W_sp (S + bss) / sqrt(L) + b_sp) + b_fin
    int n_layer_ctas, fuse_head;
    // several utterance groups of ONE evaluation launched back to back (bit 0: this launch follows another group, bit 1: another
    // group follows): the next group's layer kernel must not wait for this group's skip / head tail - see wavenet_stack3_kernel
    int chain;
    // NARROW MODELS (zero-padded to 256 channels, engine.py): only the first ks = 3 of the 4 64-channel K slabs are non-zero, so the
    // K slabs >= ks of every GEMM are skipped and GEMM1's second half (channels 128 .. 191) is issued with N = 128; the barrier
    // choreography of the dead slab is kept (arrivals without work), so every phase count is the one of the full kernel.
    // Skipped terms are exact zeros: the results are bit-identical to the padded run.
    int ks;
    CUtensorMap mapWskip, mapWsp, mapWfin;
    const float* bss; const float* b_sp; const float* b_fin; float alpha; float* out;
    int* zflags;                                          // [B * tiles_per_b], zero before the launch
};

#ifdef B2S_TLOG
#define TLOG3(l, slot) do { if (p.tlog && blockIdx.x == 2) p.tlog[(l) * 16 + (slot)] = globaltimer_ns(); } while (0)
#else
#define TLOG3(l, slot) do { } while (0)
#endif

__device__ __forceinline__ uint4 ldg_nc_u4(const void* q) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(q));
    return r;
}
__device__ __forceinline__ void st_shared_u4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// registers -> TMEM: 32 lanes x 32 consecutive fp32 columns (the mirror image of tmem_ld32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t smem_addr, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d_a(const CUtensorMap* m, uint32_t smem_addr, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_addr), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d_cg2_a(uint32_t smem_addr, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_addr),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2_a(uint32_t smem_addr, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_addr),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2_a(uint32_t smem_addr, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_addr),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
// upper 32 bits of the K-major SWIZZLE_128B matrix descriptor (stride 1024 B between 8-row groups, version 1, layout 2); the
// lower word is (address >> 4) | leading-offset 1
constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void mma2(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    const uint64_t ad = ((uint64_t)DESC_HI << 32) | a_lo, bd = ((uint64_t)DESC_HI << 32) | b_lo;
    umma_ss_cg2(tmem_d, ad, bd, idesc, accumulate);
}
// remote mbarrier arrive in the form CUTLASS' ClusterBarrier::arrive uses (default semantics).  The .release.cluster form
// makes the warp wait for ALL its outstanding global loads (the cond prefetch): measured 1-1.5 us per epilogue step.
__device__ __forceinline__ void arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// 16 bytes into ANOTHER CTA's shared memory; the store completes 16 bytes of the transaction count of an mbarrier of that CTA
// (no fence, no release: the consumer that waits on the barrier sees the data)
__device__ __forceinline__ void st_async_u4(uint32_t cluster_addr, uint4 v, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_bar)
                 : "memory");
}
// wait with cluster-scope acquire: the data guarded by the barrier was written by another CTA of the cluster (DSMEM stores)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    const uint64_t t0 = globaltimer_ns();
    uint32_t spin = 0;
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
        if ((++spin & 1023u) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
            printf("b2s: stack3 cluster mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void wait_flag3(const int* f, int want) {
    if (ld_acquire_gpu(f) >= want) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spin = 0;
    while (ld_acquire_gpu(f) < want) {
        if ((++spin & 255u) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
            printf("b2s: stack3 tile flag timeout (block %d want %d have %d)\n", blockIdx.x, want, ld_acquire_gpu(f));
            __trap();
        }
    }
}

// =====================================================================================================================
// Skip-sum / head role (wavenet.py:96-99) of the CTAs behind the layer tiles.  A pair owns FOUR tiles: slot s in {0,1} of CTA r is
// tile 4*pair + 2s + r; one cta_group::2 MMA covers the two tiles of a slot.  S_s = sum_l z_l Wskip_l^T accumulates in TMEM columns
// [256 s, +256) over all layers (z tiles arrive through zflags), then per slot: (S + bss) -> 16-bit tile in shared memory ->
// skip_projection -> relu(alpha acc + b_sp) -> 16-bit tile -> output_projection -> + b_fin -> out (fp32).
// Shared memory: [0, 128K) one head operand tile (4 swizzled K slabs) per slot, then 3 stages of {A 16 KB, B 16 KB}.  The head of
// the two slots is software-pipelined: while the tensor core runs slot 0's skip_projection the epilogue warps convert slot 1's
// skip sum, and so on (the head is the tail of every denoiser evaluation: nothing else runs beside it).
// A kernel of its own (clusters of 2, any number of pairs, no co-residency requirement) launched behind the layer kernel with
// programmatic dependent launch: it starts on the SMs the layer tiles leave idle (52 of 148 at 16 x 690 frames) once every layer
// CTA is resident, and follows the layer kernel through the z-tile flags.
// =====================================================================================================================
constexpr int SK_STAGES = 3, SK_STAGE_BYTES = 32768, SK_HTILE = 65536, SK_HBUF = 2 * SK_HTILE;      // one head operand tile per slot

constexpr int SK_SMEM_BYTES = SK_HBUF + SK_STAGES * SK_STAGE_BYTES + 256;

template <int BF16, int KS>
__global__ void __launch_bounds__(NTHREADS, 1) wavenet_skiphead3_kernel(const __grid_constant__ Stack3P p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + SK_HBUF + SK_STAGES * SK_STAGE_BYTES);
    uint64_t* empty = full + SK_STAGES;
    uint64_t* sfull = empty + SK_STAGES;       // [2] skip sum of slot s complete          (commit, both CTAs)
    uint64_t* hready = sfull + 2;              // [2 s + i] head operand tile i (0: skip sum, 1: hidden) of slot s written, both CTAs (leader, 16)
    uint64_t* hfull0 = hready + 4;             // [2] skip_projection accumulator of slot s complete (commit, both CTAs)
    uint64_t* hfull1 = hfull0 + 2;             // [2] output_projection accumulator of slot s complete
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(hfull1 + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank() & 1u, lead = 0;
    const uint16_t pmask = 3;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.mapZ);
        prefetch_tmap(&p.mapWskip);
        prefetch_tmap(&p.mapWsp);
        prefetch_tmap(&p.mapWfin);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SK_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sfull[i], 1);
            mbar_init(&hfull0[i], 1);
            mbar_init(&hfull1[i], 1);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&hready[i], EPI_WARPS * CLUSTER);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_cg2(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
#ifdef B2S_TLOG
    if (p.tlog && blockIdx.x == 0 && threadIdx.x == 0) p.tlog[MAXL * 16 + 4] = globaltimer_ns();
#endif
    // NO griddepcontrol.wait here: this kernel is released while the layer kernel runs and follows its z tiles through zflags
    // (release / acquire); everything else it reads (weights, biases) was written before the layer kernel was launched.
    // Chained groups (bit 1): the next group's layer kernel is released now, to be scheduled as this group's layer CTAs exit.
    if (p.chain & 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int L = p.L, MF = p.MF;
    const int pairidx = (int)blockIdx.x >> 1;
    int gt[2], tb[2], tt0[2];
    bool ok[2], okp[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        gt[s] = 4 * pairidx + 2 * s + (int)rank;
        const int gp = 4 * pairidx + 2 * s + 1 - (int)rank;
        tb[s] = gt[s] / p.tiles_per_b;
        tt0[s] = (gt[s] - tb[s] * p.tiles_per_b) * BM;
        ok[s] = gt[s] < p.n_layer_ctas && tt0[s] < p.T;
        okp[s] = gp < p.n_layer_ctas && (gp % p.tiles_per_b) * BM < p.T;
    }
    const uint32_t hb_a = smem_u32(smem), st_a = hb_a + SK_HBUF;
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [&] { if (++stage == SK_STAGES) { stage = 0; phase ^= 1; } };
            for (int l = 0; l < L; ++l)
                for (int half = 0; half < 2; ++half)
                    for (int s = 0; s < 2; ++s)
                        for (int kk = 0; kk < 2; ++kk) {
                            const int kb = 2 * half + kk;
                            if (KS < 4 && kb >= KS) continue;        // narrow model: this K slab is all zeros
                            mbar_wait(&empty[stage], phase ^ 1);
                            const uint32_t lb = mapa_u32(&full[stage], lead);
                            if (rank == 0) mbar_expect_tx(&full[stage], 2 * 16384 + (ok[s] ? 16384 : 0) + (okp[s] ? 16384 : 0));
                            tma_load_3d_cg2_a(st_a + stage * SK_STAGE_BYTES + 16384, &p.mapWskip, lb, kb * BK, rank * (C / 2), l);
                            if (ok[s]) {
                                if (kk == 0) {
                                    wait_flag3(p.zflags + gt[s], 2 * l + half + 1);
                                    fence_proxy_async_all();
                                }
                                tma_load_4d_cg2_a(st_a + stage * SK_STAGE_BYTES, &p.mapZ, lb, kb * BK, tt0[s], tb[s], l);
                            }
                            advance();
                        }
            for (int s = 0; s < 2; ++s)
                for (int kb = 0; kb < KS; ++kb) {            // skip_projection weights: this CTA's 128 of the 256 rows
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&full[stage], 2 * 16384);
                    tma_load_2d_cg2_a(st_a + stage * SK_STAGE_BYTES + 16384, &p.mapWsp, mapa_u32(&full[stage], lead), kb * BK, rank * (C / 2));
                    advance();
                }
            for (int s = 0; s < 2; ++s)
                for (int kb = 0; kb < KS; ++kb) {            // output_projection weights: this CTA's MF/2 of the MF rows
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&full[stage], 2 * (MF / 2) * 128);
                    tma_load_2d_cg2_a(st_a + stage * SK_STAGE_BYTES + 16384, &p.mapWfin, mapa_u32(&full[stage], lead), kb * BK, rank * (MF / 2));
                    advance();
                }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {
            const uint32_t idesc_h = make_idesc_f16(2 * BM, 256, BF16), idesc_o = make_idesc_f16(2 * BM, MF, BF16);
            const uint32_t st_lo = desc_lo(st_a), hb_lo = desc_lo(hb_a);
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [&] { if (++stage == SK_STAGES) { stage = 0; phase ^= 1; } };
            for (int l = 0; l < L; ++l)
                for (int half = 0; half < 2; ++half)
                    for (int s = 0; s < 2; ++s) {
                        for (int kk = 0; kk < 2; ++kk) {
                            if (KS < 4 && 2 * half + kk >= KS) continue;
                            mbar_wait(&full[stage], phase);
                            tc_fence_after();
                            const uint32_t a_lo = st_lo + stage * (SK_STAGE_BYTES >> 4), b_lo = a_lo + (16384 >> 4);
#pragma unroll
                            for (int k = 0; k < BK / UK; ++k) mma2(tmem_base + s * 256, a_lo + 2 * k, b_lo + 2 * k, idesc_h, (l | half | kk | k) != 0);
                            umma_commit_cg2_mcast(&empty[stage], pmask);
                            advance();
                        }
                        if (l == L - 1 && half == 1) umma_commit_cg2_mcast(&sfull[s], pmask);
#ifdef B2S_TLOG
                        if (p.tlog && blockIdx.x == 0 && half == 1 && s == 1) p.tlog[l * 16 + 9] = globaltimer_ns();
#endif
                    }
            for (int s = 0; s < 2; ++s) {
                mbar_wait(&hready[2 * s], 0);                  // (S + bss) tile of slot s, both CTAs, is in shared memory
                tc_fence_after();
                for (int kb = 0; kb < KS; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = hb_lo + s * (SK_HTILE >> 4) + kb * (ZSLAB >> 4), b_lo = st_lo + stage * (SK_STAGE_BYTES >> 4) + (16384 >> 4);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) mma2(tmem_base + s * 256, a_lo + 2 * k, b_lo + 2 * k, idesc_h, (kb | k) != 0);
                    umma_commit_cg2_mcast(&empty[stage], pmask);
                    advance();
                }
                umma_commit_cg2_mcast(&hfull0[s], pmask);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_wait(&hready[2 * s + 1], 0);              // hidden tile of slot s
                tc_fence_after();
                for (int kb = 0; kb < KS; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = hb_lo + s * (SK_HTILE >> 4) + kb * (ZSLAB >> 4), b_lo = st_lo + stage * (SK_STAGE_BYTES >> 4) + (16384 >> 4);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) mma2(tmem_base + s * 256, a_lo + 2 * k, b_lo + 2 * k, idesc_o, (kb | k) != 0);
                    umma_commit_cg2_mcast(&empty[stage], pmask);
                    advance();
                }
                umma_commit_cg2_mcast(&hfull1[s], pmask);
            }
        }
    } else if (warp >= 4) {
        const int e = warp - 4, qd = e & 3, sub = e >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
        const int row = qd * 32 + lane, sw = row & 7;
        const uint32_t hrow = hb_a + (row >> 3) * 1024 + sw * 128;
        const uint32_t lr = mapa_u32(hready, lead);
        bool valid[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) valid[s] = ok[s] && tt0[s] + row < (p.lens ? __ldg(p.lens + tb[s]) : p.T);
        // 32 columns of slot s's accumulator -> f(acc, bias) -> 16-bit -> the slot's swizzled operand tile
        auto to_tile = [&](int s, const float* bias, float alpha, bool relu) {
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {
                const int J = 4 * sub + jj;
                float acc[32];
                tmem_ld32(taddr + s * 256 + 32 * J, acc);
                tmem_ld_wait();
                uint32_t hp[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + 32 * J + 4 * i));
                    float v0 = fmaf(alpha, acc[4 * i], bv.x), v1 = fmaf(alpha, acc[4 * i + 1], bv.y);
                    float v2 = fmaf(alpha, acc[4 * i + 2], bv.z), v3 = fmaf(alpha, acc[4 * i + 3], bv.w);
                    if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
                    hp[2 * i] = valid[s] ? Half16<BF16>::pack2(v0, v1) : 0u;
                    hp[2 * i + 1] = valid[s] ? Half16<BF16>::pack2(v2, v3) : 0u;
                }
                const uint32_t slab = hrow + s * SK_HTILE + (J >> 1) * ZSLAB;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4)
                    st_shared_u4(slab + (((4 * (J & 1) + c4) ^ sw) << 4), make_uint4(hp[4 * c4], hp[4 * c4 + 1], hp[4 * c4 + 2], hp[4 * c4 + 3]));
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
        };
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            mbar_wait(&sfull[s], 0);
            tc_fence_after();
            to_tile(s, p.bss, 1.0f, false);                    // the skip sum incl. its summed biases (wavenet.py:96, before the 1/sqrt(L))
            if (lane == 0) arrive_remote(lr + 16 * s);
        }
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            mbar_wait(&hfull0[s], 0);
            tc_fence_after();
            to_tile(s, p.b_sp, p.alpha, true);                 // skip_projection + ReLU (wavenet.py:97-98)
            if (lane == 0) arrive_remote(lr + 16 * s + 8);
        }
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            mbar_wait(&hfull1[s], 0);
            tc_fence_after();
            float* orow = p.out + ((long long)tb[s] * p.T + tt0[s] + row) * MF;
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {                   // output_projection (wavenet.py:99): fp32 rows of MF values
                const int J = 2 * jj + sub;
                if (32 * J >= MF) break;
                float acc[32];
                tmem_ld32(taddr + s * 256 + 32 * J, acc);
                tmem_ld_wait();
                if (valid[s]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int col = 32 * J + 4 * i;
                        if (col < MF) {
                            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.b_fin + col));
                            *reinterpret_cast<float4*>(orow + col) = make_float4(acc[4 * i] + bv.x, acc[4 * i + 1] + bv.y, acc[4 * i + 2] + bv.z, acc[4 * i + 3] + bv.w);
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
#ifdef B2S_TLOG
    if (p.tlog && blockIdx.x == 0 && threadIdx.x == 0) p.tlog[MAXL * 16 + 5] = globaltimer_ns();
#endif
    asm volatile("griddepcontrol.wait;" ::: "memory");      // completion of this kernel implies completion of its layer kernel (chained groups)
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, 512);
    }
}

template <int BF16, int KS>
__global__ void __launch_bounds__(NTHREADS, 1) wavenet_stack3_kernel(const __grid_constant__ Stack3P p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ys = smem;
    uint8_t* zs = smem + Y_BYTES;
    uint8_t* stages = zs + Z_BYTES;
    float* etab = reinterpret_cast<float*>(stages + STAGES * STAGE_BYTES);      // [2][C]
    uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(etab) + E_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* accfull = empty + STAGES;        //     GEMM1 half accumulator complete               (commit, both CTAs)
    uint64_t* accfree = accfull + 2;           //     ... drained by both CTAs' epilogues           (leader, 16 arrivals)
    uint64_t* zready = accfree + 2;            // [4] z K slab written in both CTAs                 (leader, 16)
    uint64_t* yready = zready + 4;             //     own rows of y_m in shared memory, both CTAs   (leader, 16)
    uint64_t* halofull = yready + 1;           //     halo rows of y_m landed in both CTAs          (leader, tx)
    uint64_t* g1done = halofull + 1;           //     every GEMM1 MMA of the layer retired          (commit, both CTAs)
    uint64_t* xfull = g1done + 1;              //     residual GEMM of the layer retired            (commit, both CTAs)
    uint64_t* ydone = xfull + 1;               //     this CTA's epilogue wrote y_m                 (local, 8)
    uint64_t* zdone = ydone + 1;               //     this CTA's epilogue wrote z_l                 (local, 8)
    uint64_t* eready = zdone + 1;              // [2] step-embedding row e_m staged                 (local, 1)
    uint64_t* efree = eready + 2;              // [2] ... consumed                                  (local, 8)
    uint64_t* stemfull = efree + 2;            //     stem accumulator complete                     (commit, both CTAs)
    uint64_t* xinfull = stemfull + 1;          //     x_in tiles of both CTAs landed                (leader, tx)
    uint64_t* haloready = xinfull + 1;         //     the peer CTA's halo rows of y_m have landed       (leader, 1: the peer's relay)
    uint64_t* haloin = haloready + 1;          //     THIS CTA's halo rows written by its cluster neighbours (local, tx bytes of st.async)
    uint64_t* zfree = haloin + 1;              // [2] z ring buffer k may be overwritten: its GEMM2 K slab retired (commit) and its
                                               //     TMA store has read it (IO warp)              (both CTAs, 2)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(zfree + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A cluster = CS consecutive tiles of ONE utterance (tiles_per_b is a multiple of CS); CTAs (2k, 2k+1) form the cta_group::2
    // pairs.  Halo rows between tiles of the same cluster go through distributed shared memory, across clusters through global
    // memory (TMA store -> tile flag -> TMA load).
    const uint32_t crank = cluster_ctarank(), CS = cluster_nctarank();
    const uint32_t rank = crank & 1u, lead = crank & ~1u;
    const uint16_t pmask = (uint16_t)(3u << lead);
    const bool dsm_left = crank > 0, dsm_right = crank + 1 < CS;                  // neighbour tile inside this cluster
    const int b = blockIdx.x / p.tiles_per_b, ti = blockIdx.x - b * p.tiles_per_b, t0 = ti * BM;
    // halo sides that cross a cluster boundary inside an utterance go through global memory; at the ends of an utterance (first
    // tile, no valid tile behind) the halo rows are the conv's zero padding: zeroed once, never touched again
    const int ti_lead = ti - (int)rank;
    const bool pg_left = lead == 0 && ti_lead > 0, pg_right = lead + 2 == CS && (ti_lead + 2) * BM < p.T;
    const int pair_glob = (pg_left ? 1 : 0) + (pg_right ? 1 : 0);                // CTAs of this pair that load a halo side from global
    const bool g_left = crank == 0 && ti > 0, g_right = crank + 1 == CS && (ti + 1) * BM < p.T;
    const bool zero_left = crank == 0 && ti == 0, zero_right = crank + 1 == CS && (ti + 1) * BM >= p.T;
    const int L = p.L;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("b2s: dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.mapXin);
        prefetch_tmap(&p.mapWin);
        prefetch_tmap(&p.mapWd);
        prefetch_tmap(&p.mapWres);
        prefetch_tmap(&p.mapYe[0]);
        prefetch_tmap(&p.mapYe[1]);
        prefetch_tmap(&p.mapZ);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&accfull[i], 1);
            mbar_init(&accfree[i], EPI_WARPS * CLUSTER);
            mbar_init(&eready[i], 1);
            mbar_init(&efree[i], EPI_WARPS);
            mbar_init(&zfree[i], 2);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&zready[i], EPI_WARPS * CLUSTER);
        mbar_init(yready, EPI_WARPS * CLUSTER);
        mbar_init(halofull, 1);
        mbar_init(g1done, 1 + ((rank == 0 ? dsm_left : dsm_right) ? 1 : 0));       // own pair + the pair whose halo this CTA writes
        mbar_init(haloready, 1);
        mbar_init(haloin, 1);
        mbar_init(xfull, 1);
        mbar_init(ydone, EPI_WARPS);
        mbar_init(zdone, EPI_WARPS);
        mbar_init(stemfull, 1);
        mbar_init(xinfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_cg2(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // TMEM columns: [0,256) X (residual stream, lives across all layers), [256,512) accumulator of the stem / of one GEMM1 half

    // Programmatic dependent launch.  Normally: wait for the predecessor (the sampler update that wrote x_in) here.  A launch that
    // follows another utterance group of the SAME evaluation (chain bit 0) depends on nothing its predecessor - that group's skip /
    // head kernel - writes (its inputs were complete before the first group started, which did wait): it starts as soon as the
    // previous group's layer CTAs leave their SMs, i.e. under the previous group's 12-23 us head tail, and waits at its END instead,
    // so that "this kernel completed" still implies "everything before it completed" for whoever waits on the last one.
    if (!(p.chain & 1)) asm volatile("griddepcontrol.wait;" ::: "memory");
    // Dependent launch: the skip-sum / head kernel may start NOW.  It is scheduled only once EVERY CTA of this grid has executed this
    // instruction, i.e. is resident - so it can never take SMs this grid still needs, and this grid never waits for it.
    if (p.fuse_head) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#ifdef B2S_TLOG
    if (p.tlog && blockIdx.x == 2 && threadIdx.x == 0) { p.tlog[MAXL * 16] = globaltimer_ns(); p.tlog[MAXL * 16 + 1] = (unsigned long long)clock64(); }
#endif

    if (warp == 0) {
        // ===================== weight-tile producer (one lane; both CTAs, each fetches ITS half of every tile) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [&] { if (++stage == STAGES) { stage = 0; phase ^= 1; } };
            const uint32_t ys_a = smem_u32(ys), st_a = smem_u32(stages);
            // stem: A = x_in tile -> the own-row area of the (still idle) y slabs, B = input_projection weights
            if (rank == 0) mbar_expect_tx(xinfull, CLUSTER * p.kb_in * ZSLAB);
            for (int kb = 0; kb < p.kb_in; ++kb)
                tma_load_3d_cg2_a(ys_a + kb * YSLAB + HALO * 128, &p.mapXin, mapa_u32(xinfull, lead), kb * BK, t0, b);
            for (int kb = 0; kb < p.kb_in; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (rank == 0) mbar_expect_tx(&full[stage], CLUSTER * STAGE_BYTES);
                tma_load_2d_cg2_a(st_a + stage * STAGE_BYTES, &p.mapWin, mapa_u32(&full[stage], lead), kb * BK, rank * (C / 2));
                advance();
            }
            // The hoisted cond table (226 MB at 16 x 690 frames) and the z tiles stream through L2 once per evaluation and evict the
            // weights (30 MB), and all tiles run in lockstep: without help EVERY weight tile is an HBM miss for every CTA at the same
            // time.  Each CTA therefore pulls its 1/grid share of the NEXT layer's weights into L2 one layer ahead.
            const uint32_t wd_share = (((uint32_t)(2 * C * 3 * C * 2) + gridDim.x - 1) / gridDim.x + 15u) & ~15u;
            const uint32_t wr_share = (((uint32_t)(C * C * 2) + gridDim.x - 1) / gridDim.x + 15u) & ~15u;
            auto prefetch_weights = [&](int l) {
                const uint32_t o1 = blockIdx.x * wd_share, o2 = blockIdx.x * wr_share;
                if (o1 < (uint32_t)(2 * C * 3 * C * 2)) {
                    const uint32_t n = min(wd_share, (uint32_t)(2 * C * 3 * C * 2) - o1);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t*>(p.wd) + (size_t)l * (2 * C * 3 * C * 2) + o1), "r"(n) : "memory");
                }
                if (o2 < (uint32_t)(C * C * 2)) {
                    const uint32_t n = min(wr_share, (uint32_t)(C * C * 2) - o2);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t*>(p.wres) + (size_t)l * (C * C * 2) + o2), "r"(n) : "memory");
                }
            };
            prefetch_weights(0);
            for (int l = 0; l < L; ++l) {
                if (l + 1 < L) prefetch_weights(l + 1);
                if (l + 1 < L && t0 < p.T) {
                    // the next layer's 128 KB slab of the hoisted conditioner projection -> L2 (streams from HBM once per evaluation)
                    const uint8_t* nxt = reinterpret_cast<const uint8_t*>(p.cond) + ((l + 1) * p.cond_lstride + (long long)blockIdx.x * (BM * 2 * C)) * 2;
                    for (int i = 0; i < 8; ++i)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt + i * 16384), "r"(16384) : "memory");
                }
                auto fill_g2 = [&](int kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&full[stage], CLUSTER * STAGE_BYTES);
                    tma_load_3d_cg2_a(st_a + stage * STAGE_BYTES, &p.mapWres, mapa_u32(&full[stage], lead), kb * BK, rank * (C / 2), l);
                    advance();
                };
                for (int h = 0; h < 2; ++h) {
                    const int nrow = (h == 1 && KS == 3) ? 64 : 128;      // narrow model: half 1 is issued with N = 128
                    for (int f = 0; f < 12; ++f) {
                        if (KS < 4 && (f & 3) >= KS) continue;                      // ... and K slab 3 of every tap is all zeros
                        const int tap = (f >> 2) == 0 ? 1 : ((f >> 2) == 1 ? 0 : 2);       // centre tap first: it needs no halo
                        mbar_wait(&empty[stage], phase ^ 1);
#ifdef B2S_TLOG
                        if ((p.dbg & 1) && l > 0) { if (rank == 0) mbar_arrive(&full[stage]); advance(); continue; }
#endif
                        if (rank == 0) mbar_expect_tx(&full[stage], CLUSTER * STAGE_BYTES);
                        tma_load_3d_cg2_a(st_a + stage * STAGE_BYTES, &p.mapWd, mapa_u32(&full[stage], lead), tap * C + (f & 3) * BK,
                                          h * 256 + rank * nrow, l);
                        advance();
                    }
                }
                if (l + 1 < L)
                    for (int kb = 0; kb < KS; ++kb) fill_g2(kb);
            }
        }
    } else if (warp == 1) {
        if (rank == 1 && lane == 0) {
            // non-leader CTA: relay "my halo rows of y_m have landed" to the leader's MMA thread
            const uint32_t lr = mapa_u32(haloready, lead);
            for (int m = 0; m < L; ++m) {
                mbar_wait(haloin, (uint32_t)m & 1u);
                arrive_remote(lr);
            }
        }
        // ===================== MMA issuer (leader CTA, one lane): every MMA spans both CTAs (M = 256) =====================
        if (rank == 0 && lane == 0) {
            const uint32_t idesc_h = make_idesc_f16(2 * BM, 256, BF16);
            const uint32_t idesc_h1 = KS == 3 ? make_idesc_f16(2 * BM, 128, BF16) : idesc_h;     // narrow model: half 1 has 128 columns
            // GEMM1-retired signal: this pair's CTAs and the neighbouring CTAs that write this pair's halo rows
            const uint16_t gmask = (uint16_t)(pmask | (lead > 0 ? (1u << (lead - 1)) : 0u) | (lead + 2 < CS ? (1u << (lead + 2)) : 0u));
            const uint32_t ys_lo = desc_lo(smem_u32(ys)), zs_lo = desc_lo(smem_u32(zs)), st_lo = desc_lo(smem_u32(stages));
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [&] { if (++stage == STAGES) { stage = 0; phase ^= 1; } };
            // ---- stem: D[256..511] = x_in . W_in^T ----
            mbar_wait(xinfull, 0);
            tc_fence_after();
            for (int kb = 0; kb < p.kb_in; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a_lo = ys_lo + kb * (YSLAB >> 4) + HALO * (128 >> 4), b_lo = st_lo + stage * (STAGE_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < BK / UK; ++k) mma2(tmem_base + 256, a_lo + 2 * k, b_lo + 2 * k, idesc_h, (kb | k) != 0);
                umma_commit_cg2_mcast(&empty[stage], pmask);
                advance();
            }
            umma_commit_cg2_mcast(stemfull, pmask);
            for (int l = 0; l < L; ++l) {
                const uint32_t par = l & 1;
                const int dil = p.dil[l];
                mbar_wait(yready, par);                       // own rows of y_l are in both CTAs' shared memory (l = 0: and X_0 in TMEM)
                tc_fence_after();
                TLOG3(l, 0);
                auto g2 = [&](int kb) {                        // residual GEMM K slab kb: X += z[:, 64kb..] . (2^(l/2) Wres)^T
                    mbar_wait(&zready[kb], par);
                    tc_fence_after();
                    if (KS < 4 && kb >= KS) {                    // narrow model: nothing to multiply, keep the ring's phase count
                        umma_commit_cg2_mcast(&zfree[kb & 1], pmask);
                        return;
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = zs_lo + (kb & 1) * (ZSLAB >> 4), b_lo = st_lo + stage * (STAGE_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) mma2(tmem_base, a_lo + 2 * k, b_lo + 2 * k, idesc_h, 1u);
                    umma_commit_cg2_mcast(&empty[stage], pmask);
                    umma_commit_cg2_mcast(&zfree[kb & 1], pmask);          // ring buffer kb & 1 has been read by the tensor core
                    advance();
                };
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(accfree, (uint32_t)h);            // use 1 + 2l + h of the accumulator: the previous user has drained it
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + 256;
#pragma unroll 1
                    for (int f = 0; f < 12; ++f) {
                        const int tap = (f >> 2) == 0 ? 1 : ((f >> 2) == 1 ? 0 : 2);
                        if (h == 0 && f == 4) {                // side taps read the halo rows
                            mbar_wait(haloin, par);                // ... written by the cluster neighbours (st.async into this CTA)
                            mbar_wait(haloready, par);             // ... and into the peer CTA (relayed by its warp 1)
                            if (pair_glob) mbar_wait(halofull, par);   // ... and / or loaded from global by TMA
                            fence_proxy_async_all();
                            tc_fence_after();
                            TLOG3(l, 1);
                        }
                        if (KS < 4 && (f & 3) >= KS) continue;
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_lo = ys_lo + (f & 3) * (YSLAB >> 4) + (uint32_t)(HALO + (tap - 1) * dil) * (128 >> 4);
                        const uint32_t b_lo = st_lo + stage * (STAGE_BYTES >> 4);
                        const uint32_t idh = h ? idesc_h1 : idesc_h;
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) mma2(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idh, (f | k) != 0);
                        umma_commit_cg2_mcast(&empty[stage], pmask);
                        advance();
                    }
                    umma_commit_cg2_mcast(accfull, pmask);
                    if (h == 1) umma_commit_cg2_mcast(g1done, gmask);
                }
                if (l + 1 < L) {
                    for (int kb = 0; kb < 4; ++kb) g2(kb);
                    umma_commit_cg2_mcast(xfull, pmask);
                    TLOG3(l, 15);
                } else {                                       // last layer: no residual GEMM reads z, only the TMA stores do
                    umma_commit_cg2_mcast(&zfree[0], pmask);
                    umma_commit_cg2_mcast(&zfree[1], pmask);
                }
            }
        }
    } else if (warp == 2) {
        // ===================== step-embedding rows: e_m[c] = 2^(-m/2) Bsum_m[c] + d_m[c], double-buffered in shared memory ====
        // ... and the halo hand-off INSIDE the cluster: the d = dilation(m) edge rows of y_m go to the neighbour tiles through
        // distributed shared memory as soon as y_m is written.  This warp has nothing else to wait for, so it always stands at the
        // ydone barrier in time (the IO warp, which did this copy first, arrived 1 - 1.5 us late: its TMA stores of the z tiles take
        // ~3 us to read the ring behind the queued weight loads - in-kernel timeline - and the MMA stream stalls on the halo rows).
        const float* dv = p.dvec + (long long)b * p.d_stride;
        const uint32_t ys_a = smem_u32(ys);
        const uint32_t nl_ys = dsm_left ? mapa_u32(ys, crank - 1) : 0u, nr_ys = dsm_right ? mapa_u32(ys, crank + 1) : 0u;
        const uint32_t nl_bar = dsm_left ? mapa_u32(haloin, crank - 1) : 0u, nr_bar = dsm_right ? mapa_u32(haloin, crank + 1) : 0u;
        auto e_rows = [&](int m) {
            if (m >= 2) mbar_wait(&efree[m & 1], (uint32_t)((m - 2) >> 1) & 1u);
            const float s = exp2f(-0.5f * (float)m);
            float* dst = etab + (m & 1) * C;
#pragma unroll
            for (int i = 0; i < C / 32; ++i) {
                const int c = i * 32 + lane;
                dst[c] = fmaf(s, __ldg(p.bsum + m * C + c), __ldg(dv + m * C + c));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&eready[m & 1]);
        };
        auto halo_dsm = [&](int m) {
            mbar_wait(ydone, (uint32_t)m & 1u);                // own rows of y_m are in shared memory
            TLOG3(m, 12);
            // Only the d = dilation(m) rows next to the boundary are read by layer m's side taps: d x 128 B per slab and side
            const int d = p.dil[m];
            // arm this CTA's own halo barrier for layer m (phase m-1 is complete: GEMM1 of layer m-1 needed it)
            if (lane == 0) mbar_expect_tx(haloin, (uint32_t)((dsm_left ? 1 : 0) + (dsm_right ? 1 : 0)) * d * 128 * KS);
            if (m >= 1) mbar_wait(g1done, (uint32_t)(m - 1) & 1u);      // the neighbours no longer read the halo rows of y_{m-1}
            auto copy_side = [&](uint32_t src, uint32_t dst, uint32_t bar) {     // d rows of each slab: 8 d pieces of 16 B per slab
                for (int i0 = 0; i0 < d; i0 += 4) {
                    uint4 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = (i0 + i) * 32 + lane, s4 = idx / (8 * d), off = (idx - s4 * 8 * d) * 16;
                        if (i0 + i < d && (KS == 4 || s4 < KS))
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                                         : "r"(src + s4 * YSLAB + off));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = (i0 + i) * 32 + lane, s4 = idx / (8 * d), off = (idx - s4 * 8 * d) * 16;
                        if (i0 + i < d && (KS == 4 || s4 < KS)) st_async_u4(dst + s4 * YSLAB + off, v[i], bar);
                    }
                }
            };
            // my first d rows -> rows 144.. of the left neighbour (its +d tap); my last d rows -> rows [16 - d, 16) of the right
            // neighbour (its -d tap); both row offsets differ by 128: the swizzle phase is preserved
            if (dsm_left) copy_side(ys_a + HALO * 128, nl_ys + (HALO + BM) * 128, nl_bar);
            if (dsm_right) copy_side(ys_a + (HALO + BM - d) * 128, nr_ys + (HALO - d) * 128, nr_bar);
            TLOG3(m, 13);
        };
        e_rows(0);
        for (int m = 0; m < L; ++m) {
            if (m + 1 < L) e_rows(m + 1);                      // (efree of m - 1 was released before ydone(m - 1), which halo_dsm(m - 1) saw)
            halo_dsm(m);
        }
    } else if (warp == 3) {
        // ===================== IO warp: z ring -> z_all (TMA store), edge rows of y_m -> the neighbours' halo rows =====================
        // inside the cluster: a copy through distributed shared memory by this warp; across clusters: TMA store -> tile flag ->
        // the neighbour's TMA load
        const bool tile_ok = t0 < p.T;
        const uint32_t ys_a = smem_u32(ys), zs_a = smem_u32(zs);
        const int* fl = p.flags + b * p.tiles_per_b;
        const bool pub_left = g_left && tile_ok, pub_right = g_right;
        // the edge rows that cross a CLUSTER boundary (utterances longer than a cluster) go through global memory; the hand-off
        // inside the cluster is warp 2's
        auto publish = [&](int m) {
            if (!(pub_left || pub_right || g_left || g_right || (rank == 0 && pair_glob))) return;
            mbar_wait(ydone, (uint32_t)m & 1u);                // own rows of y_m are in shared memory
            if (lane == 0) {
                if (pub_left || pub_right) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        if (KS < 4 && s >= KS) break;
                        if (pub_left) tma_store_3d_a(&p.mapYe[m & 1], ys_a + s * YSLAB + HALO * 128, s * BK, t0, b);
                        if (pub_right) tma_store_3d_a(&p.mapYe[m & 1], ys_a + s * YSLAB + BM * 128, s * BK, t0 + BM - HALO, b);
                    }
                    tma_store_commit();
                    tma_store_wait_all();                      // writes complete
                    st_release_gpu(p.flags + blockIdx.x, m + 1);
                }
                if (g_left || g_right) {
                    if (m >= 1) mbar_wait(g1done, (uint32_t)(m - 1) & 1u);
                    if (pub_left) wait_flag3(fl + ti - 1, m + 1);
                    if (pub_right) wait_flag3(fl + ti + 1, m + 1);
                    if (pub_left || pub_right) fence_proxy_async_all();
                    TLOG3(m, 14);
                    const uint32_t lb = mapa_u32(halofull, lead);
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        if (KS < 4 && s >= KS) break;
                        if (g_left) tma_load_3d_cg2_a(ys_a + s * YSLAB, &p.mapYe[m & 1], lb, s * BK, t0 - HALO, b);
                        if (g_right) tma_load_3d_cg2_a(ys_a + s * YSLAB + (HALO + BM) * 128, &p.mapYe[m & 1], lb, s * BK, t0 + BM, b);
                    }
                }
                if (rank == 0 && pair_glob) mbar_expect_tx(halofull, pair_glob * KS * HALO * 128);
            }
            __syncwarp();
        };
        publish(0);
        for (int l = 0; l < L; ++l) {
            for (int h = 0; h < 2; ++h) {
                if (lane == 0) {
                    mbar_wait(zdone, (uint32_t)h);             // phase 2l + h: both ring buffers hold z slabs 2h, 2h + 1 of layer l
                    if (p.fuse_head && tile_ok && h == 1) {    // half 0 of this layer has reached global memory: publish it
                        tma_store_wait_all();
                        st_release_gpu(p.zflags + blockIdx.x, 2 * l + 1);
                    }
                    if (tile_ok) {
                        tma_store_4d(&p.mapZ, zs_a, (2 * h) * BK, t0, b, l);
                        if (KS == 4 || 2 * h + 1 < KS) tma_store_4d(&p.mapZ, zs_a + ZSLAB, (2 * h + 1) * BK, t0, b, l);
                        tma_store_commit();
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the stores have READ the ring
                    }
                    mbar_arrive(&zfree[0]);
                    mbar_arrive(&zfree[1]);
                }
                __syncwarp();
            }
            if (l + 1 < L) publish(l + 1);
            if (lane == 0 && p.fuse_head && tile_ok) {         // after the (time-critical) halo hand-off: publish half 1
                tma_store_wait_all();
                st_release_gpu(p.zflags + blockIdx.x, 2 * l + 2);
            }
        }
        if (lane == 0) tma_store_wait_all();
    } else {
        // ===================== epilogue: 8 warps; warp e -> TMEM lane quarter e & 3, column half e >> 2 =====================
        const int e = warp - 4, qd = e & 3, sub = e >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
        const int row = qd * 32 + lane;
        // rows at or beyond the utterance's length are the conv's zero padding: y and z are forced to 0 there (ragged batches: lens)
        const bool valid = t0 + row < (p.lens ? __ldg(p.lens + b) : p.T);
        const int sw = row & 7;
        const uint32_t zrow = smem_u32(zs) + (row >> 3) * 1024 + sw * 128;
        const uint32_t yrow = smem_u32(ys) + ((row + HALO) >> 3) * 1024 + sw * 128;       // HALO % 8 == 0: same swizzle phase
        const uint32_t et = smem_u32(etab);
        const uint32_t lead_yready = mapa_u32(yready, lead);
        const bool all_valid = __all_sync(0xffffffffu, valid);
        // y_m (own rows) = valid ? s * X + e_m[c] : 0 for this warp's four 32-column chunks -> swizzled shared-memory tile.
        // STEM: X = relu(acc + b_in) is first written to the X columns of TMEM.
        auto produce_y = [&](int m, bool stem) {
            mbar_wait(&eready[m & 1], (uint32_t)(m >> 1) & 1u);
            if (e == 0 && lane == 0 && m >= 1) TLOG3(m - 1, 6);
            const float s = exp2f(-0.5f * (float)m);
            const uint32_t eb = et + (m & 1) * (C * 4);
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {
                const int J = 4 * sub + jj;                    // columns (= channels) [32J, 32J + 32)
                if (KS < 4 && J >= 2 * KS) break;                      // narrow model: channels >= 64 ks are zero padding, never read
                float acc[32];
                tmem_ld32(taddr + (stem ? 256 : 0) + 32 * J, acc);
                tmem_ld_wait();
                if (stem) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 bi = __ldg(reinterpret_cast<const float4*>(p.b_in + 32 * J + 4 * i));
                        acc[4 * i] = fmaxf(acc[4 * i] + bi.x, 0.f);
                        acc[4 * i + 1] = fmaxf(acc[4 * i + 1] + bi.y, 0.f);
                        acc[4 * i + 2] = fmaxf(acc[4 * i + 2] + bi.z, 0.f);
                        acc[4 * i + 3] = fmaxf(acc[4 * i + 3] + bi.w, 0.f);
                    }
                    tmem_st32(taddr + 32 * J, acc);
                }
                uint32_t yp[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 ev = ld_shared_f4(eb + (32 * J + 4 * i) * 4);
                    yp[2 * i] = valid ? Half16<BF16>::pack2(fmaf(s, acc[4 * i], ev.x), fmaf(s, acc[4 * i + 1], ev.y)) : 0u;
                    yp[2 * i + 1] = valid ? Half16<BF16>::pack2(fmaf(s, acc[4 * i + 2], ev.z), fmaf(s, acc[4 * i + 3], ev.w)) : 0u;
                }
                const uint32_t slab = yrow + (J >> 1) * YSLAB;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4)
                    st_shared_u4(slab + (((4 * (J & 1) + c4) ^ sw) << 4), make_uint4(yp[4 * c4], yp[4 * c4 + 1], yp[4 * c4 + 2], yp[4 * c4 + 3]));
            }
            if (stem) tmem_st_wait();
            if (e == 0 && lane == 0 && m >= 1) TLOG3(m - 1, 7);
            fence_proxy_async_smem();                          // generic-proxy writes -> tensor core / TMA store (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&efree[m & 1]);
                mbar_arrive(ydone);
                arrive_remote(lead_yready);
                if (e == 0 && m >= 1) TLOG3(m - 1, 8);
            }
        };
        if (zero_left || zero_right) {                         // 4 slabs x 16 rows x 128 B per side = 512 x 16 B: two stores per thread
            const int tid = threadIdx.x - 128;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = tid + 256 * i, s4 = idx >> 7, off = (idx & 127) * 16;
                if (zero_left) st_shared_u4(smem_u32(ys) + s4 * YSLAB + off, make_uint4(0, 0, 0, 0));
                if (zero_right) st_shared_u4(smem_u32(ys) + s4 * YSLAB + (HALO + BM) * 128 + off, make_uint4(0, 0, 0, 0));
            }
        }
        // ---- stem epilogue ----
        mbar_wait(stemfull, 0);
        tc_fence_after();
        produce_y(0, true);
        if (lane == 0) {                                       // the stem accumulator (both GEMM1 buffers) is drained
            arrive_remote(mapa_u32(accfree, lead));
        }
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
            const uint32_t par = l & 1;
            // tile/chunk-major cond table: (chunk J of 32 packed columns, 16-byte piece i) of this tile = 128 rows x 16 B contiguous
            const uint16_t* ctile = reinterpret_cast<const uint16_t*>(p.cond) + l * p.cond_lstride + (long long)blockIdx.x * (BM * 2 * C) + row * 8;
            uint4 c[2][4];
            // Half h of GEMM1 = z K slabs 2h and 2h + 1.  ALL eight warps gate slab 2h first (pr = 0), then slab 2h + 1 (pr = 1): the
            // residual GEMM's K slab 2h is issued while the second slab is still being gated, so that after the layer's last gate
            // epilogue only ONE K slab (4 MMAs) remains before the y epilogue.  Warp (qd, sub) takes chunks 4 pr + 2 sub + {0, 1}.
            auto load_cond = [&](int h, int pr) {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        c[jj][i] = valid ? ldg_nc_u4(ctile + ((8 * h + 4 * pr + 2 * sub + jj) * 4 + i) * (BM * 8)) : make_uint4(0, 0, 0, 0);
            };
            const uint32_t lead_accfree = mapa_u32(accfree, lead);
            load_cond(0, 0);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                mbar_wait(accfull, (uint32_t)h);               // phase 2l + h
                tc_fence_after();
                if (e == 0 && lane == 0) TLOG3(l, 2 + 2 * h);
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) {
                    // narrow model: K slab 3 (half 1, pr 1) does not exist - its barrier arrivals are kept, its work is skipped
                    const bool dead = KS < 4 && 2 * h + pr >= KS;
                    float acc0[32], acc1[32];
                    if (!dead) {
                        tmem_ld32(taddr + 256 + 32 * (4 * pr + 2 * sub), acc0);
                        tmem_ld32(taddr + 256 + 32 * (4 * pr + 2 * sub + 1), acc1);
                        tmem_ld_wait();
                    }
                    if (pr == 1) {                             // the accumulator is in registers: hand it back to the MMA issuer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_remote(lead_accfree);
                    }
                    uint32_t zp[16];
#ifdef B2S_TLOG
                    if (p.dbg & 2) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) zp[i] = __float_as_uint(acc0[i]);
                    } else
#endif
                    if (!dead)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const float* acc = jj ? acc1 : acc0;
                        const uint32_t* cw = reinterpret_cast<const uint32_t*>(c[jj]);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 ca = Half16<BF16>::unpack2(cw[2 * i]), cb = Half16<BF16>::unpack2(cw[2 * i + 1]);
                            const float z0 = sigmoid_fast(acc[4 * i] + ca.x) * tanh_fast(acc[4 * i + 1] + ca.y);
                            const float z1 = sigmoid_fast(acc[4 * i + 2] + cb.x) * tanh_fast(acc[4 * i + 3] + cb.y);
                            zp[8 * jj + i] = Half16<BF16>::pack2(z0, z1);
                        }
                    }
                    if (!all_valid) {                          // only the last tile of an utterance has padding rows
#pragma unroll
                        for (int i = 0; i < 16; ++i) zp[i] = valid ? zp[i] : 0u;
                    }
                    // cond rows of the next step: the other slab of this half, the next half, or (h = 1, pr = 1) nothing: the next
                    // layer's first rows are requested at the top of the layer loop, after the y epilogue
                    if (pr == 0) { if (KS == 4 || 2 * h + 1 < KS) load_cond(h, 1); }
                    else if (h == 0) load_cond(1, 0);
                    // write 2l + h of ring buffer pr: the previous occupant's GEMM2 slab has retired and its TMA store has read it
                    if (l + h > 0) mbar_wait(&zfree[pr], (uint32_t)(2 * l + h - 1) & 1u);
                    // z channels [64 (2h + pr) + 32 sub, +32) of this row: K slab 2h + pr (ring buffer pr), 16-byte pieces 4 sub .. 4 sub + 3
                    const uint32_t slab = zrow + pr * ZSLAB;
                    if (!dead) {
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4)
                            st_shared_u4(slab + (((4 * sub + c4) ^ sw) << 4), make_uint4(zp[4 * c4], zp[4 * c4 + 1], zp[4 * c4 + 2], zp[4 * c4 + 3]));
                    }
                    fence_proxy_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_remote(mapa_u32(&zready[2 * h + pr], lead));
                }
                if (lane == 0) mbar_arrive(zdone);
                if (e == 0 && lane == 0) TLOG3(l, 3 + 2 * h);
            }
            if (l + 1 < L) {
                mbar_wait(xfull, par);
                tc_fence_after();
                if (e == 0 && lane == 0) TLOG3(l, 10);
                produce_y(l + 1, false);
                if (e == 0 && lane == 0) TLOG3(l, 11);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
#ifdef B2S_TLOG
    if (p.tlog && blockIdx.x == 2 && threadIdx.x == 0) { p.tlog[MAXL * 16 + 2] = globaltimer_ns(); p.tlog[MAXL * 16 + 3] = (unsigned long long)clock64(); }
#endif
    if (p.chain & 1) asm volatile("griddepcontrol.wait;" ::: "memory");
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, 512);
    }
}

// z_all [L, B*T, C] 16-bit as a 4-D map {C, T, B, L}, box {64, 128, 1, 1}: stores clip rows >= T of an utterance
static int make_map_z4(CUtensorMap* m, const void* base, int bf16, int T, int B, int L, int64_t layer_stride) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B2S_ERR_CUDA; }
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B, (cuuint64_t)L};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2, (cuuint64_t)layer_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)BM, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(z_all T=%d B=%d L=%d) failed with CUresult %d", T, B, L, (int)rc);
        return B2S_ERR_CUDA;
    }
    return B2S_OK;
}

// Every CTA of the launch must be resident at once (tiles wait for their neighbours): ask the driver how many clusters of THIS
// configuration fit the device as it is now (MPS / MIG / green-context limits included).  Cluster size: the largest of 8, 6, 4, 2
// that divides the tiles per utterance (halo rows inside a cluster travel through distributed shared memory); smaller if the
// device cannot co-schedule enough clusters of that size.
template <int BF16, int KS>
static int launch(const Stack3P& p, int grid, cudaStream_t st) {
    int dev = 0;
    B2S_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("b2s_tc_wavenet_stack3: device index %d out of range", dev); return B2S_ERR_UNSUPPORTED; }
    static bool configured[64] = {};
    static int max_clusters[64][5] = {};          // [dev][cluster size / 2]; 0 = not queried yet, -1 = none
    static const int forced = getenv("B2S_STACK3_CLUSTER") ? atoi(getenv("B2S_STACK3_CLUSTER")) : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    if (!configured[dev]) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(wavenet_stack3_kernel<BF16, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured[dev] = true;
    }
    int cs_used = 0, best = 0;
    for (int cs = 8; cs >= 2; cs -= 2) {
        if (p.tiles_per_b % cs || grid % cs) continue;
        if (forced && cs != forced && cs != 2) continue;
        int& n = max_clusters[dev][cs / 2];
        if (n == 0) {
            attr[0].val.clusterDim.x = cs;
            cfg.numAttrs = 1;
            int q = 0;
            if (cudaOccupancyMaxActiveClusters(&q, wavenet_stack3_kernel<BF16, 4>, &cfg) != cudaSuccess) { cudaGetLastError(); q = 0; }
            n = q > 0 ? q : -1;
        }
        if (n > best * 1) best = n > best ? n : best;
        if (n > 0 && grid <= cs * n) { cs_used = cs; break; }
    }
    if (!cs_used) {
        set_error("b2s_tc_wavenet_stack3: %d tiles do not fit the device at once (every tile must be resident; the driver reports at most "
                  "%d co-resident clusters for this kernel); split the batch by utterance", grid, best);
        return B2S_ERR_UNSUPPORTED;
    }
    static const bool verbose = getenv("B2S_STACK3_VERBOSE") != nullptr;
    if (verbose) {
        static int last_grid = -1;
        if (grid != last_grid) {
            last_grid = grid;
            fprintf(stderr, "b2s stack3: grid %d (%d layer tiles), cluster size %d; co-resident clusters reported: cs8=%d cs6=%d cs4=%d cs2=%d\n",
                    grid, p.n_layer_ctas, cs_used, max_clusters[dev][4], max_clusters[dev][3], max_clusters[dev][2], max_clusters[dev][1]);
        }
    }
    attr[0].val.clusterDim.x = cs_used;
    cfg.numAttrs = 2;
    B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg, wavenet_stack3_kernel<BF16, KS>, p));
    return B2S_OK;
}

// the skip-sum / head kernel: one CTA pair per four layer tiles, launched with programmatic dependent launch behind the layer kernel
template <int BF16, int KS>
static int launch_skiphead(const Stack3P& p, int grid, cudaStream_t st) {
    int dev = 0;
    B2S_CHECK_CUDA(cudaGetDevice(&dev));
    static bool configured[64] = {};
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        B2S_CHECK_CUDA(cudaFuncSetAttribute(wavenet_skiphead3_kernel<BF16, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES));
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SK_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    B2S_CHECK_CUDA(cudaLaunchKernelEx(&cfg, wavenet_skiphead3_kernel<BF16, KS>, p));
    return B2S_OK;
}

// largest number of tiles one launch can hold for utterances of `tiles_per_b` tiles (0 if none): what the host uses to split a batch
template <int BF16>
static int max_tiles(int tiles_per_b) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    cudaFuncSetAttribute(wavenet_stack3_kernel<BF16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tiles_per_b);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int best = 0;
    for (int cs = 8; cs >= 2; cs -= 2) {
        if (tiles_per_b % cs) continue;
        attr[0].val.clusterDim.x = cs;
        int q = 0;
        if (cudaOccupancyMaxActiveClusters(&q, wavenet_stack3_kernel<BF16, 4>, &cfg) != cudaSuccess) { cudaGetLastError(); q = 0; }
        if (q * cs > best) best = q * cs;
    }
    return best;
}

}  // namespace ws3
}  // namespace tc
}  // namespace b2s

using namespace b2s;
using namespace b2s::tc;

extern unsigned long long* g_tlog;

extern "C" int b2s_tc_wavenet_stack3_halo(void) { return ws3::HALO; }
extern "C" int b2s_tc_wavenet_stack3_max_tiles(int T, int bf16) {
    const int tpb = (ceil_div(T, ws3::BM) + 1) & ~1;
    return bf16 ? ws3::max_tiles<1>(tpb) : ws3::max_tiles<0>(tpb);
}

struct Head3 {            // operands of the fused skip sum + head (all NULL / 0: the plain stack, z_all only)
    const void* Wskip_h; const float* bss; const void* Wsp_h; const float* b_sp; const void* Wfin_h; const float* b_fin; float* out;
    int* zflags;
    int chain;
};

static int stack3_impl(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, const void* Wd_h,
                       const void* cond_h, int64_t cond_layer_stride, const void* Wres_h, const float* bsum, const float* dvec,
                       int d_stride, const int* dilations_host, int L, void* yedge0_h, void* yedge1_h, void* z_all_h,
                       int64_t z_layer_stride, int B, int T, int C, int* flags, const int* lens, int bf16, void* stream,
                       const Head3* hd) {
    B2S_CHECK_ARG(xin_h && Win_h && b_in && Wd_h && cond_h && Wres_h && bsum && dvec && dilations_host && yedge0_h && yedge1_h &&
                      z_all_h && flags, "b2s_tc_wavenet_stack3: null pointer");
    // C = 256, or 192 for a narrower model whose operands are zero-padded to the 256-channel layout (channels >= 192 all zero in the
    // weights, biases, cond table and step embeddings): the all-zero fourth K slab of every GEMM is then skipped (Stack3P::ks)
    if (C != ws3::C && C != 192) {
        set_error("b2s_tc_wavenet_stack3: specialised for %d residual channels, or 192 used channels in the 256-channel layout (got %d)", ws3::C, C);
        return B2S_ERR_UNSUPPORTED;
    }
    const int ks = C == 192 ? 3 : 4;
    C = ws3::C;                                   // every operand is laid out for 256 channels
    B2S_CHECK_ARG(L >= 1 && L <= ws3::MAXL, "b2s_tc_wavenet_stack3: 1 <= L <= %d (got %d)", ws3::MAXL, L);
    B2S_CHECK_ARG(MF > 0 && MF <= 256 && MF % 8 == 0 && ld_win % 8 == 0, "b2s_tc_wavenet_stack3: in_dims*n_feats must be a multiple of 8 and <= 256 (got %d)", MF);
    B2S_CHECK_ARG(yedge0_h != yedge1_h && d_stride % 4 == 0 && cond_layer_stride % 8 == 0, "b2s_tc_wavenet_stack3: bad strides / aliasing");
    B2S_CHECK_ARG(al16(xin_h) && al16(Win_h) && al16(b_in) && al16(Wd_h) && al16(cond_h) && al16(Wres_h) && al16(bsum) && al16(dvec) &&
                      al16(yedge0_h) && al16(yedge1_h) && al16(z_all_h), "b2s_tc_wavenet_stack3: misaligned pointer");
    if (B * T == 0) return B2S_OK;
    ws3::Stack3P p{};
    p.tiles_per_b = (ceil_div(T, ws3::BM) + 1) & ~1;
    int grid = B * p.tiles_per_b;
    p.n_layer_ctas = grid;
    for (int l = 0; l < L; ++l) {
        if (dilations_host[l] < 1 || dilations_host[l] > ws3::HALO) {
            set_error("b2s_tc_wavenet_stack3: dilation %d exceeds the resident halo of %d rows", dilations_host[l], ws3::HALO);
            return B2S_ERR_UNSUPPORTED;
        }
        p.dil[l] = dilations_host[l];
    }
    int rc = make_map_act(&p.mapXin, xin_h, bf16, MF, MF, T, B, ws3::BK, ws3::BM);
    if (rc) return rc;
    rc = make_map_w(&p.mapWin, Win_h, bf16, MF, C, ld_win, ws3::BK, C / 2);
    if (rc) return rc;
    rc = make_map_w3(&p.mapWd, Wd_h, bf16, 3 * C, 2 * C, L, ws3::BK, C / 2);
    if (rc) return rc;
    rc = make_map_w3(&p.mapWres, Wres_h, bf16, C, C, L, ws3::BK, C / 2);
    if (rc) return rc;
    rc = make_map_act(&p.mapYe[0], yedge0_h, bf16, C, C, T, B, ws3::BK, ws3::HALO);
    if (rc) return rc;
    rc = make_map_act(&p.mapYe[1], yedge1_h, bf16, C, C, T, B, ws3::BK, ws3::HALO);
    if (rc) return rc;
    B2S_CHECK_ARG(z_layer_stride >= (int64_t)B * T * C && z_layer_stride % 8 == 0, "b2s_tc_wavenet_stack3: bad z_all layer stride");
    rc = ws3::make_map_z4(&p.mapZ, z_all_h, bf16, T, B, L, z_layer_stride);
    if (rc) return rc;
    p.B = B; p.T = T; p.L = L; p.MF = MF; p.kb_in = ceil_div(MF, ws3::BK);
    p.ks = ks;
    p.wd = Wd_h; p.wres = Wres_h;
    p.cond = cond_h; p.cond_lstride = cond_layer_stride; p.b_in = b_in; p.bsum = bsum; p.dvec = dvec; p.d_stride = d_stride;
    p.flags = flags;
    p.lens = lens;
    p.tlog = g_tlog;
    p.dbg = getenv("B2S_STACK3_DBG") ? atoi(getenv("B2S_STACK3_DBG")) : 0;
    if (hd) {
        B2S_CHECK_ARG(hd->Wskip_h && hd->bss && hd->Wsp_h && hd->b_sp && hd->Wfin_h && hd->b_fin && hd->out && hd->zflags,
                      "b2s_tc_wavenet_denoiser3: null pointer");
        B2S_CHECK_ARG(MF % 16 == 0, "b2s_tc_wavenet_denoiser3: in_dims*n_feats must be a multiple of 16 (got %d)", MF);
        B2S_CHECK_ARG(al16(hd->Wskip_h) && al16(hd->bss) && al16(hd->Wsp_h) && al16(hd->b_sp) && al16(hd->Wfin_h) && al16(hd->b_fin) &&
                          al16(hd->out), "b2s_tc_wavenet_denoiser3: misaligned pointer");
        rc = make_map_w3(&p.mapWskip, hd->Wskip_h, bf16, C, C, L, ws3::BK, C / 2);
        if (rc) return rc;
        rc = make_map_w(&p.mapWsp, hd->Wsp_h, bf16, C, C, C, ws3::BK, C / 2);
        if (rc) return rc;
        rc = make_map_w(&p.mapWfin, hd->Wfin_h, bf16, C, MF, C, ws3::BK, MF / 2);
        if (rc) return rc;
        p.fuse_head = 1;
        p.chain = hd->chain & 3;
        p.bss = hd->bss; p.b_sp = hd->b_sp; p.b_fin = hd->b_fin; p.out = hd->out; p.zflags = hd->zflags;
        p.alpha = 1.0f / sqrtf((float)L);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (ks == 3) rc = bf16 ? ws3::launch<1, 3>(p, grid, st) : ws3::launch<0, 3>(p, grid, st);
    else rc = bf16 ? ws3::launch<1, 4>(p, grid, st) : ws3::launch<0, 4>(p, grid, st);
    if (rc || !hd) return rc;
    const int sgrid = 2 * ceil_div(grid, 4);
    if (ks == 3) return bf16 ? ws3::launch_skiphead<1, 3>(p, sgrid, st) : ws3::launch_skiphead<0, 3>(p, sgrid, st);
    return bf16 ? ws3::launch_skiphead<1, 4>(p, sgrid, st) : ws3::launch_skiphead<0, 4>(p, sgrid, st);
}

extern "C" int b2s_tc_wavenet_stack3(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, const void* Wd_h,
                                     const void* cond_h, int64_t cond_layer_stride, const void* Wres_h, const float* bsum,
                                     const float* dvec, int d_stride, const int* dilations_host, int L, void* yedge0_h,
                                     void* yedge1_h, void* z_all_h, int64_t z_layer_stride, int B, int T, int C, int* flags,
                                     const int* lens, int bf16, void* stream) {
    return stack3_impl(xin_h, MF, Win_h, ld_win, b_in, Wd_h, cond_h, cond_layer_stride, Wres_h, bsum, dvec, d_stride, dilations_host, L,
                       yedge0_h, yedge1_h, z_all_h, z_layer_stride, B, T, C, flags, lens, bf16, stream, nullptr);
}

/* utterances of T frames ONE b2s_tc_wavenet_denoiser3 launch can hold (layer tiles + the skip / head CTAs behind them); 0 = none */
extern "C" int b2s_tc_wavenet_denoiser3_max_utterances(int T, int bf16) {
    const int tpb = (ceil_div(T, ws3::BM) + 1) & ~1;
    return (bf16 ? ws3::max_tiles<1>(tpb) : ws3::max_tiles<0>(tpb)) / tpb;        // the skip / head kernel has no residency requirement
}

extern "C" int b2s_tc_wavenet_denoiser3(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, const void* Wd_h,
                                        const void* cond_h, int64_t cond_layer_stride, const void* Wres_h, const float* bsum,
                                        const float* dvec, int d_stride, const int* dilations_host, int L, void* yedge0_h,
                                        void* yedge1_h, void* z_all_h, int64_t z_layer_stride, const void* Wskip_h, const float* bss,
                                        const void* Wsp_h, const float* b_sp, const void* Wfin_h, const float* b_fin, float* out, int B,
                                        int T, int C, int* flags, int* zflags, const int* lens, int bf16, void* stream) {
    Head3 hd{Wskip_h, bss, Wsp_h, b_sp, Wfin_h, b_fin, out, zflags, 0};
    return stack3_impl(xin_h, MF, Win_h, ld_win, b_in, Wd_h, cond_h, cond_layer_stride, Wres_h, bsum, dvec, d_stride, dilations_host, L,
                       yedge0_h, yedge1_h, z_all_h, z_layer_stride, B, T, C, flags, lens, bf16, stream, &hd);
}

/* The same launch as one of SEVERAL utterance groups of one evaluation, issued back to back on the stream.  chain bit 0: this launch
 * follows another group of the evaluation, bit 1: another group follows.  Results are those of b2s_tc_wavenet_denoiser3; the groups'
 * skip / head tails overlap the next group's layers. */
extern "C" int b2s_tc_wavenet_denoiser3_chained(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, const void* Wd_h,
                                                const void* cond_h, int64_t cond_layer_stride, const void* Wres_h, const float* bsum,
                                                const float* dvec, int d_stride, const int* dilations_host, int L, void* yedge0_h,
                                                void* yedge1_h, void* z_all_h, int64_t z_layer_stride, const void* Wskip_h,
                                                const float* bss, const void* Wsp_h, const float* b_sp, const void* Wfin_h,
                                                const float* b_fin, float* out, int B, int T, int C, int* flags, int* zflags,
                                                const int* lens, int bf16, int chain, void* stream) {
    Head3 hd{Wskip_h, bss, Wsp_h, b_sp, Wfin_h, b_fin, out, zflags, chain};
    return stack3_impl(xin_h, MF, Win_h, ld_win, b_in, Wd_h, cond_h, cond_layer_stride, Wres_h, bsum, dvec, d_stride, dilations_host, L,
                       yedge0_h, yedge1_h, z_all_h, z_layer_stride, B, T, C, flags, lens, bf16, stream, &hd);
}
