// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences), UMMA shared-memory + instruction descriptors.
// Hand-written inline PTX; bit layouts follow the PTX ISA "tcgen05" chapter (matrix descriptor,
// instruction descriptor for .kind::f16) - see DESIGN.md for the field tables.
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "b2s_common.cuh"

namespace b2s {
namespace tc {

// ---------------------------------------------------------------------------------------------------
// shared-memory addressing
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Time-bounded wait: a lost arrival traps (launch error) after ~2 s instead of hanging the GPU box.
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spin = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spin & 1023u) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
            printf("b2s: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
                   smem_u32(bar), parity);
            __trap();
        }
    }
}

// generic-proxy writes (st.shared) -> visible to the async proxy (TMA store / UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {     // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {       // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// tcgen05: descriptors
// ---------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand tile stored as rows of 128 bytes (64 x 16-bit) with
// the 128-byte swizzle (what TMA SWIZZLE_128B writes):
//   [ 0,14) start address >> 4      [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups     [46,48) version = 1 (Blackwell)
//   [49,52) base offset = 0 (tiles are 1024 B aligned)                [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor for .kind::f16, fp32 accumulate, both operands K-major:
//   [4,6) D format = 1 (F32)   [7,10) A format   [10,13) B format  (0 = F16, 1 = BF16)
//   [15] A major = 0 (K)  [16] B major = 0 (K)   [17,23) N >> 3    [24,29) M >> 4
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int M, int N, int is_bf16) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= (uint32_t)(is_bf16 ? 1 : 0) << 7;
    d |= (uint32_t)(is_bf16 ? 1 : 0) << 10;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

// D[tmem] (+)= A[smem] . B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// tcgen05: TMEM -> registers.  32 lanes x 32 consecutive fp32 columns; thread i of the warp gets lane
// (warp_quarter * 32 + i); the warp may only touch the lane quarter (warp_id % 4).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// K-major SWIZZLE_128B descriptor whose start address is NOT 1024-byte aligned (a window that starts `r` rows into a
// TMA-written slab): `base_offset` is the descriptor's [49,52) field
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc_bo(uint32_t smem_addr, uint32_t base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7u) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

// ---------------------------------------------------------------------------------------------------
// thread-block clusters: multicast TMA, multicast tcgen05.commit, cluster barrier
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// The box lands at the SAME shared-memory offset in every CTA of `mask`, and completes bytes on the mbarrier at the
// same offset in each of them.
__device__ __forceinline__ void tma_load_2d_mcast(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}
// 3-D multicast variant (weights stacked over layers: coordinates k, n, layer)
__device__ __forceinline__ void tma_load_3d_mcast(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6}], [%2], %3;" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// ---- inter-CTA hand-off through global memory (persistent kernels): release / acquire flags + proxy fence ----
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// generic-proxy writes (other SMs' st.global, made visible by release/acquire) -> visible to this thread's TMA loads
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// arrives (once the issuing thread's earlier MMAs retire) on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

// ---------------------------------------------------------------------------------------------------
// cta_group::2 (CTA pair): one MMA spans two SMs - M = 256 (128 rows per CTA, each in its own TMEM), each CTA holds
// HALF of the B tile (N/2 rows) in its own shared memory.  Only the leader (even rank) issues MMAs / commits.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* dst_smem, uint32_t ncols) {     // whole warp, same warp id in BOTH CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// shared::cluster address of `p` (an address in THIS CTA's shared memory) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// The same arrival without release semantics: for barriers that hand back an on-chip resource the waiter does not read through memory
// (a TMEM accumulator drained with tcgen05.ld + tcgen05.wait::ld: the values are in registers, tcgen05.fence::before_thread_sync orders
// the tensor-memory reads before the arrival).  The release form makes the arriving thread wait until the warp's preceding GLOBAL stores
// are performed - for an epilogue that just wrote its tile, that is the whole store latency, per tile (ncu: "membar" stalls).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in the ISSUING CTA's shared memory, bytes complete on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_3d_cg2(void* smem, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(void* smem, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_ss_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma_commit_cg2_mcast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

// ---------------------------------------------------------------------------------------------------
// 16-bit operand types (bf16 or fp16: same tensor-core rate; fp16 has 3 more mantissa bits)
// ---------------------------------------------------------------------------------------------------
template <int BF16>
struct Half16;
template <>
struct Half16<1> {
    __device__ __forceinline__ static uint32_t pack2(float a, float b) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ __forceinline__ static float2 unpack2(uint32_t u) {
        __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
        return __bfloat1622float2(h);
    }
};
template <>
struct Half16<0> {
    __device__ __forceinline__ static uint32_t pack2(float a, float b) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ __forceinline__ static float2 unpack2(uint32_t u) {
        __half2 h = *reinterpret_cast<__half2*>(&u);
        return __half22float2(h);
    }
};

// ---------------------------------------------------------------------------------------------------
// host side: tensor-map builders (cuTensorMapEncodeTiled fetched through the runtime, no -lcuda)
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
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

// 16-bit activations [B, T, cols] (row stride ld elements) -> 3-D map {cols, T, B}, box {box_k, box_rows, 1},
// 128B swizzle, zero fill outside [0, T) per utterance.
inline int make_map_act(CUtensorMap* m, const void* base, int bf16, int cols, int ld, int T, int B, int box_k, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B2S_ERR_CUDA; }
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
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
// 16-bit weights [N, K] (row stride ldw) -> 2-D map {K, N}, box {box_k, box_rows}
inline int make_map_w(CUtensorMap* m, const void* base, int bf16, int K, int N, int ldw, int box_k, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B2S_ERR_CUDA; }
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)ldw * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
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

// 16-bit weights stacked over layers [L, N, K] -> 3-D map {K, N, L}, box {box_k, box_rows, 1}
inline int make_map_w3(CUtensorMap* m, const void* base, int bf16, int K, int N, int L, int box_k, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return B2S_ERR_CUDA; }
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)L};
    cuuint64_t gstr[2] = {(cuuint64_t)K * 2, (cuuint64_t)N * K * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                     const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(stacked weights K=%d N=%d L=%d) failed with CUresult %d", K, N, L, (int)rc);
        return B2S_ERR_CUDA;
    }
    return B2S_OK;
}

// Attributes such as the SM count or cudaFuncAttributeMaxDynamicSharedMemorySize are PER DEVICE: every cache below is keyed by the
// current device (a process may drive several GPUs).
constexpr int MAX_DEVICES = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0;
}
inline int num_sms() {
    static int n[MAX_DEVICES] = {};
    const int dev = current_device();
    if (!n[dev]) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}
// `static PerDevice configured; if (configured.first()) { cudaFuncSetAttribute(...); }` - true once per device
struct PerDevice {
    bool done[MAX_DEVICES] = {};
    bool first() {
        const int dev = current_device();
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};

inline bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

}  // namespace tc
}  // namespace b2s
