// 16-bit side of the elementwise helpers: fp32 -> bf16/fp16 casts that feed the TMA-staged
// tensor-core operands (HBM-bound, vectorised, grid-stride).
#include "b2s_tc.cuh"

namespace b2s {

template <int BF16>
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, long long n) {
    const long long n8 = n >> 3;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
        uint4 o;
        o.x = tc::Half16<BF16>::pack2(a.x, a.y);
        o.y = tc::Half16<BF16>::pack2(a.z, a.w);
        o.z = tc::Half16<BF16>::pack2(b.x, b.y);
        o.w = tc::Half16<BF16>::pack2(b.z, b.w);
        reinterpret_cast<uint4*>(out)[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        const long long i = (n8 << 3) + threadIdx.x;
        const uint32_t w = tc::Half16<BF16>::pack2(in[i], 0.f);
        out[i] = (uint16_t)(w & 0xFFFFu);
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_cast_f32_h(const float* in, void* out, int64_t n, int bf16, void* stream) {
    B2S_CHECK_ARG(n >= 0 && (n == 0 || (in && out)), "b2s_cast_f32_h: null pointer");
    B2S_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "b2s_cast_f32_h: pointers must be 16B aligned");
    if (n == 0) return B2S_OK;
    long long blocks = ((n >> 3) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (bf16) cast_kernel<1><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, (uint16_t*)out, n);
    else cast_kernel<0><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, (uint16_t*)out, n);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}
