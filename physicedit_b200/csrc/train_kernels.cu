// Row-wise piece of the attention BACKWARD used by the training path (the reference gets it from autograd through
// F.scaled_dot_product_attention, DiffSynth-Studio/diffsynth/models/qwen_image_dit.py:14-39).  The backward is seven batched tcgen05 GEMM
// launches (pe_gemm_batched with the PE_EPI_ATTN_P / PE_EPI_ATTN_DS epilogues, physicedit_b200/autograd.py) plus this HBM-bound pass:
//   delta[h, s] = sum_d dO[s, h, d] * O[s, h, d]               (fp32; the row term of the softmax Jacobian)
#include "ptx.cuh"
#include "common.cuh"

namespace pe {
namespace {

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per (token, head): lane l owns the 16-byte vector l of the head's 128 columns (lanes >= 16 idle); token-major inputs [S, >= H * 128]
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const bf16* __restrict__ d_o, long long ldd, const bf16* __restrict__ o, long long ldo,
                                                             int S, int H, float* __restrict__ delta, long long ld_delta) {
    const int lane = threadIdx.x & 31;
    const long long total = (long long)S * H;
    for (long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); i < total; i += (long long)gridDim.x * 8) {
        const long long s = i / H;
        const int h = (int)(i - s * H);
        float acc = 0.f;
        if (lane < 16) {
            const uint4 a = *reinterpret_cast<const uint4*>(d_o + s * ldd + h * 128 + lane * 8);
            const uint4 b = *reinterpret_cast<const uint4*>(o + s * ldo + h * 128 + lane * 8);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 x = unpack_bf16(aw[j]), y = unpack_bf16(bw[j]);
                acc = fmaf(x.x, y.x, acc);
                acc = fmaf(x.y, y.y, acc);
            }
        }
        acc = warp_sum_t(acc);
        if (lane == 0) delta[(long long)h * ld_delta + s] = acc;
    }
}

}  // namespace

int attention_bwd_delta_run(Handle* h, const void* d_o, int64_t ldd, const void* o, int64_t ldo, int S, int H, void* delta, int64_t ld_delta, cudaStream_t s) {
    PE_REQUIRE(h, d_o && o && delta && S > 0 && H > 0, "pe_attention_bwd_delta: null pointer or empty input");
    PE_REQUIRE(h, ldd % 8 == 0 && ldo % 8 == 0 && ldd >= (int64_t)H * 128 && ldo >= (int64_t)H * 128 && ld_delta >= S,
               "pe_attention_bwd_delta: row strides must be multiples of 8 and cover H * 128 columns, ld_delta >= S");
    PE_REQUIRE(h, ((reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(o)) & 15) == 0, "pe_attention_bwd_delta: inputs must be 16-byte aligned");
    long long blocks = ((long long)S * H + 7) / 8;
    const long long cap = (long long)h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    attn_bwd_delta_kernel<<<(int)blocks, 256, 0, s>>>(static_cast<const bf16*>(d_o), ldd, static_cast<const bf16*>(o), ldo, S, H, static_cast<float*>(delta),
                                                       ld_delta);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

}  // namespace pe
