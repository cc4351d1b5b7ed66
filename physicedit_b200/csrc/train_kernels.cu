// Row-wise pieces of the attention BACKWARD used by the training path (the reference gets them from autograd through
// F.scaled_dot_product_attention, DiffSynth-Studio/diffsynth/models/qwen_image_dit.py:14-39).  The backward of one head is composed on
// the host (physicedit_b200/autograd.py) from the tcgen05 GEMM (scores, dP, dQ, dK, dV), pe_softmax_rows (P recomputed from the scores)
// and the two HBM-bound passes below:
//   delta[r]  = sum_d dO[r, d] * O[r, d]                      (fp32; the row term of the softmax Jacobian)
//   dS[r, c]  = bf16( P[r, c] * (dP[r, c] - delta[r]) * scale ) (what the dQ / dK GEMMs consume)
#include "ptx.cuh"
#include "common.cuh"

namespace pe {
namespace {

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per row, D <= 256 (multiple of 8): lane l owns the 16-byte vector l of the row
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const bf16* __restrict__ d_o, long long ldd, const bf16* __restrict__ o, long long ldo,
                                                             int rows, int D, float* __restrict__ delta) {
    const int lane = threadIdx.x & 31;
    const int nvec = D >> 3;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * 8) {
        float acc = 0.f;
        if (lane < nvec) {
            const uint4 a = *reinterpret_cast<const uint4*>(d_o + r * ldd + lane * 8);
            const uint4 b = *reinterpret_cast<const uint4*>(o + r * ldo + lane * 8);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 x = unpack_bf16(aw[i]), y = unpack_bf16(bw[i]);
                acc = fmaf(x.x, y.x, acc);
                acc = fmaf(x.y, y.y, acc);
            }
        }
        acc = warp_sum_t(acc);
        if (lane == 0) delta[r] = acc;
    }
}

// grid-stride over 8-element column groups of every row: 16 B of P, 32 B of dP in, 16 B of dS out per thread-iteration
__global__ void __launch_bounds__(256) attn_bwd_ds_kernel(const bf16* __restrict__ p, long long ldp, const float* __restrict__ dp, long long lddp,
                                                          const float* __restrict__ delta, bf16* __restrict__ ds, long long ldds, int rows, int cols8,
                                                          float scale) {
    const long long total = (long long)rows * cols8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols8;
        const int c = (int)(i - r * cols8) * 8;
        const uint4 pv = *reinterpret_cast<const uint4*>(p + r * ldp + c);
        const float4 d0 = *reinterpret_cast<const float4*>(dp + r * lddp + c);
        const float4 d1 = *reinterpret_cast<const float4*>(dp + r * lddp + c + 4);
        const float dl = delta[r];
        const float2 p0 = unpack_bf16(pv.x), p1 = unpack_bf16(pv.y), p2 = unpack_bf16(pv.z), p3 = unpack_bf16(pv.w);
        uint4 out;
        out.x = pack_bf16(p0.x * (d0.x - dl) * scale, p0.y * (d0.y - dl) * scale);
        out.y = pack_bf16(p1.x * (d0.z - dl) * scale, p1.y * (d0.w - dl) * scale);
        out.z = pack_bf16(p2.x * (d1.x - dl) * scale, p2.y * (d1.y - dl) * scale);
        out.w = pack_bf16(p3.x * (d1.z - dl) * scale, p3.y * (d1.w - dl) * scale);
        *reinterpret_cast<uint4*>(ds + r * ldds + c) = out;
    }
}

}  // namespace

int attention_bwd_delta_run(Handle* h, const void* d_o, int64_t ldd, const void* o, int64_t ldo, int rows, int D, void* delta, cudaStream_t s) {
    PE_REQUIRE(h, d_o && o && delta && rows > 0, "pe_attention_bwd_delta: null pointer or no rows");
    PE_REQUIRE(h, D > 0 && D % 8 == 0 && D <= 256 && ldd % 8 == 0 && ldo % 8 == 0 && ldd >= D && ldo >= D,
               "pe_attention_bwd_delta: D must be a multiple of 8, <= 256, row strides multiples of 8 (D=%d)", D);
    long long blocks = ((long long)rows + 7) / 8;
    const long long cap = (long long)h->sm_count * 8;
    if (blocks > cap) blocks = cap;
    attn_bwd_delta_kernel<<<(int)blocks, 256, 0, s>>>(static_cast<const bf16*>(d_o), ldd, static_cast<const bf16*>(o), ldo, rows, D,
                                                       static_cast<float*>(delta));
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int attention_bwd_ds_run(Handle* h, const void* p, int64_t ldp, const void* dp, int64_t lddp, const void* delta, void* ds, int64_t ldds, int rows,
                         int cols, float scale, cudaStream_t s) {
    PE_REQUIRE(h, p && dp && delta && ds && rows > 0 && cols > 0, "pe_attention_bwd_ds: null pointer or empty matrix");
    PE_REQUIRE(h, cols % 8 == 0 && ldp % 8 == 0 && ldds % 8 == 0 && lddp % 4 == 0 && ldp >= cols && lddp >= cols && ldds >= cols,
               "pe_attention_bwd_ds: cols and the bf16 row strides must be multiples of 8, the fp32 stride a multiple of 4 (cols=%d)", cols);
    PE_REQUIRE(h, (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (reinterpret_cast<uintptr_t>(dp) & 15) == 0 && (reinterpret_cast<uintptr_t>(ds) & 15) == 0,
               "pe_attention_bwd_ds: buffers must be 16-byte aligned");
    const long long total = (long long)rows * (cols / 8);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    attn_bwd_ds_kernel<<<(int)blocks, 256, 0, s>>>(static_cast<const bf16*>(p), ldp, static_cast<const float*>(dp), lddp, static_cast<const float*>(delta),
                                                    static_cast<bf16*>(ds), ldds, rows, cols / 8, scale);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

}  // namespace pe
