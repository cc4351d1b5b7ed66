// HBM-bound row-wise and element-wise kernels of the DiT forward: LayerNorm+modulate, RMSNorm, the
// M=1 linears (GEMV), the timestep sinusoid, patchify / unpatchify, CFG + Euler update and the
// special-token gather / blend-scatter.  All of them are plain coalesced 16-byte-vector kernels:
// one warp owns one row (or one output feature for the GEMV) and keeps it in registers, so every
// input byte is read exactly once.  bf16 rounding points follow the reference op by op
//  so results are comparable with the reference's bf16 tensors.
#include <stdlib.h>
#include <cooperative_groups.h>
#include "ptx.cuh"
#include "common.cuh"

namespace pe {
namespace {

constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    return u;
}
// streaming 16-byte load that does not pollute L1
__device__ __forceinline__ uint4 ld_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// -------------------------------------------------------------------------------------------------
// LayerNorm (no affine, eps) followed by one of three tails.
//   MODE 0: out = bf16( bf16( bf16(n) * ops ) + shift )      (_modulate, qwen_image_dit.py:355-357)
//   MODE 1: out = bf16( n * w + b )                           (nn.LayerNorm with affine, helpers.py / DINOv2)
//   MODE 2: out = bf16( n )                                   (non-affine final LN, dinov2.py:20-24)
// kVec = number of 8-element vectors each lane owns (C <= 256 * kVec).
// -------------------------------------------------------------------------------------------------
template <int kVec, int MODE>
__global__ void __launch_bounds__(kWarpsPerCta * 32) layernorm_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int rows, int C,
                                                                       const bf16* __restrict__ p0, const bf16* __restrict__ p1, float eps,
                                                                       int split_row, const bf16* __restrict__ q0, const bf16* __restrict__ q1) {
    const int row = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= rows) return;
    if (MODE == 0 && row >= split_row) { p0 = q0; p1 = q1; }   // second token stream has its own modulation vectors
    const int lane = threadIdx.x & 31;
    const int nvec = C >> 3;
    const bf16* xr = x + (size_t)row * C;
    float v[kVec][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
            unpack8(ld_stream(xr + vi * 8), v[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[i][j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
        }
    }
    const float mean = warp_sum(s) / (float)C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        if (lane + 32 * i < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; ss += d * d; }
        }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
    bf16* orow = out + (size_t)row * C;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int vi = lane + 32 * i;
        if (vi >= nvec) continue;
        float o[8];
        if (MODE == 0) {
            // bf16(bf16(n * (1+scale)) + shift) in packed bf16x2 arithmetic: a bf16 x bf16 product is exact in fp32 and a bf16 + bf16 sum
            // is either exact in fp32 or dominated by one operand, so HMUL2 / HADD2 (one rounding each; the _rn intrinsics keep nvcc from contracting them into one HFMA2) give the bits of the
            // reference's fp32-compute-then-round ops -- without the F2F.BF16 conversions of a scalar formulation, which run on the
            // 16-lane XU pipe and made this HBM-bound kernel XU-bound (r1: 192 F2F per row-warp, 0.41 of the HBM roofline).
            const uint4 a4 = __ldg(reinterpret_cast<const uint4*>(p1 + vi * 8));   // bf16(1 + scale)
            const uint4 b4 = __ldg(reinterpret_cast<const uint4*>(p0 + vi * 8));   // shift
            const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 n2 = __floats2bfloat162_rn((v[i][2 * j] - mean) * rstd, (v[i][2 * j + 1] - mean) * rstd);
                const __nv_bfloat162 t2 = __hmul2_rn(n2, *reinterpret_cast<const __nv_bfloat162*>(&aw[j]));
                const __nv_bfloat162 o2 = __hadd2_rn(t2, *reinterpret_cast<const __nv_bfloat162*>(&bw[j]));
                ow[j] = *reinterpret_cast<const uint32_t*>(&o2);
            }
            *reinterpret_cast<uint4*>(orow + vi * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            continue;
        } else if (MODE == 1) {
            float a[8], b[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(p0 + vi * 8)), a);
            unpack8(__ldg(reinterpret_cast<const uint4*>(p1 + vi * 8)), b);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * a[j] + b[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd;
        }
        *reinterpret_cast<uint4*>(orow + vi * 8) = pack8(o);
    }
}

// -------------------------------------------------------------------------------------------------
// LN + modulate at the DiT width (C = 3072), bulk-copy staged: a persistent CTA per SM, every warp streams its rows through a
// private ring of shared-memory row buffers filled by cp.async.bulk (lane 0 issues the copy of row k+kLnSlots-1 before row k is
// normalised), so ~100 KB of loads per SM are in flight all the time regardless of what the warps compute.  The register-only
// version above issues its loads, then computes, then stores, and reached 0.48 of the HBM roofline inside the power-capped loop
// (its speed followed the SM clock).  Arithmetic and rounding points are those of layernorm_kernel<12, 0>.
// -------------------------------------------------------------------------------------------------
constexpr int kLnC = 3072;
#ifndef PE_LN_WARPS
#define PE_LN_WARPS 8
#endif
#ifndef PE_LN_SLOTS
#define PE_LN_SLOTS 3
#endif
constexpr int kLnWarps = PE_LN_WARPS;
constexpr int kLnSlots = PE_LN_SLOTS;
constexpr int kLnRowBytes = kLnC * 2;
constexpr int kLnSmem = kLnWarps * kLnSlots * kLnRowBytes + 4 * kLnRowBytes + kLnWarps * kLnSlots * 8 + 128;

__global__ void __launch_bounds__(kLnWarps * 32, 1) layernorm_modulate_bulk_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int rows,
                                                                                     const bf16* __restrict__ p0, const bf16* __restrict__ p1, float eps,
                                                                                     int split_row, const bf16* __restrict__ q0, const bf16* __restrict__ q1,
                                                                                     unsigned int* abort_flag) {
    extern __shared__ uint8_t ln_smem_raw[];
    const uint32_t base = (smem_u32(ln_smem_raw) + 127u) & ~127u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring = base + warp * (kLnSlots * kLnRowBytes);
    const uint32_t mods = base + kLnWarps * kLnSlots * kLnRowBytes;        // [shift0 | ops0 | shift1 | ops1], 6 KB each
    const uint32_t bars = mods + 4 * kLnRowBytes + (warp * kLnSlots) * 8;
    const uint8_t* gen = ln_smem_raw + (base - smem_u32(ln_smem_raw));      // generic pointer to `base`
    // modulation vectors -> shared memory once per CTA (stream 1 falls back to stream 0 when there is no second stream)
    {
        const bf16* srcs[4] = {p0, p1, q0 ? q0 : p0, q1 ? q1 : p1};
        uint4* dst = reinterpret_cast<uint4*>(const_cast<uint8_t*>(gen) + kLnWarps * kLnSlots * kLnRowBytes);
        for (int i = threadIdx.x; i < 4 * (kLnRowBytes / 16); i += blockDim.x)
            dst[i] = __ldg(reinterpret_cast<const uint4*>(srcs[i / (kLnRowBytes / 16)]) + i % (kLnRowBytes / 16));
    }
    if (lane == 0)
        for (int s = 0; s < kLnSlots; ++s) mbar_init(bars + s * 8, 1);
    fence_mbar_init();
    __syncthreads();
    const int stride = gridDim.x * kLnWarps;
    const int first = blockIdx.x * kLnWarps + warp;
    auto issue = [&](int k) {                       // k-th row of this warp -> slot k % kLnSlots
        const long long row = first + (long long)k * stride;
        if (row < rows && lane == 0) {
            const uint32_t bar = bars + (k % kLnSlots) * 8;
            fence_proxy_async_smem();               // the slot's previous contents were read through the generic proxy
            mbar_arrive_expect_tx(bar, kLnRowBytes);
            bulk_load_1d(ring + (k % kLnSlots) * kLnRowBytes, x + row * kLnC, kLnRowBytes, bar);
        }
    };
    for (int k = 0; k < kLnSlots - 1; ++k) issue(k);
    for (int k = 0;; ++k) {
        const long long row = first + (long long)k * stride;
        if (row >= rows) break;
        issue(k + kLnSlots - 1);                    // its slot was released (read into registers) in iteration k - 1
        const int slot = k % kLnSlots;
        if (!mbar_wait(bars + slot * 8, (k / kLnSlots) & 1, abort_flag, 100)) break;       // bounded like every wait in this library
        const uint4* xr = reinterpret_cast<const uint4*>(gen + (size_t)warp * (kLnSlots * kLnRowBytes) + slot * kLnRowBytes);
        float v[12][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            unpack8(xr[lane + 32 * i], v[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[i][j];
        }
        // every lane has its part of the row in registers: the slot may be overwritten by the bulk copy issued in the next iteration
        __syncwarp();
        const float mean = warp_sum(sum) / (float)kLnC;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; ss += d * d; }
        const float rstd = rsqrtf(warp_sum(ss) / (float)kLnC + eps);
        const int st = row >= split_row ? 2 : 0;
        const uint4* shift = reinterpret_cast<const uint4*>(gen + kLnWarps * kLnSlots * kLnRowBytes + st * kLnRowBytes);
        const uint4* ops = shift + kLnRowBytes / 16;
        uint4* orow = reinterpret_cast<uint4*>(out + row * kLnC);
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int vi = lane + 32 * i;
            const uint4 a4 = ops[vi], b4 = shift[vi];
            const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 n2 = __floats2bfloat162_rn((v[i][2 * j] - mean) * rstd, (v[i][2 * j + 1] - mean) * rstd);
                const __nv_bfloat162 t2 = __hmul2_rn(n2, *reinterpret_cast<const __nv_bfloat162*>(&aw[j]));
                const __nv_bfloat162 o2 = __hadd2_rn(t2, *reinterpret_cast<const __nv_bfloat162*>(&bw[j]));
                ow[j] = *reinterpret_cast<const uint32_t*>(&o2);
            }
            orow[vi] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
    }
}

template <int MODE>
int launch_layernorm(Handle* h, const void* x, void* out, int rows, int C, const void* p0, const void* p1, float eps, cudaStream_t s,
                     int split_row = 0x7fffffff, const void* r0 = nullptr, const void* r1 = nullptr) {
    PE_REQUIRE(h, rows > 0 && C > 0 && C % 8 == 0 && C <= 4096, "layernorm: need rows>0, C%%8==0, C<=4096 (rows=%d C=%d)", rows, C);
    PE_REQUIRE(h, x && out, "layernorm: null pointer");
    const dim3 grid(ceil_div(rows, kWarpsPerCta)), block(kWarpsPerCta * 32);
    const bf16* xb = static_cast<const bf16*>(x);
    bf16* ob = static_cast<bf16*>(out);
    const bf16* a = static_cast<const bf16*>(p0);
    const bf16* b = static_cast<const bf16*>(p1);
    const bf16* q0 = static_cast<const bf16*>(r0);
    const bf16* q1 = static_cast<const bf16*>(r1);
    static const bool register_only = getenv("PE_LN_REGISTER_ONLY") != nullptr;      // A/B switch for the two LN + modulate kernels
    if (MODE == 0 && C == kLnC && rows >= 4 * kLnWarps && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && !register_only) {
        static bool configured = false;
        if (!configured) {
            PE_CHECK_CUDA(h, cudaFuncSetAttribute(layernorm_modulate_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLnSmem));
            configured = true;
        }
        int ctas = ceil_div(rows, kLnWarps);
        if (ctas > h->sm_count) ctas = h->sm_count;
        layernorm_modulate_bulk_kernel<<<ctas, kLnWarps * 32, kLnSmem, s>>>(xb, ob, rows, a, b, eps, split_row, q0, q1, h->abort_flag);
        PE_CHECK_CUDA(h, cudaGetLastError());
        return PE_OK;
    }
    if (C <= 256) layernorm_kernel<1, MODE><<<grid, block, 0, s>>>(xb, ob, rows, C, a, b, eps, split_row, q0, q1);
    else if (C <= 1024) layernorm_kernel<4, MODE><<<grid, block, 0, s>>>(xb, ob, rows, C, a, b, eps, split_row, q0, q1);
    else if (C <= 3072) layernorm_kernel<12, MODE><<<grid, block, 0, s>>>(xb, ob, rows, C, a, b, eps, split_row, q0, q1);
    else layernorm_kernel<16, MODE><<<grid, block, 0, s>>>(xb, ob, rows, C, a, b, eps, split_row, q0, q1);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

// -------------------------------------------------------------------------------------------------
// RMSNorm (models/utils.py:250-257): y = bf16( bf16( x * rsqrt(mean(x^2)+eps) ) * w )
// -------------------------------------------------------------------------------------------------
template <int kVec>
__global__ void __launch_bounds__(kWarpsPerCta * 32) rmsnorm_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int rows, int C,
                                                                     const bf16* __restrict__ w, float eps) {
    const int row = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int nvec = C >> 3;
    const bf16* xr = x + (size_t)row * C;
    float v[kVec][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
            unpack8(ld_stream(xr + vi * 8), v[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
        }
    }
    const float rs = rsqrtf(warp_sum(ss) / (float)C + eps);
    bf16* orow = out + (size_t)row * C;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int vi = lane + 32 * i;
        if (vi >= nvec) continue;
        float o[8];
        if (w != nullptr) {
            float a[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(w + vi * 8)), a);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = bf16_round(v[i][j] * rs) * a[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = v[i][j] * rs;
        }
        *reinterpret_cast<uint4*>(orow + vi * 8) = pack8(o);
    }
}

// -------------------------------------------------------------------------------------------------
// GEMV for the M=1 linears: one warp per output feature, x staged (after act_in) in shared memory.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_bf16(float x) { return bf16_round(x / (1.0f + expf(-x))); }

// Each warp owns kGemvRows consecutive output features at a time, so one shared-memory read of x feeds kGemvRows dot products
// (r1: with one row per warp a batch-8 launch moved 98 KB of x from shared memory per 6 KB weight row and ran at 0.1 of the HBM
// roofline), and the weight vectors of the next 32-vector slice are in flight while the current slice is multiplied.
// The per-lane partial sums and the butterfly reduction are those of the one-row formulation, so results are bit-identical to it.
constexpr int kGemvRows = 4;

template <int kBatch, int kRows = kGemvRows, int kDepth = 1>
__global__ void __launch_bounds__(kWarpsPerCta * 32) gemv_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                  const bf16* __restrict__ bias, bf16* __restrict__ y, int N, int K,
                                                                  int act_in, int act_out, const uint8_t* __restrict__ one_plus_mask,
                                                                  const bf16* __restrict__ norm_w = nullptr, float norm_eps = 0.f,
                                                                  const bf16* __restrict__ residual = nullptr) {
    // kRows output features per warp, kDepth 16-byte vectors per feature in flight per lane (plus the same again being consumed).
    // (4, 1) streams wide matrices (N >= ~10 k: one shared-memory read of x feeds four dot products); (1, 4) is for narrow outputs
    // (the 3584-wide projections of the Qwen2.5-VL decode step, K up to 18944): four times as many warps share the rows, so all SMs
    // stream, and each lane keeps four loads of its single row in flight.
    // Optional prologues on the staged input (one-token decode of the Qwen2.5-VL text encoder: every CTA redoes them, K <= 18944 values):
    //   act_in 2   x is [kBatch][2K] = gate | up of a SwiGLU MLP: staged value = bf16(bf16(silu(gate)) * up)          (Qwen2MLP.forward)
    //   norm_w     RMSNorm in front of the linear: staged value = bf16(norm_w * bf16(x * rsqrt(mean(x^2) + eps)))     (Qwen2_5_VLRMSNorm)
    // and epilogue: residual != nullptr  ->  y = bf16(residual + bf16(acc + bias))                                     (decoder-layer skip adds)
    // The staged input is kept as bf16 (every value that lands here IS bf16-representable: raw inputs and the rounded prologue results),
    // half the shared memory of an fp32 copy: K = 18944 inputs take 37 KB instead of 74 KB per batch row, so five CTAs instead of two fit
    // an SM and one CTA's prologue overlaps the others' weight streaming (r2: the 3584 x 18944 down-projection went from 2.7 TB/s up).
    extern __shared__ __align__(16) unsigned char xs_raw[];
    bf16* xs = reinterpret_cast<bf16*>(xs_raw);   // [kBatch][K]
    __shared__ float red[kBatch][kWarpsPerCta];
    {
        const int kv = K >> 3;
        for (int i = threadIdx.x; i < kBatch * kv; i += blockDim.x) {
            const int b = i / kv, j = i - b * kv;
            uint4 val;
            if (act_in == 2) {
                float g[8], u[8], f[8];
                unpack8(*reinterpret_cast<const uint4*>(x + (size_t)b * 2 * K + j * 8), g);
                unpack8(*reinterpret_cast<const uint4*>(x + (size_t)b * 2 * K + K + j * 8), u);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = silu_bf16(g[e]) * u[e];
                val = pack8(f);
            } else {
                val = *reinterpret_cast<const uint4*>(x + (size_t)b * K + j * 8);
                if (act_in == 1) {
                    float f[8];
                    unpack8(val, f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = silu_bf16(f[e]);
                    val = pack8(f);
                }
            }
            *reinterpret_cast<uint4*>(xs + b * K + j * 8) = val;
        }
    }
    __syncthreads();
    if (norm_w != nullptr) {
        const int kv = K >> 3;
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            float ss = 0.f;
            for (int j = threadIdx.x; j < kv; j += blockDim.x) {
                float f[8];
                unpack8(*reinterpret_cast<const uint4*>(xs + b * K + j * 8), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) ss += f[e] * f[e];
            }
            ss = warp_sum(ss);
            if ((threadIdx.x & 31) == 0) red[b][threadIdx.x >> 5] = ss;
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            float tot = 0.f;
#pragma unroll
            for (int w_ = 0; w_ < kWarpsPerCta; ++w_) tot += red[b][w_];
            const float rs = rsqrtf(tot / (float)K + norm_eps);
            for (int j = threadIdx.x; j < kv; j += blockDim.x) {
                float f[8], g[8];
                unpack8(*reinterpret_cast<const uint4*>(xs + b * K + j * 8), f);
                unpack8(*reinterpret_cast<const uint4*>(norm_w + j * 8), g);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = g[e] * bf16_round(f[e] * rs);
                *reinterpret_cast<uint4*>(xs + b * K + j * 8) = pack8(f);
            }
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int nvec = K >> 3;
    // act_out 2 (pe_gemv_swiglu): w = [gate rows | up rows] (N = 2 I); a warp takes kRows / 2 gate rows and THEIR up rows, so that
    // y[b, i] = bf16(bf16(silu(gate_i)) * up_i) leaves this kernel and the down-projection needs no SwiGLU prologue
    const bool pair = act_out == 2 && kRows >= 2;
    const int half = N >> 1;
    const int rows_per_group = pair ? kRows / 2 : kRows;
    const int ngroups = ((pair ? half : N) + rows_per_group - 1) / rows_per_group;
    for (int g = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); g < ngroups; g += gridDim.x * kWarpsPerCta) {
        const int n0 = g * rows_per_group;
        const bf16* wr[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int row = pair ? (r < kRows / 2 ? min(n0 + r, half - 1) : half + min(n0 + r - kRows / 2, half - 1)) : min(n0 + r, N - 1);
            wr[r] = w + (size_t)row * K;                                               // rows past the end repeat the last row (not stored)
        }
        float acc[kRows][kBatch];
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
            for (int b = 0; b < kBatch; ++b) acc[r][b] = 0.f;
        uint4 u[kRows][kDepth], un[kRows][kDepth];
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
            for (int d = 0; d < kDepth; ++d) u[r][d] = (lane + 32 * d) < nvec ? ld_stream(wr[r] + (lane + 32 * d) * 8) : make_uint4(0, 0, 0, 0);
        for (int v0 = 0; v0 < nvec; v0 += 32 * kDepth) {
#pragma unroll
            for (int r = 0; r < kRows; ++r)
#pragma unroll
                for (int d = 0; d < kDepth; ++d) {
                    const int vn = v0 + 32 * kDepth + 32 * d + lane;
                    un[r][d] = vn < nvec ? ld_stream(wr[r] + vn * 8) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const int vi = v0 + 32 * d + lane;
                if (vi < nvec) {
                    float f[kRows][8];
#pragma unroll
                    for (int r = 0; r < kRows; ++r) unpack8(u[r][d], f[r]);
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) {
                        float xv[8];
                        unpack8(*reinterpret_cast<const uint4*>(xs + b * K + vi * 8), xv);
#pragma unroll
                        for (int r = 0; r < kRows; ++r)
                            acc[r][b] += f[r][0] * xv[0] + f[r][1] * xv[1] + f[r][2] * xv[2] + f[r][3] * xv[3] + f[r][4] * xv[4] + f[r][5] * xv[5] +
                                         f[r][6] * xv[6] + f[r][7] * xv[7];
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kRows; ++r)
#pragma unroll
                for (int d = 0; d < kDepth; ++d) u[r][d] = un[r][d];
        }
        // reduce; afterwards lane r * kBatch + b holds output (row n0 + r, batch b) and stores it
        float mine = 0.f;
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const float t = warp_sum(acc[r][b]);
                if (lane == r * kBatch + b) mine = t;
            }
        if (pair) {
            const float up = __shfl_down_sync(0xffffffffu, mine, (kRows / 2) * kBatch);      // lane (r, b) of a gate row fetches its up row's sum
            if (lane < (kRows / 2) * kBatch) {
                const int r = lane / kBatch, b = lane - r * kBatch;
                const int n = n0 + r;
                if (n < half) {
                    const float gv = bf16_round(mine + (bias ? __bfloat162float(bias[n]) : 0.f));
                    const float uv = bf16_round(up + (bias ? __bfloat162float(bias[half + n]) : 0.f));
                    y[(size_t)b * half + n] = __float2bfloat16_rn(silu_bf16(gv) * uv);
                }
            }
        } else if (lane < kRows * kBatch) {
            const int r = lane / kBatch, b = lane - r * kBatch;
            const int n = n0 + r;
            if (n < N) {
                const float bv = bias ? __bfloat162float(bias[n]) : 0.f;
                float o = bf16_round(mine + bv);
                if (act_out == 1) o = silu_bf16(o);
                if (one_plus_mask != nullptr && one_plus_mask[n]) o = bf16_round(1.0f + o);
                if (residual != nullptr) o += __bfloat162float(residual[(size_t)b * N + n]);
                y[(size_t)b * N + n] = __float2bfloat16_rn(o);
            }
        }
    }
}

// ---- split-K GEMV for long-K, narrow-N layers (the 3584 x 18944 down-projection of the Qwen2.5-VL decode step) ----------------------------
// With the whole input staged per CTA (gemv_kernel) this shape paid 38 KB (76 KB at batch 2) of shared memory and a K-long SwiGLU prologue in every
// one of 448 CTAs for only 8 weight rows each: 3.3 TB/s at batch 1, 2.0 TB/s at batch 2 (profiles/r02_gemv_bench.json).  Here a thread-block
// CLUSTER of kSplitK CTAs shares a group of rows: CTA r stages (and activates) only K-slice r, every warp accumulates its rows over that slice, the
// partial sums meet in CTA 0 through distributed shared memory in a fixed order (deterministic) and CTA 0 applies bias / residual.  Four times
// less shared memory and prologue work per CTA, all 148 SMs streaming from the first microsecond.
constexpr int kSplitK = 4;
constexpr int kSplitRows = 2;      // rows per warp
constexpr int kSplitDepth = 2;     // 16-byte vectors per row in flight per lane (plus the same again being consumed)

// capped at 64 registers = 4 CTAs per SM (batch 2 would take 75 -> 3 per SM: A/B on B200 41.1 -> 31.0 us; the same cap on the wide kernel spills and loses)
template <int kBatch>
__global__ void __cluster_dims__(kSplitK, 1, 1) __launch_bounds__(kWarpsPerCta * 32, 4)
gemv_splitk_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ bias, bf16* __restrict__ y, int N, int K, int act_in,
                   const bf16* __restrict__ residual) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int kr = (int)cluster.block_rank();
    const int Ks = K / kSplitK, k0 = kr * Ks;
    extern __shared__ __align__(16) unsigned char xs_raw[];
    bf16* xs = reinterpret_cast<bf16*>(xs_raw);                     // [kBatch][Ks]
    __shared__ float part[kWarpsPerCta][kSplitRows][kBatch];
    {
        const int kv = Ks >> 3;
        for (int i = threadIdx.x; i < kBatch * kv; i += blockDim.x) {
            const int b = i / kv, j = i - b * kv;
            uint4 val;
            if (act_in == 2) {                                      // x = gate | up: staged value = bf16(bf16(silu(gate)) * up)
                float g[8], u[8], f[8];
                unpack8(*reinterpret_cast<const uint4*>(x + (size_t)b * 2 * K + k0 + j * 8), g);
                unpack8(*reinterpret_cast<const uint4*>(x + (size_t)b * 2 * K + K + k0 + j * 8), u);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = silu_bf16(g[e]) * u[e];
                val = pack8(f);
            } else {
                val = *reinterpret_cast<const uint4*>(x + (size_t)b * K + k0 + j * 8);
                if (act_in == 1) {
                    float f[8];
                    unpack8(val, f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = silu_bf16(f[e]);
                    val = pack8(f);
                }
            }
            *reinterpret_cast<uint4*>(xs + b * Ks + j * 8) = val;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = Ks >> 3;
    const int rows_per_cluster = kWarpsPerCta * kSplitRows;
    const int ngroups = (N + rows_per_cluster - 1) / rows_per_cluster;
    const int cluster_id = blockIdx.x / kSplitK, n_clusters = gridDim.x / kSplitK;
    for (int g = cluster_id; g < ngroups; g += n_clusters) {          // same trip count in every CTA of a cluster
        const int n0 = g * rows_per_cluster + warp * kSplitRows;
        const bf16* wr[kSplitRows];
#pragma unroll
        for (int r = 0; r < kSplitRows; ++r) wr[r] = w + (size_t)min(n0 + r, N - 1) * K + k0;
        float acc[kSplitRows][kBatch];
#pragma unroll
        for (int r = 0; r < kSplitRows; ++r)
#pragma unroll
            for (int b = 0; b < kBatch; ++b) acc[r][b] = 0.f;
        uint4 u[kSplitRows][kSplitDepth], un[kSplitRows][kSplitDepth];
#pragma unroll
        for (int r = 0; r < kSplitRows; ++r)
#pragma unroll
            for (int d = 0; d < kSplitDepth; ++d) u[r][d] = (lane + 32 * d) < nvec ? ld_stream(wr[r] + (lane + 32 * d) * 8) : make_uint4(0, 0, 0, 0);
        for (int v0 = 0; v0 < nvec; v0 += 32 * kSplitDepth) {
#pragma unroll
            for (int r = 0; r < kSplitRows; ++r)
#pragma unroll
                for (int d = 0; d < kSplitDepth; ++d) {
                    const int vn = v0 + 32 * kSplitDepth + 32 * d + lane;
                    un[r][d] = vn < nvec ? ld_stream(wr[r] + vn * 8) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
            for (int d = 0; d < kSplitDepth; ++d) {
                const int vi = v0 + 32 * d + lane;
                if (vi < nvec) {
                    float f[kSplitRows][8];
#pragma unroll
                    for (int r = 0; r < kSplitRows; ++r) unpack8(u[r][d], f[r]);
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) {
                        float xv[8];
                        unpack8(*reinterpret_cast<const uint4*>(xs + b * Ks + vi * 8), xv);
#pragma unroll
                        for (int r = 0; r < kSplitRows; ++r)
                            acc[r][b] += f[r][0] * xv[0] + f[r][1] * xv[1] + f[r][2] * xv[2] + f[r][3] * xv[3] + f[r][4] * xv[4] + f[r][5] * xv[5] +
                                         f[r][6] * xv[6] + f[r][7] * xv[7];
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kSplitRows; ++r)
#pragma unroll
                for (int d = 0; d < kSplitDepth; ++d) u[r][d] = un[r][d];
        }
#pragma unroll
        for (int r = 0; r < kSplitRows; ++r)
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const float t = warp_sum(acc[r][b]);
                if (lane == 0) part[warp][r][b] = t;
            }
        cluster.sync();                                               // every CTA's partial sums are in its shared memory
        if (kr == 0 && lane < kSplitRows * kBatch) {
            const int r = lane / kBatch, b = lane - r * kBatch;
            const int n = n0 + r;
            if (n < N) {
                float sum = 0.f;
#pragma unroll
                for (int q = 0; q < kSplitK; ++q) sum += *cluster.map_shared_rank(&part[warp][r][b], q);      // K-slices in order
                const float bv = bias ? __bfloat162float(bias[n]) : 0.f;
                float o = bf16_round(sum + bv);
                if (residual != nullptr) o += __bfloat162float(residual[(size_t)b * N + n]);
                y[(size_t)b * N + n] = __float2bfloat16_rn(o);
            }
        }
        cluster.sync();                                               // CTA 0 has read them: they may be overwritten (or the CTA may exit)
    }
}

// y = bf16(act(x)) element-wise (act 1 = SiLU as the reference's nn.SiLU on a bf16 tensor).  The conditioning path applies
// SiLU(temb) once and feeds it to all 121 modulation GEMVs of a timestep batch instead of re-evaluating it in every CTA.
__global__ void act_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n, int act) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float f = __bfloat162float(x[i]);
    if (act == 1) f = silu_bf16(f);
    y[i] = __float2bfloat16_rn(f);
}

// -------------------------------------------------------------------------------------------------
// timestep sinusoid with the reference's bf16 quirks (models/utils.py:189-216)
// -------------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const bf16* __restrict__ t_in, bf16* __restrict__ out, int raw) {
    const int i = threadIdx.x;   // 0..127
    // timestep / 1000 on a bf16 CUDA tensor = bf16( float(t) * float(1/1000.) )  (ATen div-by-scalar)
    const float t0 = __bfloat162float(t_in[0]);
    const float ts = raw ? bf16_round(t0 * (float)(1.0 / 1000.0)) : t0;
    const float exponent = __fdiv_rn(__fmul_rn(-9.210340371976184f, (float)i), 128.0f);
    const float freq = bf16_round(expf(exponent));        // align_dtype_to_timestep: freqs rounded to bf16
    const float arg = __fmul_rn(1000.0f, __fmul_rn(ts, freq));
    out[i] = __float2bfloat16_rn(cosf(arg));               // flip_sin_to_cos: cos half first
    out[128 + i] = __float2bfloat16_rn(sinf(arg));
}

// -------------------------------------------------------------------------------------------------
// patchify / unpatchify:  "C (H P) (W Q) -> (H W) (C P Q)", P = Q = 2, C = 16
// -------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const bf16* __restrict__ lat, bf16* __restrict__ tok, int H8, int W8) {
    const int W2 = W8 >> 1;
    const int ntok = (H8 >> 1) * W2;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;   // (token, c)
    if (gid >= ntok * 16) return;
    const int c = gid & 15;
    const int t = gid >> 4;
    const int hh = t / W2, ww = t - hh * W2;
    const bf16* src = lat + ((size_t)c * H8 + 2 * hh) * W8 + 2 * ww;
    const uint32_t r0 = *reinterpret_cast<const uint32_t*>(src);
    const uint32_t r1 = *reinterpret_cast<const uint32_t*>(src + W8);
    *reinterpret_cast<uint2*>(tok + (size_t)t * 64 + c * 4) = make_uint2(r0, r1);
}
__global__ void unpatchify_kernel(const bf16* __restrict__ tok, long long ld, bf16* __restrict__ lat, int H8, int W8) {
    const int W2 = W8 >> 1;
    const int ntok = (H8 >> 1) * W2;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;   // (c, token): consecutive threads -> consecutive w
    if (gid >= ntok * 16) return;
    const int c = gid / ntok;
    const int t = gid - c * ntok;
    const int hh = t / W2, ww = t - hh * W2;
    const uint2 v = *reinterpret_cast<const uint2*>(tok + (size_t)t * ld + c * 4);
    bf16* dst = lat + ((size_t)c * H8 + 2 * hh) * W8 + 2 * ww;
    *reinterpret_cast<uint32_t*>(dst) = v.x;
    *reinterpret_cast<uint32_t*>(dst + W8) = v.y;
}

// -------------------------------------------------------------------------------------------------
// CFG combine + Euler update (qwen_image_physical.py:656, flow_match.py:81), every op rounded to bf16
// -------------------------------------------------------------------------------------------------
__global__ void cfg_euler_kernel(bf16* __restrict__ lat, const bf16* __restrict__ posi, const bf16* __restrict__ nega, long long n,
                                 float cfg, float dsigma) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p = __bfloat162float(posi[i]);
    float np = p;
    if (nega != nullptr) {
        const float q = __bfloat162float(nega[i]);
        np = bf16_round(q + bf16_round(cfg * bf16_round(p - q)));
    }
    lat[i] = __float2bfloat16_rn(__bfloat162float(lat[i]) + bf16_round(np * dsigma));
}

// -------------------------------------------------------------------------------------------------
// special tokens: ordered gather of the masked rows, and blend + scatter back
// -------------------------------------------------------------------------------------------------
__global__ void special_index_kernel(const uint8_t* __restrict__ mask, int T, int32_t* __restrict__ idx, int max_rows,
                                     unsigned int* __restrict__ async_err) {
    // single CTA of 1024 threads; ordered compaction via per-chunk ballot + running offset
    __shared__ int warp_cnt[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    for (int i = threadIdx.x; i <= max_rows; i += blockDim.x) idx[i] = -1;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t0 = 0; t0 < T; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        const bool m = t < T && mask[t] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, m);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; ++w) off += warp_cnt[w];
        const int pos = off + __popc(bal & ((1u << lane) - 1u));
        if (m && pos < max_rows) idx[pos] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_cnt[w];
            base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        idx[max_rows] = base;
        // more masked rows than the caller made room for: the reference handles any count (qwen_image_physical.py:1334), so dropping
        // rows would be a silent wrong answer -> raise the handle's asynchronous argument error (pe_check_async_error reports it)
        if (base > max_rows) atomicMax(async_err + 1, (unsigned int)base);
    }
}
__global__ void special_gather_rows_kernel(const bf16* __restrict__ pe, const int32_t* __restrict__ idx, int C, bf16* __restrict__ dst) {
    const int r = blockIdx.x;
    const int t = idx[r];
    const int nvec = C >> 3;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (t >= 0) u = *reinterpret_cast<const uint4*>(pe + (size_t)t * C + v * 8);
        *reinterpret_cast<uint4*>(dst + (size_t)r * C + v * 8) = u;
    }
}
__global__ void special_blend_scatter_kernel(bf16* __restrict__ pe, const int32_t* __restrict__ idx, int C, const bf16* __restrict__ pd,
                                             const bf16* __restrict__ pv, const bf16* __restrict__ t_in, float t_min, float inv_range) {
    const int r = blockIdx.x;
    const int t = idx[r];
    if (t < 0) return;
    // helpers.py:142-150 on a bf16 CUDA timestep: (t - t_min) -> bf16, / range (ATen: * float(1/range)) -> bf16, clamp
    float alpha = bf16_round(__bfloat162float(t_in[0]) - t_min);
    alpha = bf16_round(alpha * inv_range);
    alpha = fminf(fmaxf(alpha, 0.f), 1.f);
    const float oma = bf16_round(1.0f - alpha);
    const int nvec = C >> 3;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        float a[8], b[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(pd + (size_t)r * C + v * 8), a);
        unpack8(*reinterpret_cast<const uint4*>(pv + (size_t)r * C + v * 8), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = bf16_round(alpha * a[j]) + bf16_round(oma * b[j]);
        *reinterpret_cast<uint4*>(pe + (size_t)t * C + v * 8) = pack8(o);
    }
}

// x[r, :] += alpha * add[(r % period), :]     (pos-emb / frame-embedding adds of the resampler path)
__global__ void add_rows_kernel(bf16* __restrict__ x, const bf16* __restrict__ add, int rows, int C, int period, float alpha) {
    const int nvec = C >> 3;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)rows * nvec) return;
    const int r = (int)(gid / nvec), v = (int)(gid - (long long)r * nvec);
    float a[8], b[8];
    unpack8(*reinterpret_cast<const uint4*>(x + (size_t)r * C + v * 8), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(add + (size_t)(r % period) * C + v * 8)), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += alpha * b[j];
    *reinterpret_cast<uint4*>(x + (size_t)r * C + v * 8) = pack8(a);
}

}  // namespace

int layernorm_modulate_run(Handle* h, const void* x, void* out, int rows, int C, const void* shift, const void* ops, cudaStream_t s) {
    PE_REQUIRE(h, shift && ops, "pe_layernorm_modulate: shift / one_plus_scale must not be null");
    return launch_layernorm<0>(h, x, out, rows, C, shift, ops, 1e-6f, s);
}

int layernorm_modulate2_run(Handle* h, const void* x, void* out, int rows, int C, int split_row, const void* shift0, const void* ops0,
                            const void* shift1, const void* ops1, cudaStream_t s) {
    PE_REQUIRE(h, shift0 && ops0 && shift1 && ops1, "pe_layernorm_modulate2: modulation vectors must not be null");
    PE_REQUIRE(h, split_row >= 0 && split_row <= rows, "pe_layernorm_modulate2: split_row out of range");
    return launch_layernorm<0>(h, x, out, rows, C, shift0, ops0, 1e-6f, s, split_row, shift1, ops1);
}

int layernorm_affine_run(Handle* h, const void* x, void* out, int rows, int C, const void* w, const void* b, float eps, cudaStream_t s) {
    if (w == nullptr && b == nullptr) return launch_layernorm<2>(h, x, out, rows, C, nullptr, nullptr, eps, s);
    PE_REQUIRE(h, w && b, "pe_layernorm: weight and bias must both be given or both be null");
    return launch_layernorm<1>(h, x, out, rows, C, w, b, eps, s);
}

int rmsnorm_run(Handle* h, const void* x, void* out, int rows, int C, const void* w, float eps, cudaStream_t s) {
    PE_REQUIRE(h, rows > 0 && C > 0 && C % 8 == 0 && C <= 4096, "pe_rmsnorm: need rows>0, C%%8==0, C<=4096 (rows=%d C=%d)", rows, C);
    PE_REQUIRE(h, x && out, "pe_rmsnorm: null pointer");
    const dim3 grid(ceil_div(rows, kWarpsPerCta)), block(kWarpsPerCta * 32);
    const bf16* xb = static_cast<const bf16*>(x);
    bf16* ob = static_cast<bf16*>(out);
    const bf16* wb = static_cast<const bf16*>(w);
    if (C <= 256) rmsnorm_kernel<1><<<grid, block, 0, s>>>(xb, ob, rows, C, wb, eps);
    else if (C <= 1024) rmsnorm_kernel<4><<<grid, block, 0, s>>>(xb, ob, rows, C, wb, eps);
    else if (C <= 3072) rmsnorm_kernel<12><<<grid, block, 0, s>>>(xb, ob, rows, C, wb, eps);
    else rmsnorm_kernel<16><<<grid, block, 0, s>>>(xb, ob, rows, C, wb, eps);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int gemv_run(Handle* h, const void* x, const void* w, const void* bias, void* y, int batch, int N, int K, int act_in, int act_out,
             const uint8_t* one_plus_mask, cudaStream_t s, const void* norm_w, float norm_eps, const void* residual) {
    PE_REQUIRE(h, batch >= 1 && batch <= 8, "pe_gemv: batch must be 1..8 (got %d)", batch);
    PE_REQUIRE(h, act_in >= 0 && act_in <= 2 && !(act_in == 2 && norm_w != nullptr), "pe_gemv: act_in must be 0, 1 (SiLU) or 2 (SwiGLU over gate|up), and not combined with a norm");
    PE_REQUIRE(h, N > 0 && K > 0 && K % 8 == 0, "pe_gemv: N>0, K%%8==0 required (N=%d K=%d)", N, K);
    PE_REQUIRE(h, x && w && y, "pe_gemv: null pointer");
    PE_REQUIRE(h, (size_t)batch * K * 2 <= 200 * 1024, "pe_gemv: batch*K too large for shared memory");
    const bf16* xb0 = static_cast<const bf16*>(x);
    static const int split_mode = getenv("PE_GEMV_SPLITK") ? atoi(getenv("PE_GEMV_SPLITK")) : 1;            // experiments: 0 disables the split-K kernel
    if (split_mode && batch <= 2 && norm_w == nullptr && act_out == 0 && one_plus_mask == nullptr && K >= 8192 && N <= 8192 && K % (8 * kSplitK) == 0) {
        // long-K, narrow-N (the decode step's 3584 x 18944 down-projection): K split over a cluster of kSplitK CTAs
        const int clusters = ceil_div(N, kWarpsPerCta * kSplitRows);
        const size_t sm = (size_t)batch * (K / kSplitK) * sizeof(bf16);
        const bf16* wb0 = static_cast<const bf16*>(w);
        if (batch == 1)
            gemv_splitk_kernel<1><<<clusters * kSplitK, kWarpsPerCta * 32, sm, s>>>(xb0, wb0, static_cast<const bf16*>(bias), static_cast<bf16*>(y), N, K, act_in,
                                                                                   static_cast<const bf16*>(residual));
        else
            gemv_splitk_kernel<2><<<clusters * kSplitK, kWarpsPerCta * 32, sm, s>>>(xb0, wb0, static_cast<const bf16*>(bias), static_cast<bf16*>(y), N, K, act_in,
                                                                                   static_cast<const bf16*>(residual));
        PE_CHECK_CUDA(h, cudaGetLastError());
        return PE_OK;
    }
    const size_t smem = (size_t)batch * K * sizeof(bf16);
    // narrow outputs (fewer row groups than ~2 per warp slot of the machine): one row per warp, four loads in flight per lane
    static const int mode = getenv("PE_GEMV_MODE") ? atoi(getenv("PE_GEMV_MODE")) : 0;      // experiments: 1 forces the wide kernel, 2 the narrow one
    PE_REQUIRE(h, act_out != 2 || (N % 2 == 0 && residual == nullptr && one_plus_mask == nullptr), "pe_gemv_swiglu: N = 2 I, no residual");
    const bool narrow = act_out != 2 && (mode == 2 || (mode == 0 && batch <= 2 && ceil_div(N, kGemvRows) < 2 * h->sm_count * kWarpsPerCta));
    const int grid_needed = ceil_div(ceil_div(N, narrow ? 1 : kGemvRows), kWarpsPerCta);
    int grid = grid_needed;
    const bf16* xb = static_cast<const bf16*>(x);
    const bf16* wb = static_cast<const bf16*>(w);
    const bf16* bb = static_cast<const bf16*>(bias);
    bf16* yb = static_cast<bf16*>(y);
#define PE_GEMV_LAUNCH(KERN)                                                                                                  \
    {                                                                                                                        \
        if (smem > 40 * 1024) PE_CHECK_CUDA(h, cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); /* static smem counts toward the 48 KB default too */ \
        /* the row loop is grid-stride: ONE resident wave (SMs x occupancy of this instantiation), never a partial second one (r2: the batch-2 */ \
        /* wide kernel needs 80 registers -> 3 CTAs per SM; a grid of 4 per SM ran a 1/3-full second wave: 62 us instead of 48 for 272 MB) */ \
        int occ = 0;                                                                                                         \
        PE_CHECK_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, KERN, kWarpsPerCta * 32, smem));                  \
        grid = grid_needed < h->sm_count * (occ > 0 ? occ : 1) ? grid_needed : h->sm_count * (occ > 0 ? occ : 1);            \
        KERN<<<grid, kWarpsPerCta * 32, smem, s>>>(xb, wb, bb, yb, N, K, act_in, act_out, one_plus_mask,                        \
                                                   static_cast<const bf16*>(norm_w), norm_eps, static_cast<const bf16*>(residual)); \
    }
#define PE_GEMV_CASE(B)                                                                                                      \
    case B: {                                                                                                                \
        PE_GEMV_LAUNCH((gemv_kernel<B, kGemvRows, 1>))                                                                        \
        break;                                                                                                               \
    }
    // wide batch-2 GEMVs run (2 rows, depth 2) instead of (4, 1): 64 registers instead of 80 -> 4 CTAs per SM instead of 3, same per-row summation order
    // (A/B on B200: gate/up 53.3 -> 48.3 us, lm_head 180.9 -> 173.2 us).  PE_GEMV_B2=0 restores (4, 1).
    static const int b2_mode = getenv("PE_GEMV_B2") ? atoi(getenv("PE_GEMV_B2")) : 1;
    if (narrow && batch == 1) PE_GEMV_LAUNCH((gemv_kernel<1, 1, 4>))
    else if (narrow && batch == 2) PE_GEMV_LAUNCH((gemv_kernel<2, 1, 4>))
    else if (b2_mode == 1 && batch == 2 && smem <= 40 * 1024) {
        grid = 0;
        const int grid_needed2 = ceil_div(ceil_div(N, 2), kWarpsPerCta);
        int occ = 0;
        PE_CHECK_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gemv_kernel<2, 2, 2>, kWarpsPerCta * 32, smem));
        grid = grid_needed2 < h->sm_count * (occ > 0 ? occ : 1) ? grid_needed2 : h->sm_count * (occ > 0 ? occ : 1);
        gemv_kernel<2, 2, 2><<<grid, kWarpsPerCta * 32, smem, s>>>(xb, wb, bb, yb, N, K, act_in, act_out, one_plus_mask, static_cast<const bf16*>(norm_w), norm_eps,
                                                                   static_cast<const bf16*>(residual));
    }
    else switch (batch) {
        PE_GEMV_CASE(1) PE_GEMV_CASE(2) PE_GEMV_CASE(3) PE_GEMV_CASE(4) PE_GEMV_CASE(5) PE_GEMV_CASE(6) PE_GEMV_CASE(7) PE_GEMV_CASE(8)
    }
#undef PE_GEMV_LAUNCH
#undef PE_GEMV_CASE
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int act_run(Handle* h, const void* x, void* y, long long n, int act, cudaStream_t s) {
    PE_REQUIRE(h, x && y && n > 0, "pe_act: null pointer or empty tensor");
    PE_REQUIRE(h, act == 0 || act == 1, "pe_act: act must be 0 (copy) or 1 (SiLU), got %d", act);
    act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(static_cast<const bf16*>(x), static_cast<bf16*>(y), n, act);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int timestep_embedding_run(Handle* h, const void* t_in, void* out, int raw, cudaStream_t s) {
    PE_REQUIRE(h, t_in && out, "pe_timestep_embedding: null pointer");
    timestep_embedding_kernel<<<1, 128, 0, s>>>(static_cast<const bf16*>(t_in), static_cast<bf16*>(out), raw);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int patchify_run(Handle* h, const void* latents, void* tokens, int H8, int W8, cudaStream_t s) {
    PE_REQUIRE(h, latents && tokens && H8 > 0 && W8 > 0 && H8 % 2 == 0 && W8 % 2 == 0, "pe_patchify: H8, W8 must be positive and even");
    const int n = (H8 / 2) * (W8 / 2) * 16;
    patchify_kernel<<<ceil_div(n, 256), 256, 0, s>>>(static_cast<const bf16*>(latents), static_cast<bf16*>(tokens), H8, W8);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int unpatchify_run(Handle* h, const void* tokens, int64_t ld, void* latents, int H8, int W8, cudaStream_t s) {
    PE_REQUIRE(h, latents && tokens && H8 > 0 && W8 > 0 && H8 % 2 == 0 && W8 % 2 == 0, "pe_unpatchify: H8, W8 must be positive and even");
    PE_REQUIRE(h, ld >= 64 && ld % 4 == 0, "pe_unpatchify: ld must be >= 64 and a multiple of 4");
    const int n = (H8 / 2) * (W8 / 2) * 16;
    unpatchify_kernel<<<ceil_div(n, 256), 256, 0, s>>>(static_cast<const bf16*>(tokens), (long long)ld, static_cast<bf16*>(latents), H8, W8);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int cfg_euler_run(Handle* h, void* latents, const void* posi, const void* nega, int64_t n, float cfg, float dsigma, cudaStream_t s) {
    PE_REQUIRE(h, latents && posi && n > 0, "pe_cfg_euler_step: null pointer or n<=0");
    cfg_euler_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(static_cast<bf16*>(latents), static_cast<const bf16*>(posi),
                                                                static_cast<const bf16*>(nega), (long long)n, cfg, dsigma);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int special_gather_run(Handle* h, const void* prompt_emb, const uint8_t* mask, int T, int C, void* dst, int32_t* idx, int max_rows,
                       cudaStream_t s) {
    PE_REQUIRE(h, prompt_emb && mask && dst && idx, "pe_special_gather: null pointer");
    PE_REQUIRE(h, T > 0 && C > 0 && C % 8 == 0 && max_rows > 0, "pe_special_gather: bad sizes");
    special_index_kernel<<<1, 1024, 0, s>>>(mask, T, idx, max_rows, h->abort_flag);
    special_gather_rows_kernel<<<max_rows, 128, 0, s>>>(static_cast<const bf16*>(prompt_emb), idx, C, static_cast<bf16*>(dst));
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int special_blend_scatter_run(Handle* h, void* prompt_emb, const int32_t* idx, int max_rows, int C, const void* pd, const void* pv,
                              const void* t_in, float t_min, float t_max, cudaStream_t s) {
    PE_REQUIRE(h, prompt_emb && idx && pd && pv && t_in, "pe_special_blend_scatter: null pointer");
    PE_REQUIRE(h, C > 0 && C % 8 == 0 && max_rows > 0, "pe_special_blend_scatter: bad sizes");
    // Python evaluates (t_max - t_min + 1e-6) in double; ATen then multiplies by float(1/that)
    const float inv_range = (float)(1.0 / ((double)t_max - (double)t_min + 1e-6));
    special_blend_scatter_kernel<<<max_rows, 128, 0, s>>>(static_cast<bf16*>(prompt_emb), idx, C, static_cast<const bf16*>(pd),
                                                          static_cast<const bf16*>(pv), static_cast<const bf16*>(t_in), t_min, inv_range);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int add_bias_rows_run(Handle* h, void* x, const void* add, int rows, int C, int period, float alpha, cudaStream_t s) {
    PE_REQUIRE(h, x && add && rows > 0 && C > 0 && C % 8 == 0 && period > 0, "pe_add_rows: bad arguments");
    const long long n = (long long)rows * (C / 8);
    add_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(static_cast<bf16*>(x), static_cast<const bf16*>(add), rows, C, period, alpha);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

}  // namespace pe
