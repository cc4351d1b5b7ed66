// HBM-bound kernels of the QwenImageVAE encode / decode path (DiffSynth-Studio/diffsynth/models/qwen_image_vae.py) in the
// channels-last layout the implicit-GEMM convolution (gemm_sm100.cu, pe_conv2d) consumes: an activation map is bf16 [H*W, C] with
// the channels of a pixel contiguous, so a pixel is a "row" and every kernel below is a coalesced 16-byte-vector pass.
//   channel RMS norm (+SiLU)   QwenImageRMS_norm.forward :76-78, nn.SiLU :131-132,145-146
//   nearest-exact 2x upsample  QwenImageUpsample :202-215
//   space-to-depth             nn.ZeroPad2d((0,1,0,1)) + stride-2 conv of the downsample layers :246-249 (turned into a 2x2 stride-1 conv)
//   NCHW <-> NHWC              the latent de-normalisation of decode :724-725 and the normalisation of encode :712-714 ride on these
//   fp32 row softmax, transpose  the single-head attention of the mid block :186-192 (scores come from pe_gemm with PE_EPI_F32)
// bf16 rounding points follow the reference op by op.
#include "ptx.cuh"
#include "common.cuh"

namespace pe {
namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ void unpack8v(const uint4& u, float (&f)[8]) {
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8v(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    return u;
}

// ---- F.normalize(x, dim=C) * sqrt(C) * gamma (+ SiLU) -------------------------------------------------------------------------
//   n = bf16(||x||_2) clamped at 1e-12; y = bf16(x / n); y = bf16(y * sqrt(C)); y = bf16(y * gamma); act: y = bf16(silu(y))
// A pixel (row) is owned by a group of kGroup lanes (16 for C <= 128, else 32), each lane holding kVec 16-byte vectors in
// registers, so the narrow 96-channel maps still keep 24 of 32 lanes busy; every byte is read once.  Rounding to bf16 goes
// through the packed converter (two values per instruction) and integer unpacking: scalar F2F conversions, IEEE divisions and
// expf would make this pass XU-bound instead of HBM-bound.
__device__ __forceinline__ void round_pair(float& a, float& b) {
    const uint32_t p = pack_bf16(a, b);
    a = __uint_as_float(p << 16);
    b = __uint_as_float(p & 0xffff0000u);
}
template <int kGroup, int kVec, int kUnroll>
__global__ void __launch_bounds__(kWarps * 32, 4) channel_rmsnorm_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ out, long long ldo,
                                                                       long long rows, int C, const bf16* __restrict__ gamma, float scale, int act) {
    constexpr int kRowsPerWarp = 32 / kGroup;
    constexpr int kRowsPerIter = kRowsPerWarp * kUnroll;   // kUnroll independent rows per lane group: their loads are all in flight together
    const int lane = threadIdx.x & 31;
    const int gl = lane & (kGroup - 1);                 // lane inside its row group
    const int sub = lane / kGroup;                      // which of the warp's concurrent rows
    const int nvec = C >> 3;
    float g[kVec][8];
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int vi = gl + kGroup * i;
        if (vi < nvec) unpack8v(__ldg(reinterpret_cast<const uint4*>(gamma + vi * 8)), g[i]);
    }
    const long long warp_id = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long warp_stride = (long long)gridDim.x * kWarps;
    for (long long r0 = warp_id * kRowsPerIter; r0 < rows; r0 += warp_stride * kRowsPerIter) {
        uint4 raw[kUnroll][kVec];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long row = r0 + u * kRowsPerWarp + sub;
#pragma unroll
            for (int i = 0; i < kVec; ++i) {
                const int vi = gl + kGroup * i;
                raw[u][i] = make_uint4(0u, 0u, 0u, 0u);
                if (row < rows && vi < nvec) raw[u][i] = *reinterpret_cast<const uint4*>(x + row * ldx + vi * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long row = r0 + u * kRowsPerWarp + sub;
            float v[kVec][8];
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < kVec; ++i) {
                unpack8v(raw[u][i], v[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
            }
#pragma unroll
            for (int o = kGroup / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            // x / n as x * (1 / n): one reciprocal per pixel instead of an IEEE division per element (the two differ by <= 1 fp32 ulp, i.e.
            // the bf16 result differs for ~2^-15 of the elements, by one bf16 ulp)
            const float rn = __frcp_rn(fmaxf(bf16_round(sqrtf(ss)), 1e-12f));
#pragma unroll
            for (int i = 0; i < kVec; ++i) {
                const int vi = gl + kGroup * i;
                if (row < rows && vi < nvec) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = v[i][j] * rn;
#pragma unroll
                    for (int j = 0; j < 8; j += 2) round_pair(o[j], o[j + 1]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] *= scale;
#pragma unroll
                    for (int j = 0; j < 8; j += 2) round_pair(o[j], o[j + 1]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] *= g[i][j];
                    if (act) {
#pragma unroll
                        for (int j = 0; j < 8; j += 2) round_pair(o[j], o[j + 1]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = __fdividef(o[j], 1.0f + __expf(-o[j]));
                    }
                    *reinterpret_cast<uint4*>(out + row * ldo + vi * 8) = pack8v(o);
                }
            }
        }
    }
}

// ---- nearest-exact 2x upsample: out[2y+a, 2x+b, :] = in[y, x, :] ------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int cvec) {
    const long long total = (long long)H * W * cvec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cvec);
        const long long pix = i / cvec;
        const int x = (int)(pix % W);
        const int y = (int)(pix / W);
        const uint4 v = in[i];
        const long long o = ((long long)(2 * y) * (2 * W) + 2 * x) * cvec + c;
        out[o] = v;
        out[o + cvec] = v;
        out[o + (long long)2 * W * cvec] = v;
        out[o + (long long)2 * W * cvec + cvec] = v;
    }
}

// ---- space-to-depth: out[y, x, (py*2+px)*C + c] = in[2y+py, 2x+px, c]  (H, W even) -------------------------------------------
__global__ void space_to_depth_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int cvec) {
    const long long total = (long long)H * W * cvec;
    const int W2 = W >> 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cvec);
        const long long pix = i / cvec;
        const int x = (int)(pix % W);
        const int y = (int)(pix / W);
        const int ph = ((y & 1) << 1) | (x & 1);
        out[(((long long)(y >> 1) * W2 + (x >> 1)) * 4 + ph) * cvec + c] = in[i];
    }
}

// ---- layout changes at the ends of the VAE, with the latent (de)normalisation folded in -------------------------------------
// op 0: copy;  op 1 (decode :724-725): y = bf16(bf16(x / p1[c]) + p0[c]);  op 2 (encode :712-714): y = bf16(bf16(x - p0[c]) * p1[c])
__device__ __forceinline__ float apply_affine(float x, int op, const bf16* p0, const bf16* p1, int c) {
    if (op == 1) return bf16_round(bf16_round(__fdiv_rn(x, __bfloat162float(p1[c]))) + __bfloat162float(p0[c]));
    if (op == 2) return bf16_round(bf16_round(x - __bfloat162float(p0[c])) * __bfloat162float(p1[c]));
    return x;
}
__global__ void nchw_to_nhwc_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long long ldd, int C, long long HW, int op,
                                    const bf16* __restrict__ p0, const bf16* __restrict__ p1) {
    const long long total = HW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / HW);
        const long long pix = i - (long long)c * HW;           // coalesced reads along the plane
        dst[pix * ldd + c] = __float2bfloat16_rn(apply_affine(__bfloat162float(src[i]), op, p0, p1, c));
    }
}
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ src, long long lds, bf16* __restrict__ dst, int C, long long HW, int op,
                                    const bf16* __restrict__ p0, const bf16* __restrict__ p1) {
    const long long total = HW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / HW);
        const long long pix = i - (long long)c * HW;           // coalesced writes along the plane
        dst[i] = __float2bfloat16_rn(apply_affine(__bfloat162float(src[pix * lds + c]), op, p0, p1, c));
    }
}

// ---- bf16 transpose [R, C] -> [C, R] through a padded 32 x 32 shared-memory tile ---------------------------------------------
__global__ void transpose_kernel(const bf16* __restrict__ src, long long lds, bf16* __restrict__ dst, long long ldd, int R, int C) {
    __shared__ bf16 tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[i][threadIdx.x] = src[(long long)r * lds + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) dst[(long long)c * ldd + r] = tile[threadIdx.x][i];
    }
}

// ---- P = softmax(scale * S) over fp32 score rows, written as bf16 (the probabilities SDPA feeds to P.V) ----------------------
// one CTA per row; the row (<= 64 KB at 1024^2 images) is re-read from L1/L2 for the three passes.  Columns [n, n_pad) of the
// output are written as zeros so that a K-padded P.V product ignores them.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, long long lds, bf16* __restrict__ p, long long ldp, int n,
                                                           int n_pad, float scale_log2e, const unsigned char* __restrict__ mask, long long ldm,
                                                           int mask_period) {
    // mask (optional): byte [mask_period, >= n], 0 = this key is hidden from the row (EliGen entity masks: -inf in the reference's additive
    // mask, qwen_image_dit.py:493-496); row r uses mask row r % mask_period, so one mask serves every head of a batched score matrix
    __shared__ float red[8];
    __shared__ float bcast;
    const float* row = s + (long long)blockIdx.x * lds;
    bf16* prow = p + (long long)blockIdx.x * ldp;
    const unsigned char* mrow = mask ? mask + (long long)(blockIdx.x % mask_period) * ldm : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float m = -INFINITY;
    for (int i = tid; i < n; i += 256)
        if (!mrow || mrow[i]) m = fmaxf(m, row[i]);
    m = warp_max_f(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (tid == 0) { float t = red[0]; for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]); bcast = t; }
    __syncthreads();
    m = bcast * scale_log2e;
    float sum = 0.f;
    for (int i = tid; i < n; i += 256)
        if (!mrow || mrow[i]) sum += exp2f(row[i] * scale_log2e - m);
    sum = warp_sum_f(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; bcast = t; }
    __syncthreads();
    const float inv = 1.0f / bcast;
    for (int i = tid; i < n_pad; i += 256)
        prow[i] = __float2bfloat16_rn((i < n && (!mrow || mrow[i])) ? exp2f(row[i] * scale_log2e - m) * inv : 0.0f);
}

inline int grid_for(long long work_items, int threads, int sm_count) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

int channel_rmsnorm_run(Handle* h, const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int C, const void* gamma, int act,
                        cudaStream_t s) {
    PE_REQUIRE(h, x && out && gamma, "pe_channel_rmsnorm: null pointer");
    PE_REQUIRE(h, rows > 0 && C > 0 && C % 8 == 0 && C <= 512, "pe_channel_rmsnorm: C must be a multiple of 8, <= 512 (C=%d)", C);
    PE_REQUIRE(h, ldx % 8 == 0 && ldo % 8 == 0 && ldx >= C && ldo >= C, "pe_channel_rmsnorm: row strides must be multiples of 8 and >= C");
    const float scale = (float)sqrt((double)C);                 // self.scale = dim ** 0.5, applied as an fp32 scalar
    const int group = C <= 128 ? 16 : 32;
    const int rows_per_iter = (32 / group) * (C <= 256 ? 4 : 2);
    long long blocks = (rows + kWarps * rows_per_iter - 1) / (kWarps * rows_per_iter);
    const long long cap = (long long)h->sm_count * 8;
    if (blocks > cap) blocks = cap;
    const bf16* xp = static_cast<const bf16*>(x);
    bf16* op = static_cast<bf16*>(out);
    const bf16* gp = static_cast<const bf16*>(gamma);
    if (C <= 128) channel_rmsnorm_kernel<16, 1, 4><<<(int)blocks, kWarps * 32, 0, s>>>(xp, ldx, op, ldo, rows, C, gp, scale, act);
    else if (C <= 256) channel_rmsnorm_kernel<32, 1, 4><<<(int)blocks, kWarps * 32, 0, s>>>(xp, ldx, op, ldo, rows, C, gp, scale, act);
    else channel_rmsnorm_kernel<32, 2, 2><<<(int)blocks, kWarps * 32, 0, s>>>(xp, ldx, op, ldo, rows, C, gp, scale, act);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int upsample2x_run(Handle* h, const void* in, void* out, int H, int W, int C, cudaStream_t s) {
    PE_REQUIRE(h, in && out && H > 0 && W > 0 && C > 0 && C % 8 == 0, "pe_upsample2x: bad arguments (C must be a multiple of 8)");
    const long long total = (long long)H * W * (C / 8);
    upsample2x_kernel<<<grid_for(total, 256, h->sm_count), 256, 0, s>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), H, W, C / 8);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int space_to_depth_run(Handle* h, const void* in, void* out, int H, int W, int C, cudaStream_t s) {
    PE_REQUIRE(h, in && out && H > 0 && W > 0 && C > 0 && C % 8 == 0, "pe_space_to_depth: bad arguments (C must be a multiple of 8)");
    PE_REQUIRE(h, H % 2 == 0 && W % 2 == 0, "pe_space_to_depth: H and W must be even (H=%d W=%d)", H, W);
    const long long total = (long long)H * W * (C / 8);
    space_to_depth_kernel<<<grid_for(total, 256, h->sm_count), 256, 0, s>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), H, W, C / 8);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int nchw_to_nhwc_run(Handle* h, const void* src, void* dst, int64_t ldd, int C, int64_t HW, int op, const void* p0, const void* p1, cudaStream_t s) {
    PE_REQUIRE(h, src && dst && C > 0 && HW > 0 && ldd >= C, "pe_nchw_to_nhwc: bad arguments");
    PE_REQUIRE(h, op == 0 || (op >= 1 && op <= 2 && p0 && p1), "pe_nchw_to_nhwc: op 1/2 need p0 and p1");
    nchw_to_nhwc_kernel<<<grid_for(HW * C, 256, h->sm_count), 256, 0, s>>>(static_cast<const bf16*>(src), static_cast<bf16*>(dst), ldd, C, HW, op,
                                                                             static_cast<const bf16*>(p0), static_cast<const bf16*>(p1));
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int nhwc_to_nchw_run(Handle* h, const void* src, int64_t lds, void* dst, int C, int64_t HW, int op, const void* p0, const void* p1, cudaStream_t s) {
    PE_REQUIRE(h, src && dst && C > 0 && HW > 0 && lds >= C, "pe_nhwc_to_nchw: bad arguments");
    PE_REQUIRE(h, op == 0 || (op >= 1 && op <= 2 && p0 && p1), "pe_nhwc_to_nchw: op 1/2 need p0 and p1");
    nhwc_to_nchw_kernel<<<grid_for(HW * C, 256, h->sm_count), 256, 0, s>>>(static_cast<const bf16*>(src), lds, static_cast<bf16*>(dst), C, HW, op,
                                                                             static_cast<const bf16*>(p0), static_cast<const bf16*>(p1));
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int transpose_run(Handle* h, const void* src, int64_t lds, void* dst, int64_t ldd, int R, int C, cudaStream_t s) {
    PE_REQUIRE(h, src && dst && R > 0 && C > 0 && lds >= C && ldd >= R, "pe_transpose: bad arguments");
    PE_REQUIRE(h, (C + 31) / 32 <= 65535 * 32 && (R + 31) / 32 <= 65535, "pe_transpose: matrix too large");
    dim3 grid((C + 31) / 32, (R + 31) / 32);
    transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(static_cast<const bf16*>(src), lds, static_cast<bf16*>(dst), ldd, R, C);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int softmax_rows_run(Handle* h, const void* scores, int64_t lds, void* probs, int64_t ldp, int rows, int n, int n_pad, float scale, cudaStream_t s,
                     const void* mask, int64_t ldm, int mask_period) {
    PE_REQUIRE(h, scores && probs && rows > 0 && n > 0 && n_pad >= n, "pe_softmax_rows: need rows > 0 and 0 < n <= n_pad (n=%d n_pad=%d)", n, n_pad);
    PE_REQUIRE(h, lds >= n && ldp >= n_pad, "pe_softmax_rows: row strides must cover n / n_pad");
    PE_REQUIRE(h, mask == nullptr || (ldm >= n && mask_period > 0), "pe_softmax_rows_masked: mask rows must cover n, mask_period > 0");
    softmax_rows_kernel<<<rows, 256, 0, s>>>(static_cast<const float*>(scores), lds, static_cast<bf16*>(probs), ldp, n, n_pad,
                                             scale * 1.4426950408889634f, static_cast<const unsigned char*>(mask), ldm, mask_period > 0 ? mask_period : 1);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

}  // namespace pe
