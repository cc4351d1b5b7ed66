// Grouped linear layer  out = epilogue(A · Wᵀ + bias)  on the 5th-gen tensor cores.
//
// One persistent, warp-specialised kernel (bf16 operands, fp32 accumulation in TMEM):
//   warp 0      TMA producer   : A[128 x 64] / W[256 x 64] tiles, 128-byte swizzle, mbarrier ring
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma (UMMA 128x256x16, or
//                                256x256x16 on a CTA pair with cta_group::2)
//   warp 2      TMEM allocator : 512 columns = 2 accumulator stages x 256 fp32 columns
//   warps 4..11 epilogue       : tcgen05.ld -> registers -> fused epilogue -> global.  Two warps per TMEM lane
//                                quarter, each owning one 128-column half of the 256-column accumulator
//                                (for the QKV epilogue: one attention head), so the epilogue of a tile costs
//                                half as many cycles per warp and stays hidden behind the next tile's MMAs.
// The accumulator is double-buffered in TMEM, so the epilogue of tile i overlaps the main loop of
// tile i+1.  A "segment" is a token stream with its own weights (image / text stream of the
// double-stream DiT block): both streams run in ONE launch, each with its own tensor maps, which
// also gives free M-tail handling (TMA zero-fills out-of-bounds rows, stores are row-guarded).
//
// Convolution mode (pe_conv2d, the VAE): the same kernel with a different producer -- the A tile is one 128-pixel patch of a
// channels-last activation map read tap by tap through a 3-D tensor map (zero padding = TMA out-of-bounds fill, no im2col), K runs
// over (tap, 64-channel block); narrow layers trim the MMA / W box to round_up(N, 16) columns, deepen the ring, and (N <= 128) keep
// two patches' accumulators in one TMEM stage so that both share every weight box.  Replaces QwenImageCausalConv3d / nn.Conv2d in
// DiffSynth-Studio/diffsynth/models/qwen_image_vae.py:43-51,243,247-249 and the residual add :152.
//
// Reference call sites replaced: F.linear in QwenDoubleStreamAttention.forward / QwenFeedForward /
// ApproximateGELU (DiffSynth-Studio/diffsynth/models/qwen_image_dit.py:42-49,228-316), the
// gate/residual adds of QwenImageTransformerBlock.forward (:386-399), RMSNorm (models/utils.py:241-257),
// apply_rotary_emb_qwen (qwen_image_dit.py:51-57) and the three torch.cat's (:304-306).
#include <stdlib.h>
#include "ptx.cuh"
#include "common.cuh"

namespace pe {

namespace {

constexpr int kBlockK = 64;        // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kTileN = 256;
constexpr int kThreads = 384;       // 4 control warps + 8 epilogue warps
constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 12;
constexpr int kTileRegionBytes = 224 << 10;   // operand ring; full-width tiles use 192 KB of it: 4 x 48 KB (one CTA, 256 W rows) or 6 x 32 KB (CTA pair)
constexpr bool kRoundRasterDefault = true;   // deep-K layers: one rasterisation group = one round of tiles (choose_raster)
constexpr int kPanelBytes = 32 << 20;   // rasterisation: the m-tiles of a group share a sweep over n; their A panel (<= 32 MB) stays in the 126 MB L2

struct SegDev {
    CUtensorMap tmA;
    CUtensorMap tmB;
    const bf16* bias;
    bf16* out;
    const bf16* gate;
    bf16* out_k;
    bf16* out_v;
    const bf16* norm_q_w;
    const bf16* norm_k_w;
    const float2* rope;
    long long ldo;
    int M;
    int m_tiles;
    int wide_store;   // out (and out_k / out_v) 32-byte aligned, ldo and N multiples of 16: a lane stores 32 bytes (a whole L2 sector) at a time
    // head-parallel routing of the QKV epilogue (see pe_gemm_seg): heads_per_route > 0 -> head group g = head / heads_per_route is written
    // to route[which][g] (possibly peer-GPU memory) at column (head % heads_per_route) * 128
    int heads_per_route;
    bf16* route[3][8];
};

// one lane's 32-byte store (two packed 8 x bf16 groups, adjacent columns of its row)
__device__ __forceinline__ void st_global_32B(void* dst, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

struct GemmParams {
    SegDev seg[2];
    int nseg;
    int N;
    int K;
    int num_n;
    int total_m_tiles;
    int group_m;           // m-tiles per rasterisation group (balanced: ceil(total / number of groups))
    int group_n;           // n-tiles per rasterisation band (num_n: one band, the r1 order).  Deep-K layers (MLP down) split N into bands whose
                           // W panel stays in the L2 while the m-groups of the band stream past it; see gemm_run
    unsigned long long hint_a, hint_w;   // L2 eviction-priority policies of the A / W tile loads (plain 2-D loads only)
    int max_clusters;      // > 0: launch at most this many CTAs (pairs), so that one round of tiles is exactly one rasterisation group
    int num_tiles;
    int num_kb;
    int heads;             // QKV epilogue: N = 3 * heads * 128
    // implicit-GEMM convolution (pe_conv2d): segment 0's A operand is an activation map [H, W, C] read through a 3-D tensor map;
    // an m-tile is a (128 >> tile_w_log2) x (1 << tile_w_log2) patch of output pixels, the K loop runs over (tap, 64-channel block)
    int conv;              // 0: plain GEMM
    int conv_kw;           // taps per kernel row
    int conv_pad;          // zero padding on the top / left edge (bottom / right come from the TMA out-of-bounds fill)
    int kb_per_tap;        // ceil(C / 64)
    int conv_H;
    int conv_W;
    int tiles_x;
    int tile_w_log2;
    int dual_m;            // narrow conv layers (N <= 128) on one CTA: a tile is TWO stacked 128-pixel patches that share every weight box --
                           // the accumulator stage holds patch 0 in TMEM columns [0, 128) and patch 1 in [128, 256), epilogue warps 4-7 drain
                           // patch 0 and warps 8-11 patch 1.  Doubles the MMA work per k-block iteration of the producer / issuer chains.
    int patches_per_tile;  // 1, or 2 (CTA pair: one patch per CTA; dual_m: both patches in this CTA)
    int trim_n;            // issue the MMAs of a ragged last n-tile with N = round_up(N - n0, 16) instead of 256
    // batched mode (pe_gemm_batched): `batch` independent problems of one shape share a launch.  The operands are the flattened 2-D tensors of
    // segment 0; problem b reads A rows from b * a_batch_rows, W rows from b * w_batch_rows and writes out rows from b * out_batch_rows.  A tile
    // that overhangs its problem reads the neighbour's rows -- those accumulator rows / columns are never stored (row < M, col < N).
    int batch;             // 0 / 1: plain GEMM
    int m_tiles_per_batch;
    long long a_batch_rows, w_batch_rows, out_batch_rows;
    const float* vec;      // PE_EPI_ATTN_P / PE_EPI_ATTN_DS: fp32 vector of problem b at vec + b * vec_batch_stride, indexed by row or by column
    long long vec_batch_stride;
    int vec_per_column;
    float alpha;
    int num_stages;        // depth of the TMA -> MMA ring: the 192 KB tile region divided by the stage size (4 / 6 for 256-column tiles, up to
                           // kMaxStages for narrow layers, whose loads are latency- rather than bandwidth-bound)
    int stage_smem_bytes;  // shared-memory pitch of one stage: 16 KB A box + W box rounded up to 1 KB
    int stage_tx_bytes;    // bytes one pipeline stage receives (A box + W box): the W box has only round_up(N, 16) rows for narrow layers
    unsigned int* abort_flag;
};

// (y0, x0) of 128-pixel patch `sub` of m-tile `m0 / kTileM`: a tile is one patch or two vertically stacked patches (CTA pair: patch =
// CTA rank; dual_m: both patches belong to this CTA)
template <int kTileM>
__device__ __forceinline__ void conv_tile_origin(const GemmParams& p, int m0, int sub, int& y0, int& x0) {
    const int mt = m0 / kTileM;
    const int ty = mt / p.tiles_x;
    const int tile_h = 128 >> p.tile_w_log2;
    y0 = (ty * p.patches_per_tile + sub) * tile_h;
    x0 = (mt - ty * p.tiles_x) << p.tile_w_log2;
}

// columns of the n-tile at n0 that the MMAs cover: 256, or for a ragged last tile of a narrow layer round_up(N - n0, 16)
__device__ __forceinline__ int tile_n_cols(const GemmParams& p, int n0) {
    const int n_left = p.N - n0;
    return (p.trim_n && n_left < kTileN) ? ((n_left + 15) & ~15) : kTileN;
}

__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

struct Tile {
    int seg;
    int m0;   // first row of the tile inside its segment (batched mode: inside its problem)
    int n0;
    int b;    // batched mode: which problem
};

template <int kTileM>
__device__ __forceinline__ Tile decode_tile(const GemmParams& p, int t) {
    // band h = n-tiles [h * group_n, ...) (every band but the last is full, so the tiles before band h are h * total_m_tiles * group_n);
    // inside a band: m-groups of group_m m-tiles; inside a group the m index runs fastest
    int first_n = 0, band_n = p.num_n;
    if (p.group_n < p.num_n) {
        const int band_size = p.total_m_tiles * p.group_n;
        const int h = t / band_size;
        t -= h * band_size;
        first_n = h * p.group_n;
        band_n = min(p.group_n, p.num_n - first_n);
    }
    const int group_size = p.group_m * band_n;
    const int g = t / group_size;
    const int first_m = g * p.group_m;
    const int gm = min(p.group_m, p.total_m_tiles - first_m);
    const int in_group = t - g * group_size;
    int mt = first_m + in_group % gm;
    const int nt = first_n + in_group / gm;
    Tile r;
    r.seg = 0;
    r.b = 0;
    if (p.nseg > 1 && mt >= p.seg[0].m_tiles) { r.seg = 1; mt -= p.seg[0].m_tiles; }
    if (p.batch > 1) { r.b = mt / p.m_tiles_per_batch; mt -= r.b * p.m_tiles_per_batch; }
    r.m0 = mt * kTileM;
    r.n0 = nt * kTileN;
    return r;
}

__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// ---- generic per-chunk epilogues: 32 consecutive columns of one row --------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&acc)[32], const SegDev& sg, long long row, int n, int N, float row_val = 0.f,
                                               const float* col_vec = nullptr, float alpha = 1.f) {
    if (EPI == PE_EPI_ATTN_P || EPI == PE_EPI_ATTN_DS) {
        // attention backward (physicedit_b200/autograd.py): the accumulator is a tile of scores S = Q K^T (or its transpose), resp. of
        // dP = dO V^T; `row_val` / `col_vec` carry the per-query-row statistic (log2-domain LSE, resp. delta = rowsum(dO * O)).
        //   ATTN_P : out = bf16( exp2(acc * alpha - lse) )                    alpha = softmax scale * log2(e)
        //   ATTN_DS: out = bf16( P * (acc - delta) * alpha ), P read from out   alpha = softmax scale
        bf16* out_row = sg.out + row * sg.ldo;
        // ATTN_DS reads the tile of P it overwrites: all four 16-byte loads of this row's 32 columns are issued before the first store (a store
        // to out_row would otherwise order the later loads behind it -- with K = 128 these launches are epilogue-bound)
        uint4 held_p = make_uint4(0u, 0u, 0u, 0u);
        uint4 pv[4];
        if (EPI == PE_EPI_ATTN_DS) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
                pv[v] = (n + v * 8 < N) ? *reinterpret_cast<const uint4*>(out_row + n + v * 8) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int nn = n + v * 8;
            if (nn >= N) break;
            float st[8];
            if (col_vec != nullptr) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(col_vec + nn)), b = __ldg(reinterpret_cast<const float4*>(col_vec + nn + 4));
                st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w; st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) st[j] = row_val;
            }
            float x[8];
            if (EPI == PE_EPI_ATTN_P) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = exp2f(fmaf(__uint_as_float(acc[v * 8 + j]), alpha, -st[j]));
            } else {
                const uint32_t pw[4] = {pv[v].x, pv[v].y, pv[v].z, pv[v].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 pf = unpack_bf16(pw[j]);
                    x[2 * j] = pf.x * (__uint_as_float(acc[v * 8 + 2 * j]) - st[2 * j]) * alpha;
                    x[2 * j + 1] = pf.y * (__uint_as_float(acc[v * 8 + 2 * j + 1]) - st[2 * j + 1]) * alpha;
                }
            }
            uint4 o;
            o.x = pack_bf16(x[0], x[1]);
            o.y = pack_bf16(x[2], x[3]);
            o.z = pack_bf16(x[4], x[5]);
            o.w = pack_bf16(x[6], x[7]);
            // a lane owns one row: 32-byte stores (a whole L2 sector) when the layout allows -- with K = 128 these launches are bound by the
            // epilogue's store transactions, and half-filled sectors cost as much as full ones
            if (!sg.wide_store) *reinterpret_cast<uint4*>(out_row + nn) = o;
            else if ((v & 1) == 0) held_p = o;
            else st_global_32B(out_row + nn - 8, held_p, o);
        }
        return;
    }
    if (EPI == PE_EPI_F32) {
        // raw fp32 accumulators (attention scores of the VAE mid block, qwen_image_vae.py:189): out is float [M, ldo]
        // a lane owns one row: 32-byte stores (whole L2 sectors) instead of 16-byte ones; N and ldo are multiples of 8 floats
        float* o = reinterpret_cast<float*>(sg.out) + row * sg.ldo + n;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (n + v * 8 >= N) break;
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"l"(o + v * 8), "r"(acc[v * 8]), "r"(acc[v * 8 + 1]), "r"(acc[v * 8 + 2]), "r"(acc[v * 8 + 3]), "r"(acc[v * 8 + 4]),
                           "r"(acc[v * 8 + 5]), "r"(acc[v * 8 + 6]), "r"(acc[v * 8 + 7])
                         : "memory");
        }
        return;
    }
    bf16* out_row = sg.out + row * sg.ldo;
    uint4 held = make_uint4(0u, 0u, 0u, 0u);      // the even 8-column group of a pair, waiting for its odd neighbour (wide_store)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int nn = n + v * 8;
        if (nn >= N) break;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(acc[v * 8 + j]);
        if (sg.bias != nullptr) {
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(sg.bias + nn));
            const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16(bw[j]);
                x[2 * j] += f.x;
                x[2 * j + 1] += f.y;
            }
        }
        // the reference materialises the linear output in bf16 here
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = bf16_round(x[j]);

        if (EPI == PE_EPI_BIAS_GELU_SIGMOID) {
            // ApproximateGELU: x * sigmoid(1.702 * x), each op rounded to bf16 (qwen_image_dit.py:47)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = bf16_round(1.702f * x[j]);
                const float s = bf16_round(sigmoidf_fast(a));
                x[j] = x[j] * s;
            }
        } else if (EPI == PE_EPI_BIAS_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = 0.5f * x[j] * (1.0f + erff(x[j] * 0.70710678118654752f));
        } else if (EPI == PE_EPI_BIAS_SILU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = x[j] * sigmoidf_fast(x[j]);
        } else if (EPI == PE_EPI_GATE_RESIDUAL) {
            const uint4 g = __ldg(reinterpret_cast<const uint4*>(sg.gate + nn));
            const uint4 r = *reinterpret_cast<const uint4*>(out_row + nn);
            const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 gf = unpack_bf16(gw[j]);
                const float2 rf = unpack_bf16(rw[j]);
                x[2 * j]     = rf.x + bf16_round(gf.x * x[2 * j]);
                x[2 * j + 1] = rf.y + bf16_round(gf.y * x[2 * j + 1]);
            }
        }
        uint4 o;
        o.x = pack_bf16(x[0], x[1]);
        o.y = pack_bf16(x[2], x[3]);
        o.z = pack_bf16(x[4], x[5]);
        o.w = pack_bf16(x[6], x[7]);
        if (!sg.wide_store) *reinterpret_cast<uint4*>(out_row + nn) = o;
        else if ((v & 1) == 0) held = o;
        else st_global_32B(out_row + nn - 8, held, o);
    }
}

// ---- QKV epilogue: one 128-column head of one row ---------------------------------------------------
// y0 = bf16(acc+bias); q,k: y1 = bf16(y0*rsqrt(mean(y0^2)+eps)); y2 = bf16(y1*w); RoPE in fp32 -> bf16.
__device__ __forceinline__ void epilogue_qkv_head(uint32_t taddr_head, const SegDev& sg, long long row, bool row_valid,
                                                  int n_head0, int heads) {
    const int head = n_head0 >> 7;
    const int which = head / heads;             // 0 q, 1 k, 2 v
    const int col0 = (head - which * heads) << 7;
    uint32_t y0[64];                            // 128 bf16 values, packed
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t acc[32];
        tmem_ld32(taddr_head + c * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(sg.bias + n_head0 + c * 32 + v * 8));
            const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16(bw[j]);
                const float a = bf16_round(__uint_as_float(acc[v * 8 + 2 * j]) + f.x);
                const float d = bf16_round(__uint_as_float(acc[v * 8 + 2 * j + 1]) + f.y);
                ss += a * a + d * d;
                y0[c * 16 + v * 4 + j] = pack_bf16(a, d);
            }
        }
    }
    if (!row_valid) return;
    bf16* dst;
    if (sg.heads_per_route > 0) {
        const int hw = col0 >> 7, g = hw / sg.heads_per_route;
        dst = sg.route[which][g] + row * sg.ldo + ((hw - g * sg.heads_per_route) << 7);
    } else {
        dst = (which == 0 ? sg.out : (which == 1 ? sg.out_k : sg.out_v)) + row * sg.ldo + col0;
    }
    if (which == 2) {
#pragma unroll
        for (int v = 0; v < 16; v += 2) {
            uint4 o, o2;
            o.x = y0[v * 4]; o.y = y0[v * 4 + 1]; o.z = y0[v * 4 + 2]; o.w = y0[v * 4 + 3];
            o2.x = y0[v * 4 + 4]; o2.y = y0[v * 4 + 5]; o2.z = y0[v * 4 + 6]; o2.w = y0[v * 4 + 7];
            if (sg.wide_store) {
                st_global_32B(dst + v * 8, o, o2);
            } else {
                *reinterpret_cast<uint4*>(dst + v * 8) = o;
                *reinterpret_cast<uint4*>(dst + v * 8 + 8) = o2;
            }
        }
        return;
    }
    const float rs = rsqrtf(ss * (1.0f / 128.0f) + 1e-6f);
    const bf16* w = which == 0 ? sg.norm_q_w : sg.norm_k_w;
    const float4* rope = reinterpret_cast<const float4*>(sg.rope + row * 64);   // 2 (cos,sin) pairs per float4
    uint4 held = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int v = 0; v < 16; ++v) {   // 8 columns = 4 rotary pairs per iteration
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + v * 8));
        const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
        uint32_t ov[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 y = unpack_bf16(y0[v * 4 + j]);
            const float2 wf = unpack_bf16(ww[j]);
            const float a = bf16_round(bf16_round(y.x * rs) * wf.x);
            const float b = bf16_round(bf16_round(y.y * rs) * wf.y);
            const float4 cs2 = __ldg(rope + v * 2 + (j >> 1));
            const float c = (j & 1) ? cs2.z : cs2.x;
            const float s = (j & 1) ? cs2.w : cs2.y;
            ov[j] = pack_bf16(a * c - b * s, a * s + b * c);
        }
        uint4 o;
        o.x = ov[0]; o.y = ov[1]; o.z = ov[2]; o.w = ov[3];
        if (!sg.wide_store) *reinterpret_cast<uint4*>(dst + v * 8) = o;
        else if ((v & 1) == 0) held = o;
        else st_global_32B(dst + v * 8 - 8, held, o);
    }
}

template <int kCG, int EPI>
__global__ void __launch_bounds__(kThreads, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
    constexpr int kStages = kMaxStages;         // barrier slots; p.num_stages of them are in use
    constexpr int kABytes = 128 * kBlockK * 2;
    constexpr int kTileM = 128 * kCG;
    const int num_stages = p.num_stages;
    const uint32_t stage_pitch = (uint32_t)p.stage_smem_bytes;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kTileRegionBytes;
    auto a_smem = [&](int s) { return smem_base + s * stage_pitch; };
    const uint32_t b_off = p.dual_m ? 2u * kABytes : (uint32_t)kABytes;      // dual_m stages hold two A boxes in front of the W box
    auto b_smem = [&](int s) { return smem_base + s * stage_pitch + b_off; };
    auto full_bar = [&](int s) { return bar_base + s * 8; };
    auto empty_bar = [&](int s) { return bar_base + (kStages + s) * 8; };
    auto tfull_bar = [&](int s) { return bar_base + (2 * kStages + s) * 8; };
    auto tempty_bar = [&](int s) { return bar_base + (2 * kStages + 2 + s) * 8; };
    const uint32_t tmem_slot = bar_base + (2 * kStages + 4) * 8;

    const int warp = threadIdx.x >> 5;
    const uint32_t cta_rank = kCG == 2 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x / kCG;
    const int num_clusters = gridDim.x / kCG;

    if (warp == 0 && elect_one()) {
        for (int s = 0; s < p.nseg; ++s) {
            prefetch_tmap(&p.seg[s].tmA);
            prefetch_tmap(&p.seg[s].tmB);
        }
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < num_stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), kEpiWarps * kCG);   // one elected arrival per epilogue warp (of both CTAs)
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<kCG>(tmem_slot, 512);
        tmem_relinquish<kCG>();
    }
    tc_fence_before();
    if (kCG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ld_shared_u32(tmem_slot);

    if (warp == 0) {
        // ================================ TMA producer ================================
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        for (int t = cluster_id; t < p.num_tiles && ok; t += num_clusters) {
            const Tile tile = decode_tile<kTileM>(p, t);
            const SegDev& sg = p.seg[tile.seg];
            int cy0 = 0, cx0 = 0;
            if (p.conv) conv_tile_origin<kTileM>(p, tile.m0, kCG == 2 ? (int)cta_rank : 0, cy0, cx0);
            // tap (dy, dx), channel block of k-block kb: the CTA's pixel patch shifted by the tap; pixels outside the map (negative or
            // >= H / W coordinates) and channels >= C are zero-filled by the TMA unit.  The coordinates advance incrementally: this warp's
            // serial instruction chain per k-block is what bounds the narrow layers (r1: two integer divisions here cost ~350 cycles per
            // k-block, more than the MMAs of a 96-column tile -- profiles/r01_vae_conv.md)
            int cc = 0, cx = cx0 - p.conv_pad, cy = cy0 - p.conv_pad, ctap_x = 0;
            const int b_row = (kCG == 1 ? tile.n0 : tile.n0 + (int)cta_rank * (tile_n_cols(p, tile.n0) >> 1))   // a pair's CTAs supply half of the rows each
                              + (int)(tile.b * p.w_batch_rows);
            const int a_row = tile.m0 + (int)cta_rank * 128 + (int)(tile.b * p.a_batch_rows);
            const uint32_t fb = kCG == 1 ? 0u : mapa(full_bar(0), 0);      // a pair's bytes are all accounted on the leader's barriers
            for (int kb = 0; kb < p.num_kb; ++kb) {
                if (!mbar_wait(empty_bar(stage), phase ^ 1u, p.abort_flag, 1)) { ok = false; break; }
                if (elect_one()) {
                    if (kCG == 1) {
                        mbar_arrive_expect_tx(full_bar(stage), (uint32_t)p.stage_tx_bytes);
                        if (p.conv) {
                            tma_load_3d(a_smem(stage), &sg.tmA, full_bar(stage), cc, cx, cy);
                            if (p.dual_m) tma_load_3d(a_smem(stage) + kABytes, &sg.tmA, full_bar(stage), cc, cx, cy + (128 >> p.tile_w_log2));
                        } else {
                            tma_load_2d_hint(a_smem(stage), &sg.tmA, full_bar(stage), kb * kBlockK, a_row, p.hint_a);
                        }
                        tma_load_2d_hint(b_smem(stage), &sg.tmB, full_bar(stage), kb * kBlockK, b_row, p.hint_w);
                    } else {
                        const uint32_t fbs = fb + stage * 8;
                        if (leader) mbar_arrive_expect_tx(full_bar(stage), (uint32_t)p.stage_tx_bytes);
                        if (p.conv) tma_load_3d_cg2(a_smem(stage), &sg.tmA, fbs, cc, cx, cy);
                        else tma_load_2d_cg2_hint(a_smem(stage), &sg.tmA, fbs, kb * kBlockK, a_row, p.hint_a);
                        tma_load_2d_cg2_hint(b_smem(stage), &sg.tmB, fbs, kb * kBlockK, b_row, p.hint_w);
                    }
                }
                if (p.conv) {
                    cc += kBlockK;
                    if (cc >= p.kb_per_tap * kBlockK) {          // next tap
                        cc = 0;
                        ++cx;
                        if (++ctap_x == p.conv_kw) { ctap_x = 0; cx -= p.conv_kw; ++cy; }
                    }
                }
                __syncwarp();
                if (++stage == num_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (leader) {
            constexpr uint32_t idesc_full = make_idesc_bf16(kTileM, kTileN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            bool ok = true;
            for (int t = cluster_id; t < p.num_tiles && ok; t += num_clusters) {
                if (!mbar_wait<kCG == 2>(tempty_bar(acc), acc_phase ^ 1u, p.abort_flag, 2)) { ok = false; break; }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kTileN;
                uint32_t idesc = idesc_full;
                if (p.trim_n) {
                    // narrow layers (96 / 192 / 384 conv channels): do not multiply the zero-filled weight rows of a ragged n-tile
                    const int n_cols = tile_n_cols(p, decode_tile<kTileM>(p, t).n0);
                    if (n_cols < kTileN) idesc = make_idesc_bf16(kTileM, (uint32_t)n_cols, 0, 0);
                }
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    if (!mbar_wait(full_bar(stage), phase, p.abort_flag, 3)) { ok = false; break; }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t adesc = make_smem_desc_sw128(a_smem(stage), 16, 1024);
                        const uint64_t bdesc = make_smem_desc_sw128(b_smem(stage), 16, 1024);
                        // +32 bytes per UMMA_K step inside the 128-byte swizzle row (addr field is >>4).  The dual_m variant is a separate
                        // straight-line batch: testing the flag between the MMAs of the common path cost that path ~70 ns per k-block
                        // (r1: 0.33 -> 0.41 ms on the 1024^2 96->96 layer, back to 0.33 with the test hoisted)
                        if (kCG == 1 && p.dual_m) {
                            // second patch of the tile: next A box (+16 KB = +1024 in the descriptor's >>4 address field), same weights,
                            // accumulator columns [128, 256)
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                                umma_bf16<kCG>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                                umma_bf16<kCG>(d_tmem + 128, adesc + (kABytes >> 4) + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_bf16<kCG>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        if (kCG == 1) umma_commit(empty_bar(stage)); else umma_commit_cg2(empty_bar(stage), 3);
                        if (kb == p.num_kb - 1) {
                            if (kCG == 1) umma_commit(tfull_bar(acc)); else umma_commit_cg2(tfull_bar(acc), 3);
                        }
                    }
                    __syncwarp();
                    if (++stage == num_stages) { stage = 0; phase ^= 1u; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        const int ew = (warp - 4) & 3;           // == warp % 4: this warp may touch TMEM lanes [32*ew, 32*ew+32)
        const int half = (warp - 4) >> 2;        // which 128-column half of the accumulator this warp drains
        const int lane = lane_id();
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < p.num_tiles; t += num_clusters) {
            const Tile tile = decode_tile<kTileM>(p, t);
            const SegDev& sg = p.seg[tile.seg];
            if (!mbar_wait(tfull_bar(acc), acc_phase, p.abort_flag, 4)) break;
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * kTileN;
            long long row = tile.m0 + (int)cta_rank * 128 + ew * 32 + lane;
            bool row_valid = row < sg.M;
            const bool dual = kCG == 1 && p.dual_m;
            if (p.conv) {
                int cy0, cx0;
                conv_tile_origin<kTileM>(p, tile.m0, dual ? half : (int)cta_rank, cy0, cx0);
                const int r = ew * 32 + lane;
                const int yy = cy0 + (r >> p.tile_w_log2);
                const int xx = cx0 + (r & ((1 << p.tile_w_log2) - 1));
                row_valid = yy < p.conv_H && xx < p.conv_W;
                row = (long long)yy * p.conv_W + xx;
            }
            if (EPI == PE_EPI_QKV_NORM_ROPE) {
                const int n_head0 = tile.n0 + half * 128;
                if (n_head0 < p.N) epilogue_qkv_head(taddr + half * 128, sg, row, row_valid, n_head0, p.heads);
            } else {
                // each warp drains one 128-column half of the accumulator stage: columns [128 half, +128) of a 256-column tile, or, for a
                // dual_m tile, the whole (<= 128-column) result of patch `half`
                const int c0 = dual ? 0 : half * 4;
                const uint32_t tcol = dual ? (uint32_t)half * 128u : 0u;
                float row_val = 0.f;
                const float* col_vec = nullptr;
                if (EPI == PE_EPI_ATTN_P || EPI == PE_EPI_ATTN_DS) {
                    const float* vb = p.vec + tile.b * p.vec_batch_stride;
                    if (p.vec_per_column) col_vec = vb;
                    else if (row_valid) row_val = __ldg(vb + row);
                }
                const long long out_row = row + tile.b * p.out_batch_rows;
#pragma unroll 1
                for (int c = c0; c < c0 + 4; ++c) {
                    const int n = tile.n0 + c * 32;
                    if (n >= p.N) break;
                    uint32_t r[32];
                    tmem_ld32(taddr + tcol + c * 32, r);
                    tmem_ld_wait();
                    if (row_valid) epilogue_chunk<EPI>(r, sg, out_row, n, p.N, row_val, col_vec, p.alpha);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (elect_one()) {
                if (kCG == 1) mbar_arrive(tempty_bar(acc));
                else mbar_arrive_cluster(mapa(tempty_bar(acc), 0));
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    // ================================ teardown ================================
    tc_fence_before();
    if (kCG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<kCG>(tmem_base, 512);
    }
}

template <int kCG>
constexpr int gemm_smem_bytes() {
    return 1024 /*alignment slack*/ + kTileRegionBytes + 256 /*barriers: (2 * kMaxStages + 4) * 8 + TMEM slot*/;
}

template <int kCG, int EPI>
int launch_gemm(Handle* h, const GemmParams& p, cudaStream_t stream) {
    auto kern = gemm_kernel<kCG, EPI>;
    constexpr int smem = gemm_smem_bytes<kCG>();
    static bool configured = false;   // per template instance; attribute is sticky per function
    if (!configured) {
        PE_CHECK_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    int ctas = h->sm_count;
    if (kCG == 2) ctas &= ~1;
    const int max_useful = (p.max_clusters > 0 && p.max_clusters < p.num_tiles ? p.max_clusters : p.num_tiles) * kCG;
    if (ctas > max_useful) ctas = max_useful;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PE_CHECK_CUDA(h, cudaLaunchKernelEx(&cfg, kern, p));
    return PE_OK;
}

template <int kCG>
int dispatch_epilogue(Handle* h, const GemmParams& p, int epilogue, cudaStream_t stream) {
    switch (epilogue) {
        case PE_EPI_BIAS: return launch_gemm<kCG, PE_EPI_BIAS>(h, p, stream);
        case PE_EPI_BIAS_GELU_SIGMOID: return launch_gemm<kCG, PE_EPI_BIAS_GELU_SIGMOID>(h, p, stream);
        case PE_EPI_BIAS_GELU_ERF: return launch_gemm<kCG, PE_EPI_BIAS_GELU_ERF>(h, p, stream);
        case PE_EPI_GATE_RESIDUAL: return launch_gemm<kCG, PE_EPI_GATE_RESIDUAL>(h, p, stream);
        case PE_EPI_QKV_NORM_ROPE: return launch_gemm<kCG, PE_EPI_QKV_NORM_ROPE>(h, p, stream);
        case PE_EPI_BIAS_SILU: return launch_gemm<kCG, PE_EPI_BIAS_SILU>(h, p, stream);
        case PE_EPI_F32: return launch_gemm<kCG, PE_EPI_F32>(h, p, stream);
        case PE_EPI_ATTN_P: return launch_gemm<kCG, PE_EPI_ATTN_P>(h, p, stream);
        case PE_EPI_ATTN_DS: return launch_gemm<kCG, PE_EPI_ATTN_DS>(h, p, stream);
        default: return set_error(h, PE_ERR_INVALID_ARGUMENT, "pe_gemm: unknown epilogue %d", epilogue);
    }
}

}  // namespace

// A/B switch for experiments: PE_GEMM_NARROW_STORES=1 keeps the 16-byte epilogue stores (read once per process)
static bool narrow_stores() {
    static const bool v = [] { const char* e = getenv("PE_GEMM_NARROW_STORES"); return e != nullptr && e[0] == '1'; }();
    return v;
}

// Ring geometry for a launch whose CTAs each receive a 16 KB A box and a `b_rows_per_cta`-row W box per stage.
static void set_ring(GemmParams& p, int cg, int b_rows_per_cta, int a_boxes = 1) {
    const int b_bytes = b_rows_per_cta * kBlockK * 2;
    const int a_bytes = a_boxes * 128 * kBlockK * 2;
    p.stage_smem_bytes = a_bytes + ((b_bytes + 1023) & ~1023);
    p.num_stages = kTileRegionBytes / p.stage_smem_bytes;
    if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
    if (b_rows_per_cta * cg == kTileN) p.num_stages = cg == 1 ? 4 : 6;   // the DiT's full-width GEMMs: the ring they were tuned and profiled with
    static const int stage_cap = [] { const char* e = getenv("PE_GEMM_MAX_STAGES"); return e ? atoi(e) : 0; }();   // tuning aid: cap the ring depth
    if (stage_cap >= 2 && stage_cap < p.num_stages) p.num_stages = stage_cap;
    p.stage_tx_bytes = cg * (a_bytes + b_bytes);                 // a pair's boxes all land on the leader's barrier
}

// ---- rasterisation (tile index -> (m-tile, n-tile), see decode_tile) and L2 policies of the operand loads ---------------------------
// Tiles are dealt round-robin to the persistent CTAs (pairs), so ~`clusters` consecutive tile indices run concurrently and stream K in
// step: what a round of tiles reads from DRAM is (distinct m-tiles + distinct n-tiles) x tile rows x K x 2 bytes, and anything less needs
// a panel of one operand to stay in the L2 from one round to the next.
//  * wide layers (K = 3072): m-groups whose A panel is <= kPanelBytes, each swept over all n-tiles in several rounds; the panel stays
//    resident, W streams once per group.
//  * deep-K layers (MLP down-projection: K = 12288, a tile row is 6.3 MB): such a group (5 m-tiles x 12 n-tiles) is smaller than one round,
//    so nothing is re-used across rounds and the rounds straddle two groups.  There a group is made exactly one round -- group_m =
//    clusters / num_n m-tiles (6 x 12 = 72 tiles) -- and only that many CTA pairs are launched when it costs no extra round (it does
//    not at 408 tiles: 6 rounds either way).  Measured at M = 8704 (tools/gemm_raster_ab.py, profiles/r02_gemm_raster_ab.json):
//    DRAM reads 946 -> 828 MB per launch, 461.5 -> 454.9 us.
//  * measured and NOT used: evict-first on the streamed operand (the other CTAs of the round read the same rows a little later: DRAM
//    reads double, 946 -> 1709 MB), n-bands with a resident W panel (group_n < num_n: 1060 MB), one m-group for the wide layers (the
//    53 MB A panel does not stay: 303 -> 598 MB).  The knobs stay reachable through PE_GEMM_TUNE / PE_GEMM_RASTER for the A/B tool.
struct RasterTune { int on, automatic, gn, gm, ha, hw, clusters; };
static RasterTune raster_tune() {
    // tools/gemm_raster_ab.py: with PE_GEMM_TUNE in the environment, PE_GEMM_RASTER = "auto" | "gn,gm,hint_a,hint_w,clusters" is re-read on every
    // launch (gn / gm / clusters 0 = the default order's value; hints 0 normal, 1 evict-first, 2 evict-last)
    static const bool enabled = getenv("PE_GEMM_TUNE") != nullptr;
    RasterTune t = {0, 0, 0, 0, 0, 0, 0};
    if (!enabled) return t;
    const char* e = getenv("PE_GEMM_RASTER");
    if (!e) return t;
    if (e[0] == 'a') { t.on = 1; t.automatic = 1; return t; }
    if (sscanf(e, "%d,%d,%d,%d,%d", &t.gn, &t.gm, &t.ha, &t.hw, &t.clusters) == 5) t.on = 1;
    return t;
}
static unsigned long long l2_policy(int code) { return code == 1 ? kL2EvictFirst : code == 2 ? kL2EvictLast : kL2EvictNormal; }

static void choose_raster(const Handle* h, GemmParams& p, int tile_m, int cg, bool batched) {
    const long long a_tile_bytes = (long long)tile_m * p.K * 2;      // one m-tile's rows of A
    // as many m-tiles per group as keep the group's A panel within kPanelBytes, spread evenly (r1: a fixed 16 left a last group of
    // 2 m-tiles that re-streamed the whole weight matrix, and at K = 12288 a 100 MB panel that did not fit the L2 next to W)
    int gm_max = (int)(kPanelBytes / a_tile_bytes);
    if (gm_max < 1) gm_max = 1;
    const int groups = ceil_div(p.total_m_tiles, gm_max);
    p.group_m = ceil_div(p.total_m_tiles, groups);
    p.group_n = p.num_n;
    p.hint_a = p.hint_w = kL2EvictNormal;
    const RasterTune t = raster_tune();
    const bool automatic = t.on ? t.automatic != 0 : kRoundRasterDefault;
    const int clusters = (h->sm_count > cg ? h->sm_count : cg) / cg;
    if (automatic && !batched && groups > 1 && gm_max * p.num_n <= clusters) {
        // an A-panel group is smaller than one round of tiles: make a group exactly one round
        const int num_tiles = p.total_m_tiles * p.num_n;
        p.group_m = clusters / p.num_n;
        const int used = p.group_m * p.num_n;
        if (ceil_div(num_tiles, used) == ceil_div(num_tiles, clusters)) p.max_clusters = used;
    }
    if (t.on && !t.automatic) {
        if (t.gn > 0) p.group_n = t.gn < p.num_n ? t.gn : p.num_n;
        if (t.gm > 0) p.group_m = t.gm;
        p.hint_a = l2_policy(t.ha);
        p.hint_w = l2_policy(t.hw);
        p.max_clusters = t.clusters;
    }
#ifdef PE_GEMM_GROUP_M
    p.group_m = PE_GEMM_GROUP_M;      // experiments: fixed group size
#endif
}

int gemm_run(Handle* h, const pe_gemm_seg* segs, int nseg, int N, int K, int epilogue, int flags, cudaStream_t stream, const pe_gemm_batch* bt) {
    PE_REQUIRE(h, nseg >= 1 && nseg <= 2, "pe_gemm: nseg must be 1 or 2 (got %d)", nseg);
    const int batch = bt ? bt->batch : 1;
    if (bt) {
        PE_REQUIRE(h, nseg == 1 && batch >= 1 && batch <= 4096, "pe_gemm_batched: one segment, 1 <= batch <= 4096 (batch=%d)", batch);
        PE_REQUIRE(h, epilogue == PE_EPI_BIAS || epilogue == PE_EPI_F32 || epilogue == PE_EPI_ATTN_P || epilogue == PE_EPI_ATTN_DS,
                   "pe_gemm_batched: epilogue must be PE_EPI_BIAS, PE_EPI_F32, PE_EPI_ATTN_P or PE_EPI_ATTN_DS");
        PE_REQUIRE(h, bt->a_batch_rows >= 0 && bt->w_batch_rows >= 0 && bt->out_batch_rows >= 0, "pe_gemm_batched: negative batch stride");
        PE_REQUIRE(h, segs[0].bias == nullptr, "pe_gemm_batched: no bias in batched mode");
    }
    if (epilogue == PE_EPI_ATTN_P || epilogue == PE_EPI_ATTN_DS) {
        PE_REQUIRE(h, bt && bt->vec, "pe_gemm: the attention-backward epilogues need pe_gemm_batched with a statistic vector");
        PE_REQUIRE(h, (reinterpret_cast<uintptr_t>(bt->vec) & 15) == 0 && bt->vec_batch_stride % 4 == 0, "pe_gemm_batched: vec must be 16-byte aligned per problem");
    }
    PE_REQUIRE(h, N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0, "pe_gemm: N and K must be positive multiples of 8 (N=%d K=%d)", N, K);
    const int cg = (flags & PE_GEMM_FLAG_CTA_PAIR) ? 2 : 1;
    const int tile_m = 128 * cg;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.nseg = nseg;
    p.N = N;
    p.K = K;
    p.num_n = ceil_div(N, kTileN);
    p.num_kb = ceil_div(K, kBlockK);
    p.abort_flag = h->abort_flag;
    if (epilogue == PE_EPI_QKV_NORM_ROPE) {
        PE_REQUIRE(h, N % 384 == 0, "pe_gemm: QKV epilogue needs N = 3*heads*128 (N=%d)", N);
        p.heads = N / 384;
    }
    // narrow layers (PE_GEMM_FLAG_TRIM_N, N < 256): the W box has only round_up(N, 16) rows, so no zero-filled rows are written to shared memory
    const int b_box_rows = ((flags & PE_GEMM_FLAG_TRIM_N) && N < kTileN) ? ((N + 15) & ~15) : kTileN;
    set_ring(p, cg, b_box_rows / cg);
    int total_m_tiles = 0;
    for (int s = 0; s < nseg; ++s) {
        const pe_gemm_seg& in = segs[s];
        PE_REQUIRE(h, in.M > 0, "pe_gemm: segment %d has M=%d", s, in.M);
        PE_REQUIRE(h, in.a && in.w && in.out, "pe_gemm: segment %d has a null a/w/out pointer", s);
        PE_REQUIRE(h, in.lda >= K && in.lda % 8 == 0, "pe_gemm: lda must be >= K and a multiple of 8");
        PE_REQUIRE(h, in.ldo % 8 == 0, "pe_gemm: ldo must be a multiple of 8");
        PE_REQUIRE(h, (reinterpret_cast<uintptr_t>(in.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(in.w) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(in.out) & 15) == 0,
                   "pe_gemm: a / w / out must be 16-byte aligned");
        SegDev& d = p.seg[s];
        const uint64_t a_rows = (uint64_t)in.M + (bt ? (uint64_t)(batch - 1) * (uint64_t)bt->a_batch_rows : 0);
        const uint64_t w_rows = (uint64_t)N + (bt ? (uint64_t)(batch - 1) * (uint64_t)bt->w_batch_rows : 0);
        int rc = make_tmap_2d(h, &d.tmA, in.a, a_rows, (uint64_t)K, (uint64_t)in.lda, 128);
        if (rc) return rc;
        rc = make_tmap_2d(h, &d.tmB, in.w, w_rows, (uint64_t)K, (uint64_t)K, cg == 1 ? (uint32_t)b_box_rows : (uint32_t)(b_box_rows >> 1));
        if (rc) return rc;
        d.bias = static_cast<const bf16*>(in.bias);
        d.out = static_cast<bf16*>(in.out);
        d.gate = static_cast<const bf16*>(in.gate);
        d.out_k = static_cast<bf16*>(in.out_k);
        d.out_v = static_cast<bf16*>(in.out_v);
        d.norm_q_w = static_cast<const bf16*>(in.norm_q_w);
        d.norm_k_w = static_cast<const bf16*>(in.norm_k_w);
        d.rope = static_cast<const float2*>(in.rope);
        d.ldo = in.ldo;
        d.M = in.M;
        {
            uintptr_t al = reinterpret_cast<uintptr_t>(in.out) | reinterpret_cast<uintptr_t>(in.out_k) | reinterpret_cast<uintptr_t>(in.out_v);
            d.wide_store = ((al & 31) == 0 && in.ldo % 16 == 0 && N % 16 == 0 && !narrow_stores()) ? 1 : 0;
        }
        d.m_tiles = ceil_div(in.M, tile_m) * batch;
        total_m_tiles += d.m_tiles;
        if (epilogue == PE_EPI_GATE_RESIDUAL) PE_REQUIRE(h, in.gate != nullptr, "pe_gemm: gate-residual epilogue needs gate");
        if (epilogue == PE_EPI_F32) PE_REQUIRE(h, (reinterpret_cast<uintptr_t>(in.out) & 31) == 0, "pe_gemm: the fp32 output must be 32-byte aligned");
        if (epilogue == PE_EPI_QKV_NORM_ROPE)
            PE_REQUIRE(h, in.bias && in.out_k && in.out_v && in.norm_q_w && in.norm_k_w && in.rope,
                       "pe_gemm: QKV epilogue needs bias, out_k, out_v, norm weights and rope table");
        if (epilogue == PE_EPI_QKV_NORM_ROPE && in.route_ranks > 0) {
            PE_REQUIRE(h, in.route_ranks <= 8 && p.heads % in.route_ranks == 0, "pe_gemm: route_ranks must divide the head count (heads=%d ranks=%d)", p.heads, in.route_ranks);
            d.heads_per_route = p.heads / in.route_ranks;
            uintptr_t al = 0;
            for (int g = 0; g < in.route_ranks; ++g) {
                PE_REQUIRE(h, in.q_route[g] && in.k_route[g] && in.v_route[g], "pe_gemm: null route pointer (group %d)", g);
                d.route[0][g] = static_cast<bf16*>(in.q_route[g]);
                d.route[1][g] = static_cast<bf16*>(in.k_route[g]);
                d.route[2][g] = static_cast<bf16*>(in.v_route[g]);
                al |= reinterpret_cast<uintptr_t>(in.q_route[g]) | reinterpret_cast<uintptr_t>(in.k_route[g]) | reinterpret_cast<uintptr_t>(in.v_route[g]);
            }
            PE_REQUIRE(h, (al & 15) == 0, "pe_gemm: route pointers must be 16-byte aligned");
            d.wide_store = ((al & 31) == 0 && in.ldo % 16 == 0 && !narrow_stores()) ? 1 : 0;
        }
    }
    p.total_m_tiles = total_m_tiles;
    if (bt) {
        p.batch = batch;
        p.m_tiles_per_batch = total_m_tiles / batch;
        p.a_batch_rows = bt->a_batch_rows;
        p.w_batch_rows = bt->w_batch_rows;
        p.out_batch_rows = bt->out_batch_rows;
        p.vec = bt->vec;
        p.vec_batch_stride = bt->vec_batch_stride;
        p.vec_per_column = bt->vec_per_column;
        p.alpha = bt->alpha;
    }
    choose_raster(h, p, tile_m, cg, bt != nullptr);
    p.num_tiles = total_m_tiles * p.num_n;
    p.trim_n = (flags & PE_GEMM_FLAG_TRIM_N) ? 1 : 0;
    if (cg == 1) return dispatch_epilogue<1>(h, p, epilogue, stream);
    return dispatch_epilogue<2>(h, p, epilogue, stream);
}

// Implicit-GEMM 2-D convolution, stride 1, output size = input size:
//   out[y, x, n] = epilogue( sum_{dy, dx, c} x[y + dy - pad, x + dx - pad, c] * w[n, (dy*kw + dx)*cpad + c] + bias[n] )
// with zeros outside the map.  No im2col buffer: the producer warp of gemm_kernel reads one shifted pixel patch per tap straight
// from the NHWC activation map through a 3-D tensor map.
int conv2d_run(Handle* h, const pe_conv2d_desc* d, int epilogue, cudaStream_t stream) {
    PE_REQUIRE(h, d->x && d->w && d->out, "pe_conv2d: null x / w / out");
    PE_REQUIRE(h, d->H > 0 && d->W > 0 && d->C > 0 && d->N > 0, "pe_conv2d: H, W, C, N must be positive");
    PE_REQUIRE(h, d->C % 8 == 0 && d->ldx % 8 == 0 && d->ldx >= d->C, "pe_conv2d: C and ldx must be multiples of 8, ldx >= C (C=%d ldx=%lld)", d->C, (long long)d->ldx);
    PE_REQUIRE(h, d->N % 8 == 0 && d->ldo % 8 == 0, "pe_conv2d: N and ldo must be multiples of 8");
    PE_REQUIRE(h, d->kh >= 1 && d->kh <= 3 && d->kw >= 1 && d->kw <= 3 && d->pad >= 0 && d->pad <= 1, "pe_conv2d: kernel up to 3x3, pad 0 or 1");
    PE_REQUIRE(h, epilogue == PE_EPI_BIAS || epilogue == PE_EPI_GATE_RESIDUAL || epilogue == PE_EPI_BIAS_SILU,
               "pe_conv2d: epilogue must be PE_EPI_BIAS, PE_EPI_BIAS_SILU or PE_EPI_GATE_RESIDUAL");
    PE_REQUIRE(h, epilogue != PE_EPI_GATE_RESIDUAL || d->gate != nullptr, "pe_conv2d: gate-residual epilogue needs gate");
    PE_REQUIRE(h, (reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(d->out) & 15) == 0, "pe_conv2d: x / w / out must be 16-byte aligned");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.nseg = 1;
    p.N = d->N;
    p.kb_per_tap = ceil_div(d->C, kBlockK);
    const int cpad = p.kb_per_tap * kBlockK;
    p.K = d->kh * d->kw * cpad;
    p.num_n = ceil_div(d->N, kTileN);
    p.num_kb = p.K / kBlockK;
    p.abort_flag = h->abort_flag;
    p.conv = 1;
    p.conv_kw = d->kw;
    p.conv_pad = d->pad;
    p.conv_H = d->H;
    p.conv_W = d->W;
    p.tile_w_log2 = d->W > 8 ? 4 : 3;             // 8 x 16 pixel patches (16 x 8 for maps narrower than 9 pixels)
    if ((d->flags >> 4) & 7) p.tile_w_log2 = (d->flags >> 4) & 7;      // PE_CONV_FLAG_TILE_W_LOG2(n): explicit patch width 2^n (tuning / tests)
    const int cg = (d->flags & PE_CONV_FLAG_CTA_PAIR) ? 2 : 1;   // CTA pair: two stacked patches per tile, each CTA loads half of the weights
    const int tile_w = 1 << p.tile_w_log2, tile_h = 128 >> p.tile_w_log2;
    p.tiles_x = ceil_div(d->W, tile_w);
    p.trim_n = 1;
    // narrow layers on one CTA: two patches per tile share the weight boxes (see GemmParams::dual_m)
    p.dual_m = (cg == 1 && d->N <= 128 && !(d->flags & PE_CONV_FLAG_SINGLE_PATCH)) ? 1 : 0;
    p.patches_per_tile = (cg == 2 || p.dual_m) ? 2 : 1;
    SegDev& sd = p.seg[0];
    int rc = make_tmap_3d(h, &sd.tmA, d->x, (uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->ldx, (uint64_t)d->ldx * d->W,
                          64, (uint32_t)tile_w, (uint32_t)tile_h);
    if (rc) return rc;
    // W box: the rows one CTA supplies -- the whole (narrow) n-tile, or half of it on a CTA pair; never zero-filled rows, which would
    // still cost L2 -> SM bandwidth (ncu: l1tex__m_xbar2l1tex_read_bytes counts the full box)
    const int n_tile_rows = d->N < kTileN ? ((d->N + 15) & ~15) : kTileN;
    const int b_box_rows = n_tile_rows / cg;
    set_ring(p, cg, b_box_rows, p.dual_m ? 2 : 1);
    rc = make_tmap_2d(h, &sd.tmB, d->w, (uint64_t)d->N, (uint64_t)p.K, (uint64_t)p.K, (uint32_t)b_box_rows);
    if (rc) return rc;
    sd.bias = static_cast<const bf16*>(d->bias);
    sd.out = static_cast<bf16*>(d->out);
    sd.gate = static_cast<const bf16*>(d->gate);
    sd.ldo = d->ldo;
    sd.wide_store = ((reinterpret_cast<uintptr_t>(d->out) & 31) == 0 && d->ldo % 16 == 0 && d->N % 16 == 0 && !narrow_stores()) ? 1 : 0;
    sd.M = d->H * d->W;
    sd.m_tiles = p.tiles_x * ceil_div(d->H, tile_h * p.patches_per_tile);
    p.total_m_tiles = sd.m_tiles;
    {
        const long long tile_row_bytes = (long long)128 * p.patches_per_tile * p.K * 2;
        int gm_max = (int)(kPanelBytes / tile_row_bytes);
        if (gm_max < 1) gm_max = 1;
        const int groups = ceil_div(p.total_m_tiles, gm_max);
        p.group_m = ceil_div(p.total_m_tiles, groups);
    }
    p.group_n = p.num_n;
    p.hint_a = p.hint_w = kL2EvictNormal;
    p.num_tiles = p.total_m_tiles * p.num_n;
    if (cg == 1) return dispatch_epilogue<1>(h, p, epilogue, stream);
    return dispatch_epilogue<2>(h, p, epilogue, stream);
}

}  // namespace pe
