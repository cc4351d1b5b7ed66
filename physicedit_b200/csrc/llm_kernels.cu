// Kernels of the Qwen2.5-VL text-encoder path: the conditioning model that turns (prompt, edit image) into prompt_emb and
// GENERATES the "physical thinking" text (DiffSynth-Studio/diffsynth/pipelines/qwen_image_physical.py:774-800, 859-873; model wrapper
// models/qwen_image_text_encoder_withdecode.py; arithmetic = transformers' modeling_qwen2_5_vl.py, the unpinned dependency).
// The GEMMs are pe_gemm / pe_gemv; this file holds what surrounds them -- all HBM- or latency-bound, bf16 rounding points replayed:
//   pe_swiglu            down_proj(act_fn(gate_proj(x)) * up_proj(x))                 Qwen2MLP.forward / Qwen2_5_VLMLP.forward
//   pe_rope_half         q*cos + rotate_half(q)*sin                                    apply_multimodal_rotary_pos_emb / apply_rotary_pos_emb_vision
//   pe_range_attention   softmax(q k^T * scale) v with grouped KV heads and a per-query KV range (causal prefill, KV-cache decode with
//                        the length read from device memory, the vision tower's window blocks); head dim 64 / 80 / 128
//   pe_gather_rows       embed_tokens(input_ids), masked_scatter of the image embeddings
//   pe_argmax            greedy next token (first maximal index, like torch.argmax)
//   pe_kv_append         KV-cache write at a device-side position + position increment: one decode step has no host-side state, so it
//                        can be captured once in a CUDA graph and replayed per token
#include "ptx.cuh"
#include "common.cuh"

namespace pe {
namespace {

__device__ __forceinline__ float silu_bf16(float g) { return bf16_round(g / (1.0f + __expf(-g))); }

// ---- out[r, i] = bf16( bf16(silu(x[r, i])) * x[r, I + i] ) ------------------------------------------------------------------------
__global__ void swiglu_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ out, long long ldo, int rows, int I) {
    const int nvec = I >> 3;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)rows * nvec) return;
    const int r = (int)(gid / nvec), v = (int)(gid - (long long)r * nvec);
    const uint4 g4 = *reinterpret_cast<const uint4*>(x + r * ldx + v * 8);
    const uint4 u4 = *reinterpret_cast<const uint4*>(x + r * ldx + I + v * 8);
    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, uw[4] = {u4.x, u4.y, u4.z, u4.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 g = unpack_bf16(gw[j]), u = unpack_bf16(uw[j]);
        ow[j] = pack_bf16(silu_bf16(g.x) * u.x, silu_bf16(g.y) * u.y);
    }
    *reinterpret_cast<uint4*>(out + r * ldo + v * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
}

// ---- rotate-half RoPE in place on x [T, H*D] -------------------------------------------------------------------------------------
// token t uses table row (row_ptr ? *row_ptr : row0) + t of cos / sin [*, D] (fp32, both halves of D filled as HF's cat(freqs, freqs)).
// mode 0: fp32 fused (vision tower: q.float()*cos + rotate_half(q.float())*sin -> bf16)
// mode 1: the language model's bf16 op order: cos / sin rounded to bf16, a = bf16(q*cos), b = bf16(rot*sin), out = bf16(a + b)
__global__ void rope_half_kernel(bf16* __restrict__ x, long long ldx, int T, int H, int D, const float* __restrict__ cs, const float* __restrict__ sn,
                                 const int* __restrict__ row_ptr, int row0, int mode) {
    const int half = D >> 1;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)T * H * half) return;
    const int i = (int)(gid % half);
    const int hd = (int)((gid / half) % H);
    const int t = (int)(gid / ((long long)half * H));
    const long long row = (long long)(row_ptr ? *row_ptr : row0) + t;
    bf16* p = x + t * ldx + hd * D;
    const float x1 = __bfloat162float(p[i]), x2 = __bfloat162float(p[i + half]);
    float c1 = cs[row * D + i], c2 = cs[row * D + i + half], s1 = sn[row * D + i], s2 = sn[row * D + i + half];
    float o1, o2;
    if (mode == 1) {
        c1 = bf16_round(c1); c2 = bf16_round(c2); s1 = bf16_round(s1); s2 = bf16_round(s2);
        o1 = bf16_round(x1 * c1) + bf16_round(-x2 * s1);
        o2 = bf16_round(x2 * c2) + bf16_round(x1 * s2);
    } else {
        o1 = x1 * c1 - x2 * s1;
        o2 = x2 * c2 + x1 * s2;
    }
    p[i] = __float2bfloat16_rn(o1);
    p[i + half] = __float2bfloat16_rn(o2);
}

// ---- attention with grouped KV heads and a per-query KV range ---------------------------------------------------------------------
// One warp per (query, head): lanes stride over the KV rows of [lo, hi), each lane keeps a private online softmax (fp32) over its rows
// and the 32 partial results are merged at the end.  CUDA cores on purpose: the shapes here are a 1-row decode step over a KV cache,
// ~200-1500-token prefills and 64-token vision windows, once per image -- latency-bound, not the DiT's tensor-bound attention.
template <int D>
__global__ void __launch_bounds__(128) range_attention_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                                                              bf16* __restrict__ o, int H, int group, int Sq, int Skv, long long ldq, long long ldkv,
                                                              long long ldo, float scale, const int* __restrict__ kv_lo, const int* __restrict__ kv_hi,
                                                              const int* __restrict__ kv_len_ptr) {
    __shared__ float qs[4][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * 4 + warp;
    const int hd = blockIdx.y;
    if (qi >= Sq) return;
    const int hkv = hd / group;
    const bf16* qrow = q + (long long)qi * ldq + hd * D;
    for (int d = lane; d < D; d += 32) qs[warp][d] = __bfloat162float(qrow[d]) * scale;
    __syncwarp();
    const int lo = kv_lo ? kv_lo[qi] : 0;
    int hi = kv_hi ? kv_hi[qi] : (kv_len_ptr ? *kv_len_ptr : Skv);
    if (hi > Skv) hi = Skv;
    float m = -INFINITY, l = 0.f;
    float acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    for (int kj = lo + lane; kj < hi; kj += 32) {
        const bf16* krow = k + (long long)kj * ldkv + hkv * D;
        float s = 0.f;
#pragma unroll
        for (int d8 = 0; d8 < D / 8; ++d8) {
            const uint4 u = *reinterpret_cast<const uint4*>(krow + d8 * 8);
            const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), e = unpack_bf16(u.w);
            const float* qq = &qs[warp][d8 * 8];
            s += a.x * qq[0] + a.y * qq[1] + b.x * qq[2] + b.y * qq[3] + c.x * qq[4] + c.y * qq[5] + e.x * qq[6] + e.y * qq[7];
        }
        const float m_new = fmaxf(m, s);
        const float f = __expf(m - m_new);
        const float pw = __expf(s - m_new);
        l = l * f + pw;
        const bf16* vrow = v + (long long)kj * ldkv + hkv * D;
#pragma unroll
        for (int d8 = 0; d8 < D / 8; ++d8) {
            const uint4 u = *reinterpret_cast<const uint4*>(vrow + d8 * 8);
            const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), e = unpack_bf16(u.w);
            float* ac = &acc[d8 * 8];
            ac[0] = ac[0] * f + pw * a.x; ac[1] = ac[1] * f + pw * a.y; ac[2] = ac[2] * f + pw * b.x; ac[3] = ac[3] * f + pw * b.y;
            ac[4] = ac[4] * f + pw * c.x; ac[5] = ac[5] * f + pw * c.y; ac[6] = ac[6] * f + pw * e.x; ac[7] = ac[7] * f + pw * e.y;
        }
        m = m_new;
    }
    float mg = m;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, off));
    const float f = (m == -INFINITY) ? 0.f : __expf(m - mg);
    l *= f;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
    const float inv = l > 0.f ? 1.0f / l : 0.f;          // an empty range gives a zero row (never produced by the callers)
    bf16* orow = o + (long long)qi * ldo + hd * D;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float a = acc[d] * f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == (d & 31)) orow[d] = __float2bfloat16_rn(a * inv);
    }
}

// ---- one-query decode attention over a KV cache (head dim 128) ----------------------------------------------------------------------
// A CTA per query head; its kWarps warps stride over the cached rows, a warp reads one 256-byte K / V row at a time (lane = 4 channels,
// fully coalesced), every lane of a warp carries the same online-softmax state, and the warps' partial (m, l, acc) are merged through
// shared memory.  ~Skv / kWarps dependent row steps per warp instead of Skv / 32 full-row dot products per lane in the generic kernel
// (r2: 46 us -> a few us per layer at Skv ~ 450, 28 layers per token).
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32) decode_attention_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                                                                       bf16* __restrict__ o, int group, int Skv, long long ldkv, float scale,
                                                                       const int* __restrict__ kv_len_ptr) {
    constexpr int D = 128;
    __shared__ float sm_m[kWarps], sm_l[kWarps];
    __shared__ float sm_acc[kWarps][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hd = blockIdx.x, hkv = hd / group;
    int n = kv_len_ptr ? *kv_len_ptr : Skv;
    if (n > Skv) n = Skv;
    const uint2 qu = *reinterpret_cast<const uint2*>(q + hd * D + lane * 4);
    const float2 q01 = unpack_bf16(qu.x), q23 = unpack_bf16(qu.y);
    const float q0 = q01.x * scale, q1 = q01.y * scale, q2 = q23.x * scale, q3 = q23.y * scale;
    float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const bf16* kb = k + hkv * D + lane * 4;
    const bf16* vb = v + hkv * D + lane * 4;
    for (int kj = warp; kj < n; kj += kWarps) {
        const uint2 ku = __ldg(reinterpret_cast<const uint2*>(kb + (long long)kj * ldkv));
        const uint2 vu = __ldg(reinterpret_cast<const uint2*>(vb + (long long)kj * ldkv));
        const float2 k01 = unpack_bf16(ku.x), k23 = unpack_bf16(ku.y);
        float s = q0 * k01.x + q1 * k01.y + q2 * k23.x + q3 * k23.y;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const float m_new = fmaxf(m, s);
        const float f = __expf(m - m_new), pw = __expf(s - m_new);
        const float2 v01 = unpack_bf16(vu.x), v23 = unpack_bf16(vu.y);
        l = l * f + pw;
        a0 = a0 * f + pw * v01.x; a1 = a1 * f + pw * v01.y; a2 = a2 * f + pw * v23.x; a3 = a3 * f + pw * v23.y;
        m = m_new;
    }
    if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
    sm_acc[warp][lane * 4 + 0] = a0; sm_acc[warp][lane * 4 + 1] = a1; sm_acc[warp][lane * 4 + 2] = a2; sm_acc[warp][lane * 4 + 3] = a3;
    __syncthreads();
    if (warp == 0) {
        float mg = -INFINITY;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) mg = fmaxf(mg, sm_m[w]);
        float lt = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const float f = sm_m[w] == -INFINITY ? 0.f : __expf(sm_m[w] - mg);
            lt += f * sm_l[w];
            r0 += f * sm_acc[w][lane * 4 + 0]; r1 += f * sm_acc[w][lane * 4 + 1]; r2 += f * sm_acc[w][lane * 4 + 2]; r3 += f * sm_acc[w][lane * 4 + 3];
        }
        const float inv = lt > 0.f ? 1.0f / lt : 0.f;
        uint2 ou;
        ou.x = pack_bf16(r0 * inv, r1 * inv);
        ou.y = pack_bf16(r2 * inv, r3 * inv);
        *reinterpret_cast<uint2*>(o + hd * D + lane * 4) = ou;
    }
}

// ---- the whole attention part of one decode step in ONE launch for up to 8 requests: rope (bf16 op order) of the new q / k heads, KV-cache
// append, attention of the new query over the cache + itself.  grid = (query heads, requests).  Every CTA rotates its own q head and the k head
// of its KV group (7 query heads share one: the rotation is 128 elements), reads the cached rows [0, n) like decode_attention_kernel and takes the
// NEW row from registers (as row n, by the warp n % kWarps that would have read it: same values, same order -> bit-identical to
// rope_kv_append_kernel + decode_attention_kernel); the first CTA of a KV group writes the rotated k and v into the cache.
struct DecodeReqDev {
    const bf16* qkv;      // [ (Hq + 2 Hkv) * 128 ] = q | k | v of the new token (not modified)
    bf16* cache_k;        // [cap, ldc]
    bf16* cache_v;
    bf16* out;            // [Hq * 128]
    const int* ctr;       // ctr[0] = cache rows before the append, ctr[1] = rope table row
    int cap;              // cache capacity (rows)
};
struct DecodeBatchDev {
    DecodeReqDev r[8];
};
__device__ __forceinline__ void rope4(const float (&x)[4], const float (&xp)[4], const float* __restrict__ cs, const float* __restrict__ sn, long long row,
                                      int D, int c0, bool hi, float (&o)[4]) {
    // rotate-half with the language model's rounding points (rope_half_kernel mode 1): channel c < D/2: bf16(bf16(x c) + bf16(-x' s)), c >= D/2:
    // bf16(bf16(x c) + bf16(x' s)); x' = the partner channel c +- D/2 (held by lane +- 16)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float c = bf16_round(cs[row * D + c0 + i]), sv = bf16_round(sn[row * D + c0 + i]);
        o[i] = bf16_round(bf16_round(x[i] * c) + bf16_round((hi ? xp[i] : -xp[i]) * sv));
    }
}
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32) decode_attention_fused_kernel(const DecodeBatchDev batch, int Hq, int Hkv, long long ldc,
                                                                             const float* __restrict__ cs, const float* __restrict__ sn, float scale) {
    constexpr int D = 128;
    __shared__ float sm_m[kWarps], sm_l[kWarps];
    __shared__ float sm_acc[kWarps][D];
    const DecodeReqDev& rq = batch.r[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hd = blockIdx.x, group = Hq / Hkv, hkv = hd / group;
    const int n_prev = min(rq.ctr[0], rq.cap - 1);
    const long long row = rq.ctr[1];
    const int c0 = lane * 4;
    const bool hi = lane >= 16;
    auto load4 = [&](const bf16* p, float (&f)[4]) {
        const uint2 u = *reinterpret_cast<const uint2*>(p + c0);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    };
    float qx[4], qp[4], kx[4], kp[4], vx[4], qr[4], kr[4];
    load4(rq.qkv + hd * D, qx);
    load4(rq.qkv + (Hq + hkv) * D, kx);
    load4(rq.qkv + (Hq + Hkv + hkv) * D, vx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        qp[i] = __shfl_xor_sync(0xffffffffu, qx[i], 16);
        kp[i] = __shfl_xor_sync(0xffffffffu, kx[i], 16);
    }
    rope4(qx, qp, cs, sn, row, D, c0, hi, qr);
    rope4(kx, kp, cs, sn, row, D, c0, hi, kr);
    if (hd % group == 0 && warp == 0) {                       // the KV group's first CTA appends the new row
        uint2 ku, vu;
        ku.x = pack_bf16(kr[0], kr[1]); ku.y = pack_bf16(kr[2], kr[3]);
        vu.x = pack_bf16(vx[0], vx[1]); vu.y = pack_bf16(vx[2], vx[3]);
        *reinterpret_cast<uint2*>(rq.cache_k + (long long)n_prev * ldc + hkv * D + c0) = ku;
        *reinterpret_cast<uint2*>(rq.cache_v + (long long)n_prev * ldc + hkv * D + c0) = vu;
    }
    const float q0 = qr[0] * scale, q1 = qr[1] * scale, q2 = qr[2] * scale, q3 = qr[3] * scale;
    float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const bf16* kb = rq.cache_k + hkv * D + c0;
    const bf16* vb = rq.cache_v + hkv * D + c0;
    auto step = [&](float k0, float k1, float k2, float k3, float v0, float v1, float v2, float v3) {
        float s = q0 * k0 + q1 * k1 + q2 * k2 + q3 * k3;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const float m_new = fmaxf(m, s);
        const float f = __expf(m - m_new), pw = __expf(s - m_new);
        l = l * f + pw;
        a0 = a0 * f + pw * v0; a1 = a1 * f + pw * v1; a2 = a2 * f + pw * v2; a3 = a3 * f + pw * v3;
        m = m_new;
    };
    for (int kj = warp; kj < n_prev; kj += kWarps) {
        const uint2 ku = __ldg(reinterpret_cast<const uint2*>(kb + (long long)kj * ldc));
        const uint2 vu = __ldg(reinterpret_cast<const uint2*>(vb + (long long)kj * ldc));
        const float2 k01 = unpack_bf16(ku.x), k23 = unpack_bf16(ku.y), v01 = unpack_bf16(vu.x), v23 = unpack_bf16(vu.y);
        step(k01.x, k01.y, k23.x, k23.y, v01.x, v01.y, v23.x, v23.y);
    }
    if (warp == n_prev % kWarps) step(kr[0], kr[1], kr[2], kr[3], vx[0], vx[1], vx[2], vx[3]);      // the new token's own row, from registers
    if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
    sm_acc[warp][c0 + 0] = a0; sm_acc[warp][c0 + 1] = a1; sm_acc[warp][c0 + 2] = a2; sm_acc[warp][c0 + 3] = a3;
    __syncthreads();
    if (warp == 0) {
        float mg = -INFINITY;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) mg = fmaxf(mg, sm_m[w]);
        float lt = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const float f = sm_m[w] == -INFINITY ? 0.f : __expf(sm_m[w] - mg);
            lt += f * sm_l[w];
            r0 += f * sm_acc[w][c0 + 0]; r1 += f * sm_acc[w][c0 + 1]; r2 += f * sm_acc[w][c0 + 2]; r3 += f * sm_acc[w][c0 + 3];
        }
        const float inv = lt > 0.f ? 1.0f / lt : 0.f;
        uint2 ou;
        ou.x = pack_bf16(r0 * inv, r1 * inv);
        ou.y = pack_bf16(r2 * inv, r3 * inv);
        *reinterpret_cast<uint2*>(rq.out + hd * D + c0) = ou;
    }
}

// ---- out[i, :] = table[ids[i], :] ; ids < 0 leave the row untouched (used to scatter image embeddings into the token stream) ----------
__global__ void gather_rows_kernel(const bf16* __restrict__ table, long long ldt, const long long* __restrict__ ids, bf16* __restrict__ out,
                                   long long ldo, int n, int C) {
    const int r = blockIdx.x;
    if (r >= n) return;
    const long long id = ids[r];
    if (id < 0) return;
    const int nvec = C >> 3;
    for (int vv = threadIdx.x; vv < nvec; vv += blockDim.x)
        *reinterpret_cast<uint4*>(out + r * ldo + vv * 8) = __ldg(reinterpret_cast<const uint4*>(table + id * ldt + vv * 8));
}

// ---- first index of the maximum of a bf16 vector (torch.argmax semantics) -> int64; one CTA -----------------------------------------
__global__ void __launch_bounds__(1024) argmax_kernel(const bf16* __restrict__ x, int n, long long* __restrict__ out, long long* __restrict__ log,
                                                      const int* __restrict__ log_pos) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float val = __bfloat162float(x[i]);
        if (val > best || (val == best && i < bi)) { best = val; bi = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) {
            out[0] = bi;
            if (log != nullptr) log[log_pos ? *log_pos : 0] = bi;       // generated-token log, indexed by the device-side step counter
        }
    }
}

// ---- KV cache append at a device-side position, then advance the counters -----------------------------------------------------------
// cache_k / cache_v [max_len, C]; k_new / v_new [C]; pos[0] = number of cached rows (row to write), pos[1] = rope table row, pos[2] = step.
__global__ void kv_append_kernel(const bf16* __restrict__ k_new, const bf16* __restrict__ v_new, bf16* __restrict__ cache_k, bf16* __restrict__ cache_v,
                                 long long ldc, int C, const int* __restrict__ pos) {
    const long long row = pos[0];
    for (int i = threadIdx.x; i < (C >> 3); i += blockDim.x) {
        *reinterpret_cast<uint4*>(cache_k + row * ldc + i * 8) = *reinterpret_cast<const uint4*>(k_new + i * 8);
        *reinterpret_cast<uint4*>(cache_v + row * ldc + i * 8) = *reinterpret_cast<const uint4*>(v_new + i * 8);
    }
}
// rope (bf16 op order, mode 1 of rope_half_kernel) on the q and k heads of ONE decode row qkv = [q | k | v], then the KV-cache append of
// the rotated k and of v -- one launch instead of three per layer and request.
__global__ void rope_kv_append_kernel(bf16* __restrict__ qkv, int Hq, int Hkv, int D, const float* __restrict__ cs, const float* __restrict__ sn,
                                      bf16* __restrict__ cache_k, bf16* __restrict__ cache_v, long long ldc, const int* __restrict__ ctr) {
    const int half = D >> 1;
    const int n_rot = (Hq + Hkv) * half, n_v = Hkv * D;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const long long cache_row = ctr[0], row = ctr[1];
    if (gid < n_rot) {
        const int i = gid % half, hd = gid / half;
        bf16* p = qkv + hd * D;
        const float x1 = __bfloat162float(p[i]), x2 = __bfloat162float(p[i + half]);
        const float c1 = bf16_round(cs[row * D + i]), c2 = bf16_round(cs[row * D + i + half]);
        const float s1 = bf16_round(sn[row * D + i]), s2 = bf16_round(sn[row * D + i + half]);
        const bf16 o1 = __float2bfloat16_rn(bf16_round(x1 * c1) + bf16_round(-x2 * s1));
        const bf16 o2 = __float2bfloat16_rn(bf16_round(x2 * c2) + bf16_round(x1 * s2));
        p[i] = o1;
        p[i + half] = o2;
        if (hd >= Hq) {
            bf16* ck = cache_k + cache_row * ldc + (hd - Hq) * D;
            ck[i] = o1;
            ck[i + half] = o2;
        }
    } else if (gid < n_rot + n_v) {
        const int j = gid - n_rot;
        cache_v[cache_row * ldc + j] = qkv[(Hq + Hkv) * D + j];
    }
}
__global__ void advance_kernel(int* __restrict__ pos, int n) {
    if (threadIdx.x < n) pos[threadIdx.x] += 1;
}

}  // namespace

int swiglu_run(Handle* h, const void* x, int64_t ldx, void* out, int64_t ldo, int rows, int I, cudaStream_t s) {
    PE_REQUIRE(h, x && out && rows > 0 && I > 0 && I % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "pe_swiglu: bad arguments (rows=%d I=%d)", rows, I);
    const long long n = (long long)rows * (I / 8);
    swiglu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(static_cast<const bf16*>(x), ldx, static_cast<bf16*>(out), ldo, rows, I);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int rope_half_run(Handle* h, void* x, int64_t ldx, int T, int H, int D, const float* cs, const float* sn, const int* row_ptr, int row0, int mode,
                  cudaStream_t s) {
    PE_REQUIRE(h, x && cs && sn && T > 0 && H > 0 && D > 0 && D % 2 == 0, "pe_rope_half: bad arguments (T=%d H=%d D=%d)", T, H, D);
    PE_REQUIRE(h, mode == 0 || mode == 1, "pe_rope_half: mode must be 0 (fp32 fused) or 1 (bf16 op order)");
    const long long n = (long long)T * H * (D / 2);
    rope_half_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(static_cast<bf16*>(x), ldx, T, H, D, cs, sn, row_ptr, row0, mode);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int range_attention_run(Handle* h, const void* q, const void* k, const void* v, void* o, int H, int Hkv, int Sq, int Skv, int D, int64_t ldq,
                        int64_t ldkv, int64_t ldo, float scale, const int* kv_lo, const int* kv_hi, const int* kv_len_ptr, cudaStream_t s) {
    PE_REQUIRE(h, q && k && v && o, "pe_range_attention: null pointer");
    PE_REQUIRE(h, H > 0 && Hkv > 0 && H % Hkv == 0 && Sq > 0 && Skv > 0, "pe_range_attention: bad head / sequence sizes (H=%d Hkv=%d Sq=%d Skv=%d)", H, Hkv, Sq, Skv);
    PE_REQUIRE(h, D == 64 || D == 80 || D == 128, "pe_range_attention: head dim must be 64, 80 or 128 (got %d)", D);
    PE_REQUIRE(h, ldq % 8 == 0 && ldkv % 8 == 0, "pe_range_attention: ldq / ldkv must be multiples of 8");
    const dim3 grid(ceil_div(Sq, 4), H);
    const bf16 *qb = static_cast<const bf16*>(q), *kb = static_cast<const bf16*>(k), *vb = static_cast<const bf16*>(v);
    bf16* ob = static_cast<bf16*>(o);
    const int group = H / Hkv;
    if (Sq == 1 && D == 128 && kv_lo == nullptr && kv_hi == nullptr) {          // one-token decode over the KV cache
        decode_attention_kernel<16><<<H, 16 * 32, 0, s>>>(qb, kb, vb, ob, group, Skv, ldkv, scale, kv_len_ptr);
        PE_CHECK_CUDA(h, cudaGetLastError());
        return PE_OK;
    }
    if (D == 64) range_attention_kernel<64><<<grid, 128, 0, s>>>(qb, kb, vb, ob, H, group, Sq, Skv, ldq, ldkv, ldo, scale, kv_lo, kv_hi, kv_len_ptr);
    else if (D == 80) range_attention_kernel<80><<<grid, 128, 0, s>>>(qb, kb, vb, ob, H, group, Sq, Skv, ldq, ldkv, ldo, scale, kv_lo, kv_hi, kv_len_ptr);
    else range_attention_kernel<128><<<grid, 128, 0, s>>>(qb, kb, vb, ob, H, group, Sq, Skv, ldq, ldkv, ldo, scale, kv_lo, kv_hi, kv_len_ptr);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int gather_rows_run(Handle* h, const void* table, int64_t ldt, const int64_t* ids, void* out, int64_t ldo, int n, int C, cudaStream_t s) {
    PE_REQUIRE(h, table && ids && out && n > 0 && C > 0 && C % 8 == 0 && ldt % 8 == 0 && ldo % 8 == 0, "pe_gather_rows: bad arguments (n=%d C=%d)", n, C);
    gather_rows_kernel<<<n, 128, 0, s>>>(static_cast<const bf16*>(table), ldt, reinterpret_cast<const long long*>(ids), static_cast<bf16*>(out), ldo, n, C);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int argmax_run(Handle* h, const void* x, int n, int64_t* out, int64_t* log, const int* log_pos, cudaStream_t s) {
    PE_REQUIRE(h, x && out && n > 0, "pe_argmax: bad arguments (n=%d)", n);
    argmax_kernel<<<1, 1024, 0, s>>>(static_cast<const bf16*>(x), n, reinterpret_cast<long long*>(out), reinterpret_cast<long long*>(log), log_pos);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int kv_append_run(Handle* h, const void* k_new, const void* v_new, void* cache_k, void* cache_v, int64_t ldc, int C, const int* pos, cudaStream_t s) {
    PE_REQUIRE(h, k_new && v_new && cache_k && cache_v && pos && C > 0 && C % 8 == 0 && ldc % 8 == 0, "pe_kv_append: bad arguments (C=%d)", C);
    kv_append_kernel<<<1, 128, 0, s>>>(static_cast<const bf16*>(k_new), static_cast<const bf16*>(v_new), static_cast<bf16*>(cache_k),
                                       static_cast<bf16*>(cache_v), ldc, C, pos);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int rope_kv_append_run(Handle* h, void* qkv, int Hq, int Hkv, int D, const float* cs, const float* sn, void* cache_k, void* cache_v, int64_t ldc,
                       const int* ctr, cudaStream_t s) {
    PE_REQUIRE(h, qkv && cs && sn && cache_k && cache_v && ctr && Hq > 0 && Hkv > 0 && D > 0 && D % 2 == 0, "pe_rope_kv_append: bad arguments");
    const int n = (Hq + Hkv) * (D / 2) + Hkv * D;
    rope_kv_append_kernel<<<(n + 255) / 256, 256, 0, s>>>(static_cast<bf16*>(qkv), Hq, Hkv, D, cs, sn, static_cast<bf16*>(cache_k), static_cast<bf16*>(cache_v), ldc, ctr);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int decode_attention_fused_run(Handle* h, const pe_decode_req* reqs, int n_req, int Hq, int Hkv, int D, int64_t ldc, const float* cs,
                               const float* sn, float scale, cudaStream_t s) {
    PE_REQUIRE(h, reqs && cs && sn && n_req >= 1 && n_req <= 8, "pe_decode_attention_fused: 1..8 requests (got %d)", n_req);
    PE_REQUIRE(h, D == 128 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0 && ldc % 8 == 0 && ldc >= (int64_t)Hkv * D,
               "pe_decode_attention_fused: head dim 128, Hq a multiple of Hkv, ldc >= Hkv * 128 (Hq=%d Hkv=%d D=%d)", Hq, Hkv, D);
    DecodeBatchDev b;
    memset(&b, 0, sizeof(b));
    for (int i = 0; i < n_req; ++i) {
        PE_REQUIRE(h, reqs[i].qkv && reqs[i].cache_k && reqs[i].cache_v && reqs[i].out && reqs[i].counters && reqs[i].cache_rows > 0,
                   "pe_decode_attention_fused: request %d has a null pointer or an empty cache", i);
        b.r[i] = DecodeReqDev{static_cast<const bf16*>(reqs[i].qkv), static_cast<bf16*>(reqs[i].cache_k), static_cast<bf16*>(reqs[i].cache_v),
                              static_cast<bf16*>(reqs[i].out), reqs[i].counters, (int)reqs[i].cache_rows};
    }
    decode_attention_fused_kernel<16><<<dim3(Hq, n_req), 16 * 32, 0, s>>>(b, Hq, Hkv, ldc, cs, sn, scale);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

int advance_run(Handle* h, int* pos, int n, cudaStream_t s) {
    PE_REQUIRE(h, pos && n > 0 && n <= 32, "pe_advance: bad arguments");
    advance_kernel<<<1, 32, 0, s>>>(pos, n);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

}  // namespace pe
