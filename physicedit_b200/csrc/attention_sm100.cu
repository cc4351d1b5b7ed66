// Joint (text+image) non-causal attention forward, head dim 128, on tcgen05 / TMEM / TMA.
//
// Replaces qwen_image_flash_attention -> F.scaled_dot_product_attention
// (DiffSynth-Studio/diffsynth/models/qwen_image_dit.py:14-39) on q,k,v produced by
// QwenDoubleStreamAttention.forward (:274-316).  S x S is never materialised.
//
// One persistent CTA per SM; a work item is (head, block of kQT*128 query rows).
//   warp 0          TMA producer : Q tiles once per item, then K_0,V_0,K_1,V_1,... through a ring of
//                                  32 KB buffers (each tile = two [128 x 64] 128B-swizzled halves)
//   warp 1          MMA issuer   : S_q = Q_q K_j^T  (SS, 128x128x16 x8)   -> TMEM  (fp32, 128 columns)
//                                  O_q += P_q V_j   (P from TMEM ("TS") or from smem, V MN-major)
//   warp 2          TMEM allocator
//   warps 4..4+4kQT softmax      : one warpgroup per query tile, one thread per query row:
//                                  tcgen05.ld S -> running max / exp2 / row sum -> P (bf16) written
//                                  over S in TMEM (or to swizzled smem) -> mbarrier -> PV MMA.
//                                  The O accumulator is rescaled lazily (only when the running max
//                                  grows by more than 2^8), so the common KV step never touches O.
// With kQT = 2 the two query tiles ping-pong: the tensor core computes S_1 / PV_1 while the
// softmax warpgroup of tile 0 works, and vice versa.
#include <type_traits>
#include "ptx.cuh"
#include "common.cuh"

namespace pe {
namespace {

constexpr int kTile = 128;           // query rows per tile, kv rows per tile, head dim
constexpr int kHalfBytes = 128 * 64 * 2;   // one [128 x 64] bf16 swizzled half = 16 KB
constexpr int kTileBytes = 2 * kHalfBytes; // 32 KB

// Optional in-kernel timeline (build with -DPE_ATTN_TRACE): CTA 0 logs (event id, step, clock) triples of its first work
// item into the handle's workspace; tools/attn_trace.py turns them into a per-step latency breakdown.
#ifdef PE_ATTN_TRACE
// fire-and-forget stores into a per-role region (role 0 = MMA issuer, 1 / 2 = softmax warp 0 of tile 0 / 1); the slot
// counter lives in a register, so a trace point costs a clock read and one store (no round trip).
#define PE_TRACE_DECL(role) unsigned int trace_n = 0; const int trace_role = (role);
#define PE_TRACE(ev, step)                                                                                   \
    do {                                                                                                     \
        if (blockIdx.x == 0 && lane_id() == 0 && p.trace != nullptr && trace_n < 2000u) {                    \
            long long* tp = p.trace + 2 + trace_role * 4000 + 2 * trace_n;                                   \
            tp[0] = (static_cast<long long>(ev) << 32) | static_cast<unsigned int>(step);                    \
            tp[1] = clock64();                                                                               \
            ++trace_n;                                                                                       \
        }                                                                                                    \
    } while (0)
#else
#define PE_TRACE_DECL(role)
#define PE_TRACE(ev, step) do {} while (0)
#endif

struct AttnParams {
    CUtensorMap tmQ, tmK, tmV;
    bf16* o;
    long long ldo;
    int S;
    int H;
    int n_qblk;      // blocks of kQT*128 query rows
    int n_items;     // H * n_qblk
    int n_kv;        // ceil(S / 128)
    float scale_log2;
    unsigned int* abort_flag;
    long long* trace;
    int* item_flags;  // kernel 4: per work item, the sequence number of the launch whose trailing reference overflowed on it
    int seq;          // this launch's sequence number
    int only_flagged; // kernel 1 as the fix-up pass of kernel 4: process only items with item_flags[item] == seq
    // output routed by query row (pe_attention_fwd_routed, kernel 1): rows < route_end[i] (first match) go to route_o[i] + row * ldo + head * 128
    int n_route;
    int route_end[8];
    bf16* route_o[8];
    float* lse;      // kernel 1, optional: lse[head * S + row] = log2 of the row's softmax denominator, scale included (for the backward)
};

__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// ---- lean MMA issue: descriptors as (lo, hi) 32-bit words so that stepping through a tile is one uniform add per operand ----
// smem matrix descriptor, 128B swizzle: lo = start address >> 4 | (LBO >> 4) << 16 ; hi = SBO >> 4 | version(1) << 14 | swizzle(2) << 29
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t kDescLboK = (16u >> 4) << 16;                          // K-major operands (Q, K, P-in-smem): LBO unused
constexpr uint32_t kDescLboV = (static_cast<uint32_t>(kHalfBytes) >> 4) << 16;   // MN-major V: 16 KB between the two 64-column halves
__device__ __forceinline__ uint32_t desc_lo_k(uint32_t addr) { return ((addr & 0x3ffffu) >> 4) | kDescLboK; }
__device__ __forceinline__ uint32_t desc_lo_v(uint32_t addr) { return ((addr & 0x3ffffu) >> 4) | kDescLboV; }
template <bool kAcc>
__device__ __forceinline__ void umma_ss_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(kAcc ? 1u : 0u) : "memory");
}
__device__ __forceinline__ void umma_ts_lohi(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss_lohi_acc(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(acc) : "memory");
}

// ---- softmax helpers (one thread = one query row; a chunk = 32 consecutive kv columns) ------------------------------
#ifndef PE_ATTN_POLY_EVERY
#define PE_ATTN_POLY_EVERY 0      // >0: 1 of every N packed pairs takes the FMA-pipe exp2 instead of MUFU (measured: no gain on B200 with one softmax warp per SMSP)
#endif
// 2^x on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax on [-0.5, 0.5], rel. error ~1e-4, far below
// bf16's 2^-9): takes load off the MUFU unit, which is the co-bottleneck of the softmax on B200 (16 ex2/clk/SM).
__device__ __forceinline__ float exp2_fma(float x) {
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;                 // 1.5 * 2^23: round-to-nearest integer lands in the low mantissa bits
    const float f = x - (t - 12582912.0f);           // f in [-0.5, 0.5]
    float pz = fmaf(f, 0.05550410866f, 0.24022650696f);
    pz = fmaf(pz, f, 0.69314718056f);
    pz = fmaf(pz, f, 1.0f);
    return __int_as_float(__float_as_int(pz) + (__float_as_int(t) << 23));
}
// ---- packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2): two elements per FMA-pipe issue slot ----
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
#ifndef PE_ATTN_ROLES_LAST
#define PE_ATTN_ROLES_LAST 0      // 1: softmax warps are warps 0.., TMA / MMA / allocator warps come last (highest warp id on their SMSP).
                                  // r2 A/B on B200: no effect on kernel 1 (0.8073 vs 0.8075 ms), kernel 4 slower (0.845 vs 0.815) -> off
#endif
#ifndef PE_ATTN_DBG
#define PE_ATTN_DBG 0             // timing experiments only (wrong results): 1 = no exp, 2 = no P store, 4 = no max exchange, 8 = no S load
#endif
#ifndef PE_ATTN_POLY_PAIRS
#define PE_ATTN_POLY_PAIRS 0      // of every 8 (element-pairs), how many take the FMA-pipe exp2 instead of MUFU (r1: 0..4 measured, no gain)
#endif
// (p0, p1) = 2^(x0, x1) for a packed pair on the FMA pipe: Cody-Waite split + degree-3 polynomial, all in f32x2
__device__ __forceinline__ void exp2_fma_pair(uint64_t x2, float& p0, float& p1) {
    float x0, x1;
    upk2(x2, x0, x1);
    x2 = pk2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
    const uint64_t magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f), neg1 = pk2(-1.0f, -1.0f);
    const uint64_t t2 = fadd2(x2, magic);                  // integer part lands in the low mantissa bits
    const uint64_t n2 = fadd2(t2, nmagic);
    const uint64_t f2 = ffma2(n2, neg1, x2);               // f = x - n in [-0.5, 0.5]
    uint64_t pz = ffma2(f2, pk2(0.05550410866f, 0.05550410866f), pk2(0.24022650696f, 0.24022650696f));
    pz = ffma2(pz, f2, pk2(0.69314718056f, 0.69314718056f));
    pz = ffma2(pz, f2, pk2(1.0f, 1.0f));
    float z0, z1, t0, t1;
    upk2(pz, z0, z1);
    upk2(t2, t0, t1);
    p0 = __int_as_float(__float_as_int(z0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(z1) + (__float_as_int(t1) << 23));
}
__device__ __forceinline__ void row_max_chunk(const uint32_t (&r)[32], int col0, int kv_valid, bool full, float& mx) {
    if (full) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (col0 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(r[i]));
    }
}
__device__ __forceinline__ void softmax_chunk(const uint32_t (&r)[32], int col0, int kv_valid, bool full, float scale_log2, float m,
                                              float& lsum, float& mx, uint32_t (&pk)[16]) {
    if (full) {
        float s0 = 0.f, s1 = 0.f, m0 = mx, m1 = mx;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(r[2 * i + 2]), __uint_as_float(r[2 * i + 3])));
        }
        mx = fmaxf(m0, m1);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float x0 = fmaf(__uint_as_float(r[2 * i]), scale_log2, -m);
            const float x1 = fmaf(__uint_as_float(r[2 * i + 1]), scale_log2, -m);
            float p0, p1;
            if (PE_ATTN_POLY_EVERY > 0 && (i % (PE_ATTN_POLY_EVERY > 0 ? PE_ATTN_POLY_EVERY : 1)) == 1) { p0 = exp2_fma(x0); p1 = exp2_fma(x1); }
            else { p0 = ex2(x0); p1 = ex2(x1); }
            s0 += p0;
            s1 += p1;
            pk[i] = pack_bf16(p0, p1);
        }
        lsum += s0 + s1;
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float p0 = ex2(fmaf(__uint_as_float(r[2 * i]), scale_log2, -m));
            float p1 = ex2(fmaf(__uint_as_float(r[2 * i + 1]), scale_log2, -m));
            if (col0 + 2 * i >= kv_valid) p0 = 0.f; else mx = fmaxf(mx, __uint_as_float(r[2 * i]));
            if (col0 + 2 * i + 1 >= kv_valid) p1 = 0.f; else mx = fmaxf(mx, __uint_as_float(r[2 * i + 1]));
            lsum += p0 + p1;
            pk[i] = pack_bf16(p0, p1);
        }
    }
}
// O[row, :] *= f for this thread's row (128 fp32 columns in TMEM); warp-collective
__device__ __forceinline__ void scale_o_rows(uint32_t o_addr, float f, uint32_t (&tmp)[32]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        tmem_ld32(o_addr + c * 32, tmp);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) tmp[i] = __float_as_uint(__uint_as_float(tmp[i]) * f);
        tmem_st32(o_addr + c * 32, tmp);
    }
    tmem_st_wait();
}
// P chunk c (32 kv columns, 16 packed words) -> TMEM (overlaying the consumed S columns [16c, 16c+16)) or swizzled smem
template <bool kPTmem>
__device__ __forceinline__ void store_p_chunk(const uint32_t (&pk)[16], int c, uint32_t p_row, int row_in_tile) {
    if (kPTmem) {
        tmem_st16(p_row + c * 16, pk);
    } else {
        // K-major 128B-swizzled [128 x 64] halves: row r at r*128, 16-byte chunk index ^ (r & 7)
        const uint32_t rowb = p_row + (c >> 1) * kHalfBytes;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int chunk = (c & 1) * 4 + v;
            st_shared_v4(rowb + ((chunk ^ (row_in_tile & 7)) << 4), pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
        }
    }
}

template <int kQT, bool kPTmem>
struct AttnCfg {
    static constexpr int kKV = kPTmem ? (kQT == 2 ? 5 : 6) : (kQT == 2 ? 3 : 5);     // K/V ring depth (32 KB each)
    static constexpr int kPBytes = kPTmem ? 0 : kQT * kTileBytes;
    static constexpr int kSmemData = kQT * kTileBytes + kKV * kTileBytes + kPBytes;
    static constexpr int kSmem = 1024 + kSmemData + 256;
    static constexpr int kThreads = 128 + 128 * kQT;
    static constexpr int kTmemCols = kQT == 2 ? 512 : 256;
};

template <int kQT, bool kPTmem>
__global__ void __launch_bounds__(AttnCfg<kQT, kPTmem>::kThreads, 1) attention_kernel(const __grid_constant__ AttnParams p) {
    using Cfg = AttnCfg<kQT, kPTmem>;
    constexpr int kKV = Cfg::kKV;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    auto q_smem = [&](int q) { return smem_base + q * kTileBytes; };
    auto kv_smem = [&](int s) { return smem_base + (kQT + s) * kTileBytes; };
    auto p_smem = [&](int q) { return smem_base + (kQT + kKV + q) * kTileBytes; };
    const uint32_t bar_base = smem_base + Cfg::kSmemData;
    // barrier map (8 bytes each)
    const uint32_t q_full = bar_base, q_empty = bar_base + 8;
    auto kv_full = [&](int s) { return bar_base + 16 + s * 8; };
    auto kv_empty = [&](int s) { return bar_base + 16 + (kKV + s) * 8; };
    auto s_full = [&](int q) { return bar_base + 16 + (2 * kKV + q) * 8; };
    auto p_full = [&](int q) { return bar_base + 16 + (2 * kKV + 2 + q) * 8; };
    auto pv_done = [&](int q) { return bar_base + 16 + (2 * kKV + 4 + q) * 8; };
    auto o_empty = [&](int q) { return bar_base + 16 + (2 * kKV + 6 + q) * 8; };
    const uint32_t tmem_slot = bar_base + 16 + (2 * kKV + 8) * 8;

    // Warp roles: softmax warps FIRST, the TMA / MMA / TMEM-allocator warps LAST.  The SM sub-partition arbiter favours the highest
    // warp id among eligible warps, and every instruction of the single MMA-issuer warp is on the tensor pipe's critical path
    // (UTCHMMA issue blocks on a shallow queue, so the pipe idles whenever the issuer is held up): with the issuer as warp 1 it
    // competed with -- and lost to -- the softmax warps of its sub-partition (r2 trace: ~700 cycles of turn-around per KV step).
    constexpr int kSmWarps = PE_ATTN_ROLES_LAST ? 4 * kQT : 0;
    const int warp_raw = threadIdx.x >> 5;
    const int warp = PE_ATTN_ROLES_LAST ? (warp_raw >= kSmWarps ? warp_raw - kSmWarps : warp_raw + 4) : warp_raw;   // logical: 0 TMA, 1 MMA, 2 alloc, 4.. softmax
    const int lane = lane_id();

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&p.tmQ);
        prefetch_tmap(&p.tmK);
        prefetch_tmap(&p.tmV);
    }
    if (warp == 1 && elect_one()) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < kKV; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
        for (int q = 0; q < kQT; ++q) {
            mbar_init(s_full(q), 1);
            mbar_init(p_full(q), 4);
            mbar_init(pv_done(q), 1);
            mbar_init(o_empty(q), 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ld_shared_u32(tmem_slot);
    auto s_tmem = [&](int q) { return tmem_base + q * 128; };
    auto o_tmem = [&](int q) { return tmem_base + kQT * 128 + q * 128; };

    if (warp == 0) {
        // ======================================= TMA producer =======================================
        uint32_t n = 0;          // ring sequence number (K_0, V_0, K_1, V_1, ...), continues across items
        uint32_t it = 0;
        bool ok = true;
        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x, ++it) {
            if (p.only_flagged && p.item_flags[item] != p.seq) { --it; continue; }     // fix-up pass of kernel 4: every role skips alike
            const int head = item / p.n_qblk;
            const int qb = item - head * p.n_qblk;
            const int col0 = head * kTile;
            if (!mbar_wait(q_empty, (it & 1u) ^ 1u, p.abort_flag, 10)) break;
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, kQT * kTileBytes);
#pragma unroll
                for (int q = 0; q < kQT; ++q) {
                    const int row0 = (qb * kQT + q) * kTile;
                    tma_load_2d(q_smem(q), &p.tmQ, q_full, col0, row0);
                    tma_load_2d(q_smem(q) + kHalfBytes, &p.tmQ, q_full, col0 + 64, row0);
                }
            }
            __syncwarp();
            for (int j = 0; j < p.n_kv && ok; ++j) {
#pragma unroll
                for (int kv = 0; kv < 2; ++kv, ++n) {
                    const int slot = n % kKV;
                    const uint32_t ph = (n / kKV) & 1u;
                    if (!mbar_wait(kv_empty(slot), ph ^ 1u, p.abort_flag, 11)) { ok = false; break; }
                    if (elect_one()) {
                        const CUtensorMap* tm = kv == 0 ? &p.tmK : &p.tmV;
                        mbar_arrive_expect_tx(kv_full(slot), kTileBytes);
                        tma_load_2d(kv_smem(slot), tm, kv_full(slot), col0, j * kTile);
                        tma_load_2d(kv_smem(slot) + kHalfBytes, tm, kv_full(slot), col0 + 64, j * kTile);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ======================================= MMA issuer =======================================
        // One elected lane issues everything; the warp stays converged (the uniform datapath that feeds UTCHMMA needs it).
        // r1 elimination runs showed this warp's own serial time line (blocking UTCHMMA issue + the instructions between
        // batches) is what caps the kernel, so the per-batch overhead is kept minimal: the leader is elected once, PV_q(j) and
        // S_q(j+1) go out as ONE batch, and every descriptor is "base word + small constant".
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);     // Q (K-major) x K (K-major)
        constexpr uint32_t idesc_o = make_idesc_bf16(128, 128, 0, 1);     // P (K-major) x V (MN-major)
        const bool leader = elect_one();
        uint32_t n = 0, it = 0;
        uint32_t p_phase[2] = {0, 0};
        bool ok = true;
        PE_TRACE_DECL(0)

        auto issue_s = [&](int q, uint32_t k_base) {
            // S_q = Q_q K^T : 8 k-steps of 16 head-dim elements; +32 B inside a swizzle row, +16 KB per half
            const uint32_t qlo = desc_lo_k(q_smem(q)), klo = desc_lo_k(k_base);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint32_t off16 = ((kk >> 2) * kHalfBytes + (kk & 3) * 32) >> 4;
                if (kk == 0) umma_ss_lohi<false>(s_tmem(q), qlo + off16, klo + off16, idesc_s);
                else umma_ss_lohi<true>(s_tmem(q), qlo + off16, klo + off16, idesc_s);
            }
            umma_commit(s_full(q));
        };
        auto issue_pv = [&](int q, uint32_t v_base, bool accumulate, bool last) {
            // O_q (+)= P_q V : 8 k-steps of 16 kv rows; V rows are 128 B apart, 16 rows = 2 KB
            const uint32_t vlo = desc_lo_v(v_base);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint32_t acc = (accumulate || kk != 0) ? 1u : 0u;
                if (kPTmem) {
                    umma_ts_lohi(o_tmem(q), s_tmem(q) + kk * 8, vlo + kk * (2048 >> 4), idesc_o, acc);
                } else {
                    const uint32_t off16 = ((kk >> 2) * kHalfBytes + (kk & 3) * 32) >> 4;
                    umma_ss_lohi_acc(o_tmem(q), desc_lo_k(p_smem(q)) + off16, vlo + kk * (2048 >> 4), idesc_o, acc);
                }
            }
            // only the item's last PV is ever waited for (by the epilogue): one commit per item keeps every phase of the barrier
            // observed by its waiter (compute-sanitizer synccheck flags arrivals on phases nobody waits for)
            if (last) umma_commit(pv_done(q));
        };

        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x, ++it) {
            if (p.only_flagged && p.item_flags[item] != p.seq) { --it; continue; }
            if (!mbar_wait(q_full, it & 1u, p.abort_flag, 20)) break;
            // K_0
            uint32_t k_slot = n % kKV;
            if (!mbar_wait(kv_full(k_slot), (n / kKV) & 1u, p.abort_flag, 21)) break;
            ++n;
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int q = 0; q < kQT; ++q) issue_s(q, kv_smem(k_slot));
                umma_commit(kv_empty(k_slot));
                if (p.n_kv == 1) umma_commit(q_empty);
            }
            __syncwarp();
            for (int j = 0; j < p.n_kv && ok; ++j) {
                const uint32_t v_slot = n % kKV;
                if (!mbar_wait(kv_full(v_slot), (n / kKV) & 1u, p.abort_flag, 22)) { ok = false; break; }
                ++n;
                const bool more = j + 1 < p.n_kv;
                if (more) {
                    k_slot = n % kKV;
                    if (!mbar_wait(kv_full(k_slot), (n / kKV) & 1u, p.abort_flag, 23)) { ok = false; break; }
                    ++n;
                }
#pragma unroll
                for (int q = 0; q < kQT; ++q) {
                    if (j == 0) {
                        // previous item's epilogue must have drained O_q
                        if (!mbar_wait(o_empty(q), (it & 1u) ^ 1u, p.abort_flag, 24)) { ok = false; break; }
                    }
                    PE_TRACE(10 + q, j);
                    if (!mbar_wait(p_full(q), p_phase[q], p.abort_flag, 25)) { ok = false; break; }
                    p_phase[q] ^= 1u;
                    tc_fence_after();
                    PE_TRACE(12 + q, j);
                    if (leader) {
                        issue_pv(q, kv_smem(v_slot), j > 0, !more);
                        if (more) issue_s(q, kv_smem(k_slot));
                        if (q == kQT - 1) {
                            umma_commit(kv_empty(v_slot));
                            if (more) umma_commit(kv_empty(k_slot));
                            if (j + 2 == p.n_kv) umma_commit(q_empty);   // last S MMAs of this item were just issued
                        }
                    }
                    __syncwarp();
                    PE_TRACE(14 + q, j);
                }
                if (!ok) break;
            }
        }
    } else if (warp >= 4) {
        // ======================================= softmax / correction / epilogue =======================================
        const int q = (warp - 4) >> 2;
        const int wq = warp & 3;
        const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
        const uint32_t s_addr = s_tmem(q) + lane_off;
        const uint32_t o_addr = o_tmem(q) + lane_off;
        const int row_in_tile = wq * 32 + lane;
        uint32_t s_phase = 0, item_par = 0;
        bool ok = true;
        PE_TRACE_DECL(1 + q)
        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x) {
            if (p.only_flagged && p.item_flags[item] != p.seq) continue;
            const int head = item / p.n_qblk;
            const int qb = item - head * p.n_qblk;
            // Softmax state.  m_ref is the exponent reference (log2 domain) of everything accumulated in O and l so far.
            // It trails the true running max: a KV step exponentiates against the max of the PREVIOUS steps (one pass over
            // S instead of two), and the reference is only moved -- with a deferred O rescale -- when the max grew by more
            // than 2^8.  A stale (lower) reference scales P, O and l by the same factor, which the final O/l cancels; a guard
            // re-does the step exactly in the (practically unreachable) case where the jump could overflow.
            float m_ref = -INFINITY;
            float l = 0.f;
            float f_pending = 1.0f;        // O must still be multiplied by this (applied after the previous PV finished)
            for (int j = 0; j < p.n_kv; ++j) {
                if (wq == 0) PE_TRACE(20 + q, j);
                if (!mbar_wait(s_full(q), s_phase, p.abort_flag, 30)) { ok = false; break; }
                s_phase ^= 1u;
                tc_fence_after();
                if (wq == 0) PE_TRACE(22 + q, j);
                const int kv_valid = p.S - j * kTile;     // >= 128 except possibly on the last tile
                const bool full = kv_valid >= kTile;
                uint32_t ra[32], rb[32];
                if (j == 0) {
                    // first tile: exact row max (two passes), TMEM loads software-pipelined
                    float mx = -INFINITY;
                    tmem_ld32(s_addr, ra);
                    tmem_ld_wait();
                    tmem_ld32(s_addr + 32, rb);
                    row_max_chunk(ra, 0, kv_valid, full, mx);
                    tmem_ld_wait();
                    tmem_ld32(s_addr + 64, ra);
                    row_max_chunk(rb, 32, kv_valid, full, mx);
                    tmem_ld_wait();
                    tmem_ld32(s_addr + 96, rb);
                    row_max_chunk(ra, 64, kv_valid, full, mx);
                    tmem_ld_wait();
                    row_max_chunk(rb, 96, kv_valid, full, mx);
                    m_ref = mx * p.scale_log2;
                }
                tmem_ld32(s_addr, ra);       // first chunk is fetched while we wait for PV(j-1)
                if (j > 0) {
                    // PV(j-1) reads P (which we are about to overwrite) and updates O.  It was issued before S(j) and the
                    // tensor pipe executes in order, so the s_full(j) arrival above already implies PV(j-1) has completed:
                    // no separate wait on pv_done is needed inside the loop.
                    if (__any_sync(0xffffffffu, f_pending != 1.0f)) {
                        tmem_ld_wait();
                        scale_o_rows(o_addr, f_pending, rb);
                    }
                }
                f_pending = 1.0f;
                if (wq == 0) PE_TRACE(24 + q, j);
                // ---- single pass: P = exp2(S*scale - m_ref), row sum, row max of this tile ----
                float lsum = 0.f, mx = -INFINITY;
                uint32_t pk[64];
                tmem_ld_wait();
                tmem_ld32(s_addr + 32, rb);
                softmax_chunk(ra, 0, kv_valid, full, p.scale_log2, m_ref, lsum, mx, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
                tmem_ld_wait();
                tmem_ld32(s_addr + 64, ra);
                softmax_chunk(rb, 32, kv_valid, full, p.scale_log2, m_ref, lsum, mx, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
                tmem_ld_wait();
                tmem_ld32(s_addr + 96, rb);
                softmax_chunk(ra, 64, kv_valid, full, p.scale_log2, m_ref, lsum, mx, *reinterpret_cast<uint32_t(*)[16]>(&pk[32]));
                tmem_ld_wait();
                softmax_chunk(rb, 96, kv_valid, full, p.scale_log2, m_ref, lsum, mx, *reinterpret_cast<uint32_t(*)[16]>(&pk[48]));
                const float mx_scaled = mx * p.scale_log2;
                if (__any_sync(0xffffffffu, mx_scaled - m_ref > 100.0f)) {
                    // overflow guard (never taken for sane logits): move the reference to the true max and redo the tile.
                    // S is still intact in TMEM because P has not been stored yet.
                    const float m_new = fmaxf(m_ref, mx_scaled);
                    const float f = ex2(m_ref - m_new);
                    if (j > 0) scale_o_rows(o_addr, f, rb);
                    l *= f;
                    m_ref = m_new;
                    lsum = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        tmem_ld32(s_addr + c * 32, ra);
                        tmem_ld_wait();
                        softmax_chunk(ra, c * 32, kv_valid, full, p.scale_log2, m_ref, lsum, mx, *reinterpret_cast<uint32_t(*)[16]>(&pk[c * 16]));
                    }
                }
                if (wq == 0) PE_TRACE(26 + q, j);
                const uint32_t p_row = kPTmem ? s_addr : p_smem(q) + row_in_tile * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c) store_p_chunk<kPTmem>(*reinterpret_cast<uint32_t(*)[16]>(&pk[c * 16]), c, p_row, row_in_tile);
                if (kPTmem) tmem_st_wait(); else fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(q));
                if (wq == 0) PE_TRACE(28 + q, j);
                l += lsum;
                // lazy reference update for the following steps (O is rescaled after PV(j) has finished)
                if (mx_scaled > m_ref + 8.0f) {
                    f_pending = ex2(m_ref - mx_scaled);
                    l *= f_pending;
                    m_ref = mx_scaled;
                }
            }
            if (!ok) break;

            // ---- epilogue: O / l -> bf16 -> global ----
            if (!mbar_wait(pv_done(q), item_par, p.abort_flag, 32)) break;     // committed once per item, after its last PV
            item_par ^= 1u;
            tc_fence_after();
            const float inv = f_pending / l;     // includes the O rescale still pending from the last step
            const long long row = (long long)(qb * kQT + q) * kTile + row_in_tile;
            if (p.lse != nullptr && row < p.S) p.lse[(long long)head * p.S + row] = m_ref + log2f(l);     // l is relative to the current m_ref
            bf16* obase = p.o;
            if (p.n_route > 0) {                      // sequence-parallel mode: the row's owner rank gets it (possibly over NVLink)
                obase = p.route_o[p.n_route - 1];
#pragma unroll 1
                for (int i = p.n_route - 2; i >= 0; --i)
                    if (row < p.route_end[i]) obase = p.route_o[i];
            }
            bf16* orow = obase + row * p.ldo + head * kTile;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld32(o_addr + c * 32, r);
                tmem_ld_wait();
                if (row < p.S) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(r[8 * v]) * inv, __uint_as_float(r[8 * v + 1]) * inv);
                        o.y = pack_bf16(__uint_as_float(r[8 * v + 2]) * inv, __uint_as_float(r[8 * v + 3]) * inv);
                        o.z = pack_bf16(__uint_as_float(r[8 * v + 4]) * inv, __uint_as_float(r[8 * v + 5]) * inv);
                        o.w = pack_bf16(__uint_as_float(r[8 * v + 6]) * inv, __uint_as_float(r[8 * v + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + c * 32 + v * 8) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty(q));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, Cfg::kTmemCols);
    }
}

template <int kQT, bool kPTmem>
int launch_attention(Handle* h, AttnParams& p, cudaStream_t stream) {
    using Cfg = AttnCfg<kQT, kPTmem>;
    auto kern = attention_kernel<kQT, kPTmem>;
    static bool configured = false;
    if (!configured) {
        PE_CHECK_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
        configured = true;
    }
    p.n_qblk = ceil_div(p.S, kTile * kQT);
    p.n_items = p.H * p.n_qblk;
    int ctas = h->sm_count;
    if (ctas > p.n_items) ctas = p.n_items;
    kern<<<ctas, Cfg::kThreads, Cfg::kSmem, stream>>>(p);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

// -------------------------------------------------------------------------------------------------
// attention_kernel2: same tensor-side pipeline (two ping-ponged 128-row query tiles, P in TMEM, 32 KB K/V ring), but the
// softmax of each tile is spread over TWO warpgroups: a thread owns one query row x 64 kv columns and keeps that half row of
// S in registers (one TMEM read pass).  Two softmax warps per SMSP per tile give the thread-level parallelism a single warp
// lacks (r1 timeline: one warp per SMSP needed ~2150 cycles per 128x128 tile, twice the MUFU floor).  The exact row max is
// formed every step: the two half-row owners exchange their maxima through shared memory around a 64-thread named barrier.
//   warps 4..11  tile 0 (4..7: kv columns 0-63, 8..11: columns 64-127), warps 12..19 tile 1
// P half h overlays S columns [64h, 64h+32) of its own half, so a thread only overwrites S values it has already consumed.
// -------------------------------------------------------------------------------------------------
constexpr int kA2KV = 4;                                    // K/V ring depth
constexpr int kA2Threads = 128 + 512;
constexpr int kA2SmemData = 2 * kTileBytes + kA2KV * kTileBytes;
constexpr int kA2Exch = 2 * 2 * 2 * 128 * 4 + 2 * 2 * 128 * 4;   // row-max exchange (2 slots) + row-sum exchange
constexpr int kA2Smem = 1024 + kA2SmemData + kA2Exch + 256;

__device__ __forceinline__ void tmem_ld16_(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8_(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8_(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(kA2Threads, 1) attention_kernel2(const __grid_constant__ AttnParams p) {
    constexpr int kKV = kA2KV;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    auto q_smem = [&](int q) { return smem_base + q * kTileBytes; };
    auto kv_smem = [&](int s) { return smem_base + (2 + s) * kTileBytes; };
    const uint32_t exch_base = smem_base + kA2SmemData;                 // [slot][q][half][128] f32 row max, then [q][half][128] f32 row sum
    const uint32_t lsum_base = exch_base + 2 * 2 * 2 * 128 * 4;
    const uint32_t bar_base = exch_base + kA2Exch;
    const uint32_t q_full = bar_base, q_empty = bar_base + 8;
    auto kv_full = [&](int s) { return bar_base + 16 + s * 8; };
    auto kv_empty = [&](int s) { return bar_base + 16 + (kKV + s) * 8; };
    auto s_full = [&](int q) { return bar_base + 16 + (2 * kKV + q) * 8; };
    auto p_full = [&](int q) { return bar_base + 16 + (2 * kKV + 2 + q) * 8; };
    auto pv_done = [&](int q) { return bar_base + 16 + (2 * kKV + 4 + q) * 8; };
    auto o_empty = [&](int q) { return bar_base + 16 + (2 * kKV + 6 + q) * 8; };
    const uint32_t tmem_slot = bar_base + 16 + (2 * kKV + 8) * 8;

    const int warp = threadIdx.x >> 5;
    const int lane = lane_id();

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&p.tmQ);
        prefetch_tmap(&p.tmK);
        prefetch_tmap(&p.tmV);
    }
    if (warp == 1 && elect_one()) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < kKV; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
        for (int q = 0; q < 2; ++q) {
            mbar_init(s_full(q), 1);
            mbar_init(p_full(q), 8);       // 8 softmax warps per tile
            mbar_init(pv_done(q), 1);
            mbar_init(o_empty(q), 8);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ld_shared_u32(tmem_slot);
    auto s_tmem = [&](int q) { return tmem_base + q * 128; };
    auto o_tmem = [&](int q) { return tmem_base + 256 + q * 128; };

    if (warp == 0) {
        // ======================================= TMA producer =======================================
        uint32_t n = 0, it = 0;
        bool ok = true;
        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x, ++it) {
            const int head = item / p.n_qblk;
            const int qb = item - head * p.n_qblk;
            const int col0 = head * kTile;
            if (!mbar_wait(q_empty, (it & 1u) ^ 1u, p.abort_flag, 40)) break;
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int row0 = (qb * 2 + q) * kTile;
                    tma_load_2d(q_smem(q), &p.tmQ, q_full, col0, row0);
                    tma_load_2d(q_smem(q) + kHalfBytes, &p.tmQ, q_full, col0 + 64, row0);
                }
            }
            __syncwarp();
            for (int j = 0; j < p.n_kv && ok; ++j) {
#pragma unroll
                for (int kv = 0; kv < 2; ++kv, ++n) {
                    const int slot = n % kKV;
                    const uint32_t ph = (n / kKV) & 1u;
                    if (!mbar_wait(kv_empty(slot), ph ^ 1u, p.abort_flag, 41)) { ok = false; break; }
                    if (elect_one()) {
                        const CUtensorMap* tm = kv == 0 ? &p.tmK : &p.tmV;
                        mbar_arrive_expect_tx(kv_full(slot), kTileBytes);
                        tma_load_2d(kv_smem(slot), tm, kv_full(slot), col0, j * kTile);
                        tma_load_2d(kv_smem(slot) + kHalfBytes, tm, kv_full(slot), col0 + 64, j * kTile);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ======================================= MMA issuer =======================================
        // (A variant that ran this whole loop on one lane of a diverged warp was measured 35 % slower per batch:
        //  the uniform datapath that feeds UTCHMMA needs the converged warp.)
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_o = make_idesc_bf16(128, 128, 0, 1);
        uint32_t n = 0, it = 0;
        uint32_t p_phase[2] = {0, 0};
        bool ok = true;
        PE_TRACE_DECL(0)
        auto issue_s = [&](int q, uint32_t k_base) {
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t off = (kk >> 2) * kHalfBytes + (kk & 3) * 32;
                    umma_bf16<1>(s_tmem(q), make_smem_desc_sw128(q_smem(q) + off, 16, 1024),
                                 make_smem_desc_sw128(k_base + off, 16, 1024), idesc_s, kk != 0 ? 1u : 0u);
                }
                umma_commit(s_full(q));
            }
            __syncwarp();
        };
        auto issue_pv = [&](int q, uint32_t v_base, bool accumulate, bool last) {
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint64_t bdesc = make_smem_desc_sw128(v_base + kk * 2048, kHalfBytes, 1024);
                    // P: kv columns 0-63 are packed in TMEM columns [0,32), kv columns 64-127 in TMEM columns [64,96)
                    const uint32_t a_tmem = s_tmem(q) + (kk >> 2) * 64 + (kk & 3) * 8;
                    umma_bf16_ts(o_tmem(q), a_tmem, bdesc, idesc_o, (accumulate || kk != 0) ? 1u : 0u);
                }
                if (last) umma_commit(pv_done(q));       // once per item (see attention_kernel)
            }
            __syncwarp();
        };
        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x, ++it) {
            if (!mbar_wait(q_full, it & 1u, p.abort_flag, 50)) break;
            uint32_t k_slot = n % kKV;
            if (!mbar_wait(kv_full(k_slot), (n / kKV) & 1u, p.abort_flag, 51)) break;
            ++n;
            tc_fence_after();
#pragma unroll
            for (int q = 0; q < 2; ++q) issue_s(q, kv_smem(k_slot));
            if (elect_one()) {
                umma_commit(kv_empty(k_slot));
                if (p.n_kv == 1) umma_commit(q_empty);
            }
            __syncwarp();
            for (int j = 0; j < p.n_kv && ok; ++j) {
                const uint32_t v_slot = n % kKV;
                if (!mbar_wait(kv_full(v_slot), (n / kKV) & 1u, p.abort_flag, 52)) { ok = false; break; }
                ++n;
                const bool more = j + 1 < p.n_kv;
                if (more) {
                    k_slot = n % kKV;
                    if (!mbar_wait(kv_full(k_slot), (n / kKV) & 1u, p.abort_flag, 53)) { ok = false; break; }
                    ++n;
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (j == 0) {
                        if (!mbar_wait(o_empty(q), (it & 1u) ^ 1u, p.abort_flag, 54)) { ok = false; break; }
                    }
                    PE_TRACE(10 + q, j);
                    if (!mbar_wait(p_full(q), p_phase[q], p.abort_flag, 55)) { ok = false; break; }
                    p_phase[q] ^= 1u;
                    tc_fence_after();
                    PE_TRACE(12 + q, j);
                    issue_pv(q, kv_smem(v_slot), j > 0, !more);
                    if (more) issue_s(q, kv_smem(k_slot));
                    PE_TRACE(14 + q, j);
                }
                if (!ok) break;
                if (elect_one()) {
                    umma_commit(kv_empty(v_slot));
                    if (more) umma_commit(kv_empty(k_slot));
                    if (j + 2 == p.n_kv) umma_commit(q_empty);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ======================================= softmax / correction / epilogue =======================================
        const int sw = warp - 4;
        const int q = sw >> 3;                 // query tile
        const int half = (sw >> 2) & 1;        // which 64 kv columns of every S tile (and which 64 columns of O)
        const int wq = sw & 3;                 // == warp % 4: TMEM lane quarter
        const int row_in_tile = wq * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
        const uint32_t s_addr = s_tmem(q) + lane_off + half * 64;    // own S columns; P overlays the first 32 of them
        const uint32_t o_addr = o_tmem(q) + lane_off + half * 64;
        const int pair_bar = 1 + q * 4 + wq;                           // named barrier shared with the warp owning the other half
        auto mx_slot = [&](int slot, int hf) { return exch_base + (((slot * 2 + q) * 2 + hf) * 128 + row_in_tile) * 4; };
        const uint32_t l_own = lsum_base + ((q * 2 + half) * 128 + row_in_tile) * 4;
        const uint32_t l_other = lsum_base + ((q * 2 + (half ^ 1)) * 128 + row_in_tile) * 4;
        uint32_t s_phase = 0, item_par = 0;
        bool ok = true;
        const bool tr = (wq == 0 && half == 0);
        PE_TRACE_DECL(1 + q)
        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x) {
            const int head = item / p.n_qblk;
            const int qb = item - head * p.n_qblk;
            float m_ref = -INFINITY;          // exponent reference (log2 domain) of O and l; trails the running max by < 2^8
            float l = 0.f;                    // row sum over this thread's kv columns only
            for (int j = 0; j < p.n_kv; ++j) {
                if (tr) PE_TRACE(20 + q, j);
                if (!mbar_wait(s_full(q), s_phase, p.abort_flag, 60)) { ok = false; break; }
                s_phase ^= 1u;
                tc_fence_after();
                if (tr) PE_TRACE(22 + q, j);
                const int kv_valid = p.S - j * kTile - half * 64;      // valid columns among this thread's 64
                uint32_t r[64];
                if (PE_ATTN_DBG & 8) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) r[i] = __float_as_uint((float)((i * 7 + lane + j) & 15));
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_ld16_(s_addr + c * 16, &r[c * 16]);
                    tmem_ld_wait();
                }
                if (tr) PE_TRACE(30 + q, j);
                // ---- exact row max: own 64 columns, then exchange with the owner of the other half ----
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                if (kv_valid >= 64) {
#pragma unroll
                    for (int i = 0; i < 64; i += 8) {
                        m0 = fmaxf(m0, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
                        m1 = fmaxf(m1, fmaxf(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
                        m2 = fmaxf(m2, fmaxf(__uint_as_float(r[i + 4]), __uint_as_float(r[i + 5])));
                        m3 = fmaxf(m3, fmaxf(__uint_as_float(r[i + 6]), __uint_as_float(r[i + 7])));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        if (i < kv_valid) m0 = fmaxf(m0, __uint_as_float(r[i]));
                }
                const float mx_own = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                st_shared_f32(mx_slot(j & 1, half), mx_own);
                if (!(PE_ATTN_DBG & 4)) named_bar_sync(pair_bar, 64);
                const float mx_scaled = fmaxf(mx_own, ld_shared_f32(mx_slot(j & 1, half ^ 1))) * p.scale_log2;
                if (tr) PE_TRACE(24 + q, j);
                // ---- lazy reference update: both owners of a row take the same decision from the same numbers ----
                float f = 1.0f;
                if (mx_scaled > m_ref + 8.0f) {
                    f = ex2(m_ref - mx_scaled);       // 0 on the first tile (m_ref = -inf)
                    m_ref = mx_scaled;
                    l *= f;
                }
                // O rescale (own 64 columns).  PV(j-1) was issued before S(j) on the in-order tensor pipe, so the s_full(j)
                // arrival implies it has completed.  On the first step O is uninitialised and PV(0) overwrites it.
                if (j > 0 && __any_sync(0xffffffffu, f != 1.0f)) {
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        uint32_t t[8];
                        tmem_ld8_(o_addr + c * 8, t);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * f);
                        tmem_st8_(o_addr + c * 8, t);
                    }
                    tmem_st_wait();
                }
                // ---- P = exp2(S*scale - m_ref) from registers, row sum, P (bf16) over the consumed S columns ----
                float s0 = 0.f, s1 = 0.f;
                if (kv_valid >= 64) {
                    const uint64_t scale2 = pk2(p.scale_log2, p.scale_log2), negm2 = pk2(-m_ref, -m_ref);
                    uint64_t sum2 = pk2(0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int e = c * 16 + 2 * i;
                            const uint64_t x2 = ffma2(pk2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), scale2, negm2);
                            float p0, p1;
                            // pairs 1, 3, 6 (then 4, 0, ...) of every 8 go to the FMA pipe: spread so MUFU and FMA work interleave
                            constexpr int kOrder[8] = {1, 3, 6, 4, 0, 7, 2, 5};
                            bool poly = false;
#pragma unroll
                            for (int t = 0; t < PE_ATTN_POLY_PAIRS; ++t) poly = poly || (kOrder[t] == i);
                            if (poly) {
                                exp2_fma_pair(x2, p0, p1);
                            } else {
                                float x0, x1;
                                upk2(x2, x0, x1);
                                if (PE_ATTN_DBG & 1) { p0 = x0; p1 = x1; } else {
                                p0 = ex2(x0);
                                p1 = ex2(x1); }
                            }
                            sum2 = fadd2(sum2, pk2(p0, p1));
                            pk[i] = pack_bf16(p0, p1);
                        }
                        if (!(PE_ATTN_DBG & 2)) tmem_st8_(s_addr + c * 8, pk); else if (pk[0] == 0x12345u) st_shared_f32(l_own, 1.f);
                    }
                    upk2(sum2, s0, s1);
                } else {
                    // ragged last KV tile: columns beyond the sequence contribute nothing
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int e = c * 16 + 2 * i;
                            float p0 = ex2(fmaf(__uint_as_float(r[e]), p.scale_log2, -m_ref));
                            float p1 = ex2(fmaf(__uint_as_float(r[e + 1]), p.scale_log2, -m_ref));
                            if (e >= kv_valid) p0 = 0.f;
                            if (e + 1 >= kv_valid) p1 = 0.f;
                            s0 += p0;
                            s1 += p1;
                            pk[i] = pack_bf16(p0, p1);
                        }
                        tmem_st8_(s_addr + c * 8, pk);
                    }
                }
                l += s0 + s1;
                if (tr) PE_TRACE(26 + q, j);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(q));
                if (tr) PE_TRACE(28 + q, j);
            }
            if (!ok) break;
            // ---- epilogue: O / l -> bf16 -> global (own 64 columns of the head) ----
            if (!mbar_wait(pv_done(q), item_par, p.abort_flag, 62)) break;
            item_par ^= 1u;
            tc_fence_after();
            st_shared_f32(l_own, l);
            named_bar_sync(pair_bar, 64);
            const float inv = 1.0f / (l + ld_shared_f32(l_other));
            const long long row = (long long)(qb * 2 + q) * kTile + row_in_tile;
            bf16* orow = p.o + row * p.ldo + head * kTile + half * 64;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t t[16];
                tmem_ld16_(o_addr + c * 16, t);
                tmem_ld_wait();
                if (row < p.S) {
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(t[8 * v]) * inv, __uint_as_float(t[8 * v + 1]) * inv);
                        o.y = pack_bf16(__uint_as_float(t[8 * v + 2]) * inv, __uint_as_float(t[8 * v + 3]) * inv);
                        o.z = pack_bf16(__uint_as_float(t[8 * v + 4]) * inv, __uint_as_float(t[8 * v + 5]) * inv);
                        o.w = pack_bf16(__uint_as_float(t[8 * v + 6]) * inv, __uint_as_float(t[8 * v + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + c * 16 + v * 8) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty(q));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

int launch_attention2(Handle* h, AttnParams& p, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        PE_CHECK_CUDA(h, cudaFuncSetAttribute(attention_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, kA2Smem));
        configured = true;
    }
    p.n_qblk = ceil_div(p.S, kTile * 2);
    p.n_items = p.H * p.n_qblk;
    int ctas = h->sm_count;
    if (ctas > p.n_items) ctas = p.n_items;
    attention_kernel2<<<ctas, kA2Threads, kA2Smem, stream>>>(p);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

// -------------------------------------------------------------------------------------------------
// attention_kernel4 (PE_ATTN_FLAG_HALF_ROW): attention_kernel2's layout -- two warps per SM sub-partition per query tile, a thread owns
// one query row x 64 kv columns held in registers -- WITHOUT its per-step exchange: the exponent reference trails the row max by one
// KV step (see the comment in the softmax section), so the two owners of a row never wait for each other inside the loop, and
// PE_A4_POLY_PAIRS of every 8 element pairs take their exp2 on the FMA pipe (packed f32x2 Cody-Waite + cubic) to relieve MUFU.EX2.
// r2 trace of kernel 1 (tools/attn_trace.py): the per-tile chain is  softmax 2190  ->  99  ->  PV+S MMAs 1025  ->  364  cycles, and
// a single warp per SMSP runs its pass at 26 cycles per element pair against the 16-cycle MUFU bound; two warps per SMSP reach it.
// -------------------------------------------------------------------------------------------------
#ifndef PE_A4_POLY_PAIRS
#define PE_A4_POLY_PAIRS 3
#endif
__global__ void __launch_bounds__(kA2Threads, 1) attention_kernel4(const __grid_constant__ AttnParams p) {
    constexpr int kKV = kA2KV;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    auto q_smem = [&](int q) { return smem_base + q * kTileBytes; };
    auto kv_smem = [&](int s) { return smem_base + (2 + s) * kTileBytes; };
    const uint32_t exch_base = smem_base + kA2SmemData;                 // [slot][q][half][128] f32 row max, then [q][half][128] f32 row sum
    const uint32_t lsum_base = exch_base + 2 * 2 * 2 * 128 * 4;
    const uint32_t bar_base = exch_base + kA2Exch;
    const uint32_t q_full = bar_base, q_empty = bar_base + 8;
    auto kv_full = [&](int s) { return bar_base + 16 + s * 8; };
    auto kv_empty = [&](int s) { return bar_base + 16 + (kKV + s) * 8; };
    auto s_full = [&](int q) { return bar_base + 16 + (2 * kKV + q) * 8; };
    auto p_full = [&](int q) { return bar_base + 16 + (2 * kKV + 2 + q) * 8; };
    auto pv_done = [&](int q) { return bar_base + 16 + (2 * kKV + 4 + q) * 8; };
    auto o_empty = [&](int q) { return bar_base + 16 + (2 * kKV + 6 + q) * 8; };
    const uint32_t tmem_slot = bar_base + 16 + (2 * kKV + 8) * 8;

    // softmax warps first, role warps last: the MMA issuer must be the highest warp id of its SM sub-partition (see attention_kernel)
    const int warp_raw = threadIdx.x >> 5;
    const int warp = PE_ATTN_ROLES_LAST ? (warp_raw >= 16 ? warp_raw - 16 : warp_raw + 4) : warp_raw;
    const int lane = lane_id();

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&p.tmQ);
        prefetch_tmap(&p.tmK);
        prefetch_tmap(&p.tmV);
    }
    if (warp == 1 && elect_one()) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < kKV; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
        for (int q = 0; q < 2; ++q) {
            mbar_init(s_full(q), 1);
            mbar_init(p_full(q), 8);       // 8 softmax warps per tile
            mbar_init(pv_done(q), 1);
            mbar_init(o_empty(q), 8);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ld_shared_u32(tmem_slot);
    auto s_tmem = [&](int q) { return tmem_base + q * 128; };
    auto o_tmem = [&](int q) { return tmem_base + 256 + q * 128; };

    if (warp == 0) {
        // ======================================= TMA producer =======================================
        uint32_t n = 0, it = 0;
        bool ok = true;
        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x, ++it) {
            const int head = item / p.n_qblk;
            const int qb = item - head * p.n_qblk;
            const int col0 = head * kTile;
            if (!mbar_wait(q_empty, (it & 1u) ^ 1u, p.abort_flag, 40)) break;
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int row0 = (qb * 2 + q) * kTile;
                    tma_load_2d(q_smem(q), &p.tmQ, q_full, col0, row0);
                    tma_load_2d(q_smem(q) + kHalfBytes, &p.tmQ, q_full, col0 + 64, row0);
                }
            }
            __syncwarp();
            for (int j = 0; j < p.n_kv && ok; ++j) {
#pragma unroll
                for (int kv = 0; kv < 2; ++kv, ++n) {
                    const int slot = n % kKV;
                    const uint32_t ph = (n / kKV) & 1u;
                    if (!mbar_wait(kv_empty(slot), ph ^ 1u, p.abort_flag, 41)) { ok = false; break; }
                    if (elect_one()) {
                        const CUtensorMap* tm = kv == 0 ? &p.tmK : &p.tmV;
                        mbar_arrive_expect_tx(kv_full(slot), kTileBytes);
                        tma_load_2d(kv_smem(slot), tm, kv_full(slot), col0, j * kTile);
                        tma_load_2d(kv_smem(slot) + kHalfBytes, tm, kv_full(slot), col0 + 64, j * kTile);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ======================================= MMA issuer =======================================
        // In this kernel the softmax link of the per-tile chain is short enough that THIS warp's own serial time line decides the
        // period (r2 trace: ~740 cycles between the end of tile 1's batch and the start of the wait for P0 -- two commits and two
        // kv_full polls of 200-280 cycles each, although those barriers completed long before; SYNCS round trips share the SMSP's MIO
        // queue with the MUFU-heavy softmax warps).  So: the leader is elected once, descriptors are base word + constant (as in
        // kernel 1), and every barrier the NEXT batch needs is probed with a non-blocking test_wait issued BEFORE the current batch's
        // 16 blocking UTCHMMA issues -- the probe's round trip overlaps the ~1000 cycles of issue, and the blocking wait is only taken
        // when the probe said "not yet".
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_o = make_idesc_bf16(128, 128, 0, 1);
        const bool leader = elect_one();
        uint32_t n = 0, it = 0;
        uint32_t p_phase[2] = {0, 0};
        bool ok = true;
        PE_TRACE_DECL(0)
        auto issue_s = [&](int q, uint32_t k_base) {
            const uint32_t qlo = desc_lo_k(q_smem(q)), klo = desc_lo_k(k_base);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint32_t off16 = ((kk >> 2) * kHalfBytes + (kk & 3) * 32) >> 4;
                if (kk == 0) umma_ss_lohi<false>(s_tmem(q), qlo + off16, klo + off16, idesc_s);
                else umma_ss_lohi<true>(s_tmem(q), qlo + off16, klo + off16, idesc_s);
            }
            umma_commit(s_full(q));
        };
        auto issue_pv = [&](int q, uint32_t v_base, bool accumulate, bool last) {
            const uint32_t vlo = desc_lo_v(v_base);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                // P: kv columns 0-63 are packed in TMEM columns [0,32), kv columns 64-127 in TMEM columns [64,96)
                const uint32_t a_tmem = s_tmem(q) + (kk >> 2) * 64 + (kk & 3) * 8;
                umma_ts_lohi(o_tmem(q), a_tmem, vlo + kk * (2048 >> 4), idesc_o, (accumulate || kk != 0) ? 1u : 0u);
            }
            if (last) umma_commit(pv_done(q));       // once per item (see attention_kernel)
        };
        auto probe = [&](uint32_t bar, uint32_t parity) { return __all_sync(0xffffffffu, mbar_test_wait(bar, parity)); };
        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x, ++it) {
            if (!mbar_wait(q_full, it & 1u, p.abort_flag, 50)) break;
            uint32_t k_slot = n % kKV;
            if (!mbar_wait(kv_full(k_slot), (n / kKV) & 1u, p.abort_flag, 51)) break;
            ++n;
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int q = 0; q < 2; ++q) issue_s(q, kv_smem(k_slot));
                umma_commit(kv_empty(k_slot));
                if (p.n_kv == 1) umma_commit(q_empty);
            }
            __syncwarp();
            bool v_ready = false, k_ready = false, p_ready[2] = {false, false};
            for (int j = 0; j < p.n_kv && ok; ++j) {
                const uint32_t v_slot = n % kKV;
                if (!v_ready && !mbar_wait(kv_full(v_slot), (n / kKV) & 1u, p.abort_flag, 52)) { ok = false; break; }
                ++n;
                const bool more = j + 1 < p.n_kv;
                if (more) {
                    k_slot = n % kKV;
                    if (!k_ready && !mbar_wait(kv_full(k_slot), (n / kKV) & 1u, p.abort_flag, 53)) { ok = false; break; }
                    ++n;
                }
                v_ready = k_ready = false;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (j == 0) {
                        if (!mbar_wait(o_empty(q), (it & 1u) ^ 1u, p.abort_flag, 54)) { ok = false; break; }
                    }
                    PE_TRACE(10 + q, j);
                    if (!p_ready[q] && !mbar_wait(p_full(q), p_phase[q], p.abort_flag, 55)) { ok = false; break; }
                    p_ready[q] = false;
                    p_phase[q] ^= 1u;
                    tc_fence_after();
                    PE_TRACE(12 + q, j);
                    // probes for what follows this batch, issued before its MMAs
                    if (q == 0) {
                        p_ready[1] = probe(p_full(1), p_phase[1]);
                    } else if (more) {
                        v_ready = probe(kv_full(n % kKV), (n / kKV) & 1u);
                        if (j + 2 < p.n_kv) k_ready = probe(kv_full((n + 1) % kKV), ((n + 1) / kKV) & 1u);
                        p_ready[0] = probe(p_full(0), p_phase[0]);
                    }
                    if (leader) {
                        issue_pv(q, kv_smem(v_slot), j > 0, !more);
                        if (more) issue_s(q, kv_smem(k_slot));
                        if (q == 1) {
                            umma_commit(kv_empty(v_slot));
                            if (more) umma_commit(kv_empty(k_slot));
                            if (j + 2 == p.n_kv) umma_commit(q_empty);
                        }
                    }
                    __syncwarp();
                    PE_TRACE(14 + q, j);
                }
                if (!ok) break;
            }
        }
    } else if (warp >= 4) {
        // ======================================= softmax / correction / epilogue =======================================
        const int sw = warp - 4;
        const int q = sw >> 3;                 // query tile
        const int half = (sw >> 2) & 1;        // which 64 kv columns of every S tile (and which 64 columns of O)
        const int wq = sw & 3;                 // == warp % 4: TMEM lane quarter
        const int row_in_tile = wq * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
        const uint32_t s_addr = s_tmem(q) + lane_off + half * 64;    // own S columns; P overlays the first 32 of them
        const uint32_t o_addr = o_tmem(q) + lane_off + half * 64;
        const int pair_bar = 1 + q * 4 + wq;                           // named barrier shared with the warp owning the other half
        auto mx_slot = [&](int slot, int hf) { return exch_base + (((slot * 2 + q) * 2 + hf) * 128 + row_in_tile) * 4; };
        const uint32_t l_own = lsum_base + ((q * 2 + half) * 128 + row_in_tile) * 4;
        const uint32_t l_other = lsum_base + ((q * 2 + (half ^ 1)) * 128 + row_in_tile) * 4;
        uint32_t s_phase = 0, item_par = 0;
        bool ok = true;
        const bool tr = (wq == 0 && half == 0);
        PE_TRACE_DECL(1 + q)
        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x) {
            const int head = item / p.n_qblk;
            const int qb = item - head * p.n_qblk;
            // Exponent reference (log2 domain) of everything accumulated in O and l.  It TRAILS the running row max by one KV step:
            // step j exponentiates against the max of steps < j, which both owners of a row know without waiting for each other --
            // each leaves its half-row max of step j in shared memory before arriving on p_full(j) and reads the partner's after
            // s_full(j+1), an arrival that transitively follows the partner's p_full(j) arrival.  So the only per-step
            // synchronisation of the two half-row owners is the pair of mbarriers they use anyway.  A stale (lower) reference
            // scales P, O and l by one common factor that the final O / l cancels; it is moved (with an O rescale) when the max
            // grew by more than 2^8.  A step whose logits jump more than 2^100 above the reference could overflow: it raises
            // the item's flag and the item is redone by the exact kernel (launch_attention4).
            float m_ref = -INFINITY;
            float l = 0.f;                    // row sum over this thread's kv columns only
            float mx_prev = -INFINITY;        // own half-row max (raw logits) of the previous step
            bool ovf = false;
            for (int j = 0; j < p.n_kv; ++j) {
                if (tr) PE_TRACE(20 + q, j);
                if (!mbar_wait(s_full(q), s_phase, p.abort_flag, 60)) { ok = false; break; }
                s_phase ^= 1u;
                tc_fence_after();
                if (tr) PE_TRACE(22 + q, j);
                const int kv_valid = p.S - j * kTile - half * 64;      // valid columns among this thread's 64
                uint32_t r[64];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld16_(s_addr + c * 16, &r[c * 16]);
                if (j == 0) {
                    // first tile of the item: exact row max, exchanged once through the pair's named barrier
                    tmem_ld_wait();
                    float m0 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        if (i < kv_valid) m0 = fmaxf(m0, __uint_as_float(r[i]));
                    st_shared_f32(mx_slot(1, half), m0);
                    named_bar_sync(pair_bar, 64);
                    m_ref = fmaxf(m0, ld_shared_f32(mx_slot(1, half ^ 1))) * p.scale_log2;
                } else {
                    const float mxs = fmaxf(mx_prev, ld_shared_f32(mx_slot((j - 1) & 1, half ^ 1))) * p.scale_log2;
                    float f = 1.0f;
                    if (mxs > m_ref + 8.0f) {
                        f = ex2(m_ref - mxs);
                        m_ref = mxs;
                        l *= f;
                    }
                    // O rescale (own 64 columns).  PV(j-1) was issued before S(j) on the in-order tensor pipe, so the s_full(j)
                    // arrival implies it has completed.
                    if (__any_sync(0xffffffffu, f != 1.0f)) {
#pragma unroll 1
                        for (int c = 0; c < 8; ++c) {
                            uint32_t t[8];
                            tmem_ld8_(o_addr + c * 8, t);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 8; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * f);
                            tmem_st8_(o_addr + c * 8, t);
                        }
                        tmem_st_wait();
                    }
                    tmem_ld_wait();
                }
                if (tr) PE_TRACE(24 + q, j);
                // ---- P = exp2(S*scale - m_ref) from registers, half-row sum and max, P (bf16) over the consumed S columns ----
                float s0 = 0.f, s1 = 0.f, mx_own = -INFINITY;
                if (kv_valid >= 64) {
                    const uint64_t scale2 = pk2(p.scale_log2, p.scale_log2), negm2 = pk2(-m_ref, -m_ref);
                    uint64_t sum2 = pk2(0.f, 0.f);
                    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int e = c * 16 + 2 * i;
                            const float a0 = __uint_as_float(r[e]), a1 = __uint_as_float(r[e + 1]);
                            if (i & 1) m1 = fmaxf(m1, fmaxf(a0, a1)); else m0 = fmaxf(m0, fmaxf(a0, a1));
                            const uint64_t x2 = ffma2(pk2(a0, a1), scale2, negm2);
                            float p0, p1;
                            // PE_A4_POLY_PAIRS of every 8 pairs take the FMA-pipe exp2 (Cody-Waite + cubic, rel. error 1e-4 << bf16's
                            // 2^-9) instead of MUFU.EX2 (16 / clk / SM, the co-bottleneck): spread so that MUFU and FMA work interleave
                            constexpr int kOrder[8] = {1, 4, 6, 3, 0, 7, 2, 5};
                            bool poly = false;
#pragma unroll
                            for (int t = 0; t < PE_A4_POLY_PAIRS; ++t) poly = poly || (kOrder[t] == i);
                            if (poly) {
                                exp2_fma_pair(x2, p0, p1);
                            } else {
                                float x0, x1;
                                upk2(x2, x0, x1);
                                p0 = ex2(x0);
                                p1 = ex2(x1);
                            }
                            sum2 = fadd2(sum2, pk2(p0, p1));
                            pk[i] = pack_bf16(p0, p1);
                        }
                        tmem_st8_(s_addr + c * 8, pk);
                    }
                    upk2(sum2, s0, s1);
                    mx_own = fmaxf(m0, m1);
                } else {
                    // ragged last KV tile: columns beyond the sequence contribute nothing
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int e = c * 16 + 2 * i;
                            float p0 = ex2(fmaf(__uint_as_float(r[e]), p.scale_log2, -m_ref));
                            float p1 = ex2(fmaf(__uint_as_float(r[e + 1]), p.scale_log2, -m_ref));
                            if (e >= kv_valid) p0 = 0.f; else mx_own = fmaxf(mx_own, __uint_as_float(r[e]));
                            if (e + 1 >= kv_valid) p1 = 0.f; else mx_own = fmaxf(mx_own, __uint_as_float(r[e + 1]));
                            s0 += p0;
                            s1 += p1;
                            pk[i] = pack_bf16(p0, p1);
                        }
                        tmem_st8_(s_addr + c * 8, pk);
                    }
                }
                l += s0 + s1;
                ovf = ovf || (mx_own * p.scale_log2 - m_ref > 100.0f);
                mx_prev = mx_own;
                st_shared_f32(mx_slot(j & 1, half), mx_own);
                if (tr) PE_TRACE(26 + q, j);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(q));
                if (tr) PE_TRACE(28 + q, j);
            }
            if (!ok) break;
            // ---- epilogue: O / l -> bf16 -> global (own 64 columns of the head) ----
            if (!mbar_wait(pv_done(q), item_par, p.abort_flag, 62)) break;
            item_par ^= 1u;
            tc_fence_after();
            st_shared_f32(l_own, l);
            named_bar_sync(pair_bar, 64);
            const float inv = 1.0f / (l + ld_shared_f32(l_other));
            if (__any_sync(0xffffffffu, ovf) && lane == 0) p.item_flags[item] = p.seq;      // redo this item exactly (benign race: same value)
            const long long row = (long long)(qb * 2 + q) * kTile + row_in_tile;
            bf16* orow = p.o + row * p.ldo + head * kTile + half * 64;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t t[16];
                tmem_ld16_(o_addr + c * 16, t);
                tmem_ld_wait();
                if (row < p.S) {
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(t[8 * v]) * inv, __uint_as_float(t[8 * v + 1]) * inv);
                        o.y = pack_bf16(__uint_as_float(t[8 * v + 2]) * inv, __uint_as_float(t[8 * v + 3]) * inv);
                        o.z = pack_bf16(__uint_as_float(t[8 * v + 4]) * inv, __uint_as_float(t[8 * v + 5]) * inv);
                        o.w = pack_bf16(__uint_as_float(t[8 * v + 6]) * inv, __uint_as_float(t[8 * v + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + c * 16 + v * 8) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty(q));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

int launch_attention4(Handle* h, AttnParams& p, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        PE_CHECK_CUDA(h, cudaFuncSetAttribute(attention_kernel4, cudaFuncAttributeMaxDynamicSharedMemorySize, kA2Smem));
        configured = true;
    }
    p.n_qblk = ceil_div(p.S, kTile * 2);
    p.n_items = p.H * p.n_qblk;
    int ctas = h->sm_count;
    if (ctas > p.n_items) ctas = p.n_items;
    // per-item overflow flags live in the handle's workspace (behind the first 256 KB used by trace builds): an item whose logits
    // jumped more than 2^100 above its trailing reference writes this launch's sequence number there and is redone by the exact
    // kernel (two-pass first tile, in-step redo) in a second launch that skips every other item -- normally all of them.
    // (8 flag regions used round-robin: launches that run concurrently on two streams -- the two CFG branches -- never share one)
    p.seq = ++h->attn_seq;
    if (p.seq == 0) p.seq = ++h->attn_seq;
    constexpr size_t kFlagRegion = 32 << 10;
    PE_REQUIRE(h, (size_t)p.n_items * sizeof(int) <= kFlagRegion && (256 << 10) + 8 * kFlagRegion <= h->workspace_bytes,
               "pe_attention_fwd: too many work items (%d) for the overflow-flag region", p.n_items);
    p.item_flags = reinterpret_cast<int*>(static_cast<char*>(h->workspace) + (256 << 10) + (static_cast<unsigned>(p.seq) & 7u) * kFlagRegion);
    attention_kernel4<<<ctas, kA2Threads, kA2Smem, stream>>>(p);
    PE_CHECK_CUDA(h, cudaGetLastError());
    p.only_flagged = 1;
    return launch_attention<2, true>(h, p, stream);
}

// -------------------------------------------------------------------------------------------------
// attention_kernel3 (PE_ATTN_FLAG_KV64): the softmax is taken OFF the tensor pipe's latency chain.
// In attention_kernel the chain of one query tile is  softmax(j) -> P hand-off -> PV(j) -> S(j+1) -> hand-off -> softmax(j+1):
// S(j+1) cannot be issued before PV(j) has consumed P(j), because P overlays S and TMEM (512 columns = S0,S1,O0,O1) has no
// room for a second S buffer.  Here a KV step is 64 rows, so an S tile is 128 x 64 fp32 = 64 columns and every query tile owns
// TWO S buffers (TMEM: S00,S01,S10,S11 = 4 x 64 columns, O0,O1 = 2 x 128).  S(j+2) is issued together with PV(j), i.e. one whole
// step ahead of the softmax, so the softmax warpgroups run back to back (they only wait when the tensor pipe is the slower side)
// and both tiles' softmax run concurrently (two busy warps per SMSP).  Tensor work per 64-row step and tile: PV = 4 MMAs
// 128x128x16 (P from TMEM), S = 8 MMAs 128x64x16.
//   warp 0      TMA producer: Q once per item; a ring of 32 KB stages, stage(j) = [V_j | K_{j+2}], prologue stage = [K_0 | K_1]
//   warp 1 / 3  MMA issuer of query tile 0 / 1 (r1 trace: one issuer warp's own instruction stream -- waits, commits, 24 MMAs per
//               step -- took 2400 cycles per step against 1280 cycles of tensor work; with one issuer per tile each has the whole
//               step for half of it.  TMEM hazards are per tile, so each tile's program order on its own warp is all that matters)
//   warp 2      TMEM allocator
//   warps 4-7 / 8-11  softmax + epilogue of query tile 0 / 1 (one thread per query row).
// Softmax arithmetic (trailing reference, lazy O rescale, overflow guard) is the one of attention_kernel; the O rescale now
// waits explicitly for PV(j-1), because the s_full(j) arrival no longer implies it: S(j+1) was issued right after PV(j-1), so its
// s_full arrival does (non-consuming early wait); on the last step a pv_done commit (made for the last two steps only) does.
// -------------------------------------------------------------------------------------------------
#ifndef PE_A3_DBG
#define PE_A3_DBG 0      // timing experiments only (wrong results): 1 = no S MMAs, 2 = no PV MMAs
#endif
#ifndef PE_A3_SKEW
#define PE_A3_SKEW 0     // cycles by which tile 1's softmax warps are held back at the start of every work item
#endif
constexpr int kA3Rows = 64;                                  // kv rows per step
constexpr int kA3Half = kA3Rows * 64 * 2;                    // one [64 x 64] bf16 swizzled half = 8 KB
constexpr int kA3Tile = 2 * kA3Half;                         // one K_j or V_j tile: 16 KB
constexpr int kA3Stage = 2 * kA3Tile;                        // [V_j | K_{j+2}] = 32 KB
constexpr int kA3Ring = 4;
constexpr int kA3Threads = 128 + 256;
constexpr int kA3SmemData = 2 * kTileBytes + kA3Ring * kA3Stage;
constexpr int kA3Smem = 1024 + kA3SmemData + 512;
constexpr uint32_t kDescLboV64 = (static_cast<uint32_t>(kA3Half) >> 4) << 16;

__global__ void __launch_bounds__(kA3Threads, 1) attention_kernel3(const __grid_constant__ AttnParams p) {
    constexpr int kR = kA3Ring;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    auto q_smem = [&](int q) { return smem_base + q * kTileBytes; };
    auto st_smem = [&](int s) { return smem_base + 2 * kTileBytes + s * kA3Stage; };
    const uint32_t bar_base = smem_base + kA3SmemData;
    const uint32_t q_full = bar_base, q_empty = bar_base + 8;
    auto st_full = [&](int s) { return bar_base + 16 + s * 8; };
    auto st_empty = [&](int s) { return bar_base + 16 + (kR + s) * 8; };
    auto s_full = [&](int q, int b) { return bar_base + 16 + (2 * kR + q * 2 + b) * 8; };
    auto p_full = [&](int q, int b) { return bar_base + 16 + (2 * kR + 4 + q * 2 + b) * 8; };
    // pv_done(q, 0): PV(n-2) of the item has completed; pv_done(q, 1): PV(n-1) has.  One commit per item each, so the phase
    // parity is the item parity and a waiter can never be more than one phase behind (a single barrier taking both commits
    // aliased when both PVs had completed before the epilogue's first wait -- found as a timeout with a slower softmax).
    auto pv_done = [&](int q, int w) { return bar_base + 16 + (2 * kR + 8 + q * 2 + w) * 8; };
    auto o_empty = [&](int q) { return bar_base + 16 + (2 * kR + 12 + q) * 8; };
    const uint32_t tmem_slot = bar_base + 16 + (2 * kR + 14) * 8;

    const int warp = threadIdx.x >> 5;
    const int lane = lane_id();

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&p.tmQ);
        prefetch_tmap(&p.tmK);
        prefetch_tmap(&p.tmV);
    }
    if (warp == 1 && elect_one()) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 2);                 // both issuers
        for (int s = 0; s < kR; ++s) { mbar_init(st_full(s), 1); mbar_init(st_empty(s), 2); }
        for (int q = 0; q < 2; ++q) {
            for (int b = 0; b < 2; ++b) { mbar_init(s_full(q, b), 1); mbar_init(p_full(q, b), 4); }
            mbar_init(pv_done(q, 0), 1);
            mbar_init(pv_done(q, 1), 1);
            mbar_init(o_empty(q), 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ld_shared_u32(tmem_slot);
    auto s_tmem = [&](int q, int b) { return tmem_base + q * 128 + b * 64; };
    auto o_tmem = [&](int q) { return tmem_base + 256 + q * 128; };
    const int n_kv = p.n_kv;                 // 64-row steps

    if (warp == 0) {
        // ======================================= TMA producer =======================================
        uint32_t n = 0, it = 0;              // n: stage sequence number, continues across items
        bool ok = true;
        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x, ++it) {
            const int head = item / p.n_qblk;
            const int qb = item - head * p.n_qblk;
            const int col0 = head * kTile;
            if (!mbar_wait(q_empty, (it & 1u) ^ 1u, p.abort_flag, 70)) break;
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int row0 = (qb * 2 + q) * kTile;
                    tma_load_2d(q_smem(q), &p.tmQ, q_full, col0, row0);
                    tma_load_2d(q_smem(q) + kHalfBytes, &p.tmQ, q_full, col0 + 64, row0);
                }
            }
            __syncwarp();
            // stage = [first | second]; a negative step means "nothing in this half"
            auto load = [&](const CUtensorMap* tm_a, int ja, int jb) -> bool {
                const int slot = n % kR;
                const uint32_t ph = (n / kR) & 1u;
                if (!mbar_wait(st_empty(slot), ph ^ 1u, p.abort_flag, 71)) return false;
                if (elect_one()) {
                    const uint32_t dst = st_smem(slot);
                    mbar_arrive_expect_tx(st_full(slot), jb >= 0 ? kA3Stage : kA3Tile);
                    tma_load_2d(dst, tm_a, st_full(slot), col0, ja * kA3Rows);
                    tma_load_2d(dst + kA3Half, tm_a, st_full(slot), col0 + 64, ja * kA3Rows);
                    if (jb >= 0) {
                        tma_load_2d(dst + kA3Tile, &p.tmK, st_full(slot), col0, jb * kA3Rows);
                        tma_load_2d(dst + kA3Tile + kA3Half, &p.tmK, st_full(slot), col0 + 64, jb * kA3Rows);
                    }
                }
                __syncwarp();
                ++n;
                return true;
            };
            ok = load(&p.tmK, 0, n_kv > 1 ? 1 : -1);
            for (int j = 0; j < n_kv && ok; ++j) ok = load(&p.tmV, j, j + 2 < n_kv ? j + 2 : -1);
        }
    } else if (warp == 1 || warp == 3) {
        // ======================================= MMA issuer of query tile q =======================================
        const int q = warp >> 1;
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);      // Q (K-major) x K_j (K-major, 64 rows)
        constexpr uint32_t idesc_o = make_idesc_bf16(128, 128, 0, 1);     // P (TMEM) x V_j (MN-major, 64 rows)
        const bool leader = elect_one();
        const uint32_t qlo = desc_lo_k(q_smem(q));
        const uint32_t o_acc = o_tmem(q);
        uint32_t n = 0, it = 0;
        uint32_t p_phase = 0;                 // bit b = parity of p_full(q, b)
        bool ok = true;
        PE_TRACE_DECL(0)

        auto issue_s = [&](int b, uint32_t k_base) {
            const uint32_t klo = desc_lo_k(k_base);
            const uint32_t d = s_tmem(q, b);
#pragma unroll
            for (int kk = 0; kk < ((PE_A3_DBG & 1) ? 0 : 8); ++kk) {
                const uint32_t qoff = ((kk >> 2) * kHalfBytes + (kk & 3) * 32) >> 4;
                const uint32_t koff = ((kk >> 2) * kA3Half + (kk & 3) * 32) >> 4;
                if (kk == 0) umma_ss_lohi<false>(d, qlo + qoff, klo + koff, idesc_s);
                else umma_ss_lohi<true>(d, qlo + qoff, klo + koff, idesc_s);
            }
            umma_commit(s_full(q, b));
        };
        auto issue_pv = [&](int b, uint32_t v_base, bool accumulate) {
            const uint32_t vlo = ((v_base & 0x3ffffu) >> 4) | kDescLboV64;
            const uint32_t a = s_tmem(q, b);
#pragma unroll
            for (int kk = 0; kk < ((PE_A3_DBG & 2) ? 0 : 4); ++kk)
                umma_ts_lohi(o_acc, a + kk * 8, vlo + kk * (2048 >> 4), idesc_o, (accumulate || kk != 0) ? 1u : 0u);
        };

        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x, ++it) {
            if (!mbar_wait(q_full, it & 1u, p.abort_flag, 80)) break;
            {
                // prologue stage [K_0 | K_1]: S(q, 0) and S(q, 1) fill both S buffers
                const uint32_t slot = n % kR;
                if (!mbar_wait(st_full(slot), (n / kR) & 1u, p.abort_flag, 81)) break;
                ++n;
                tc_fence_after();
                if (leader) {
                    issue_s(0, st_smem(slot));
                    if (n_kv > 1) issue_s(1, st_smem(slot) + kA3Tile);
                    umma_commit(st_empty(slot));
                    if (n_kv <= 2) umma_commit(q_empty);              // these were the item's last S MMAs
                }
                __syncwarp();
            }
            // previous item's epilogue must have drained O_q before PV(0) overwrites it
            if (!mbar_wait(o_empty(q), (it & 1u) ^ 1u, p.abort_flag, 84)) break;
            for (int j = 0; j < n_kv; ++j) {
                const int b = j & 1;
                const uint32_t slot = n % kR;
                if (!mbar_wait(st_full(slot), (n / kR) & 1u, p.abort_flag, 82)) { ok = false; break; }
                ++n;
                if (q == 0) PE_TRACE(10, j);
                if (!mbar_wait(p_full(q, b), (p_phase >> b) & 1u, p.abort_flag, 85)) { ok = false; break; }
                p_phase ^= 1u << b;
                tc_fence_after();
                if (q == 0) PE_TRACE(12, j);
                if (leader) {
                    issue_pv(b, st_smem(slot), j > 0);
                    if (j + 2 >= n_kv) umma_commit(pv_done(q, j + 2 - n_kv));   // only the last two PVs are ever waited for
                    if (q == 0) PE_TRACE(40, j);
                    if (j + 2 < n_kv) issue_s(b, st_smem(slot) + kA3Tile);
                    umma_commit(st_empty(slot));
                    if (j + 3 == n_kv) umma_commit(q_empty);            // the item's last S MMAs were just issued
                }
                __syncwarp();
                if (q == 0) PE_TRACE(14, j);
            }
        }
    } else if (warp >= 4) {
        // ======================================= softmax / correction / epilogue =======================================
        const int q = (warp - 4) >> 2;
        const int wq = warp & 3;
        const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
        const uint32_t o_addr = o_tmem(q) + lane_off;
        const int row_in_tile = wq * 32 + lane;
        uint32_t s_phase = 0;                 // bit b = parity of s_full(q, b)
        uint32_t item_par = 0;                // parity of this CTA's item count = phase parity of pv_done(q, *)
        bool ok = true;
        PE_TRACE_DECL(1 + q)
        for (int item = blockIdx.x; item < p.n_items && ok; item += gridDim.x) {
            const int head = item / p.n_qblk;
            const int qb = item - head * p.n_qblk;
            float m_ref = -INFINITY;          // exponent reference (log2 domain) of O and l; trails the running max
            float l = 0.f;
            float f_pending = 1.0f;           // O must still be multiplied by this before the next PV
            bool s_ready = false;             // answer of the non-blocking probe of s_full for the coming step
            // One KV step.  kFirst / kFull are compile-time so that the steady-state loop (full 64-row steps after the first) is one
            // compact straight-line body: with the first-step, ragged-tail and redo code inlined into a single loop the executed path
            // was spread over 28 KB of code and ~9 % of the softmax warps' samples were instruction-fetch stalls (r1 ncu source view).
            auto kv_step = [&](auto first_tag, auto full_tag, const int j) -> bool {
                constexpr bool kFirst = decltype(first_tag)::value;
                constexpr bool kFull = decltype(full_tag)::value;
                const int b = j & 1;
                if (wq == 0) PE_TRACE(20 + q, j);
                if (kFirst || !__all_sync(0xffffffffu, s_ready)) {
                    if (!mbar_wait(s_full(q, b), (s_phase >> b) & 1u, p.abort_flag, 90)) return false;
                }
                s_phase ^= 1u << b;
                tc_fence_after();
#if PE_A3_SKEW > 0
                if (kFirst && q == 1) {
                    const long long t_skew = clock64() + PE_A3_SKEW;
                    while (clock64() < t_skew) {}
                }
#endif
                if (wq == 0) PE_TRACE(22 + q, j);
                const uint32_t s_addr = s_tmem(q, b) + lane_off;
                const int kv_valid = kFull ? kA3Rows : p.S - j * kA3Rows;
                uint32_t ra[32], rb[32], pk[32];
                tmem_ld32(s_addr, ra);
                tmem_ld32(s_addr + 32, rb);
                tmem_ld_wait();
                if (wq == 0) PE_TRACE(30 + q, j);
                if (kFirst) {
                    // first step: exact row max
                    float mx0 = -INFINITY;
                    row_max_chunk(ra, 0, kv_valid, kFull, mx0);
                    row_max_chunk(rb, 32, kv_valid, kFull, mx0);
                    m_ref = mx0 * p.scale_log2;
                }
                float f_apply = f_pending;
                f_pending = 1.0f;
                float lsum = 0.f, mx = -INFINITY;
                softmax_chunk(ra, 0, kv_valid, kFull, p.scale_log2, m_ref, lsum, mx, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
                softmax_chunk(rb, 32, kv_valid, kFull, p.scale_log2, m_ref, lsum, mx, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
                // Is S(j+1) there?  (It was issued a whole step ago.)  Asked here, answered at the top of the next step, so the
                // ~130-cycle latency of a barrier poll is off the critical path (r1 ncu: 13 % of the softmax warps' samples).
                s_ready = (j + 1 < n_kv) && mbar_test_wait(s_full(q, b ^ 1), (s_phase >> (b ^ 1)) & 1u);
                const float mx_scaled = mx * p.scale_log2;
                const bool jump = mx_scaled - m_ref > 100.0f;
                if (__any_sync(0xffffffffu, jump || f_apply != 1.0f)) {
                    // rare path: overflow guard (redo the step against the true max; S is still in registers) and / or the
                    // deferred O rescale.  Both touch O, which PV(j-1) may still be updating: wait for it first.
                    if (__any_sync(0xffffffffu, jump)) {
                        const float m_new = fmaxf(m_ref, mx_scaled);
                        const float f = ex2(m_ref - m_new);
                        f_apply *= f;
                        l *= f;
                        m_ref = m_new;
                        lsum = 0.f;
                        softmax_chunk(ra, 0, kv_valid, kFull, p.scale_log2, m_ref, lsum, mx, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
                        softmax_chunk(rb, 32, kv_valid, kFull, p.scale_log2, m_ref, lsum, mx, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
                    }
                    if (!kFirst) {
                        bool done;
                        if (j + 1 < n_kv) done = mbar_wait(s_full(q, b ^ 1), (s_phase >> (b ^ 1)) & 1u, p.abort_flag, 91);   // S(j+1) follows PV(j-1)
                        else done = mbar_wait(pv_done(q, 0), item_par, p.abort_flag, 94);                                     // j = n-1: PV(n-2)
                        if (!done) return false;
                        tc_fence_after();
                        scale_o_rows(o_addr, f_apply, ra);
                    }
                }
                if (wq == 0) PE_TRACE(26 + q, j);
                tmem_st16(s_addr, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
                tmem_st16(s_addr + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(q, b));
                if (wq == 0) PE_TRACE(28 + q, j);
                l += lsum;
                if (mx_scaled > m_ref + 8.0f) {
                    f_pending = ex2(m_ref - mx_scaled);
                    l *= f_pending;
                    m_ref = mx_scaled;
                }
                return true;
            };
            using T = std::true_type;
            using F = std::false_type;
            const int n_full = p.S / kA3Rows;         // steps with all 64 kv rows valid (n_kv - 1 or n_kv)
            ok = n_full >= 1 ? kv_step(T{}, T{}, 0) : kv_step(T{}, F{}, 0);
#pragma unroll 1
            for (int j = 1; j < n_full && ok; ++j) ok = kv_step(F{}, T{}, j);
            if (ok && n_full >= 1 && n_full < n_kv) ok = kv_step(F{}, F{}, n_full);
            if (!ok) break;

            // ---- epilogue: O / l -> bf16 -> global ----
            if (!mbar_wait(pv_done(q, 1), item_par, p.abort_flag, 92)) break;      // PV(n-1), hence every PV of the item, has completed
            item_par ^= 1u;
            tc_fence_after();
            const float inv = f_pending / l;
            const long long row = (long long)(qb * 2 + q) * kTile + row_in_tile;
            bf16* orow = p.o + row * p.ldo + head * kTile;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld32(o_addr + c * 32, r);
                tmem_ld_wait();
                if (row < p.S) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(r[8 * v]) * inv, __uint_as_float(r[8 * v + 1]) * inv);
                        o.y = pack_bf16(__uint_as_float(r[8 * v + 2]) * inv, __uint_as_float(r[8 * v + 3]) * inv);
                        o.z = pack_bf16(__uint_as_float(r[8 * v + 4]) * inv, __uint_as_float(r[8 * v + 5]) * inv);
                        o.w = pack_bf16(__uint_as_float(r[8 * v + 6]) * inv, __uint_as_float(r[8 * v + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + c * 32 + v * 8) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty(q));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

int launch_attention3(Handle* h, AttnParams& p, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        PE_CHECK_CUDA(h, cudaFuncSetAttribute(attention_kernel3, cudaFuncAttributeMaxDynamicSharedMemorySize, kA3Smem));
        configured = true;
    }
    p.n_qblk = ceil_div(p.S, kTile * 2);
    p.n_items = p.H * p.n_qblk;
    p.n_kv = ceil_div(p.S, kA3Rows);
    int ctas = h->sm_count;
    if (ctas > p.n_items) ctas = p.n_items;
    attention_kernel3<<<ctas, kA3Threads, kA3Smem, stream>>>(p);
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

// -------------------------------------------------------------------------------------------------
// small generic attention (CUDA cores) for the training-path encoders whose sequences are tiny:
// DINOv2 ViT-B (261 tokens, 12 heads x 64; transformers modeling_dinov2_with_registers.py:174-254)
// and the perceiver resampler (64 latent queries over <= 10304 media+latent keys, 8 heads x 64;
// helpers.py:21-65).  One warp per query row; lanes stride over keys with a private online softmax
// and are merged at the end.
// -------------------------------------------------------------------------------------------------
// kStage: the (batch, head)'s K and V rows are staged once per CTA in shared memory (row pitch D + 8 elements: the lanes' 16-byte reads of 32
// different rows then fall on distinct banks) and each of the CTA's 8 warps walks kSmallQPerWarp queries over them -- for the DINOv2 shapes
// (261 keys) every query warp used to re-read K / V from L2 (1.2 GB of L2 traffic per ViT layer: 350 us for 1.25 GFLOP).
constexpr int kSmallQPerWarp = 4;
template <int D, bool kStage>
__global__ void __launch_bounds__(kStage ? 256 : 128) small_attention_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                                                                             bf16* __restrict__ o, int H, int Sq, int Skv, long long ldq, long long ldkv,
                                                                             long long ldo, float scale) {
    constexpr int kWarps = kStage ? 8 : 4;
    constexpr int kPitch = D + 8;
    __shared__ float qs[kWarps][D];
    extern __shared__ __align__(16) uint8_t kv_stage_raw[];
    bf16* ks = reinterpret_cast<bf16*>(kv_stage_raw);
    bf16* vs = ks + (size_t)Skv * kPitch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hd = blockIdx.y, b = blockIdx.z;
    if (kStage) {
        const int vec_per_row = D / 8;
        for (int i = threadIdx.x; i < Skv * vec_per_row; i += blockDim.x) {
            const int r = i / vec_per_row, c = i - r * vec_per_row;
            const long long g = ((long long)b * Skv + r) * ldkv + hd * D + c * 8;
            *reinterpret_cast<uint4*>(ks + (size_t)r * kPitch + c * 8) = *reinterpret_cast<const uint4*>(k + g);
            *reinterpret_cast<uint4*>(vs + (size_t)r * kPitch + c * 8) = *reinterpret_cast<const uint4*>(v + g);
        }
        __syncthreads();
    }
    for (int qq_ = 0; qq_ < (kStage ? kSmallQPerWarp : 1); ++qq_) {
    const int qi = kStage ? (blockIdx.x * kWarps + warp) * kSmallQPerWarp + qq_ : blockIdx.x * 4 + warp;
    if (qi >= Sq) break;
    const bf16* qrow = q + ((long long)b * Sq + qi) * ldq + hd * D;
    __syncwarp();
    for (int d = lane; d < D; d += 32) qs[warp][d] = __bfloat162float(qrow[d]) * scale;
    __syncwarp();
    float m = -INFINITY, l = 0.f;
    float acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    for (int kj = lane; kj < Skv; kj += 32) {
        const bf16* krow = kStage ? ks + (size_t)kj * kPitch : k + ((long long)b * Skv + kj) * ldkv + hd * D;
        float s = 0.f;
#pragma unroll
        for (int d8 = 0; d8 < D / 8; ++d8) {
            const uint4 u = *reinterpret_cast<const uint4*>(krow + d8 * 8);
            const float2 a = unpack_bf16(u.x), bq = unpack_bf16(u.y), c = unpack_bf16(u.z), e = unpack_bf16(u.w);
            const float* qq = &qs[warp][d8 * 8];
            s += a.x * qq[0] + a.y * qq[1] + bq.x * qq[2] + bq.y * qq[3] + c.x * qq[4] + c.y * qq[5] + e.x * qq[6] + e.y * qq[7];
        }
        const float m_new = fmaxf(m, s);
        const float f = __expf(m - m_new);
        const float pw = __expf(s - m_new);
        l = l * f + pw;
        const bf16* vrow = kStage ? vs + (size_t)kj * kPitch : v + ((long long)b * Skv + kj) * ldkv + hd * D;
#pragma unroll
        for (int d8 = 0; d8 < D / 8; ++d8) {
            const uint4 u = *reinterpret_cast<const uint4*>(vrow + d8 * 8);
            const float2 a = unpack_bf16(u.x), bq = unpack_bf16(u.y), c = unpack_bf16(u.z), e = unpack_bf16(u.w);
            float* ac = &acc[d8 * 8];
            ac[0] = ac[0] * f + pw * a.x; ac[1] = ac[1] * f + pw * a.y; ac[2] = ac[2] * f + pw * bq.x; ac[3] = ac[3] * f + pw * bq.y;
            ac[4] = ac[4] * f + pw * c.x; ac[5] = ac[5] * f + pw * c.y; ac[6] = ac[6] * f + pw * e.x; ac[7] = ac[7] * f + pw * e.y;
        }
        m = m_new;
    }
    // merge the 32 partial softmaxes
    float mg = m;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, off));
    const float f = (m == -INFINITY) ? 0.f : __expf(m - mg);
    l *= f;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
    const float inv = 1.0f / l;
    bf16* orow = o + ((long long)b * Sq + qi) * ldo + hd * D;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float a = acc[d] * f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == (d & 31)) orow[d] = __float2bfloat16_rn(a * inv);
    }
    }   // queries of this warp
}

}  // namespace

int attention_run(Handle* h, const void* q, const void* k, const void* v, void* o, int S, int H, int64_t ld, float scale,
                  int flags, cudaStream_t stream, float* lse) {
    PE_REQUIRE(h, q && k && v && o, "pe_attention_fwd: null pointer");
    PE_REQUIRE(h, lse == nullptr || (flags & ~3) == 0, "pe_attention_fwd_lse: only the default kernel (flags 0..3) writes the row statistics");
    PE_REQUIRE(h, S > 0 && H > 0, "pe_attention_fwd: S and H must be positive (S=%d H=%d)", S, H);
    PE_REQUIRE(h, ld >= (int64_t)H * 128 && ld % 8 == 0, "pe_attention_fwd: ld must be >= H*128 and a multiple of 8");
    PE_REQUIRE(h, ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                    reinterpret_cast<uintptr_t>(o)) & 15) == 0, "pe_attention_fwd: q/k/v/o must be 16-byte aligned");
    AttnParams p;
    memset(&p, 0, sizeof(p));
    int rc = make_tmap_2d(h, &p.tmQ, q, (uint64_t)S, (uint64_t)H * 128, (uint64_t)ld, 128);
    if (rc) return rc;
    const uint32_t kv_box_rows = (flags & PE_ATTN_FLAG_KV64) ? 64 : 128;
    rc = make_tmap_2d(h, &p.tmK, k, (uint64_t)S, (uint64_t)H * 128, (uint64_t)ld, kv_box_rows);
    if (rc) return rc;
    rc = make_tmap_2d(h, &p.tmV, v, (uint64_t)S, (uint64_t)H * 128, (uint64_t)ld, kv_box_rows);
    if (rc) return rc;
    p.o = static_cast<bf16*>(o);
    p.lse = lse;
    p.ldo = ld;
    p.S = S;
    p.H = H;
    p.n_kv = ceil_div(S, kTile);
    p.scale_log2 = scale * 1.4426950408889634f;
    p.abort_flag = h->abort_flag;
#ifdef PE_ATTN_TRACE
    p.trace = static_cast<long long*>(h->workspace);
    if (p.trace) cudaMemsetAsync(p.trace, 0, 200000, stream);
#endif
    if (flags & PE_ATTN_FLAG_KV64)
        return launch_attention3(h, p, stream);          // 64-row KV steps, double-buffered S: softmax off the MMA latency chain
    if (flags & PE_ATTN_FLAG_HALF_ROW)
        return launch_attention4(h, p, stream);          // half-row threads, trailing reference (no per-step exchange), exp2 split MUFU / FMA
    if (flags & PE_ATTN_FLAG_SPLIT_ROW_SOFTMAX)
        return launch_attention2(h, p, stream);          // split-row softmax: exact max every step, two warps per SMSP per tile
    const bool one_tile = (flags & PE_ATTN_FLAG_SINGLE_Q_TILE) != 0;
    const bool p_smem = (flags & PE_ATTN_FLAG_P_VIA_SMEM) != 0;
    if (one_tile) return p_smem ? launch_attention<1, false>(h, p, stream) : launch_attention<1, true>(h, p, stream);
    return p_smem ? launch_attention<2, false>(h, p, stream) : launch_attention<2, true>(h, p, stream);
}

int attention_routed_run(Handle* h, const void* q, const void* k, const void* v, int S, int H, int64_t ld, float scale, int flags, int n_route,
                         const int32_t* route_end, void* const* o_route, int64_t ldo, cudaStream_t stream) {
    PE_REQUIRE(h, q && k && v && route_end && o_route, "pe_attention_fwd_routed: null pointer");
    PE_REQUIRE(h, S > 0 && H > 0 && n_route >= 1 && n_route <= 8, "pe_attention_fwd_routed: bad sizes (S=%d H=%d n_route=%d)", S, H, n_route);
    PE_REQUIRE(h, ld >= (int64_t)H * 128 && ld % 8 == 0 && ldo % 8 == 0, "pe_attention_fwd_routed: ld must be >= H*128, ld / ldo multiples of 8");
    PE_REQUIRE(h, (flags & ~3) == 0, "pe_attention_fwd_routed: only the default kernel (flags 0..3) supports routed output");
    PE_REQUIRE(h, route_end[n_route - 1] >= S, "pe_attention_fwd_routed: the last route must cover row S-1");
    AttnParams p;
    memset(&p, 0, sizeof(p));
    int rc = make_tmap_2d(h, &p.tmQ, q, (uint64_t)S, (uint64_t)H * 128, (uint64_t)ld, 128);
    if (rc) return rc;
    rc = make_tmap_2d(h, &p.tmK, k, (uint64_t)S, (uint64_t)H * 128, (uint64_t)ld, 128);
    if (rc) return rc;
    rc = make_tmap_2d(h, &p.tmV, v, (uint64_t)S, (uint64_t)H * 128, (uint64_t)ld, 128);
    if (rc) return rc;
    p.n_route = n_route;
    for (int i = 0; i < n_route; ++i) {
        PE_REQUIRE(h, o_route[i] != nullptr && (reinterpret_cast<uintptr_t>(o_route[i]) & 15) == 0, "pe_attention_fwd_routed: route %d: null or unaligned pointer", i);
        p.route_end[i] = route_end[i];
        p.route_o[i] = static_cast<bf16*>(o_route[i]);
    }
    p.o = p.route_o[0];
    p.ldo = ldo;
    p.S = S;
    p.H = H;
    p.n_kv = ceil_div(S, kTile);
    p.scale_log2 = scale * 1.4426950408889634f;
    p.abort_flag = h->abort_flag;
    const bool one_tile = (flags & PE_ATTN_FLAG_SINGLE_Q_TILE) != 0;
    const bool p_smem = (flags & PE_ATTN_FLAG_P_VIA_SMEM) != 0;
    if (one_tile) return p_smem ? launch_attention<1, false>(h, p, stream) : launch_attention<1, true>(h, p, stream);
    return p_smem ? launch_attention<2, false>(h, p, stream) : launch_attention<2, true>(h, p, stream);
}

int small_attention_run(Handle* h, const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Skv, int D,
                        int64_t ldq, int64_t ldkv, int64_t ldo, float scale, cudaStream_t s) {
    PE_REQUIRE(h, q && k && v && o, "pe_small_attention: null pointer");
    PE_REQUIRE(h, B > 0 && H > 0 && Sq > 0 && Skv > 0, "pe_small_attention: sizes must be positive");
    PE_REQUIRE(h, D == 64 || D == 128, "pe_small_attention: head dim must be 64 or 128 (got %d)", D);
    PE_REQUIRE(h, ldq % 8 == 0 && ldkv % 8 == 0, "pe_small_attention: ldq / ldkv must be multiples of 8");
    const bf16 *qb = static_cast<const bf16*>(q), *kb = static_cast<const bf16*>(k), *vb = static_cast<const bf16*>(v);
    bf16* ob = static_cast<bf16*>(o);
    const size_t stage_bytes = (size_t)2 * Skv * (D + 8) * sizeof(bf16);
    if (stage_bytes <= 160 * 1024 && Sq >= 32) {
        // K / V of one (batch, head) fit in shared memory (DINOv2: 261 keys x 64 -> 75 KB): stage them once per 32 queries
        const dim3 grid(ceil_div(Sq, 8 * kSmallQPerWarp), H, B);
        if (D == 64) {
            PE_CHECK_CUDA(h, cudaFuncSetAttribute(small_attention_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
            small_attention_kernel<64, true><<<grid, 256, stage_bytes, s>>>(qb, kb, vb, ob, H, Sq, Skv, ldq, ldkv, ldo, scale);
        } else {
            PE_CHECK_CUDA(h, cudaFuncSetAttribute(small_attention_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
            small_attention_kernel<128, true><<<grid, 256, stage_bytes, s>>>(qb, kb, vb, ob, H, Sq, Skv, ldq, ldkv, ldo, scale);
        }
    } else {
        const dim3 grid(ceil_div(Sq, 4), H, B);
        if (D == 64) small_attention_kernel<64, false><<<grid, 128, 0, s>>>(qb, kb, vb, ob, H, Sq, Skv, ldq, ldkv, ldo, scale);
        else small_attention_kernel<128, false><<<grid, 128, 0, s>>>(qb, kb, vb, ob, H, Sq, Skv, ldq, ldkv, ldo, scale);
    }
    PE_CHECK_CUDA(h, cudaGetLastError());
    return PE_OK;
}

}  // namespace pe
