// Microbenchmark 4: TMEM load/store throughput and the softmax step (LDTM 64 columns -> exp pass -> STTM 32 columns) without
// any MMA or barrier traffic.  148 CTAs, 4 or 8 warps (1 or 2 per SMSP; warp w uses TMEM lane quarter w % 4).
//   V=0 LDTM only (2 x 32x32b.x32 + wait per iteration)     V=1 LDTM + STTM (2 x x16) + waits
//   V=2 the full step: LDTM, scalar exp pass (fmaf, ex2, fadd, pack, fmax3), STTM     V=3 same, LDTM of the next step issued before the pass (prefetch)
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<unsigned*>(&v); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void pass(const uint32_t (&r)[32], float c, float m, float& s0, float& s1, float& mx, uint32_t* pk) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float a = __uint_as_float(r[2 * i]), b = __uint_as_float(r[2 * i + 1]);
        mx = fmaxf(mx, fmaxf(a, b));
        const float p0 = ex2(fmaf(a, c, -m)), p1 = ex2(fmaf(b, c, -m));
        s0 += p0; s1 += p1;
        pk[i] = pack(p0, p1);
    }
}

template <int V>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* cyc, int iters, float c, float m0, float dm) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;     // second warp of an SMSP: other columns
    uint32_t ra[32], rb[32], pk[32];
    // initialise the TMEM region this warp reads
#pragma unroll
    for (int i = 0; i < 32; ++i) pk[i] = __float_as_uint((float)((threadIdx.x * 7 + i * 3) & 15) - 20.f);
    for (int c0 = 0; c0 < 128; c0 += 16) tmem_st16(base + c0, pk);
    st_wait();
    float s0 = 0.f, s1 = 0.f, mx = -1e30f, m = m0;
    unsigned sink = 0;
    __syncthreads();
    long long t0 = clock64();
    if (V == 3 || V == 7 || V == 5 || V == 8) { tmem_ld32(base, ra); tmem_ld32(base + 32, rb); if (V == 5 || V == 8) ld_wait(); }
    for (int it = 0; it < iters; ++it) {
        const uint32_t a = base + (it & 1) * 64;
        if (V == 0) {
            tmem_ld32(a, ra); tmem_ld32(a + 32, rb); ld_wait();
            sink ^= ra[it & 31] ^ rb[(it * 3) & 31];
        } else if (V == 1) {
            tmem_ld32(a, ra); tmem_ld32(a + 32, rb); ld_wait();
            tmem_st16(a, ra); tmem_st16(a + 16, rb); st_wait();
        } else if (V == 2) {
            tmem_ld32(a, ra); tmem_ld32(a + 32, rb); ld_wait();
            pass(ra, c, m, s0, s1, mx, pk);
            pass(rb, c, m, s0, s1, mx, pk + 16);
            tmem_st16(a, pk); tmem_st16(a + 16, pk + 16); st_wait();
        } else if (V == 4) {
            tmem_ld32(a, ra); tmem_ld32(a + 32, rb); ld_wait();
            pass(ra, c, m, s0, s1, mx, pk);
            pass(rb, c, m, s0, s1, mx, pk + 16);
            sink ^= pk[0] ^ pk[31];
        } else if (V == 5) {
            pass(ra, c, m, s0, s1, mx, pk);
            pass(rb, c, m, s0, s1, mx, pk + 16);
            tmem_st16(a, pk); tmem_st16(a + 16, pk + 16); st_wait();
        } else if (V == 6) {
            tmem_ld32(a, ra); tmem_ld32(a + 32, rb); ld_wait();
            pass(ra, c, m, s0, s1, mx, pk);
            pass(rb, c, m, s0, s1, mx, pk + 16);
            tmem_st16(a, pk); tmem_st16(a + 16, pk + 16);
        } else if (V == 7) {
            ld_wait();
            pass(ra, c, m, s0, s1, mx, pk);
            pass(rb, c, m, s0, s1, mx, pk + 16);
            tmem_st16(a, pk); tmem_st16(a + 16, pk + 16);
            tmem_ld32(base + ((it + 1) & 1) * 64, ra);
            tmem_ld32(base + ((it + 1) & 1) * 64 + 32, rb);
            st_wait();
        } else if (V == 8) {
            pass(ra, c, m, s0, s1, mx, pk);
            pass(rb, c, m, s0, s1, mx, pk + 16);
            sink ^= pk[0] ^ pk[31];
        } else {
            ld_wait();
            pass(ra, c, m, s0, s1, mx, pk);
            tmem_ld32(base + ((it + 1) & 1) * 64, ra);
            pass(rb, c, m, s0, s1, mx, pk + 16);
            tmem_ld32(base + ((it + 1) & 1) * 64 + 32, rb);
            tmem_st16(a, pk); tmem_st16(a + 16, pk + 16); st_wait();
        }
        m += dm;
    }
    ld_wait();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + mx + __uint_as_float(sink ^ pk[3] ^ ra[1] ^ rb[2]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
template <int V>
void run(int warps, const char* name) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) k<V><<<148, warps * 32>>>(out, cyc, iters, 0.1275f, -3.f, 0.001f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-34s warps/SMSP=%d  cycles per 64-column step = %.1f  (%.2f per warp-element per SMSP; TMEM read %.0f B/clk/SM) %s\n", name, warps / 4,
           avg / iters, avg / iters / 64.0 / (warps / 4.0), warps * 32 * 64 * 4.0 / (avg / iters), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8}) {
        run<0>(w, "0 LDTM only"); run<1>(w, "1 LDTM + STTM"); run<2>(w, "2 LDTM, exp pass, STTM"); run<3>(w, "3 same with LDTM prefetch");
        run<4>(w, "4 LDTM + pass (no store)"); run<5>(w, "5 pass + STTM + wait (no load)"); run<6>(w, "6 LDTM, pass, STTM, no st wait");
        run<7>(w, "7 pass, STTM, LDTM next, st wait"); run<8>(w, "8 pass only (static inputs)");
    }
    return 0;
}
