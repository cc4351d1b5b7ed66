// Microbenchmark (round 2): issue rate per SM sub-partition of the instructions a softmax pass can be built from, to decide how
// the exponentials should be split between the MUFU and the FMA / ALU pipes:
//   MUFU.EX2 (f32), MUFU.EX2.F16 / .BF16 (is the half-precision form any faster?), FFMA, FFMA2 (f32x2), FADD2, FMNMX3, F2FP pack,
//   SHL+IADD (exponent insertion of a polynomial exp2), and complete per-pair recipes (MUFU path, polynomial path, mixes).
// 32 independent register chains per thread so nothing is latency-bound; inputs come from memory so nothing is folded.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_rates.cu -o pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 32
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

template <int MODE>
__global__ void k(const float* in, float* out, long long* cyc, int iters) {
    float r[N];
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = in[(threadIdx.x * N + i) & 1023];
    const float c = in[1], m = in[2];
    const uint64_t c2 = pk2(c, c), m2 = pk2(-m, -m);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            float a = r[i], b = r[i + 1];
            if (MODE == 0) {            // MUFU.EX2 f32
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
            } else if (MODE == 1) {     // MUFU.EX2.F16 on both halves of a packed word (2 MUFU ops)
                uint32_t u = __float_as_uint(a), v = __float_as_uint(b);
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u));
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v));
                a = __uint_as_float(u); b = __uint_as_float(v);
            } else if (MODE == 2) {     // scalar MUFU.EX2.F16 (one op)
                unsigned short h = (unsigned short)__float_as_uint(a), g = (unsigned short)__float_as_uint(b);
                asm volatile("ex2.approx.f16 %0, %0;" : "+h"(h));
                asm volatile("ex2.approx.f16 %0, %0;" : "+h"(g));
                a = __uint_as_float(h); b = __uint_as_float(g);
            } else if (MODE == 3) {     // FFMA 3-register
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(c), "f"(m));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b) : "f"(c), "f"(m));
            } else if (MODE == 4) {     // FFMA2
                uint64_t x = pk2(a, b);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(c2), "l"(m2));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(c2), "l"(m2));
                upk2(x, a, b);
            } else if (MODE == 5) {     // FADD2
                uint64_t x = pk2(a, b);
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(m2));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(m2));
                upk2(x, a, b);
            } else if (MODE == 6) {     // FMNMX3-ish: max of three
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(b) : "f"(a), "f"(m));
            } else if (MODE == 7) {     // F2FP pack
                uint32_t u, v;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a), "f"(b));
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(v) : "f"(b), "f"(a));
                a = __uint_as_float(u); b = __uint_as_float(v);
            } else if (MODE == 8) {     // SHL + IADD (exponent insertion)
                uint32_t u = __float_as_uint(a), v = __float_as_uint(b);
                asm volatile("shl.b32 %0, %0, 23;" : "+r"(u));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(v) : "r"(u));
                a = __uint_as_float(u); b = __uint_as_float(v);
            } else if (MODE == 9) {     // full MUFU recipe for a pair: FFMA2, 2 EX2, FADD2 (sum), pack
                uint64_t x = pk2(a, b);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(c2), "l"(m2));
                upk2(x, a, b);
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
                uint32_t u;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(b), "f"(a));
                a = a + __uint_as_float(u);
            } else if (MODE == 10 || MODE == 11 || MODE == 12) {   // mixes: every 4th / 3rd / 2nd pair takes the polynomial path
                const int every = MODE == 10 ? 4 : (MODE == 11 ? 3 : 2);
                uint64_t x = pk2(a, b);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(c2), "l"(m2));
                if (((i >> 1) % every) == 0) {
                    const uint64_t magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f), neg1 = pk2(-1.0f, -1.0f);
                    uint64_t t2, n2, f2, pz;
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(t2) : "l"(x), "l"(magic));
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(n2) : "l"(t2), "l"(nmagic));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f2) : "l"(n2), "l"(neg1), "l"(x));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pz) : "l"(f2), "l"(pk2(0.05550410866f, 0.05550410866f)), "l"(pk2(0.24022650696f, 0.24022650696f)));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pz) : "l"(f2), "l"(pk2(0.69314718056f, 0.69314718056f)));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pz) : "l"(f2), "l"(pk2(1.0f, 1.0f)));
                    float z0, z1, t0_, t1_;
                    upk2(pz, z0, z1); upk2(t2, t0_, t1_);
                    a = __int_as_float(__float_as_int(z0) + (__float_as_int(t0_) << 23));
                    b = __int_as_float(__float_as_int(z1) + (__float_as_int(t1_) << 23));
                } else {
                    upk2(x, a, b);
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
                }
                uint32_t u;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(b), "f"(a));
                a = a + __uint_as_float(u);
            }
            r[i] = a; r[i + 1] = b;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int ops_per_pair) {
    float *in, *out; long long* cyc;
    cudaMalloc(&in, 1024 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    float hin[1024]; for (int i = 0; i < 1024; ++i) hin[i] = -0.01f * (float)(i % 97) - 0.5f; hin[1] = 0.999f; hin[2] = 0.001f;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    const int iters = 1000;
    printf("%-44s", name);
    for (int warps : {4, 8, 16}) {
        k<MODE><<<148, warps * 32>>>(in, out, cyc, iters);
        k<MODE><<<148, warps * 32>>>(in, out, cyc, iters);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        // cycles per warp-level PAIR of elements per SMSP
        printf("  %dw/SMSP: %6.2f cyc/pair", warps / 4, avg / ((warps / 4.0) * iters * (N / 2)));
    }
    printf("   (%d counted ops per pair)\n", ops_per_pair);
    cudaFree(in); cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("MUFU.EX2 f32 x2", 2);
    run<1>("ex2.approx.f16x2 x2 (4 halves)", 2);
    run<2>("ex2.approx.f16 scalar x2", 2);
    run<3>("FFMA x2", 2);
    run<4>("FFMA2 x2 (4 elements)", 2);
    run<5>("FADD2 x2 (4 elements)", 2);
    run<6>("FMNMX3 x2", 2);
    run<7>("F2FP.BF16 pack x2", 2);
    run<8>("SHL + IADD", 2);
    run<9>("pair recipe: FFMA2, 2 EX2, pack, FADD", 5);
    run<10>("pair recipe, 1 of 4 pairs polynomial", 0);
    run<11>("pair recipe, 1 of 3 pairs polynomial", 0);
    run<12>("pair recipe, 1 of 2 pairs polynomial", 0);
    return 0;
}
