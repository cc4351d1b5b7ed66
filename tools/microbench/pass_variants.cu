// Microbenchmark 5: source-level variants of the 64-element softmax pass (what schedule does ptxas produce, what does it cost?).
// Inputs live in registers (loaded once from TMEM), the reference m moves every iteration.  Cycles per warp-element per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<unsigned*>(&v); }
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }


__device__ __forceinline__ float exp2_fma(float x) {
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float p = fmaf(f, 0.05550410866f, 0.24022650696f);
    p = fmaf(p, f, 0.69314718056f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ void exp2_fma_pair(uint64_t x2, float& p0, float& p1) {
    float x0, x1; upk2(x2, x0, x1);
    x2 = pk2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
    const uint64_t magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f), neg1 = pk2(-1.0f, -1.0f);
    const uint64_t t2 = fadd2(x2, magic), n2 = fadd2(t2, nmagic), f2 = ffma2(n2, neg1, x2);
    uint64_t pz = ffma2(f2, pk2(0.05550410866f, 0.05550410866f), pk2(0.24022650696f, 0.24022650696f));
    pz = ffma2(pz, f2, pk2(0.69314718056f, 0.69314718056f));
    pz = ffma2(pz, f2, pk2(1.0f, 1.0f));
    float z0, z1, t0, t1; upk2(pz, z0, z1); upk2(t2, t0, t1);
    p0 = __int_as_float(__float_as_int(z0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(z1) + (__float_as_int(t1) << 23));
}
// poly without the clamp (inputs are known to be > -125 after the reference subtraction in all but pathological rows; the
// kernel clamps once per pair with one packed max instead)
__device__ __forceinline__ void exp2_fma_pair_nc(uint64_t x2, float& p0, float& p1) {
    const uint64_t magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f), neg1 = pk2(-1.0f, -1.0f);
    const uint64_t t2 = fadd2(x2, magic), n2 = fadd2(t2, nmagic), f2 = ffma2(n2, neg1, x2);
    uint64_t pz = ffma2(f2, pk2(0.05550410866f, 0.05550410866f), pk2(0.24022650696f, 0.24022650696f));
    pz = ffma2(pz, f2, pk2(0.69314718056f, 0.69314718056f));
    pz = ffma2(pz, f2, pk2(1.0f, 1.0f));
    float z0, z1, t0, t1; upk2(pz, z0, z1); upk2(t2, t0, t1);
    p0 = __int_as_float(__float_as_int(z0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(z1) + (__float_as_int(t1) << 23));
}

// every variant: r[64] raw scores -> pk[32] packed bf16 P, row sum into s, running max of the raw scores into mx
template <int PV>
__device__ __forceinline__ void pass64(const float (&r)[64], float c, float m, float& s, float& mx, uint32_t (&pk)[32]) {
    if (PV == 0) {                       // the kernel's current form: 2 sum chains
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            mx = fmaxf(mx, fmaxf(r[2 * i], r[2 * i + 1]));
            const float p0 = ex2(fmaf(r[2 * i], c, -m)), p1 = ex2(fmaf(r[2 * i + 1], c, -m));
            s0 += p0; s1 += p1;
            pk[i] = pack(p0, p1);
        }
        s += s0 + s1;
    } else if (PV == 1) {                // packed ffma2 / fadd2
        const uint64_t c2 = pk2(c, c), nm2 = pk2(-m, -m);
        uint64_t sum2 = pk2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            mx = fmaxf(mx, fmaxf(r[2 * i], r[2 * i + 1]));
            float x0, x1;
            upk2(ffma2(pk2(r[2 * i], r[2 * i + 1]), c2, nm2), x0, x1);
            const float p0 = ex2(x0), p1 = ex2(x1);
            sum2 = fadd2(sum2, pk2(p0, p1));
            pk[i] = pack(p0, p1);
        }
        float a, b; upk2(sum2, a, b);
        s += a + b;
    } else if (PV == 2) {                // 8 sum chains
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            mx = fmaxf(mx, fmaxf(r[2 * i], r[2 * i + 1]));
            const float p0 = ex2(fmaf(r[2 * i], c, -m)), p1 = ex2(fmaf(r[2 * i + 1], c, -m));
            acc[(2 * i) & 7] += p0; acc[(2 * i + 1) & 7] += p1;
            pk[i] = pack(p0, p1);
        }
        s += ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    } else if (PV == 3) {                // all exponentials first, then the consumers
        float p[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) p[i] = ex2(fmaf(r[i], c, -m));
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            mx = fmaxf(mx, fmaxf(r[2 * i], r[2 * i + 1]));
            s0 += p[2 * i]; s1 += p[2 * i + 1];
            pk[i] = pack(p[2 * i], p[2 * i + 1]);
        }
        s += s0 + s1;
    } else if (PV == 4) {                // the row sum taken from the ROUNDED bf16 pairs with packed adds (tree), fewer fp32 adds
        float s0 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            mx = fmaxf(mx, fmaxf(fmaxf(r[2 * i], r[2 * i + 1]), fmaxf(r[2 * i + 2], r[2 * i + 3])));
            const float p0 = ex2(fmaf(r[2 * i], c, -m)), p1 = ex2(fmaf(r[2 * i + 1], c, -m));
            const float p2 = ex2(fmaf(r[2 * i + 2], c, -m)), p3 = ex2(fmaf(r[2 * i + 3], c, -m));
            s0 += (p0 + p1) + (p2 + p3);
            pk[i] = pack(p0, p1);
            pk[i + 1] = pack(p2, p3);
        }
        s += s0;
    } else if (PV == 5) {                // volatile MUFUs in groups of 8, consumers one group behind
        float p[64];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int g = 0; g < 9; ++g) {
            if (g < 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[g * 8 + i] = ex2v(fmaf(r[g * 8 + i], c, -m));
            }
            if (g > 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = (g - 1) * 8 + 2 * i;
                    mx = fmaxf(mx, fmaxf(r[e], r[e + 1]));
                    s0 += p[e]; s1 += p[e + 1];
                    pk[e / 2] = pack(p[e], p[e + 1]);
                }
            }
        }
        s += s0 + s1;
    } else if (PV == 6) {                // no row sum at all (lower bound if the sum moved elsewhere)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            mx = fmaxf(mx, fmaxf(r[2 * i], r[2 * i + 1]));
            const float p0 = ex2(fmaf(r[2 * i], c, -m)), p1 = ex2(fmaf(r[2 * i + 1], c, -m));
            pk[i] = pack(p0, p1);
        }
    } else if (PV >= 10 && PV < 20) {    // scalar poly: (PV-10) of every 8 elements on the FMA pipe
        constexpr int K = PV - 10;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            mx = fmaxf(mx, fmaxf(r[2 * i], r[2 * i + 1]));
            const float x0 = fmaf(r[2 * i], c, -m), x1 = fmaf(r[2 * i + 1], c, -m);
            const bool poly0 = (((2 * i) & 7) * K) / 8 != ((((2 * i) & 7) + 1) * K) / 8;
            const bool poly1 = (((2 * i + 1) & 7) * K) / 8 != ((((2 * i + 1) & 7) + 1) * K) / 8;
            const float p0 = poly0 ? exp2_fma(x0) : ex2(x0), p1 = poly1 ? exp2_fma(x1) : ex2(x1);
            s0 += p0; s1 += p1;
            pk[i] = pack(p0, p1);
        }
        s += s0 + s1;
    } else if (PV >= 20 && PV < 40) {    // packed: (PV-20 or PV-30) of every 8 PAIRS on the FMA pipe (30+: without the clamp)
        constexpr int K = PV >= 30 ? PV - 30 : PV - 20;
        const uint64_t c2 = pk2(c, c), nm2 = pk2(-m, -m);
        uint64_t sum2 = pk2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            mx = fmaxf(mx, fmaxf(r[2 * i], r[2 * i + 1]));
            const uint64_t x2 = ffma2(pk2(r[2 * i], r[2 * i + 1]), c2, nm2);
            const bool poly = ((i & 7) * K) / 8 != (((i & 7) + 1) * K) / 8;
            float p0, p1;
            if (poly) { if (PV >= 30) exp2_fma_pair_nc(x2, p0, p1); else exp2_fma_pair(x2, p0, p1); }
            else { float x0, x1; upk2(x2, x0, x1); p0 = ex2(x0); p1 = ex2(x1); }
            sum2 = fadd2(sum2, pk2(p0, p1));
            pk[i] = pack(p0, p1);
        }
        float a, b; upk2(sum2, a, b);
        s += a + b;
    } else if (PV == 7) {                // no max, no sum: FFMA, MUFU, pack only
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float p0 = ex2(fmaf(r[2 * i], c, -m)), p1 = ex2(fmaf(r[2 * i + 1], c, -m));
            pk[i] = pack(p0, p1);
        }
    }
}

__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// SPIN: 0 none, 1 = one extra warp per SMSP spinning with all 32 lanes on mbarrier.try_wait, 2 = the same with one lane only,
// 3 = all lanes, with a __nanosleep(64) back-off between polls
template <int PV, int SPIN>
__global__ void __launch_bounds__(384, 1) k(float* out, long long* cyc, int iters, float c, float m0, float dm, const float* in, int nwork) {
    __shared__ unsigned long long bar;
    __shared__ volatile int done;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(1) : "memory");
        done = 0;
    }
    __syncthreads();
    if ((int)(threadIdx.x >> 5) >= nwork) {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        if (SPIN == 2 && (threadIdx.x & 31) != 0) return;
        while (!done) {
            if (try_wait(b, 0)) break;
            if (SPIN == 3) __nanosleep(64);
        }
        return;
    }
    float r[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = in[(threadIdx.x * 64 + i) & 4095];
    uint32_t pk[32];
    float s = 0.f, mx = -1e30f, m = m0;
    unsigned sink = 0;
    asm volatile("bar.sync 1, %0;" ::"r"(nwork * 32) : "memory");      // workers only (the spinner warps never arrive)
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        pass64<PV>(r, c, m, s, mx, pk);
#pragma unroll
        for (int i = 0; i < 32; ++i) sink += pk[i];          // stands in for the STTM consuming every packed word (IADD: ALU pipe)
        m += dm;
        mx += dm;                                            // keeps the max chain loop-variant
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + mx + __uint_as_float(sink);
    done = 1;
}
template <int PV, int SPIN>
void run(int warps, const char* name, const float* in) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    const int total = warps + (SPIN ? 4 : 0);
    for (int rep = 0; rep < 2; ++rep) k<PV, SPIN><<<148, total * 32>>>(out, cyc, iters, 0.1275f, -3.f, 0.001f, in, warps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-44s spin=%d warps/SMSP=%d  cycles per 64-element step = %7.1f  (%.2f per warp-element per SMSP) %s\n", name, SPIN, warps / 4, avg / iters,
           avg / iters / 64.0 / (warps / 4.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = (float)((i * 7) & 15) - 20.f;
    float* in; cudaMalloc(&in, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w : {4, 8}) {
        run<0, 0>(w, "0 current (2 sum chains)", in); run<0, 1>(w, "0 current (2 sum chains)", in); run<0, 2>(w, "0 current (2 sum chains)", in);
        run<0, 3>(w, "0 current (2 sum chains)", in);
        run<33, 0>(w, "33 packed poly 3/8, no clamp", in); run<33, 1>(w, "33 packed poly 3/8, no clamp", in);
    }
    return 0;
}
