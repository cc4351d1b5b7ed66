// Microbenchmark 2: softmax inner pass with part of the exponentials on the FMA pipe (scalar and packed f32x2 variants).
// Prints cycles per warp-element per SM sub-partition for 1 / 2 / 4 warps per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
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
// PACKED: 0 scalar, 1 f32x2.  POLY: number of pairs (of 16 per 32 elements) that use the FMA-pipe exp2.
template <int PACKED, int POLY>
__global__ void k(float* out, long long* cyc, int iters, float c, float m) {
    float r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = (float)((threadIdx.x * 7 + i * 3) & 15) - 20.f;
    float s0 = 0.f, s1 = 0.f;
    uint64_t sum2 = pk2(0.f, 0.f);
    unsigned sink = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint64_t c2 = pk2(c, c), nm2 = pk2(-m, -m);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const bool poly = ((i * POLY) / 16) != (((i + 1) * POLY) / 16);      // POLY of the 16 pairs, evenly spread
            float p0, p1;
            if (PACKED) {
                const uint64_t x2 = ffma2(pk2(r[2 * i], r[2 * i + 1]), c2, nm2);
                if (poly) exp2_fma_pair(x2, p0, p1);
                else { float x0, x1; upk2(x2, x0, x1); p0 = ex2(x0); p1 = ex2(x1); }
                sum2 = fadd2(sum2, pk2(p0, p1));
            } else {
                const float x0 = fmaf(r[2 * i], c, -m), x1 = fmaf(r[2 * i + 1], c, -m);
                if (poly) { p0 = exp2_fma(x0); p1 = exp2_fma(x1); } else { p0 = ex2(x0); p1 = ex2(x1); }
                s0 += p0; s1 += p1;
            }
            sink ^= pack(p0, p1);
        }
        m += 1e-7f;
        r[0] += 1e-7f * (float)(sink & 1);            // loop-carried dependence (static index: r[] must stay in registers)
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float a, b; upk2(sum2, a, b);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + a + b + __uint_as_float(sink);
}
template <int PACKED, int POLY>
void run(int warps) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) k<PACKED, POLY><<<148, warps * 32>>>(out, cyc, iters, 0.1275f, -3.f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%s poly %2d/16  warps/SMSP=%d  cycles per warp-element per SMSP = %.2f\n", PACKED ? "f32x2 " : "scalar", POLY, warps / 4,
           avg / ((warps / 4.0) * iters * 32));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<0, 0>(w); run<0, 4>(w); run<0, 6>(w); run<0, 8>(w);
        run<1, 0>(w); run<1, 4>(w); run<1, 6>(w); run<1, 8>(w); run<1, 10>(w);
    }
    return 0;
}
