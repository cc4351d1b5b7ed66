// Microbenchmark 3: throughput of the softmax inner pass variants with inputs that change every iteration (the running
// reference m moves by a representable amount, so nothing is loop-invariant).  Prints cycles per warp-element per SMSP.
//   0 scalar: fmaf, ex2, fadd, pack, fmax3          1 packed: ffma2, 2x ex2, fadd2, pack, fmax3
//   2 packed, consumers lag 8 elements behind the MUFUs (explicit software pipeline)
//   3 f16x2 exponentials: ffma2 -> cvt.f16x2 -> ex2.approx.f16x2 (one MUFU per pair?) -> P stays f16x2, sum via hfma2 into f16x2 partials
//   4 MUFU.EX2 only      5 FFMA only     6 FFMA2 only      7 ex2.f16x2 only
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<unsigned*>(&v); }
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned cvt_f16x2(float lo, float hi) { unsigned r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ unsigned ex2_h2(unsigned x) { unsigned y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned hadd2(unsigned a, unsigned b) { unsigned y; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b)); return y; }

template <int V>
__global__ void k(float* out, long long* cyc, int iters, float c, float m0, float dm) {
    float r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = (float)((threadIdx.x * 7 + i * 3) & 15) - 20.f;
    float s0 = 0.f, s1 = 0.f, mx = -1e30f, m = m0;
    uint64_t sum2 = pk2(0.f, 0.f);
    unsigned sink = 0, hs = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint64_t c2 = pk2(c, c), nm2 = pk2(-m, -m);
        if (V == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float p0 = ex2(fmaf(r[2 * i], c, -m)), p1 = ex2(fmaf(r[2 * i + 1], c, -m));
                s0 += p0; s1 += p1;
                mx = fmaxf(mx, fmaxf(p0, p1));
                sink ^= pack(p0, p1);
            }
        } else if (V == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x0, x1;
                upk2(ffma2(pk2(r[2 * i], r[2 * i + 1]), c2, nm2), x0, x1);
                const float p0 = ex2(x0), p1 = ex2(x1);
                sum2 = fadd2(sum2, pk2(p0, p1));
                mx = fmaxf(mx, fmaxf(p0, p1));
                sink ^= pack(p0, p1);
            }
        } else if (V == 2) {
            float x[32], p[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) upk2(ffma2(pk2(r[2 * i], r[2 * i + 1]), c2, nm2), x[2 * i], x[2 * i + 1]);
#pragma unroll
            for (int g = 0; g < 5; ++g) {
                if (g < 4) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) p[g * 8 + i] = ex2(x[g * 8 + i]);
                }
                if (g > 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int e = (g - 1) * 8 + 2 * i;
                        sum2 = fadd2(sum2, pk2(p[e], p[e + 1]));
                        mx = fmaxf(mx, fmaxf(p[e], p[e + 1]));
                        sink ^= pack(p[e], p[e + 1]);
                    }
                }
            }
        } else if (V == 3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x0, x1;
                upk2(ffma2(pk2(r[2 * i], r[2 * i + 1]), c2, nm2), x0, x1);
                const unsigned ph = ex2_h2(cvt_f16x2(x0, x1));
                hs = hadd2(hs, ph);
                sink ^= ph;
            }
        } else if (V == 4) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sink ^= __float_as_uint(ex2(r[i] + m));
        } else if (V == 5) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sink ^= __float_as_uint(fmaf(r[i], c, -m));
        } else if (V == 6) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { float x0, x1; upk2(ffma2(pk2(r[2 * i], r[2 * i + 1]), c2, nm2), x0, x1); sink ^= __float_as_uint(x0) ^ __float_as_uint(x1); }
        } else if (V == 7) {
#pragma unroll
            for (int i = 0; i < 16; ++i) sink ^= ex2_h2(__float_as_uint(r[2 * i]) ^ __float_as_uint(m));
        }
        m += dm;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float a, b; upk2(sum2, a, b);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + a + b + mx + __uint_as_float(sink ^ hs);
}
template <int V>
void run(int warps, const char* name) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) k<V><<<148, warps * 32>>>(out, cyc, iters, 0.1275f, -3.f, 0.001f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-28s warps/SMSP=%d  cycles per warp-element per SMSP = %.2f\n", name, warps / 4, avg / ((warps / 4.0) * iters * 32));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<0>(w, "0 scalar pass"); run<1>(w, "1 packed pass"); run<2>(w, "2 packed, lagged consumers"); run<3>(w, "3 f16x2 exp pass");
        run<4>(w, "4 MUFU.EX2 only"); run<5>(w, "5 FFMA only"); run<6>(w, "6 FFMA2 only (per element)"); run<7>(w, "7 ex2.f16x2 only (per elem)");
    }
    return 0;
}
