// Microbenchmark: throughput of the softmax inner pass (FFMA -> MUFU.EX2 -> FADD -> F2FP.BF16 pack) per SM sub-partition,
// for 1 / 2 / 4 warps per SMSP and with parts of the mix removed.  nvcc -arch=sm_100a -O3 exp_pass.cu -o exp_pass
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<unsigned*>(&v); }
template <int MODE>   // 0 full, 1 no MUFU, 2 no FADD, 3 no pack, 4 MUFU only
__global__ void k(float* out, long long* cyc, int iters, float c, float m) {
    float r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = (float)((threadIdx.x * 7 + i * 3) & 15) - 20.f;
    float s0 = 0.f, s1 = 0.f;
    unsigned sink = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float x0 = MODE == 4 ? r[2 * i] : fmaf(r[2 * i], c, -m), x1 = MODE == 4 ? r[2 * i + 1] : fmaf(r[2 * i + 1], c, -m);
            float p0 = MODE == 1 ? x0 : ex2(x0), p1 = MODE == 1 ? x1 : ex2(x1);
            if (MODE != 2 && MODE != 4) { s0 += p0; s1 += p1; }
            if (MODE != 3 && MODE != 4) sink ^= pack(p0, p1); else sink ^= __float_as_uint(p0) ^ __float_as_uint(p1);
        }
        m += 1e-7f;
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] += 1e-7f * (float)(sink & 1);   // keep a loop-carried dependence so nothing is hoisted
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + __uint_as_float(sink);
}
template <int MODE>
void run(const char* name, int warps) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    k<MODE><<<148, warps * 32>>>(out, cyc, iters, 0.1275f, -3.f);
    k<MODE><<<148, warps * 32>>>(out, cyc, iters, 0.1275f, -3.f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // warp-elements processed per SMSP = (warps/4) * iters * 32
    printf("%-10s warps/SMSP=%d  cycles per warp-element per SMSP = %.2f\n", name, warps / 4, avg / ((warps / 4.0) * iters * 32));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<0>("full", w); run<1>("no_mufu", w); run<2>("no_fadd", w); run<3>("no_pack", w); run<4>("mufu_only", w);
    }
    return 0;
}
