// Microbenchmark 6: cost of the synchronisation primitives the attention pipeline uses between warps.
//   (a) mbarrier.try_wait on a phase that completed long ago (all 32 lanes / one lane + shfl)
//   (b) the same followed by tcgen05.fence::after_thread_sync
//   (c) tcgen05.fence::before_thread_sync + __syncwarp + lane-0 mbarrier.arrive
//   (d) ping-pong between two warps through two mbarriers: one-way hand-off latency
//   (e) __any_sync / __all_sync
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void wait_all(uint32_t bar, uint32_t parity) { while (!try_wait(bar, parity)) {} __all_sync(0xffffffffu, 1); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

__global__ void k(long long* out, int iters) {
    __shared__ unsigned long long bars[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(s32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive(s32(&bars[0]));     // phase 0 of bars[0] is complete from now on
    __syncthreads();
    long long t0, t1;
    unsigned acc = 0;
    if (warp == 0) {
        // (a) all lanes
        t0 = clock64();
        for (int i = 0; i < iters; ++i) { wait_all(s32(&bars[0]), 0); }
        t1 = clock64();
        if (lane == 0) out[0] = (t1 - t0) / iters;
        // (a2) one lane + shfl
        t0 = clock64();
        for (int i = 0; i < iters; ++i) { unsigned ok = 1; if (lane == 0) { while (!try_wait(s32(&bars[0]), 0)) {} } acc += __shfl_sync(0xffffffffu, ok, 0); }
        t1 = clock64();
        if (lane == 0) out[1] = (t1 - t0) / iters;
        // (b) wait + fence after
        t0 = clock64();
        for (int i = 0; i < iters; ++i) { wait_all(s32(&bars[0]), 0); fence_after(); }
        t1 = clock64();
        if (lane == 0) out[2] = (t1 - t0) / iters;
        // (c) fence before + syncwarp + arrive (on a barrier nobody waits for; count 1 -> phases just flip)
        t0 = clock64();
        for (int i = 0; i < iters; ++i) { fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(s32(&bars[1])); }
        t1 = clock64();
        if (lane == 0) out[3] = (t1 - t0) / iters;
        // (e) votes
        t0 = clock64();
        for (int i = 0; i < iters; ++i) { acc += __any_sync(0xffffffffu, (acc + i) & 1); }
        t1 = clock64();
        if (lane == 0) out[4] = (t1 - t0) / iters;
    }
    __syncthreads();
    // (d) ping-pong warp 0 <-> warp 1 (different SMSPs)
    if (warp < 2) {
        const uint32_t mine = s32(&bars[2 + warp]), other = s32(&bars[2 + (warp ^ 1)]);
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (warp == 0) { __syncwarp(); if (lane == 0) mbar_arrive(other); wait_all(mine, i & 1); }
            else { wait_all(mine, i & 1); __syncwarp(); if (lane == 0) mbar_arrive(other); }
        }
        t1 = clock64();
        if (threadIdx.x == 0) out[5] = (t1 - t0) / iters;    // round trip = 2 hand-offs
    }
    if (acc == 0x7fffffff) out[7] = acc;
}
int main() {
    long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    k<<<1, 128>>>(d, 2000); k<<<1, 128>>>(d, 2000);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("%s\ntry_wait on a completed phase, 32 lanes + vote : %lld cycles\n", cudaGetErrorString(e), h[0]);
    printf("try_wait on a completed phase, 1 lane + shfl    : %lld cycles\n", h[1]);
    printf("try_wait (32 lanes) + tcgen05.fence::after      : %lld cycles\n", h[2]);
    printf("tcgen05.fence::before + syncwarp + arrive        : %lld cycles\n", h[3]);
    printf("__any_sync                                       : %lld cycles\n", h[4]);
    printf("ping-pong round trip (2 hand-offs)               : %lld cycles\n", h[5]);
    return 0;
}
