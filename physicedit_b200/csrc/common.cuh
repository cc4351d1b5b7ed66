// Shared host-side plumbing for libpe_b200: error codes, the handle, tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/pe_b200.h"

namespace pe {

typedef __nv_bfloat16 bf16;

struct Handle {
    int device = 0;
    int sm_count = 0;
    unsigned int* abort_flag = nullptr;      // device word: set by a kernel whose pipeline timed out
    char last_error[512] = {0};
    // lazily resolved driver entry point (no link-time dependency on libcuda)
    void* encode_tiled = nullptr;
    // scratch owned by the handle (tile counters, split reductions ...)
    void* workspace = nullptr;
    size_t workspace_bytes = 0;
    int attn_seq = 0;                        // launch counter of the half-row attention kernel (names its overflow flags)
};

int set_error(Handle* h, int code, const char* fmt, ...);

#define PE_CHECK_CUDA(h, expr)                                                                   \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return pe::set_error((h), PE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);                    \
    } while (0)

#define PE_REQUIRE(h, cond, ...)                                                                 \
    do {                                                                                         \
        if (!(cond)) return pe::set_error((h), PE_ERR_INVALID_ARGUMENT, __VA_ARGS__);            \
    } while (0)

// Encode a 2-D bf16 row-major tensor map with 128-byte swizzle.
//   rows x cols matrix, `ld` elements between rows, box = box_rows x 64 columns.
int make_tmap_2d(Handle* h, CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols = 64);

// 3-D bf16 tensor map over an NHWC activation map: dims (C, W, H), `ld_w` / `ld_h` elements between pixels / rows,
// box = box_c x box_w x box_h, 128-byte swizzle (box_c = 64), out-of-bounds elements read as zero.
int make_tmap_3d(Handle* h, CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t ld_w, uint64_t ld_h,
                 uint32_t box_c, uint32_t box_w, uint32_t box_h);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace pe
