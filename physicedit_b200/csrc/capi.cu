// C ABI of libpe_b200 (see include/pe_b200.h): handle life cycle, error reporting, tensor-map
// encoding, and the extern "C" entry points that forward to the kernels' launchers.
#include <stdarg.h>
#include <new>
#include "common.cuh"

namespace pe {

int set_error(Handle* h, int code, const char* fmt, ...) {
    if (h != nullptr) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->last_error, sizeof(h->last_error), fmt, ap);
        va_end(ap);
    }
    return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_2d(Handle* h, CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols) {
    if (h->encode_tiled == nullptr) return set_error(h, PE_ERR_NOT_INITIALIZED, "cuTensorMapEncodeTiled not resolved");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * sizeof(bf16)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(h->encode_tiled)(
        out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(h, PE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu box=%ux%u",
                         (int)r, base, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    return PE_OK;
}

int make_tmap_3d(Handle* h, CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t ld_w, uint64_t ld_h,
                 uint32_t box_c, uint32_t box_w, uint32_t box_h) {
    if (h->encode_tiled == nullptr) return set_error(h, PE_ERR_NOT_INITIALIZED, "cuTensorMapEncodeTiled not resolved");
    cuuint64_t dims[3] = {C, W, H};
    cuuint64_t strides[2] = {ld_w * sizeof(bf16), ld_h * sizeof(bf16)};
    cuuint32_t box[3] = {box_c, box_w, box_h};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(h->encode_tiled)(
        out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(h, PE_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d): base=%p C=%llu W=%llu H=%llu ld_w=%llu box=%ux%ux%u",
                         (int)r, base, (unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)ld_w,
                         box_c, box_w, box_h);
    return PE_OK;
}

// launchers implemented in the kernel translation units
int decode_attention_fused_run(Handle* h, const pe_decode_req* reqs, int n_req, int Hq, int Hkv, int D, int64_t ldc, const float* cs,
                               const float* sn, float scale, cudaStream_t s);
int gemm_run(Handle* h, const pe_gemm_seg* segs, int nseg, int N, int K, int epilogue, int flags, cudaStream_t stream, const pe_gemm_batch* bt = nullptr);
int attention_run(Handle* h, const void* q, const void* k, const void* v, void* o, int S, int H, int64_t ld, float scale,
                  int flags, cudaStream_t stream, float* lse = nullptr);
int attention_routed_run(Handle* h, const void* q, const void* k, const void* v, int S, int H, int64_t ld, float scale, int flags, int n_route,
                         const int32_t* route_end, void* const* o_route, int64_t ldo, cudaStream_t stream);
int layernorm_modulate_run(Handle* h, const void* x, void* out, int rows, int C, const void* shift, const void* ops, cudaStream_t s);
int layernorm_modulate2_run(Handle* h, const void* x, void* out, int rows, int C, int split_row, const void* shift0, const void* ops0,
                            const void* shift1, const void* ops1, cudaStream_t s);
int rmsnorm_run(Handle* h, const void* x, void* out, int rows, int C, const void* w, float eps, cudaStream_t s);
int gemv_run(Handle* h, const void* x, const void* w, const void* bias, void* y, int batch, int N, int K, int act_in,
             int act_out, const uint8_t* one_plus_mask, cudaStream_t s, const void* norm_w = nullptr, float norm_eps = 0.f,
             const void* residual = nullptr);
int act_run(Handle* h, const void* x, void* y, long long n, int act, cudaStream_t s);
int timestep_embedding_run(Handle* h, const void* t_in, void* out, int raw, cudaStream_t s);
int patchify_run(Handle* h, const void* latents, void* tokens, int H8, int W8, cudaStream_t s);
int unpatchify_run(Handle* h, const void* tokens, int64_t ld, void* latents, int H8, int W8, cudaStream_t s);
int cfg_euler_run(Handle* h, void* latents, const void* posi, const void* nega, int64_t n, float cfg, float dsigma, cudaStream_t s);
int special_gather_run(Handle* h, const void* prompt_emb, const uint8_t* mask, int T, int C, void* dst, int32_t* idx,
                       int max_rows, cudaStream_t s);
int special_blend_scatter_run(Handle* h, void* prompt_emb, const int32_t* idx, int max_rows, int C, const void* pd,
                              const void* pv, const void* t_in, float t_min, float t_max, cudaStream_t s);
int small_attention_run(Handle* h, const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Skv, int D,
                        int64_t ldq, int64_t ldkv, int64_t ldo, float scale, cudaStream_t s);
int layernorm_affine_run(Handle* h, const void* x, void* out, int rows, int C, const void* w, const void* b, float eps, cudaStream_t s);
int add_bias_rows_run(Handle* h, void* x, const void* add, int rows, int C, int period, float alpha, cudaStream_t s);
int conv2d_run(Handle* h, const pe_conv2d_desc* d, int epilogue, cudaStream_t stream);
int channel_rmsnorm_run(Handle* h, const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int C, const void* gamma, int act,
                        cudaStream_t s);
int upsample2x_run(Handle* h, const void* in, void* out, int H, int W, int C, cudaStream_t s);
int space_to_depth_run(Handle* h, const void* in, void* out, int H, int W, int C, cudaStream_t s);
int nchw_to_nhwc_run(Handle* h, const void* src, void* dst, int64_t ldd, int C, int64_t HW, int op, const void* p0, const void* p1, cudaStream_t s);
int nhwc_to_nchw_run(Handle* h, const void* src, int64_t lds, void* dst, int C, int64_t HW, int op, const void* p0, const void* p1, cudaStream_t s);
int transpose_run(Handle* h, const void* src, int64_t lds, void* dst, int64_t ldd, int R, int C, cudaStream_t s);
int softmax_rows_run(Handle* h, const void* scores, int64_t lds, void* probs, int64_t ldp, int rows, int n, int n_pad, float scale, cudaStream_t s,
                     const void* mask = nullptr, int64_t ldm = 0, int mask_period = 0);
int attention_bwd_delta_run(Handle* h, const void* d_o, int64_t ldd, const void* o, int64_t ldo, int S, int H, void* delta, int64_t ld_delta, cudaStream_t s);
int swiglu_run(Handle* h, const void* x, int64_t ldx, void* out, int64_t ldo, int rows, int I, cudaStream_t s);
int rope_half_run(Handle* h, void* x, int64_t ldx, int T, int H, int D, const float* cs, const float* sn, const int* row_ptr, int row0, int mode,
                  cudaStream_t s);
int range_attention_run(Handle* h, const void* q, const void* k, const void* v, void* o, int H, int Hkv, int Sq, int Skv, int D, int64_t ldq,
                        int64_t ldkv, int64_t ldo, float scale, const int* kv_lo, const int* kv_hi, const int* kv_len_ptr, cudaStream_t s);
int gather_rows_run(Handle* h, const void* table, int64_t ldt, const int64_t* ids, void* out, int64_t ldo, int n, int C, cudaStream_t s);
int argmax_run(Handle* h, const void* x, int n, int64_t* out, int64_t* log, const int* log_pos, cudaStream_t s);
int kv_append_run(Handle* h, const void* k_new, const void* v_new, void* cache_k, void* cache_v, int64_t ldc, int C, const int* pos, cudaStream_t s);
int advance_run(Handle* h, int* pos, int n, cudaStream_t s);
int rope_kv_append_run(Handle* h, void* qkv, int Hq, int Hkv, int D, const float* cs, const float* sn, void* cache_k, void* cache_v, int64_t ldc,
                       const int* ctr, cudaStream_t s);

}  // namespace pe

using pe::Handle;

extern "C" {

int pe_abi_version(void) { return PE_B200_ABI_VERSION; }

int pe_create(pe_handle_t* out, int device) {
    if (out == nullptr) return PE_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return PE_ERR_UNSUPPORTED_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PE_ERR_CUDA;
    if (prop.major != 10) return PE_ERR_UNSUPPORTED_DEVICE;   // sm_100a SASS only; no fallback path exists
    if (cudaSetDevice(device) != cudaSuccess) return PE_ERR_CUDA;
    Handle* h = new (std::nothrow) Handle();
    if (h == nullptr) return PE_ERR_OUT_OF_MEMORY;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &h->encode_tiled, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || h->encode_tiled == nullptr) {
        delete h;
        return PE_ERR_CUDA;
    }
    if (cudaMalloc(&h->abort_flag, 256) != cudaSuccess || cudaMemset(h->abort_flag, 0, 256) != cudaSuccess) {
        delete h;
        return PE_ERR_OUT_OF_MEMORY;
    }
    h->workspace_bytes = 1 << 20;
    if (cudaMalloc(&h->workspace, h->workspace_bytes) != cudaSuccess || cudaMemset(h->workspace, 0, h->workspace_bytes) != cudaSuccess) {
        cudaFree(h->abort_flag);
        delete h;
        return PE_ERR_OUT_OF_MEMORY;
    }
    *out = reinterpret_cast<pe_handle_t>(h);
    return PE_OK;
}

int pe_destroy(pe_handle_t hh) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    if (h == nullptr) return PE_ERR_INVALID_ARGUMENT;
    if (h->abort_flag) cudaFree(h->abort_flag);
    if (h->workspace) cudaFree(h->workspace);
    delete h;
    return PE_OK;
}

const char* pe_last_error(pe_handle_t hh) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    return h ? h->last_error : "null handle";
}

int pe_sm_count(pe_handle_t hh) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    return h ? h->sm_count : PE_ERR_INVALID_ARGUMENT;
}

int pe_workspace(pe_handle_t hh, void** ptr, size_t* bytes) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    if (h == nullptr || ptr == nullptr || bytes == nullptr) return PE_ERR_INVALID_ARGUMENT;
    *ptr = h->workspace;
    *bytes = h->workspace_bytes;
    return PE_OK;
}

int pe_check_async_error(pe_handle_t hh, void* stream, unsigned int* diag) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    if (h == nullptr) return PE_ERR_INVALID_ARGUMENT;
    PE_CHECK_CUDA(h, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    unsigned int w[2] = {0, 0};       // [0] pipeline time-out (site, block), [1] data-dependent argument error (special-token count)
    PE_CHECK_CUDA(h, cudaMemcpy(w, h->abort_flag, sizeof(w), cudaMemcpyDeviceToHost));
    const unsigned int v = w[0];
    if (diag) *diag = v != 0 ? v : w[1];
    if (v != 0 || w[1] != 0) cudaMemset(h->abort_flag, 0, sizeof(w));
    if (v != 0)
        return pe::set_error(h, PE_ERR_KERNEL_TIMEOUT, "kernel pipeline wait timed out: site=%u block=%u", (v >> 16) & 0x7fffu, v & 0xffffu);
    if (w[1] != 0)
        return pe::set_error(h, PE_ERR_INVALID_ARGUMENT, "pe_special_gather: the mask selects %u rows but the destination holds fewer "
                             "(rows beyond it were NOT processed)", w[1]);
    return PE_OK;
}

#define PE_H(hh)                                        \
    Handle* h = reinterpret_cast<Handle*>(hh);          \
    if (h == nullptr) return PE_ERR_INVALID_ARGUMENT;

int pe_gemm(pe_handle_t hh, const pe_gemm_seg* segs, int nseg, int N, int K, int epilogue, int flags, void* stream) {
    PE_H(hh);
    PE_REQUIRE(h, segs != nullptr, "pe_gemm: segs is null");
    return pe::gemm_run(h, segs, nseg, N, K, epilogue, flags, static_cast<cudaStream_t>(stream));
}

int pe_gemm_batched(pe_handle_t hh, const pe_gemm_seg* seg, const pe_gemm_batch* batch, int N, int K, int epilogue, int flags, void* stream) {
    PE_H(hh);
    PE_REQUIRE(h, seg != nullptr && batch != nullptr, "pe_gemm_batched: null descriptor");
    return pe::gemm_run(h, seg, 1, N, K, epilogue, flags, static_cast<cudaStream_t>(stream), batch);
}

int pe_attention_fwd_lse(pe_handle_t hh, const void* q, const void* k, const void* v, void* o, int S, int H, int64_t ld, float scale, int flags,
                         float* lse, void* stream) {
    PE_H(hh);
    PE_REQUIRE(h, lse != nullptr, "pe_attention_fwd_lse: lse is null");
    return pe::attention_run(h, q, k, v, o, S, H, ld, scale, flags, static_cast<cudaStream_t>(stream), lse);
}

int pe_attention_fwd(pe_handle_t hh, const void* q, const void* k, const void* v, void* o, int S, int H, int64_t ld,
                     float scale, int flags, void* stream) {
    PE_H(hh);
    return pe::attention_run(h, q, k, v, o, S, H, ld, scale, flags, static_cast<cudaStream_t>(stream));
}

int pe_attention_fwd_routed(pe_handle_t hh, const void* q, const void* k, const void* v, int S, int H, int64_t ld, float scale, int flags,
                            int n_route, const int32_t* route_end, void* const* o_route, int64_t ldo, void* stream) {
    PE_H(hh);
    return pe::attention_routed_run(h, q, k, v, S, H, ld, scale, flags, n_route, route_end, o_route, ldo, static_cast<cudaStream_t>(stream));
}

int pe_small_attention(pe_handle_t hh, const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Skv,
                       int D, int64_t ldq, int64_t ldkv, int64_t ldo, float scale, void* stream) {
    PE_H(hh);
    return pe::small_attention_run(h, q, k, v, o, B, H, Sq, Skv, D, ldq, ldkv, ldo, scale, static_cast<cudaStream_t>(stream));
}

int pe_layernorm_modulate(pe_handle_t hh, const void* x, void* out, int rows, int C, const void* shift,
                          const void* one_plus_scale, void* stream) {
    PE_H(hh);
    return pe::layernorm_modulate_run(h, x, out, rows, C, shift, one_plus_scale, static_cast<cudaStream_t>(stream));
}

int pe_layernorm_modulate2(pe_handle_t hh, const void* x, void* out, int rows, int C, int split_row, const void* shift0,
                           const void* one_plus_scale0, const void* shift1, const void* one_plus_scale1, void* stream) {
    PE_H(hh);
    return pe::layernorm_modulate2_run(h, x, out, rows, C, split_row, shift0, one_plus_scale0, shift1, one_plus_scale1,
                                       static_cast<cudaStream_t>(stream));
}

int pe_layernorm(pe_handle_t hh, const void* x, void* out, int rows, int C, const void* w, const void* b, float eps, void* stream) {
    PE_H(hh);
    return pe::layernorm_affine_run(h, x, out, rows, C, w, b, eps, static_cast<cudaStream_t>(stream));
}

int pe_rmsnorm(pe_handle_t hh, const void* x, void* out, int rows, int C, const void* w, float eps, void* stream) {
    PE_H(hh);
    return pe::rmsnorm_run(h, x, out, rows, C, w, eps, static_cast<cudaStream_t>(stream));
}

int pe_add_rows(pe_handle_t hh, void* x, const void* add, int rows, int C, int period, float alpha, void* stream) {
    PE_H(hh);
    return pe::add_bias_rows_run(h, x, add, rows, C, period, alpha, static_cast<cudaStream_t>(stream));
}

int pe_gemv(pe_handle_t hh, const void* x, const void* w, const void* bias, void* y, int batch, int N, int K, int act_in,
            int act_out, const uint8_t* one_plus_mask, void* stream) {
    PE_H(hh);
    return pe::gemv_run(h, x, w, bias, y, batch, N, K, act_in, act_out, one_plus_mask, static_cast<cudaStream_t>(stream));
}

int pe_gemv_fused(pe_handle_t hh, const void* x, const void* w, const void* bias, void* y, int batch, int N, int K, int act_in,
                  const void* norm_w, float norm_eps, const void* residual, void* stream) {
    PE_H(hh);
    return pe::gemv_run(h, x, w, bias, y, batch, N, K, act_in, 0, nullptr, static_cast<cudaStream_t>(stream), norm_w, norm_eps, residual);
}

int pe_decode_attention_fused(pe_handle_t hh, const pe_decode_req* reqs, int n_req, int Hq, int Hkv, int D, int64_t ldc,
                              const float* cos_table, const float* sin_table, float scale, void* stream) {
    PE_H(hh);
    return pe::decode_attention_fused_run(h, reqs, n_req, Hq, Hkv, D, ldc, cos_table, sin_table, scale, static_cast<cudaStream_t>(stream));
}

int pe_gemv_swiglu(pe_handle_t hh, const void* x, const void* w, const void* bias, void* y, int batch, int I, int K, const void* norm_w, float norm_eps,
                   void* stream) {
    PE_H(hh);
    PE_REQUIRE(h, I > 0, "pe_gemv_swiglu: I must be positive");
    return pe::gemv_run(h, x, w, bias, y, batch, 2 * I, K, 0, 2, nullptr, static_cast<cudaStream_t>(stream), norm_w, norm_eps, nullptr);
}

int pe_act(pe_handle_t hh, const void* x, void* y, int64_t n, int act, void* stream) {
    PE_H(hh);
    return pe::act_run(h, x, y, (long long)n, act, static_cast<cudaStream_t>(stream));
}

int pe_timestep_embedding(pe_handle_t hh, const void* t_in, void* out, int raw, void* stream) {
    PE_H(hh);
    return pe::timestep_embedding_run(h, t_in, out, raw, static_cast<cudaStream_t>(stream));
}

int pe_patchify(pe_handle_t hh, const void* latents, void* tokens, int H8, int W8, void* stream) {
    PE_H(hh);
    return pe::patchify_run(h, latents, tokens, H8, W8, static_cast<cudaStream_t>(stream));
}

int pe_unpatchify(pe_handle_t hh, const void* tokens, int64_t ld, void* latents, int H8, int W8, void* stream) {
    PE_H(hh);
    return pe::unpatchify_run(h, tokens, ld, latents, H8, W8, static_cast<cudaStream_t>(stream));
}

int pe_cfg_euler_step(pe_handle_t hh, void* latents, const void* posi, const void* nega, int64_t n, float cfg_scale,
                      float dsigma, void* stream) {
    PE_H(hh);
    return pe::cfg_euler_run(h, latents, posi, nega, n, cfg_scale, dsigma, static_cast<cudaStream_t>(stream));
}

int pe_special_gather(pe_handle_t hh, const void* prompt_emb, const uint8_t* mask, int T, int C, void* dst, int32_t* idx,
                      int max_rows, void* stream) {
    PE_H(hh);
    return pe::special_gather_run(h, prompt_emb, mask, T, C, dst, idx, max_rows, static_cast<cudaStream_t>(stream));
}

int pe_special_blend_scatter(pe_handle_t hh, void* prompt_emb, const int32_t* idx, int max_rows, int C, const void* pred_dino,
                             const void* pred_vae, const void* t_in, float t_min, float t_max, void* stream) {
    PE_H(hh);
    return pe::special_blend_scatter_run(h, prompt_emb, idx, max_rows, C, pred_dino, pred_vae, t_in, t_min, t_max,
                                         static_cast<cudaStream_t>(stream));
}

int pe_conv2d(pe_handle_t hh, const pe_conv2d_desc* desc, int epilogue, void* stream) {
    PE_H(hh);
    PE_REQUIRE(h, desc != nullptr, "pe_conv2d: desc is null");
    return pe::conv2d_run(h, desc, epilogue, static_cast<cudaStream_t>(stream));
}

int pe_channel_rmsnorm(pe_handle_t hh, const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int C, const void* gamma, int act,
                       void* stream) {
    PE_H(hh);
    return pe::channel_rmsnorm_run(h, x, ldx, out, ldo, rows, C, gamma, act, static_cast<cudaStream_t>(stream));
}

int pe_upsample2x(pe_handle_t hh, const void* in, void* out, int H, int W, int C, void* stream) {
    PE_H(hh);
    return pe::upsample2x_run(h, in, out, H, W, C, static_cast<cudaStream_t>(stream));
}

int pe_space_to_depth(pe_handle_t hh, const void* in, void* out, int H, int W, int C, void* stream) {
    PE_H(hh);
    return pe::space_to_depth_run(h, in, out, H, W, C, static_cast<cudaStream_t>(stream));
}

int pe_nchw_to_nhwc(pe_handle_t hh, const void* src, void* dst, int64_t ldd, int C, int64_t HW, int op, const void* p0, const void* p1,
                    void* stream) {
    PE_H(hh);
    return pe::nchw_to_nhwc_run(h, src, dst, ldd, C, HW, op, p0, p1, static_cast<cudaStream_t>(stream));
}

int pe_nhwc_to_nchw(pe_handle_t hh, const void* src, int64_t lds, void* dst, int C, int64_t HW, int op, const void* p0, const void* p1,
                    void* stream) {
    PE_H(hh);
    return pe::nhwc_to_nchw_run(h, src, lds, dst, C, HW, op, p0, p1, static_cast<cudaStream_t>(stream));
}

int pe_transpose(pe_handle_t hh, const void* src, int64_t lds, void* dst, int64_t ldd, int R, int C, void* stream) {
    PE_H(hh);
    return pe::transpose_run(h, src, lds, dst, ldd, R, C, static_cast<cudaStream_t>(stream));
}

int pe_softmax_rows(pe_handle_t hh, const void* scores, int64_t lds, void* probs, int64_t ldp, int rows, int n, int n_pad, float scale,
                    void* stream) {
    PE_H(hh);
    return pe::softmax_rows_run(h, scores, lds, probs, ldp, rows, n, n_pad, scale, static_cast<cudaStream_t>(stream));
}

int pe_attention_bwd_delta(pe_handle_t hh, const void* d_o, int64_t ldd, const void* o, int64_t ldo, int S, int H, void* delta, int64_t ld_delta,
                           void* stream) {
    PE_H(hh);
    return pe::attention_bwd_delta_run(h, d_o, ldd, o, ldo, S, H, delta, ld_delta, static_cast<cudaStream_t>(stream));
}

int pe_softmax_rows_masked(pe_handle_t hh, const void* scores, int64_t lds, void* probs, int64_t ldp, int rows, int n, int n_pad, float scale,
                           const void* mask, int64_t ldm, int mask_period, void* stream) {
    PE_H(hh);
    PE_REQUIRE(h, mask != nullptr, "pe_softmax_rows_masked: mask is null");
    return pe::softmax_rows_run(h, scores, lds, probs, ldp, rows, n, n_pad, scale, static_cast<cudaStream_t>(stream), mask, ldm, mask_period);
}

int pe_swiglu(pe_handle_t hh, const void* x, int64_t ldx, void* out, int64_t ldo, int rows, int I, void* stream) {
    PE_H(hh);
    return pe::swiglu_run(h, x, ldx, out, ldo, rows, I, static_cast<cudaStream_t>(stream));
}

int pe_rope_half(pe_handle_t hh, void* x, int64_t ldx, int T, int H, int D, const float* cos_table, const float* sin_table, const int32_t* row_ptr,
                 int row0, int mode, void* stream) {
    PE_H(hh);
    return pe::rope_half_run(h, x, ldx, T, H, D, cos_table, sin_table, row_ptr, row0, mode, static_cast<cudaStream_t>(stream));
}

int pe_range_attention(pe_handle_t hh, const void* q, const void* k, const void* v, void* o, int H, int Hkv, int Sq, int Skv, int D, int64_t ldq,
                       int64_t ldkv, int64_t ldo, float scale, const int32_t* kv_lo, const int32_t* kv_hi, const int32_t* kv_len_ptr, void* stream) {
    PE_H(hh);
    return pe::range_attention_run(h, q, k, v, o, H, Hkv, Sq, Skv, D, ldq, ldkv, ldo, scale, kv_lo, kv_hi, kv_len_ptr, static_cast<cudaStream_t>(stream));
}

int pe_gather_rows(pe_handle_t hh, const void* table, int64_t ldt, const int64_t* ids, void* out, int64_t ldo, int n, int C, void* stream) {
    PE_H(hh);
    return pe::gather_rows_run(h, table, ldt, ids, out, ldo, n, C, static_cast<cudaStream_t>(stream));
}

int pe_argmax(pe_handle_t hh, const void* x, int n, int64_t* out, int64_t* log, const int32_t* log_pos, void* stream) {
    PE_H(hh);
    return pe::argmax_run(h, x, n, out, log, log_pos, static_cast<cudaStream_t>(stream));
}

int pe_kv_append(pe_handle_t hh, const void* k_new, const void* v_new, void* cache_k, void* cache_v, int64_t ldc, int C, const int32_t* pos,
                 void* stream) {
    PE_H(hh);
    return pe::kv_append_run(h, k_new, v_new, cache_k, cache_v, ldc, C, pos, static_cast<cudaStream_t>(stream));
}

int pe_rope_kv_append(pe_handle_t hh, void* qkv, int Hq, int Hkv, int D, const float* cos_table, const float* sin_table, void* cache_k, void* cache_v,
                      int64_t ldc, const int32_t* counters, void* stream) {
    PE_H(hh);
    return pe::rope_kv_append_run(h, qkv, Hq, Hkv, D, cos_table, sin_table, cache_k, cache_v, ldc, counters, static_cast<cudaStream_t>(stream));
}

int pe_advance(pe_handle_t hh, int32_t* counters, int n, void* stream) {
    PE_H(hh);
    return pe::advance_run(h, counters, n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
