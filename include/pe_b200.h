/*
 * libpe_b200 -- C ABI of the B200-native PhysicEdit / Qwen-Image-Edit DiT hot path.
 *
 * The reference (liangbingzhao/PhysicEdit @ 5f239b9) is pure Python/PyTorch: its "FFI" for this path
 * is the set of ATen/cuBLAS/SDPA library calls made from
 *     DiffSynth-Studio/diffsynth/pipelines/qwen_image_physical.py:1302-1403  (model_fn_qwen_image)
 *     DiffSynth-Studio/diffsynth/models/qwen_image_dit.py:14-57,228-401      (attention, MLP, block)
 *     DiffSynth-Studio/diffsynth/models/utils.py:189-309                     (timestep emb, RMSNorm, AdaLN)
 *     DiffSynth-Studio/diffsynth/pipelines/helpers.py:123-164                (VisualThinkingDualAdapter)
 *     DiffSynth-Studio/diffsynth/schedulers/flow_match.py:72-82              (Euler step)
 * Every entry point below names the reference call site(s) it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  All tensor pointers are DEVICE pointers owned by the
 *     caller; `stream` is a cudaStream_t passed as void*.  The library never allocates tensors and
 *     never synchronises the stream: calls enqueue work and return.
 *   - every function returns 0 (PE_OK) or a negative pe_status; pe_last_error(h) gives the text.
 *     Nothing throws across the ABI.  There is no CPU fallback: without a CUDA device of compute
 *     capability 10.x pe_create fails with PE_ERR_UNSUPPORTED_DEVICE.
 *   - activations / weights are bf16 (row-major, innermost stride 1) unless a parameter says otherwise;
 *     accumulation is fp32.  Rounding points follow SURVEY.md Appendix B.
 *   - one handle per device per host thread; functions are re-entrant across handles.
 */
#ifndef PE_B200_H_
#define PE_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PE_B200_ABI_VERSION 2

typedef enum pe_status {
    PE_OK = 0,
    PE_ERR_INVALID_ARGUMENT = -1,
    PE_ERR_CUDA = -2,
    PE_ERR_UNSUPPORTED_DEVICE = -3,
    PE_ERR_KERNEL_TIMEOUT = -4,   /* a bounded in-kernel wait expired (see pe_check_async_error) */
    PE_ERR_OUT_OF_MEMORY = -5,
    PE_ERR_NOT_INITIALIZED = -6
} pe_status;

typedef struct pe_handle_s* pe_handle_t;

/* ------------------------------------------------------------------------------------------- */
/* life cycle                                                                                   */
/* ------------------------------------------------------------------------------------------- */
int pe_abi_version(void);
/* device: CUDA ordinal.  Fails (PE_ERR_UNSUPPORTED_DEVICE) unless the device is sm_100. */
int pe_create(pe_handle_t* out, int device);
int pe_destroy(pe_handle_t h);
const char* pe_last_error(pe_handle_t h);
/* Synchronises `stream` and reports whether any kernel since the last check hit a bounded-wait
 * timeout (PE_ERR_KERNEL_TIMEOUT; diagnostic word in *diag if non-NULL).  Test/debug helper. */
int pe_check_async_error(pe_handle_t h, void* stream, unsigned int* diag);
/* handle-owned device scratch (1 MiB): diagnostics of trace builds (-DPE_ATTN_TRACE) land here. */
int pe_workspace(pe_handle_t h, void** ptr, size_t* bytes);
/* number of SMs of the handle's device (grid sizing information for callers / benchmarks). */
int pe_sm_count(pe_handle_t h);

/* ------------------------------------------------------------------------------------------- */
/* grouped linear layer with fused epilogue  (tcgen05 / TMEM / TMA)                             */
/*   replaces F.linear at qwen_image_dit.py:45,240,259-264,312-313 and the elementwise ops      */
/*   around them (ApproximateGELU :47, gate*x+residual :386-399, RMSNorm utils.py:250-257,       */
/*   apply_rotary_emb_qwen :51-57, torch.cat :304-306).                                          */
/* ------------------------------------------------------------------------------------------- */
typedef enum pe_epilogue {
    PE_EPI_BIAS = 0,           /* out = bf16(acc + bias)                                         */
    PE_EPI_BIAS_GELU_SIGMOID = 2, /* h=bf16(acc+bias); out = h * sigmoid(1.702 h)  (ApproximateGELU) */
    PE_EPI_BIAS_GELU_ERF = 3,  /* h=bf16(acc+bias); out = gelu_erf(h)   (nn.GELU, helpers.py:127) */
    PE_EPI_GATE_RESIDUAL = 4,  /* o=bf16(acc+bias); out = residual + gate[n]*o  (in place on out) */
    PE_EPI_QKV_NORM_ROPE = 5,  /* N = 3*H*128: per-head RMSNorm(q,k)*w, RoPE(q,k), v passthrough;
                                  writes q/k/v into three [M, H*128] buffers                      */
    PE_EPI_BIAS_SILU = 6,      /* out = silu(bf16(acc + bias))   (timestep MLP)                   */
    PE_EPI_F32 = 7,            /* out = acc as float [M, ldo] (ldo in floats; no bias): attention scores of the VAE mid block */
    PE_EPI_ATTN_P = 8,         /* pe_gemm_batched only: out = bf16(exp2(acc * alpha - vec[i]))          (attention backward: P from Q K^T)  */
    PE_EPI_ATTN_DS = 9         /* pe_gemm_batched only: out = bf16(out * (acc - vec[i]) * alpha), in place (attention backward: dS from dO V^T) */
} pe_epilogue;

/* One segment (= token stream with its own weights) of a grouped GEMM. */
typedef struct pe_gemm_seg {
    const void* a;        /* bf16 [M, K], row stride lda elements                                 */
    int64_t lda;
    const void* w;        /* bf16 [N, K] contiguous (nn.Linear.weight layout)                     */
    const void* bias;     /* bf16 [N] or NULL                                                     */
    void* out;            /* bf16 [M, N] row stride ldo; for QKV: q buffer                        */
    int64_t ldo;
    int32_t M;
    int32_t _pad0;
    const void* gate;     /* PE_EPI_GATE_RESIDUAL: bf16 [N]                                       */
    void* out_k;          /* PE_EPI_QKV_NORM_ROPE: k buffer, same ldo                             */
    void* out_v;          /* PE_EPI_QKV_NORM_ROPE: v buffer, same ldo                             */
    const void* norm_q_w; /* PE_EPI_QKV_NORM_ROPE: bf16 [128]                                     */
    const void* norm_k_w; /* PE_EPI_QKV_NORM_ROPE: bf16 [128]                                     */
    const void* rope;     /* PE_EPI_QKV_NORM_ROPE: float2 (cos,sin) [M, 64]                       */
    /* PE_EPI_QKV_NORM_ROPE, head-parallel (Ulysses) mode: route_ranks > 0 splits the heads into route_ranks equal groups; group g's q / k / v
     * go to q_route[g] / k_route[g] / v_route[g] (buffers [*, heads/route_ranks * 128], row stride ldo, already offset to this segment's first
     * row) instead of out / out_k / out_v.  The pointers may be PEER-GPU memory mapped into this process (NVLink P2P stores straight from the
     * epilogue: the all-to-all of a sequence-parallel attention costs no extra pass). */
    void* q_route[8];
    void* k_route[8];
    void* v_route[8];
    int32_t route_ranks;
    int32_t _pad1;
} pe_gemm_seg;

#define PE_GEMM_FLAG_CTA_PAIR 1   /* use cta_group::2 (256-row tiles on an SM pair) */
#define PE_GEMM_FLAG_TRIM_N   2   /* narrow outputs (N % 256 != 0): issue the last n-tile's MMAs with N = round_up(N - n0, 16) */

int pe_gemm(pe_handle_t h, const pe_gemm_seg* segs, int nseg, int N, int K, int epilogue, int flags, void* stream);

/* `batch` independent products of one shape in ONE launch (training path, SURVEY 8f3: the per-head matrix products of the attention backward that
 * the reference gets from autograd through F.scaled_dot_product_attention, models/qwen_image_dit.py:14-39).  seg->a / seg->w / seg->out are the
 * flattened operands of all problems: problem b reads A rows [b * a_batch_rows, + M), W rows [b * w_batch_rows, + N) and writes out rows
 * [b * out_batch_rows, + M); seg->bias must be null.  Epilogues: PE_EPI_BIAS (plain bf16 store), PE_EPI_F32, and the two attention-backward
 * ones, which take a per-problem fp32 statistic `vec + b * vec_batch_stride` indexed by the output row (vec_per_column = 0) or column (= 1):
 *   PE_EPI_ATTN_P : out = bf16(exp2(acc * alpha - vec[i]))   -- P (or P^T) from the scores, vec = log2-domain LSE of pe_attention_fwd_lse
 *   PE_EPI_ATTN_DS: out = bf16(out * (acc - vec[i]) * alpha) -- dS (or dS^T) in place over P, vec = delta = rowsum(dO * O)                      */
typedef struct pe_gemm_batch {
    int32_t batch;
    int32_t vec_per_column;
    int64_t a_batch_rows;
    int64_t w_batch_rows;
    int64_t out_batch_rows;
    const float* vec;
    int64_t vec_batch_stride;
    float alpha;
    int32_t _pad;
} pe_gemm_batch;
int pe_gemm_batched(pe_handle_t h, const pe_gemm_seg* seg, const pe_gemm_batch* batch, int N, int K, int epilogue, int flags, void* stream);

/* ------------------------------------------------------------------------------------------- */
/* joint (text+image) non-causal attention, head dim 128                                        */
/*   replaces qwen_image_flash_attention / F.scaled_dot_product_attention (qwen_image_dit.py:37) */
/*   q,k,v,o: bf16 [S, H*128] token-major (row stride ld elements); scale = 1/sqrt(128).         */
/* ------------------------------------------------------------------------------------------- */
#define PE_ATTN_FLAG_SINGLE_Q_TILE 1  /* one 128-row query tile per CTA instead of two ping-ponged tiles  */
#define PE_ATTN_FLAG_P_VIA_SMEM    2  /* stage P through shared memory (SS MMA) instead of TMEM (TS MMA)   */
#define PE_ATTN_FLAG_SPLIT_ROW_SOFTMAX 8 /* attention_kernel2: a query row is shared by two threads (64 kv columns each), exact row
                                            max every step; r1: same speed isolated, 2 % slower inside the denoise loop -> not the default */
#define PE_ATTN_FLAG_KV64 16           /* attention_kernel3: 64-row KV steps with two S buffers per query tile in TMEM, so S(j+2) is issued
                                            one step ahead and the softmax never waits on the PV -> S latency chain */
#define PE_ATTN_FLAG_HALF_ROW 32        /* attention_kernel4: two threads per query row like kernel2, but the exponent reference trails the row
                                            max by one KV step (no per-step exchange between the two owners) and part of the exponentials
                                            run on the FMA pipe; a step whose logits jump > 2^100 over the reference is redone exactly */
int pe_attention_fwd(pe_handle_t h, const void* q, const void* k, const void* v, void* o,
                     int S, int H, int64_t ld, float scale, int flags, void* stream);
/* Same kernel (flags 0 / 1 / 2 only), output ROUTED by query row: rows [route_end[i-1], route_end[i]) are written to
 * o_route[i] + row * ldo + head * 128 (route_end[-1] = 0, route_end[n_route-1] >= S).  In the sequence-parallel mode rank r computes its
 * H = heads / ranks heads for ALL rows and writes every row into the attention buffer of the rank that owns it (peer-mapped pointers:
 * NVLink P2P stores from the epilogue), i.e. the second all-to-all of Ulysses attention is fused into this kernel. */
/* pe_attention_fwd that also returns the row statistics the backward needs: lse[head * S + row] = log2(sum_j exp2(scale * log2(e) * s_j))
 * (fp32, log2 domain, softmax scale included).  flags: PE_ATTN_FLAG_* of the default kernel only (0 .. 3). */
int pe_attention_fwd_lse(pe_handle_t h, const void* q, const void* k, const void* v, void* o, int S, int H, int64_t ld, float scale, int flags,
                         float* lse, void* stream);

int pe_attention_fwd_routed(pe_handle_t h, const void* q, const void* k, const void* v, int S, int H, int64_t ld, float scale, int flags,
                            int n_route, const int32_t* route_end, void* const* o_route, int64_t ldo, void* stream);

/* small generic attention for the training-path encoders (DINOv2 ViT-B: 261 tokens x 12 heads x 64,
 * transformers modeling_dinov2_with_registers.py:174-254; perceiver resampler: 64 latent queries over
 * <= 10304 keys, 8 heads x 64, helpers.py:21-65).  q: [B, Sq, *] row stride ldq, head hd at column
 * hd*D; k, v: [B, Skv, *] row stride ldkv; o: [B, Sq, *] row stride ldo.  D in {64, 128}. */
int pe_small_attention(pe_handle_t h, const void* q, const void* k, const void* v, void* o, int B, int H,
                       int Sq, int Skv, int D, int64_t ldq, int64_t ldkv, int64_t ldo, float scale, void* stream);

/* ------------------------------------------------------------------------------------------- */
/* row-wise / elementwise kernels (HBM-bound)                                                   */
/* ------------------------------------------------------------------------------------------- */
/* out[r,:] = bf16(bf16(LN(x[r,:])) * one_plus_scale) + shift ; LN non-affine, eps 1e-6
 *   replaces F.layer_norm + _modulate (qwen_image_dit.py:355-357,378-383; utils.py:306-308).
 *   shift / one_plus_scale: bf16 [C] (one_plus_scale already holds bf16(1+scale)). */
int pe_layernorm_modulate(pe_handle_t h, const void* x, void* out, int rows, int C,
                          const void* shift, const void* one_plus_scale, void* stream);
/* same over a joint [text; image] token buffer: rows [0, split_row) use (shift0, one_plus_scale0) (txt_mod),
 * rows [split_row, rows) use (shift1, one_plus_scale1) (img_mod) -- one launch for both streams of
 * QwenImageTransformerBlock.forward (qwen_image_dit.py:378-383, 389-393). */
int pe_layernorm_modulate2(pe_handle_t h, const void* x, void* out, int rows, int C, int split_row,
                           const void* shift0, const void* one_plus_scale0,
                           const void* shift1, const void* one_plus_scale1, void* stream);
/* nn.LayerNorm: out = bf16(LN(x) * w + b) (helpers.py:13,27-28,98; DINOv2 blocks); w = b = NULL gives the
 * non-affine LN of Dinov2withNorm (dinov2.py:20-24). */
int pe_layernorm(pe_handle_t h, const void* x, void* out, int rows, int C, const void* w, const void* b,
                 float eps, void* stream);
/* x[r,:] += alpha * add[r % period, :]  (positional / frame-index embedding adds, helpers.py:88,
 * qwen_image_physical.py:1074-1116; alpha=-1 gives the "middle - source" delta). */
int pe_add_rows(pe_handle_t h, void* x, const void* add, int rows, int C, int period, float alpha, void* stream);
/* out = bf16(bf16(x*rsqrt(mean(x^2)+eps)) * w)     (RMSNorm, utils.py:250-257; txt_norm) */
int pe_rmsnorm(pe_handle_t h, const void* x, void* out, int rows, int C, const void* w, float eps, void* stream);

/* y[b, n] = act_out( sum_k act_in(x[b,k]) * W[n,k] + bias[n] ),  b < batch <= 8.  HBM-bound GEMV.
 *   replaces the M=1 linears: img_mod/txt_mod (qwen_image_dit.py:335,349), norm_out.linear
 *   (utils.py:305), timestep_embedder (utils.py:266-270).
 *   act_in / act_out: 0 none, 1 SiLU.  If one_plus_mask != NULL it is a uint8 [N] mask: where 1 the
 *   stored value is bf16(1 + y) (pre-computes the "1 + scale" of _modulate). */
int pe_gemv(pe_handle_t h, const void* x, const void* w, const void* bias, void* y, int batch, int N, int K,
            int act_in, int act_out, const uint8_t* one_plus_mask, void* stream);

/* y = bf16(act(x)) element-wise on n bf16 values; act: 0 copy, 1 SiLU (the nn.SiLU in front of img_mod / txt_mod / norm_out.linear,
 * qwen_image_dit.py:333,347, utils.py:305, materialised once per timestep instead of inside every GEMV). */
int pe_act(pe_handle_t h, const void* x, void* y, int64_t n, int act, void* stream);

/* sinusoidal timestep embedding with the reference's bf16 quirks (utils.py:189-216):
 *   t_in: bf16 [1].  raw != 0: t_in is the loop's bf16(t) and ts = bf16(t_in/1000) is formed on device with
 *   ATen's CUDA rounding (qwen_image_physical.py:1342); raw == 0: t_in already holds ts (TimestepEmbeddings.forward). Then
 *   out[0:128]=cos(1000*ts*f_i), out[128:256]=sin(...), f_i = bf16(exp(-ln(1e4) i/128)); out bf16 [256]. */
int pe_timestep_embedding(pe_handle_t h, const void* t_in, void* out, int raw, void* stream);

/* patchify: latents bf16 [16, H8, W8] -> tokens [ (H8/2)*(W8/2), 64 ], channel order (c, p, q)
 *   replaces rearrange "B C (H P) (W Q) -> B (H W) (C P Q)" (qwen_image_physical.py:1344,1354). */
int pe_patchify(pe_handle_t h, const void* latents, void* tokens, int H8, int W8, void* stream);
/* inverse (qwen_image_physical.py:1402); tokens row stride ld elements. */
int pe_unpatchify(pe_handle_t h, const void* tokens, int64_t ld, void* latents, int H8, int W8, void* stream);

/* CFG combine + Euler update, bf16 roundings as the reference (qwen_image_physical.py:656-661,
 * flow_match.py:72-82):  np = nega + cfg*(posi-nega);  latents = latents + np * dsigma. */
int pe_cfg_euler_step(pe_handle_t h, void* latents, const void* posi, const void* nega, int64_t n,
                      float cfg_scale, float dsigma, void* stream);

/* special-token adapter plumbing (qwen_image_physical.py:1333-1336, helpers.py:142-164):
 *   gather: rows of prompt_emb [T, C] where mask[t]!=0 -> dst [max_rows, C] (zero padded),
 *           row indices -> idx int32 [max_rows] (-1 unused), count -> idx[max_rows].
 *   blend_scatter: prompt_emb[idx[i], :] = bf16(alpha*dino[i]) + bf16((1-alpha)*vae[i]) ... with the
 *           reference's bf16 op order; alpha computed on device from t (bf16 [1]), t_min, t_max. */
int pe_special_gather(pe_handle_t h, const void* prompt_emb, const uint8_t* mask, int T, int C,
                      void* dst, int32_t* idx, int max_rows, void* stream);
int pe_special_blend_scatter(pe_handle_t h, void* prompt_emb, const int32_t* idx, int max_rows, int C,
                             const void* pred_dino, const void* pred_vae, const void* t_in,
                             float t_min, float t_max, void* stream);

/* ------------------------------------------------------------------------------------------- */
/* QwenImageVAE encode / decode (DiffSynth-Studio/diffsynth/models/qwen_image_vae.py; called at  */
/* pipelines/qwen_image_physical.py:665,1273,1298,1092,1106).  Activation maps are channels-last */
/* bf16 [H*W, C] (pixel stride ld elements).  At T = 1 without a feature cache every             */
/* QwenImageCausalConv3d (:8-51) is a 2-D convolution with the last temporal slice of its kernel.*/
/* ------------------------------------------------------------------------------------------- */
/* Implicit-GEMM convolution on the tensor cores (same kernel as pe_gemm; the A operand is read tap by tap
 * from the activation map through a 3-D TMA tensor map, zero padding = TMA out-of-bounds fill, no im2col):
 *   out[y,x,n] = epi( sum_{dy<kh, dx<kw, c<C} x[y+dy-pad, x+dx-pad, c] * w[n, (dy*kw+dx)*cpad + c] + bias[n] ),
 *   cpad = round_up(C, 64), stride 1, output H x W = input H x W.
 * Replaces F.conv3d / F.conv2d at qwen_image_vae.py:51 (conv1/conv2/conv_in/conv_out), :243 (upsample conv) and, after
 * pe_space_to_depth with a 2x2 kernel and pad 0, the stride-2 downsample conv :247-249.
 * epilogue: PE_EPI_BIAS | PE_EPI_BIAS_SILU | PE_EPI_GATE_RESIDUAL (out += gate[n] * bf16(acc+bias), in place; gate = ones
 * gives the residual add of QwenImageResidualBlock.forward :152). */
typedef struct pe_conv2d_desc {
    const void* x;      /* bf16 [H, W, >=C], pixel stride ldx elements                      */
    int64_t ldx;
    const void* w;      /* bf16 [N, kh*kw*cpad] contiguous, tap-major then channel          */
    const void* bias;   /* bf16 [N] or NULL                                                  */
    void* out;          /* bf16 [H*W, >=N], pixel stride ldo elements                        */
    int64_t ldo;
    const void* gate;   /* PE_EPI_GATE_RESIDUAL: bf16 [N]                                    */
    int32_t H, W, C, N;
    int32_t kh, kw, pad;
    int32_t flags;      /* PE_CONV_FLAG_*                                                   */
} pe_conv2d_desc;
#define PE_CONV_FLAG_TILE_W_LOG2(n) (((n) & 7) << 4)   /* explicit pixel-patch width 2^n (3..7; patch height 128 >> n); 0 = automatic */
#define PE_CONV_FLAG_SINGLE_PATCH 2  /* N <= 128 on one CTA: do NOT pair two 128-pixel patches per tile (the default pairs them so that both
                                        share every weight box and each k-block iteration carries twice the MMA work) */
#define PE_CONV_FLAG_CTA_PAIR 1   /* cta_group::2: a tile is two stacked 128-pixel patches on an SM pair, each CTA loads half of the
                                     weight rows (halves the L2 -> SM weight traffic that bounds the narrow layers) */
int pe_conv2d(pe_handle_t h, const pe_conv2d_desc* desc, int epilogue, void* stream);

/* QwenImageRMS_norm (:76-78) over the channels of every pixel, optionally followed by nn.SiLU:
 *   n = max(bf16(||x||_2), 1e-12); y = bf16(bf16(bf16(x / n) * sqrt(C)) * gamma[c]); act != 0: y = bf16(silu(y)). */
int pe_channel_rmsnorm(pe_handle_t h, const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int C,
                       const void* gamma, int act, void* stream);
/* QwenImageUpsample (:202-215, nearest-exact, scale 2): in [H, W, C] -> out [2H, 2W, C], both contiguous. */
int pe_upsample2x(pe_handle_t h, const void* in, void* out, int H, int W, int C, void* stream);
/* in [H, W, C] -> out [H/2, W/2, 4C], out[y,x,(py*2+px)*C+c] = in[2y+py, 2x+px, c]: with it the ZeroPad2d((0,1,0,1)) + 3x3 stride-2
 * convolution of the downsample layers (:246-249) becomes a 2x2 stride-1 pe_conv2d with pad 0. */
int pe_space_to_depth(pe_handle_t h, const void* in, void* out, int H, int W, int C, void* stream);
/* layout changes at the ends of the VAE with the latent (de)normalisation folded in:
 *   op 0 copy; op 1 y = bf16(bf16(x / p1[c]) + p0[c]) (decode :724-725, p0 = mean, p1 = 1/std);
 *   op 2 y = bf16(bf16(x - p0[c]) * p1[c]) (encode :712-714).  p0, p1: bf16 [C]. */
int pe_nchw_to_nhwc(pe_handle_t h, const void* src, void* dst, int64_t ldd, int C, int64_t HW, int op,
                    const void* p0, const void* p1, void* stream);
int pe_nhwc_to_nchw(pe_handle_t h, const void* src, int64_t lds, void* dst, int C, int64_t HW, int op,
                    const void* p0, const void* p1, void* stream);
/* bf16 [R, C] (row stride lds) -> [C, R] (row stride ldd): V^T for the P.V product of the mid-block attention. */
int pe_transpose(pe_handle_t h, const void* src, int64_t lds, void* dst, int64_t ldd, int R, int C, void* stream);
/* probs[r, 0:n] = bf16(softmax(scale * scores[r, 0:n])), probs[r, n:n_pad] = 0; scores float [rows, lds]
 * (F.scaled_dot_product_attention of QwenImageAttentionBlock :189, single head of dim 384; the zero columns let the
 * P.V product run with its K dimension padded to a multiple of 8). */
int pe_softmax_rows(pe_handle_t h, const void* scores, int64_t lds, void* probs, int64_t ldp, int rows, int n, int n_pad,
                    float scale, void* stream);
/* The same with a key mask (EliGen entity control, SURVEY 8f5: the additive 0 / -inf mask of QwenImageDiT.process_entity_masks,
 * models/qwen_image_dit.py:433-498, handed to F.scaled_dot_product_attention :36): mask is a byte matrix [mask_period, ldm], 0 = hidden;
 * row r of the scores uses mask row r % mask_period (one mask for all heads of a batched score matrix). */
int pe_softmax_rows_masked(pe_handle_t h, const void* scores, int64_t lds, void* probs, int64_t ldp, int rows, int n, int n_pad,
                           float scale, const void* mask, int64_t ldm, int mask_period, void* stream);

/* ------------------------------------------------------------------------------------------- */
/* training path: the row term of the attention backward                            */
/*   The reference differentiates F.scaled_dot_product_attention (models/qwen_image_dit.py:14-39) */
/*   with autograd (pipelines/qwen_image_physical.py:313-329, scripts/train/train_physicedit.py   */
/*   :648-652).  Here the backward is pe_attention_fwd_lse (row statistics) + seven               */
/*   pe_gemm_batched launches (PE_EPI_ATTN_P / PE_EPI_ATTN_DS) + this pass.                        */
/* ------------------------------------------------------------------------------------------- */
/* delta[h * ld_delta + s] = sum_d dO[s, h*128 + d] * O[s, h*128 + d] (fp32); dO, O bf16 token-major [S, >= H*128]. */
int pe_attention_bwd_delta(pe_handle_t h, const void* d_o, int64_t ldd, const void* o, int64_t ldo, int S, int H, void* delta, int64_t ld_delta,
                           void* stream);

/* ------------------------------------------------------------------------------------------- */
/* Qwen2.5-VL text-encoder path: edit_forward prefill and greedy generate            */
/*   call sites: pipelines/qwen_image_physical.py:774-800 (prompt_emb), :859-873 (generate);      */
/*   model wrapper models/qwen_image_text_encoder_withdecode.py:147-275; arithmetic = transformers */
/*   modeling_qwen2_5_vl.py (un-pinned third-party dependency; installed 5.5.0).  The linears run  */
/*   on pe_gemm / pe_gemv, the norms on pe_rmsnorm; the entry points below are the rest.           */
/* ------------------------------------------------------------------------------------------- */
/* M <= 8 linear of the one-token decode step with its neighbours fused in (every CTA redoes the prologue on the <= 18944 inputs):
 *   act_in 0 / 1: x [batch, K] (1: SiLU first);  act_in 2: x [batch, 2K] = gate | up, input = bf16(bf16(silu(gate)) * up) (Qwen2MLP);
 *   norm_w != NULL: input = bf16(norm_w * bf16(x * rsqrt(mean(x^2) + eps))) (Qwen2_5_VLRMSNorm :66-71 in front of q/k/v and gate/up);
 *   residual != NULL: y = bf16(residual + bf16(acc + bias)) (Qwen2_5_VLDecoderLayer.forward :818, :824; may alias y).
 * Same arithmetic and rounding points as pe_rmsnorm / pe_swiglu / pe_gemv / pe_add_rows run one after the other. */
int pe_gemv_fused(pe_handle_t h, const void* x, const void* w, const void* bias, void* y, int batch, int N, int K, int act_in,
                  const void* norm_w, float norm_eps, const void* residual, void* stream);
/* The gate / up projections of the decode step and act_fn(gate) * up in ONE launch (Qwen2MLP.forward modeling_qwen2_5_vl.py:622-624): w = [gate_proj
 * rows | up_proj rows] [2 I, K], x [batch, K] with the optional RMSNorm prologue of pe_gemv_fused, y[b, i] = bf16(bf16(silu(bf16(g_i))) * bf16(u_i)),
 * [batch, I].  Bit-identical to pe_gemv_fused into a gate|up buffer followed by pe_swiglu; the down-projection then needs no SwiGLU prologue. */
int pe_gemv_swiglu(pe_handle_t h, const void* x, const void* w, const void* bias, void* y, int batch, int I, int K, const void* norm_w, float norm_eps,
                   void* stream);
/* out[r, i] = bf16(bf16(silu(x[r, i])) * x[r, I + i]): act_fn(gate_proj(x)) * up_proj(x) on a fused [rows, 2I] gate|up buffer
 * (Qwen2MLP.forward modeling_qwen2_5_vl.py:622-624, Qwen2_5_VLMLP.forward :87-88). */
int pe_swiglu(pe_handle_t h, const void* x, int64_t ldx, void* out, int64_t ldo, int rows, int I, void* stream);
/* rotate-half RoPE in place on x [T, H*D] (row stride ldx): token t takes row (row_ptr ? *row_ptr : row0) + t of the fp32 tables
 * cos / sin [*, D].  mode 0: fp32 arithmetic (apply_rotary_pos_emb_vision :159-171); mode 1: the language model's bf16 op order with
 * bf16-rounded cos / sin (apply_multimodal_rotary_pos_emb :627-668 on bf16 tensors).  row_ptr is a DEVICE pointer so that a decode step
 * captured in a CUDA graph reads its position at replay time. */
int pe_rope_half(pe_handle_t h, void* x, int64_t ldx, int T, int H, int D, const float* cos_table, const float* sin_table,
                 const int32_t* row_ptr, int row0, int mode, void* stream);
/* o[i, hd] = softmax(scale * q[i, hd] . k[lo_i:hi_i, hd / (H/Hkv)]^T) v[lo_i:hi_i, ...]: grouped KV heads (repeat_kv :174-183) and a
 * per-query KV range instead of an additive mask.  kv_lo / kv_hi: device int32 [Sq] or NULL (0 / the KV length); kv_len_ptr: device
 * int32 holding the KV length when kv_hi is NULL (KV-cache decode), else Skv.  D in {64, 80, 128}.  Replaces the SDPA calls of
 * Qwen2_5_VLAttention.forward :739-751 (causal) and Qwen2_5_VLVisionAttention.forward :262-283 (per-window / full chunks). */
int pe_range_attention(pe_handle_t h, const void* q, const void* k, const void* v, void* o, int H, int Hkv, int Sq, int Skv, int D,
                       int64_t ldq, int64_t ldkv, int64_t ldo, float scale, const int32_t* kv_lo, const int32_t* kv_hi,
                       const int32_t* kv_len_ptr, void* stream);
/* out[i, :] = table[ids[i], :] for ids[i] >= 0 (negative: row left as is): embed_tokens (:874) and the masked_scatter of the image
 * embeddings into the token stream (:1311-1316).  ids: device int64 [n]. */
int pe_gather_rows(pe_handle_t h, const void* table, int64_t ldt, const int64_t* ids, void* out, int64_t ldo, int n, int C, void* stream);
/* out[0] = first index of max(x[0:n]) (torch.argmax; greedy decoding); optionally log[*log_pos] = out[0] (device-side token log). */
int pe_argmax(pe_handle_t h, const void* x, int n, int64_t* out, int64_t* log, const int32_t* log_pos, void* stream);
/* cache_k[pos[0], :] = k_new, cache_v[pos[0], :] = v_new (DynamicCache.update of one decode step, position read on the device). */
int pe_kv_append(pe_handle_t h, const void* k_new, const void* v_new, void* cache_k, void* cache_v, int64_t ldc, int C,
                 const int32_t* pos, void* stream);
/* One decode row qkv = [q (Hq heads) | k (Hkv heads) | v]: pe_rope_half (mode 1) on the q and k heads at table row counters[1], then
 * cache_k[counters[0], :] = rotated k, cache_v[counters[0], :] = v.  Fuses two pe_rope_half and one pe_kv_append launch. */
int pe_rope_kv_append(pe_handle_t h, void* qkv, int Hq, int Hkv, int D, const float* cos_table, const float* sin_table, void* cache_k,
                      void* cache_v, int64_t ldc, const int32_t* counters, void* stream);
/* The attention part of ONE decode step for up to 8 requests in ONE launch (the per-layer body of Qwen2_5_VLAttention.forward for a single new
 * token, modeling_qwen2_5_vl.py:690-760, with DynamicCache.update): pe_rope_kv_append + the one-query pe_range_attention of every request.
 * qkv = [q (Hq heads) | k (Hkv) | v (Hkv)] of the new token (left untouched), counters[0] = cache rows before the append, counters[1] = rope
 * table row; out [Hq * 128].  Bit-identical to the two-launch sequence. */
typedef struct pe_decode_req {
    const void* qkv;
    void* cache_k;
    void* cache_v;
    void* out;
    const int32_t* counters;
    int64_t cache_rows;   /* capacity of this request's cache */
} pe_decode_req;
int pe_decode_attention_fused(pe_handle_t h, const pe_decode_req* reqs, int n_req, int Hq, int Hkv, int D, int64_t ldc,
                              const float* cos_table, const float* sin_table, float scale, void* stream);
/* counters[0:n] += 1 (KV length, rope row and step counters of the captured decode step). */
int pe_advance(pe_handle_t h, int32_t* counters, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PE_B200_H_ */
