#!/usr/bin/env python
"""Small launches of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): shapes with ragged tails."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from physicedit_b200 import native as nv
nat = nv.Native.get(0)
dev = "cuda"
bf = dict(device=dev, dtype=torch.bfloat16)
torch.manual_seed(0)
if os.environ.get("SANITIZE_ONLY_R2"):
    # ---- round-2 kernels: batched GEMM + attention-backward epilogues, forward with row statistics, delta pass, masked softmax, cluster split-K GEMV,
    # SwiGLU-epilogue GEMV, staged small attention, LLM row kernels, routed attention / QKV epilogue (local pointers) ----
    B, M, N, K = 3, 200, 144, 128
    Mp = 208
    a = torch.randn(B * Mp, K, **bf); w = torch.randn(B * N, K, **bf) / 11
    out = torch.zeros(B * Mp, N, **bf); o32 = torch.zeros(B * Mp, N, device=dev)
    vec = torch.randn(B, 224, device=dev)
    kw = dict(batch=B, M=M, N=N, K=K, a_batch_rows=Mp, w_batch_rows=N, out_batch_rows=Mp)
    nat.gemm_batched(a, w, out, **kw)
    nat.gemm_batched(a, w, o32, epilogue=nv.EPI_F32, **kw)
    for col in (False, True):
        nat.gemm_batched(a, w, out, epilogue=nv.EPI_ATTN_P, vec=vec, vec_batch_stride=224, vec_per_column=col, alpha=0.3, **kw)
        nat.gemm_batched(a, w, out, epilogue=nv.EPI_ATTN_DS, vec=vec, vec_batch_stride=224, vec_per_column=col, alpha=0.3, **kw)
    S, H = 330, 2
    q, k, v = (torch.randn(S, H * 128, **bf) for _ in range(3))
    o = torch.empty_like(q); lse = torch.empty(H, S, device=dev)
    nat.attention_lse(q, k, v, o, lse, H, 1 / math.sqrt(128))
    delta = torch.zeros(H, 336, device=dev)
    nat.attention_bwd_delta(o, q, delta, H)
    nat.attention_routed(q, k, v, H, 1 / math.sqrt(128), [128, 330], [o.data_ptr(), o.data_ptr()], H * 128, 0)
    sc = torch.randn(2 * 72, 72, device=dev); pr = torch.empty(2 * 72, 72, **bf); mk = (torch.rand(72, 72, device=dev) > 0.3).to(torch.uint8); mk.fill_diagonal_(1)
    nat.softmax_rows(sc, pr, 70, 0.1, mask=mk)
    for batch in (1, 2):
        wl = torch.randn(50, 8192, **bf) / 90; xl = torch.randn(batch, 2 * 8192, **bf); yl = torch.zeros(batch, 50, **bf)
        nat.gemv_fused(xl, wl, None, yl, act_in=2, residual=yl)                    # cluster split-K kernel (K >= 8192), ragged N
        wg = torch.randn(2 * 1000, 512, **bf) / 22; xg = torch.randn(batch, 512, **bf); yg = torch.zeros(batch, 1000, **bf)
        nat.gemv_swiglu(xg, wg, None, yg, norm_w=torch.ones(512, **bf))
    qs, ks_, vs_ = torch.randn(2 * 70, 3 * 64, **bf), torch.randn(2 * 261, 3 * 64, **bf), torch.randn(2 * 261, 3 * 64, **bf)
    os_ = torch.empty(2 * 70, 3 * 64, **bf)
    nat.small_attention(qs, ks_, vs_, os_, 2, 3, 70, 261, 64, 0.125)               # staged variant (Sq >= 32, K/V fit in shared memory)
    nat.check_async()
    torch.cuda.synchronize()
    print("sanitize targets (r2) done")
    sys.exit(0)
if not os.environ.get("SANITIZE_ONLY_VAE"):      # the DiT kernel families (sanitized in r1 sessions 2-3; set the variable to run only the VAE additions)
    # attention: all kernels, ragged S
    for flags in (0, 16, 8):
        S, H = 330, 2
        q, k, v = (torch.randn(S, H * 128, device=dev).bfloat16() for _ in range(3))
        o = torch.empty_like(q)
        nat.attention(q, k, v, o, H, 1 / math.sqrt(128), flags)
    # GEMMs: single CTA and CTA pair, two segments, ragged M
    for cg in (0, nv.GEMM_FLAG_CTA_PAIR):
        a1, a2 = torch.randn(300, 192, **bf), torch.randn(77, 192, **bf)
        w = torch.randn(264, 192, **bf)
        b = torch.randn(264, **bf)
        o1, o2 = torch.empty(300, 264, **bf), torch.empty(77, 264, **bf)
        nat.gemm([dict(a=a1, w=w, bias=b, out=o1), dict(a=a2, w=w, bias=b, out=o2)], 264, 192, nv.EPI_BIAS_GELU_SIGMOID, cg)
    # LN + modulate: bulk-copy kernel (C = 3072, rows >= 32) and the register kernel (few rows)
    for rows in (77, 5):
        x = torch.randn(rows, 3072, **bf); out = torch.empty_like(x)
        sh = [torch.randn(3072, **bf) for _ in range(4)]
        nat.layernorm_modulate2(x, out, rows // 2, sh[0], sh[1], sh[2], sh[3])
    # GEMV + act
    t = torch.randn(3, 3072, **bf); ta = torch.empty_like(t)
    nat.act(t, ta, 1)
    w = torch.randn(1030, 3072, **bf); y = torch.empty(3, 1030, **bf)
    nat.gemv(ta, w, torch.zeros(1030, **bf), y, 0, 0, None)
# VAE path: implicit-GEMM conv (one CTA / CTA pair, ragged patches, trimmed n-tile, residual epilogue), narrow GEMM with the fp32 epilogue,
# channel RMS-norm, layout kernels, softmax, transpose
for pair in (0, 1):
    H, W, C, N = 13, 21, 96, 96
    x = torch.randn(H * W, C, **bf); w = torch.randn(N, 9 * 128, **bf) * 0.03; b = torch.randn(N, **bf)
    out = torch.randn(H * W, N, **bf)
    nat.conv2d(x, H, W, C, w, b, out, N, 3, 3, 1, nv.EPI_GATE_RESIDUAL, gate=torch.ones(N, **bf), flags=pair)
    w2 = torch.randn(384, 4 * 384, **bf) * 0.03
    x2 = torch.randn(12 * 20, 384, **bf); o2 = torch.empty(12 * 20, 384, **bf)
    nat.conv2d(x2, 12, 20, 384, w2, torch.zeros(384, **bf), o2, 384, 2, 2, 0, nv.EPI_BIAS, flags=pair)
    a = torch.randn(130, 96, **bf); wq = torch.randn(72, 96, **bf); sc = torch.empty(130, 72, device=dev, dtype=torch.float32)
    nat.gemm([dict(a=a, w=wq, bias=None, out=sc)], 72, 96, nv.EPI_F32, nv.GEMM_FLAG_TRIM_N | pair)
for C in (96, 192, 384):
    x = torch.randn(333, C, **bf); o = torch.empty_like(x)
    nat.channel_rmsnorm(x, o, C, torch.ones(C, **bf), True)
x = torch.randn(10 * 14, 96, **bf)
nat.upsample2x(x, torch.empty(4 * 140, 96, **bf), 10, 14, 96)
nat.space_to_depth(x, torch.empty(35, 384, **bf), 10, 14, 96)
lat = torch.randn(16, 6, 10, **bf); z = torch.zeros(60, 64, **bf); p0 = torch.randn(16, **bf); p1 = torch.rand(16, **bf) + 0.5
nat.nchw_to_nhwc(lat, z, 16, 1, p0, p1)
nat.nhwc_to_nchw(z, torch.empty(16, 6, 10, **bf), 16, 2, p0, p1)
s32 = torch.randn(61, 64, device=dev); pr = torch.empty(61, 64, **bf)
nat.softmax_rows(s32, pr, 61, 0.05)
vt = torch.zeros(384, 64, **bf)
nat.transpose(torch.randn(61, 384, **bf), vt[:, :61])
nat.check_async()
torch.cuda.synchronize()
print("sanitize targets done")
