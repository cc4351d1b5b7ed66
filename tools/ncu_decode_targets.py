#!/usr/bin/env python
"""One launch each of the decode step's MLP kernels at the 7B shapes, batch 2, cold L2 (for `ncu --set full -k regex:gemv`): the gate / up GEMV with the
SwiGLU epilogue (37888 x 3584) and the cluster split-K down-projection (3584 x 18944); plus the batched GEMM with the attention-backward P epilogue."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from physicedit_b200 import native as nv
nat = nv.Native.get(0)
dev = "cuda"
bf = dict(device=dev, dtype=torch.bfloat16)
I, K = 18944, 3584
wgu = torch.randn(2 * I, K, **bf) / 60; wd = torch.randn(K, I, **bf) / 140
x = torch.randn(2, K, **bf); nw = torch.ones(K, **bf); hm = torch.empty(2, I, **bf); y = torch.zeros(2, K, **bf)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
S, H = 3632, 24
Sp = (S + 15) // 16 * 16
qh, kh = torch.randn(H * Sp, 128, **bf), torch.randn(H * Sp, 128, **bf)
P = torch.empty(H * Sp, Sp, **bf); lse = torch.randn(H, Sp, device=dev) + 12
for i in range(2):
    flush.zero_()
    nat.gemv_swiglu(x, wgu, None, hm, norm_w=nw)
    flush.zero_()
    nat.gemv_fused(hm, wd, None, y, residual=y)
    flush.zero_()
    nat.gemm_batched(qh, kh, P, batch=H, M=S, N=Sp, K=128, a_batch_rows=Sp, w_batch_rows=Sp, out_batch_rows=Sp, epilogue=nv.EPI_ATTN_P, vec=lse, vec_batch_stride=Sp,
                     alpha=0.1275)
nat.check_async()
torch.cuda.synchronize()
print("done")
