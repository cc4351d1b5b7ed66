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
nat.check_async()
torch.cuda.synchronize()
print("sanitize targets done")
