#!/usr/bin/env python
"""Launches pe_conv2d at the VAE's two dominant layer geometries (for `ncu --set full -k regex:gemm_kernel`):
1024x1024 96->96 and 512x512 192->192, 3x3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from physicedit_b200 import native as nv
nat = nv.Native.get(0)
flags = int(os.environ.get("CONV_FLAGS", "0"))
for (H, W, C, N) in ((1024, 1024, 96, 96), (512, 512, 192, 192)):
    cpad = (C + 63) // 64 * 64
    x = torch.randn(H * W, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, 9 * cpad, device="cuda") * 0.03).to(torch.bfloat16)
    b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(H * W, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        nat.conv2d(x, H, W, C, w, b, out, N, 3, 3, 1, 0, flags=flags)
nat.check_async()
print("done")
