#!/usr/bin/env python
"""Launches the four block GEMMs once each at the 1024^2 shapes with the L2 flushed in between (for ncu DRAM-traffic captures):
QKV (N=9216, K=3072), out-projection (N=3072, K=3072), MLP up (N=12288, K=3072), MLP down (N=3072, K=12288); image + text segments."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from physicedit_b200 import native as nv
nat = nv.Native.get(0)
dev = "cuda"
Si, St, D = 8192, 512, 3072
bf = dict(device=dev, dtype=torch.bfloat16)
def rnd(*s, sc=1.0): return (torch.randn(*s, device=dev) * sc).bfloat16()
x_i, x_t = rnd(Si, D), rnd(St, D)
h_i, h_t = rnd(Si, 4 * D), rnd(St, 4 * D)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def gemm(name, N, K, a_i, a_t, epi):
    w_i, w_t = rnd(N, K, sc=1 / 55), rnd(N, K, sc=1 / 55)
    b = torch.zeros(N, **bf)
    o_i, o_t = torch.zeros(Si, N, **bf), torch.zeros(St, N, **bf)
    segs = [dict(a=a_i, w=w_i, bias=b, out=o_i), dict(a=a_t, w=w_t, bias=b, out=o_t)]
    if epi == nv.EPI_GATE_RESIDUAL:
        g = rnd(N)
        for s in segs: s["gate"] = g
    for _ in range(2):
        flush.zero_()
        nat.gemm(segs, N, K, epi, nv.GEMM_FLAG_CTA_PAIR)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): nat.gemm(segs, N, K, epi, nv.GEMM_FLAG_CTA_PAIR)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms*1e3:.1f} us  {2*(Si+St)*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
gemm("out  N=3072  K=3072 ", D, D, x_i, x_t, nv.EPI_GATE_RESIDUAL)
gemm("up   N=12288 K=3072 ", 4 * D, D, x_i, x_t, nv.EPI_BIAS_GELU_SIGMOID)
gemm("down N=3072  K=12288", D, 4 * D, h_i, h_t, nv.EPI_GATE_RESIDUAL)
gemm("qkvN N=9216  K=3072 (bias epilogue)", 3 * D, D, x_i, x_t, nv.EPI_BIAS)
nat.check_async()
