#!/usr/bin/env python
"""Clean timing of the decode-step GEMV shapes of the Qwen2.5-VL text encoder: 28 distinct weight matrices per shape (so L2 never helps),
one CUDA graph of 28 launches, replayed; prints us per launch and GB/s.  PE_GEMV_MODE=1/2 forces the wide / narrow kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from physicedit_b200 import native as nv
nat = nv.Native.get(0)
dev = "cuda"
res = {}
for name, N, K, fused in (("gate_up_swiglu", 37888, 3584, "gu_swiglu"), ("qkv", 4608, 3584, "norm"), ("o", 3584, 3584, "res"), ("gate_up", 37888, 3584, "norm"), ("down", 3584, 18944, "swiglu"), ("down_plain", 3584, 18944, "res"), ("lm_head", 152064, 3584, "norm")):
    for batch in (1, 2):
        L = 28 if name != "lm_head" else 4
        ws = [torch.randn(N, K, device=dev).bfloat16() for _ in range(L)]
        x = torch.randn(batch, K * (2 if fused == "swiglu" else 1), device=dev).bfloat16()
        nw = torch.ones(K, device=dev).bfloat16()
        y = torch.zeros(batch, N // (2 if fused == "gu_swiglu" else 1), device=dev).bfloat16()
        def run():
            for w in ws:
                if fused == "gu_swiglu": nat.gemv_swiglu(x, w, None, y, norm_w=nw)
                elif fused == "norm": nat.gemv_fused(x, w, None, y, norm_w=nw)
                elif fused == "res": nat.gemv_fused(x, w, None, y, residual=y)
                else: nat.gemv_fused(x, w, None, y, act_in=2, residual=y)
        run(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                run()
        torch.cuda.current_stream().wait_stream(s)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (10 * L)
        res[f"{name}_b{batch}"] = (round(us, 1), round(N * K * 2 / us / 1e3, 0))
        del ws
print(json.dumps(res))
