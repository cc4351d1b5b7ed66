#!/usr/bin/env python
"""A/B of the joint-attention kernel against the library kernels in the image, same box, same inputs.

    python tools/attn_ab.py [--out gpurun_out/attn_ab.json]

For S = 8704 (1024^2 step, T = 512) and S = 20992 (2048^2), 24 heads x 128, bf16, non-causal:
  pe_attention_fwd (libpe_b200, flags 0 and 16)  vs  F.scaled_dot_product_attention forced to the cuDNN backend, to the flash
  backend, and flash_attn 2.8.x's flash_attn_func.  CUDA events around 20 back-to-back launches after 5 warm-up launches
  ("burst": boost clocks, q/k/v L2-resident as they are in the loop right after the QKV GEMM) and around a 3-second loop
  ("sustained": the power-capped clock the denoise loop runs at).  Library kernels are the yardstick, not the product path.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def sustained(fn, seconds=3.0):
    fn(); torch.cuda.synchronize()
    t_end = time.time() + seconds
    n, ms = 0, 0.0
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms, n = ms + e0.elapsed_time(e1), n + 50
    # the last third of the window is the power-capped steady state
    return ms / n


def sm_clock():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
        return out.stdout.strip().splitlines()[0]
    except Exception:  # noqa: BLE001
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "attn_ab.json"))
    ap.add_argument("--flags", default="0,16")
    ap.add_argument("--no-sustained", action="store_true")
    args = ap.parse_args()
    from torch.nn.attention import SDPBackend, sdpa_kernel
    import torch.nn.functional as F
    from physicedit_b200 import native as nv
    nat = nv.Native.get(0)
    H, D = 24, 128
    res = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__, "cases": {}}
    try:
        import flash_attn
        from flash_attn import flash_attn_func
        res["flash_attn"] = flash_attn.__version__
    except Exception as e:  # noqa: BLE001
        flash_attn_func = None
        res["flash_attn"] = f"unavailable: {e}"
    res["cudnn"] = torch.backends.cudnn.version()
    for S in (8704, 20992):
        g = torch.Generator(device="cuda").manual_seed(S)
        q, k, v = (torch.randn(S, H * D, device="cuda", generator=g).bfloat16() for _ in range(3))
        o = torch.empty_like(q)
        flops = 4.0 * S * S * D * H
        case = {}
        # token-major [S, H*D] is what the DiT holds; the libraries get the same memory as strided [1, H, S, D] / [1, S, H, D] views
        q4, k4, v4 = (t.view(1, S, H, D).transpose(1, 2) for t in (q, k, v))
        ref = None
        impls = {}
        for fl in [int(x) for x in args.flags.split(",")]:
            impls[f"pe_attention_fwd(flags={fl})"] = (lambda fl=fl: nat.attention(q, k, v, o, H, 1 / math.sqrt(D), fl))
        for name, be in (("sdpa_cudnn", SDPBackend.CUDNN_ATTENTION), ("sdpa_flash", SDPBackend.FLASH_ATTENTION)):
            def run(be=be):
                with sdpa_kernel([be]):
                    return F.scaled_dot_product_attention(q4, k4, v4)
            impls[name] = run
        if flash_attn_func is not None:
            qf, kf, vf = (t.view(1, S, H, D) for t in (q, k, v))
            impls["flash_attn_func"] = lambda: flash_attn_func(qf, kf, vf)
        for name, fn in impls.items():
            try:
                out = fn()
                torch.cuda.synchronize()
                if name.startswith("pe_"):
                    nat.check_async()
                    got = o.clone()
                else:
                    got = out.transpose(1, 2).reshape(S, H * D) if out.shape[1] == H else out.reshape(S, H * D)
                if ref is None:
                    rows = torch.arange(0, S, max(1, S // 256), device="cuda")[:256]
                    ref = (rows, torch.stack([torch.softmax(q[rows, h * D:(h + 1) * D].float() @ k[:, h * D:(h + 1) * D].float().t() / math.sqrt(D), -1)
                                              @ v[:, h * D:(h + 1) * D].float() for h in range(H)], 1).reshape(len(rows), H * D))
                err = ((got[ref[0]].float() - ref[1]).norm() / ref[1].norm()).item()
                ms = timed(fn)
                entry = {"ms_burst": round(ms, 4), "tflops_burst": round(flops / ms / 1e9, 1), "rel_l2_vs_fp32": err, "clock_after_burst": sm_clock()}
                if not args.no_sustained:
                    ms_s = sustained(fn)
                    entry.update(ms_sustained=round(ms_s, 4), tflops_sustained=round(flops / ms_s / 1e9, 1), clock_after_sustained=sm_clock())
                case[name] = entry
            except Exception as e:  # noqa: BLE001
                case[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            print(S, name, case[name], flush=True)
        res["cases"][f"S={S}"] = case
        del q, k, v, o
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
