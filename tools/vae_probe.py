"""GPU probe: times QwenImageVAE.encode / .decode at 1024 x 1024 on the native path, with a per-kernel-family CUDA-event
breakdown, and writes gpurun_out/vae_probe.json.  Usage: python tools/vae_probe.py [H W]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from physicedit_b200 import native as nv  # noqa: E402
from physicedit_b200.vae import QwenImageVAE  # noqa: E402


def synth_state_dict(m, seed=0):
    g = torch.Generator("cpu").manual_seed(seed)
    sd = {}
    for k, v in m.state_dict().items():
        if k.endswith("gamma"):
            t = 1 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("weight"):
            t = (torch.rand(v.shape, generator=g) * 2 - 1) * 1.7 / (v.shape[1] * v.shape[-1] * v.shape[-2]) ** 0.5
        else:
            t = (torch.rand(v.shape, generator=g) * 2 - 1) * 0.05
        sd[k] = t.to(torch.bfloat16)
    return sd


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    W = int(sys.argv[2]) if len(sys.argv) > 2 else H
    nat = nv.Native.get(0)
    with torch.device("meta"):
        m = QwenImageVAE()
    m.load_state_dict(synth_state_dict(m), assign=True)
    m = m.to("cuda").eval()
    img = (torch.rand(1, 3, H, W, device="cuda") * 2 - 1).to(torch.bfloat16)
    lat = torch.randn(1, 16, H // 8, W // 8, device="cuda").to(torch.bfloat16)
    res = {"H": H, "W": W}
    for name, fn, arg in (("decode", m.decode, lat), ("encode", m.encode, img)):
        for _ in range(2):
            out = fn(arg)
        nat.check_async()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = nat.launches
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            out = fn(arg)
        e1.record()
        torch.cuda.synchronize()
        res[name] = {"ms": e0.elapsed_time(e1) / 5, "launches": (nat.launches - n0) // 5, "finite": bool(torch.isfinite(out.float()).all())}
        nat.prof = {}
        fn(arg)
        torch.cuda.synchronize()
        res[name]["by_kernel_ms"] = {k: {"launches": v[0], "ms": round(v[1], 4)} for k, v in sorted(nat.profile_summary().items(), key=lambda kv: -kv[1][1])}
        nat.prof = None
    # algorithmic FLOPs of the decoder convolutions at this size (2*HW*9*Cin*Cout per 3x3 layer) for the roofline line
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "vae_probe.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
