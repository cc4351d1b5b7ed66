"""GPU probe: pe_conv2d timing over patch shapes / CTA pairing at the VAE's layer geometries -> gpurun_out/conv_probe.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from physicedit_b200 import native as nv  # noqa: E402


def main():
    nat = nv.Native.get(0)
    res = []
    for (H, W, C, N) in ((1024, 1024, 96, 96), (1024, 1024, 96, 8), (512, 512, 192, 192), (256, 256, 384, 384)):
        cpad = (C + 63) // 64 * 64
        x = torch.randn(H * W, C, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, 9 * cpad, device="cuda") * 0.03).to(torch.bfloat16)
        b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(H * W, N, device="cuda", dtype=torch.bfloat16)
        ref = None
        for pair in (0, 1, 2):
            for lg in (4,):
                flags = pair | (lg << 4)
                for _ in range(2):
                    nat.conv2d(x, H, W, C, w, b, out, N, 3, 3, 1, 0, flags=flags)
                nat.check_async()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    nat.conv2d(x, H, W, C, w, b, out, N, 3, 3, 1, 0, flags=flags)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                if ref is None:
                    ref = out.clone()
                same = bool(torch.equal(ref, out))
                res.append(dict(H=H, W=W, C=C, N=N, pair=pair, tile_w=1 << lg, ms=round(ms, 4), tflops=round(2 * H * W * 9 * C * N / ms / 1e9, 1), same=same))
                print(res[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "conv_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
