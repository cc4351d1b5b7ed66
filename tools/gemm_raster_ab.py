#!/usr/bin/env python
"""A/B of the GEMM tile rasterisation and the L2 policies of the operand loads (gemm_sm100.cu::choose_raster) at the 1024^2 block shapes.

  python tools/gemm_raster_ab.py                 # bit-equality against the default order, burst timing, up -> down loop timing
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_kernel \
      --csv --log-file gpurun_out/raster_ncu.csv python tools/gemm_raster_ab.py --ncu
                                                 # per variant one launch with the L2 flushed and one right after it (launch order = the
                                                 # order of `plan()` printed on stdout)
The variants are chosen per launch through PE_GEMM_RASTER (read by the library when PE_GEMM_TUNE is set): "gn,gm,hint_a,hint_w,clusters"."""
import json, os, sys
os.environ["PE_GEMM_TUNE"] = "1"
os.environ["PE_GEMM_RASTER"] = "0,0,0,0,0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from physicedit_b200 import native as nv

NCU = "--ncu" in sys.argv
nat = nv.Native.get(0)
dev = "cuda"
Si, St, D = 8192, 512, 3072
torch.manual_seed(0)
def rnd(*s, sc=1.0): return (torch.randn(*s, device=dev) * sc).bfloat16()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

DOWN = [("base", "0,0,0,0,0"), ("base+hints", "0,0,1,2,0"), ("m6 c72", "0,6,0,0,72"), ("m6 c72 +hints", "0,6,1,2,72"),
        ("band6 m12 c72", "6,12,0,0,72"), ("band6 m12 c72 +hints", "6,12,1,2,72"), ("band6 m12 c74 +hints", "6,12,1,2,0"),
        ("band6 m12 c72 A-first", "6,12,1,0,72"), ("band4 m18 c72 +hints", "4,18,1,2,72"),
        ("m6 c74", "0,6,0,0,0"), ("m6 c72 W-last", "0,6,0,2,72"), ("base W-last", "0,0,0,2,0"), ("band6 m12 c72 W-last", "6,12,0,2,72"),
        ("auto", "auto")]
WIDE = [("base", "0,0,0,0,0"), ("base A-last W-first", "0,0,2,1,0"), ("one group", "0,34,0,0,0"), ("one group A-last W-first", "0,34,2,1,0"),
        ("base A-last", "0,0,2,0,0"), ("base W-last", "0,0,0,2,0"), ("auto", "auto")]


class Gemm:
    def __init__(self, name, N, K, epi):
        self.name, self.N, self.K, self.epi = name, N, K, epi
        self.a = [rnd(Si, K), rnd(St, K)]
        self.w = [rnd(N, K, sc=1 / 55), rnd(N, K, sc=1 / 55)]
        self.b = rnd(N, sc=0.1)
        self.g = rnd(N)
        self.res = [rnd(Si, N), rnd(St, N)]
        self.o = [torch.empty(Si, N, device=dev, dtype=torch.bfloat16), torch.empty(St, N, device=dev, dtype=torch.bfloat16)]

    def reset(self):
        for o, r in zip(self.o, self.res): o.copy_(r)

    def launch(self, variant=None):
        if variant is not None: os.environ["PE_GEMM_RASTER"] = variant
        segs = [dict(a=a, w=w, bias=self.b, out=o) for a, w, o in zip(self.a, self.w, self.o)]
        if self.epi == nv.EPI_GATE_RESIDUAL:
            for s in segs: s["gate"] = self.g
        nat.gemm(segs, self.N, self.K, self.epi, nv.GEMM_FLAG_CTA_PAIR)

    def flops(self): return 2 * (Si + St) * self.N * self.K


def timed(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


down = Gemm("down N=3072 K=12288", D, 4 * D, nv.EPI_GATE_RESIDUAL)
up = Gemm("up N=12288 K=3072", 4 * D, D, nv.EPI_BIAS_GELU_SIGMOID)
qkv = Gemm("qkv-shape N=9216 K=3072", 3 * D, D, nv.EPI_BIAS)
plan = [(down, DOWN), (up, WIDE), (qkv, WIDE)]

if NCU:
    order = []
    for g, variants in plan:
        g.reset(); g.launch("0,0,0,0,0"); torch.cuda.synchronize()          # first-launch work (function attributes) outside the captures
        order.append((g.name, "warm-up", "-"))
        for label, v in variants:
            flush.zero_(); torch.cuda.synchronize()
            g.launch(v); order.append((g.name, label, "cold"))
            g.launch(v); order.append((g.name, label, "after itself"))
            torch.cuda.synchronize()
    nat.check_async()
    print(json.dumps({"launch_order": order}))
    sys.exit(0)

out = {"shapes": "M = 8192 image + 512 text rows (two segments, two weight matrices), CTA pairs", "results": []}
for g, variants in plan:
    g.reset(); g.launch("0,0,0,0,0"); torch.cuda.synchronize()
    ref = [o.clone() for o in g.o]
    # sanity of the default order against a plain matmul (fp32 accumulate) on the text segment
    if g.epi in (nv.EPI_BIAS, nv.EPI_GATE_RESIDUAL):
        want = (g.a[1].float() @ g.w[1].float().t() + g.b.float())
        if g.epi == nv.EPI_GATE_RESIDUAL: want = g.res[1].float() + g.g.float() * want
        err = ((ref[1].float() - want).norm() / want.norm()).item()
        out.setdefault("default_vs_matmul_rel_l2", {})[g.name] = err
        print(json.dumps({"gemm": g.name, "default_vs_matmul_rel_l2": err}), flush=True)
    for label, v in variants:
        g.reset(); g.launch(v); torch.cuda.synchronize()
        same = all(torch.equal(o, r) for o, r in zip(g.o, ref))
        for _ in range(3): g.launch(v)
        burst = timed(lambda: g.launch(v), 20)
        rec = {"gemm": g.name, "variant": label, "raster": v, "bit_identical_to_default": same, "burst_us": round(burst * 1e3, 1),
               "burst_tflops": round(g.flops() / burst / 1e9, 1)}
        out["results"].append(rec)
        print(json.dumps(rec), flush=True)
# the in-loop pattern: up-projection writes the 214 MB hidden, the down-projection reads it (weights cold in between: two blocks' worth)
up2 = Gemm("up (second block)", 4 * D, D, nv.EPI_BIAS_GELU_SIGMOID)
down2 = Gemm("down (second block)", D, 4 * D, nv.EPI_GATE_RESIDUAL)
down.a, down2.a = up.o, up2.o                   # the down-projection reads what the up-projection just wrote
for label, v in DOWN:
    def pair():
        up.launch("0,0,0,0,0"); down.launch(v); up2.launch("0,0,0,0,0"); down2.launch(v)
    for _ in range(10): pair()
    ms = timed(pair, 150) / 2
    rec = {"loop": "up -> down, two alternating weight sets, 300 pairs (~0.3 s at the sustained clock)", "down_variant": label, "us_per_pair": round(ms * 1e3, 1)}
    out["results"].append(rec)
    print(json.dumps(rec), flush=True)
nat.check_async()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_raster_ab.json", "w"), indent=1)
