"""Isolated timing of pe_attention_fwd at the 1024^2 shape (S=8704, 24 heads).  Usage: attn_time.py [flags ...]"""
import sys, math, os, torch
sys.path.insert(0, os.getcwd())
from physicedit_b200 import native as nv
nat = nv.Native.get(0)
S, H = int(os.environ.get("ATTN_S", "8704")), 24
q, k, v = (torch.randn(S, H*128, device="cuda").bfloat16() for _ in range(3))
o = torch.empty_like(q)
for flags in [int(a) for a in sys.argv[1:]] or [0]:
    for _ in range(3): nat.attention(q, k, v, o, H, 1/math.sqrt(128), flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): nat.attention(q, k, v, o, H, 1/math.sqrt(128), flags)
    e1.record(); torch.cuda.synchronize()
    nat.check_async()
    ms = e0.elapsed_time(e1)/10
    print(os.environ.get("PE_B200_LIB","default").split("/")[-1], "flags", flags, "S", S, round(ms,4), "ms", round(4*S*S*128*H/ms/1e9,1), "TF", flush=True)
