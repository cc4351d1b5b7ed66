"""Localises an attention-kernel error: per d-half / per query tile / per KV-step relative errors against fp32 softmax."""
import sys, math, os, torch
sys.path.insert(0, os.getcwd())
from physicedit_b200 import native as nv
nat = nv.Native.get(0)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 16
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()
for S in [int(a) for a in sys.argv[2:]] or [64, 128, 192, 256, 320]:
    torch.manual_seed(S)
    q, k, v = (torch.randn(S, 128, device="cuda").bfloat16() for _ in range(3))
    o = torch.zeros_like(q)
    nat.attention(q, k, v, o, 1, 1 / math.sqrt(128), flags)
    nat.check_async()
    P = torch.softmax(q.float() @ k.float().t() / math.sqrt(128), dim=-1)
    ref = P @ v.float()
    line = f"S={S}: all {rel(o, ref):.3e} | d<64 {rel(o[:, :64], ref[:, :64]):.3e} d>=64 {rel(o[:, 64:], ref[:, 64:]):.3e}"
    for t in range(0, S, 128):
        line += f" | rows {t}.. {rel(o[t:t+128], ref[t:t+128]):.3e}"
    print(line)
    # which KV steps are represented?  least squares: o ~ sum_j a_j * (P[:, j-th 64 block] @ v[block])
    nb = (S + 63) // 64
    parts = torch.stack([(P[:, j*64:(j+1)*64] @ v[j*64:(j+1)*64].float()).flatten() for j in range(nb)], 1)
    sol = torch.linalg.lstsq(parts, o.float().flatten()[:, None]).solution.flatten()
    print("   per-KV-block weights (1 = correct):", [round(x, 3) for x in sol.tolist()])
    # within block 0: weight of each 16-row k-step
    parts = torch.stack([(P[:, j*16:(j+1)*16] @ v[j*16:(j+1)*16].float()).flatten() for j in range(min(S, 128) // 16)], 1)
    rest = (P[:, min(S, 128):] @ v[min(S, 128):].float()).flatten() if S > 128 else 0
    sol = torch.linalg.lstsq(parts, (o.float().flatten() - rest)[:, None]).solution.flatten()
    print("   per-16-row weights in the first 128 kv rows:", [round(x, 3) for x in sol.tolist()])
