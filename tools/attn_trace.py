#!/usr/bin/env python
"""Per-step timeline of the attention kernel (needs a -DPE_ATTN_TRACE build selected with PE_B200_LIB).
Prints, for CTA 0's first work item, the cycle stamps of the MMA-issuer and softmax events of each KV step."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from physicedit_b200 import native as nv
nat = nv.Native.get(0)
S, H = 8704, 24
q, k, v = (torch.randn(S, H * 128, device="cuda").bfloat16() for _ in range(3))
o = torch.empty_like(q)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(2):
    nat.attention(q, k, v, o, H, 1 / math.sqrt(128), flags)
nat.check_async()
w = nat.workspace_read(2 + 3 * 4000)
ev = []
for role in range(3):
    for i in range(2000):
        a, b = int(w[2 + role * 4000 + 2 * i]), int(w[3 + role * 4000 + 2 * i])
        if b == 0:
            break
        ev.append((a >> 32, a & 0xffffffff, b))
names = {10: "mma:wait_p0", 11: "mma:wait_p1", 12: "mma:got_p0", 13: "mma:got_p1", 14: "mma:issued0", 15: "mma:issued1",
         20: "sm0:wait_s", 21: "sm1:wait_s", 22: "sm0:got_s", 23: "sm1:got_s", 24: "sm0:max_ok", 25: "sm1:max_ok", 26: "sm0:pass_done", 27: "sm1:pass_done",
         28: "sm0:arrived", 29: "sm1:arrived", 30: "sm0:loaded", 31: "sm1:loaded"}
t0 = min(e[2] for e in ev)
first = [e for e in ev if e[1] < 68]
# keep only the first item's events: the first occurrence of each (event, step)
seen, rows = set(), []
for e in sorted(first, key=lambda x: x[2]):
    if (e[0], e[1]) in seen:
        continue
    seen.add((e[0], e[1]))
    rows.append(e)
for step in (3, 4, 5, 20, 21):
    print(f"--- step {step}")
    for e in rows:
        if e[1] == step:
            print(f"   {str(names.get(e[0], e[0])):16s} {e[2] - t0:9d}")
# summary: mean per-step period and component durations over steps 8..60
def T(evn, st):
    for e in rows:
        if e[0] == evn and e[1] == st:
            return e[2]
    return None
per, sm, wait_s, mma_wait, issue = [], [], [], [], []
for st in range(8, 60):
    if T(22, st) and T(22, st + 1):
        per.append(T(22, st + 1) - T(22, st))
    if T(22, st) and T(28, st):
        sm.append(T(28, st) - T(22, st))
    if T(20, st) and T(22, st):
        wait_s.append(T(22, st) - T(20, st))
    if T(10, st) and T(12, st):
        mma_wait.append(T(12, st) - T(10, st))
    if T(12, st) and T(14, st):
        issue.append(T(14, st) - T(12, st))
mean = lambda x: sum(x) / max(len(x), 1)
print(f"period/step {mean(per):.0f} cyc | softmax0 busy {mean(sm):.0f} | softmax0 waits for S {mean(wait_s):.0f} | mma waits for P0 {mean(mma_wait):.0f} | mma issue PV0+S0 {mean(issue):.0f}")
# finer MMA-issuer events (attention_kernel3): 40/41 = PV issued, 42/43 = S issued
for st in (20, 21):
    print(f"--- issuer step {st}:", [(names.get(e[0], e[0]), e[2] - t0) for e in rows if e[1] == st and (e[0] < 20 or e[0] >= 40)])
