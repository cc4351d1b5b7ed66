#!/usr/bin/env python
"""GPU bring-up probe: runs every kernel variant of libpe_b200 in its OWN subprocess with a timeout
(a wedged variant must not take the others down), compares with a torch fp32 computation of the same
op on the GPU, times the big shapes with CUDA events and appends one JSON line per case to
gpurun_out/probe.jsonl.

    python tools/gpu_probe.py all            # driver: spawns the cases below
    python tools/gpu_probe.py case <name>    # one case in this process
"""
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def rel_l2(a, b):
    import torch
    a = a.float(); b = b.float()
    return (torch.linalg.vector_norm(a - b) / (torch.linalg.vector_norm(b) + 1e-30)).item()


def time_cuda(fn, iters=10, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ------------------------------------------------------------------------------------------------
def case_gemm(name, M, N, K, epi="bias", cg=1, M2=0, timing=False):
    import torch
    from physicedit_b200 import native as nv
    nat = nv.Native.get(0)
    torch.manual_seed(0)
    dev = "cuda"
    segs_in = []
    for m in [M] + ([M2] if M2 else []):
        a = (torch.randn(m, K, device=dev) * 1.0).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        b = (torch.randn(N, device=dev) * 0.5).bfloat16()
        segs_in.append((a, w, b))
    flags = nv.GEMM_FLAG_CTA_PAIR if cg == 2 else 0
    res = {"name": name, "M": M, "N": N, "K": K, "epi": epi, "cg": cg}
    errs = []
    segs = []
    refs = []
    H = N // 384 if epi == "qkv" else 0
    for (a, w, b) in segs_in:
        m = a.shape[0]
        y = (a.float() @ w.float().t() + b.float())
        yb = y.bfloat16()
        if epi == "bias":
            out = torch.zeros(m, N, device=dev, dtype=torch.bfloat16)
            segs.append(dict(a=a, w=w, bias=b, out=out)); refs.append((out, yb, y))
        elif epi == "gelu":
            out = torch.zeros(m, N, device=dev, dtype=torch.bfloat16)
            ref = yb * torch.sigmoid(1.702 * yb)
            segs.append(dict(a=a, w=w, bias=b, out=out)); refs.append((out, ref, y * torch.sigmoid(1.702 * y)))
        elif epi == "gelu_erf":
            out = torch.zeros(m, N, device=dev, dtype=torch.bfloat16)
            ref = torch.nn.functional.gelu(yb)
            segs.append(dict(a=a, w=w, bias=b, out=out)); refs.append((out, ref, torch.nn.functional.gelu(y)))
        elif epi == "silu":
            out = torch.zeros(m, N, device=dev, dtype=torch.bfloat16)
            ref = torch.nn.functional.silu(yb)
            segs.append(dict(a=a, w=w, bias=b, out=out)); refs.append((out, ref, torch.nn.functional.silu(y)))
        elif epi == "gate":
            resid = torch.randn(m, N, device=dev).bfloat16()
            gate = torch.randn(N, device=dev).bfloat16()
            out = resid.clone()
            ref = resid + gate * yb
            segs.append(dict(a=a, w=w, bias=b, out=out, gate=gate)); refs.append((out, ref, resid.float() + gate.float() * y))
        elif epi == "qkv":
            d = H * 128
            oq = torch.zeros(m, d, device=dev, dtype=torch.bfloat16); ok = torch.zeros_like(oq); ov = torch.zeros_like(oq)
            nq = (1 + 0.1 * torch.randn(128, device=dev)).bfloat16(); nk = (1 + 0.1 * torch.randn(128, device=dev)).bfloat16()
            ang = torch.rand(m, 64, device=dev) * 6.28
            rope = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()   # float2 (cos, sin) [m, 64]
            segs.append(dict(a=a, w=w, bias=b, out=oq, out_k=ok, out_v=ov, norm_q_w=nq, norm_k_w=nk, rope=rope))

            def normrope(t, wn):
                t = t.view(m, H, 128)
                var = t.float().square().mean(-1, keepdim=True)
                t = (t * torch.rsqrt(var + 1e-6)).to(torch.bfloat16) * wn
                tc = torch.view_as_complex(t.float().reshape(m, H, 64, 2))
                fc = torch.view_as_complex(rope).unsqueeze(1)
                return torch.view_as_real(tc * fc).flatten(2).to(torch.bfloat16).reshape(m, d)
            q_ref = normrope(yb[:, :d], nq); k_ref = normrope(yb[:, d:2 * d], nk); v_ref = yb[:, 2 * d:]
            refs.append((oq, q_ref, q_ref)); refs.append((ok, k_ref, k_ref)); refs.append((ov, v_ref, v_ref))
    epi_code = {"bias": nv.EPI_BIAS, "gelu": nv.EPI_BIAS_GELU_SIGMOID, "gelu_erf": nv.EPI_BIAS_GELU_ERF, "silu": nv.EPI_BIAS_SILU,
                "gate": nv.EPI_GATE_RESIDUAL, "qkv": nv.EPI_QKV_NORM_ROPE}[epi]
    nat.gemm(segs, N, K, epi_code, flags)
    nat.check_async()
    for out, ref_b, ref_f in refs:
        errs.append(rel_l2(out, ref_f))
        res.setdefault("max_abs", []).append((out.float() - ref_b.float()).abs().max().item())
        res.setdefault("mismatch_frac", []).append((out != ref_b).float().mean().item())
    res["rel_l2"] = errs
    res["ok"] = bool(all(e < 6e-3 for e in errs))
    if timing:
        if epi == "gate":
            pass  # in-place accumulation on out is fine for timing
        ms = time_cuda(lambda: nat.gemm(segs, N, K, epi_code, flags))
        nat.check_async()
        flops = 2.0 * sum(a.shape[0] for a, _, _ in segs_in) * N * K
        res["ms"] = ms
        res["tflops"] = flops / ms / 1e9
        a, w, b = segs_in[0]
        ms_t = time_cuda(lambda: torch.nn.functional.linear(a, w, b))
        res["torch_ms_seg0"] = ms_t
        res["torch_tflops_seg0"] = 2.0 * a.shape[0] * N * K / ms_t / 1e9
    return res


def case_attn(name, S, H, flags=0, timing=False):
    import torch
    from physicedit_b200 import native as nv
    nat = nv.Native.get(0)
    torch.manual_seed(0)
    dev = "cuda"
    d = H * 128
    q = torch.randn(S, d, device=dev).bfloat16()
    k = torch.randn(S, d, device=dev).bfloat16()
    v = torch.randn(S, d, device=dev).bfloat16()
    o = torch.zeros(S, d, device=dev, dtype=torch.bfloat16)
    scale = 1.0 / math.sqrt(128)
    nat.attention(q, k, v, o, H, scale, flags)
    nat.check_async()
    qh = q.view(S, H, 128).transpose(0, 1).float(); kh = k.view(S, H, 128).transpose(0, 1).float(); vh = v.view(S, H, 128).transpose(0, 1).float()
    hs = min(H, 4)   # reference on a few heads (fp32, exact softmax)
    ref = torch.softmax(qh[:hs] @ kh[:hs].transpose(1, 2) * scale, dim=-1) @ vh[:hs]
    ref = ref.transpose(0, 1).reshape(S, hs * 128)
    err = rel_l2(o[:, :hs * 128], ref)
    res = {"name": name, "S": S, "H": H, "flags": flags, "rel_l2": err, "ok": bool(err < 1e-2),
           "max_abs": (o[:, :hs * 128].float() - ref).abs().max().item()}
    if H > hs:
        # remaining heads against torch SDPA (bf16)
        sd = torch.nn.functional.scaled_dot_product_attention(q.view(S, H, 128).transpose(0, 1)[None], k.view(S, H, 128).transpose(0, 1)[None],
                                                              v.view(S, H, 128).transpose(0, 1)[None])[0].transpose(0, 1).reshape(S, d)
        res["rel_l2_vs_sdpa_all_heads"] = rel_l2(o, sd)
        res["ok"] = res["ok"] and res["rel_l2_vs_sdpa_all_heads"] < 1.5e-2
    if timing:
        ms = time_cuda(lambda: nat.attention(q, k, v, o, H, scale, flags))
        nat.check_async()
        res["ms"] = ms
        res["tflops"] = 4.0 * S * S * 128 * H / ms / 1e9
        q4 = q.view(1, S, H, 128).transpose(1, 2); k4 = k.view(1, S, H, 128).transpose(1, 2); v4 = v.view(1, S, H, 128).transpose(1, 2)
        ms_t = time_cuda(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4))
        res["torch_sdpa_ms"] = ms_t
        res["torch_sdpa_tflops"] = 4.0 * S * S * 128 * H / ms_t / 1e9
    return res


def case_rowwise(name):
    import torch
    from physicedit_b200 import native as nv
    nat = nv.Native.get(0)
    torch.manual_seed(0)
    dev = "cuda"
    res = {"name": name, "sub": {}}
    ok_all = True

    def rec(k, got, ref, exact=False, tol=1e-2):
        nonlocal ok_all
        mism = (got != ref).float().mean().item()
        e = rel_l2(got, ref)
        ok = (mism == 0.0) if exact else (e < tol)
        res["sub"][k] = {"rel_l2": e, "mismatch_frac": mism, "ok": bool(ok)}
        ok_all = ok_all and ok

    # layernorm + modulate
    for rows, C in ((1000, 3072), (77, 3584), (13, 768), (5, 64)):
        x = (torch.randn(rows, C, device=dev) * 2 + 0.3).bfloat16()
        shift = torch.randn(C, device=dev).bfloat16(); scale = torch.randn(C, device=dev).bfloat16()
        ops = 1 + scale
        out = torch.empty_like(x)
        nat.layernorm_modulate(x, out, shift, ops)
        ref = torch.nn.functional.layer_norm(x, (C,), eps=1e-6) * (1 + scale) + shift
        rec(f"ln_mod_{rows}x{C}", out, ref, tol=4e-3)
        w = torch.randn(C, device=dev).bfloat16(); b = torch.randn(C, device=dev).bfloat16()
        nat.layernorm(x, out, w, b, 1e-5)
        rec(f"ln_affine_{rows}x{C}", out, torch.nn.functional.layer_norm(x, (C,), w, b, 1e-5), tol=4e-3)
        nat.layernorm(x, out, None, None, 1e-6)
        rec(f"ln_plain_{rows}x{C}", out, torch.nn.functional.layer_norm(x, (C,), eps=1e-6), tol=4e-3)
        nat.rmsnorm(x, out, w, 1e-6)
        var = x.float().square().mean(-1, keepdim=True)
        rec(f"rms_{rows}x{C}", out, (x * torch.rsqrt(var + 1e-6)).to(torch.bfloat16) * w, tol=4e-3)
    # gemv
    for batch, N, K, ai, ao in ((1, 18432, 3072, 1, 0), (1, 3072, 256, 0, 1), (2, 6144, 3072, 1, 0), (1, 1000, 3584, 0, 0)):
        x = torch.randn(batch, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16(); b = torch.randn(N, device=dev).bfloat16()
        y = torch.empty(batch, N, device=dev, dtype=torch.bfloat16)
        mask = torch.zeros(N, dtype=torch.uint8, device=dev); mask[N // 6: N // 3] = 1
        nat.gemv(x, w, b, y, ai, ao, mask)
        xi = torch.nn.functional.silu(x) if ai else x
        ref = torch.nn.functional.linear(xi.float(), w.float(), b.float()).bfloat16()
        if ao:
            ref = torch.nn.functional.silu(ref)
        ref = torch.where(mask.bool()[None], 1 + ref, ref)
        rec(f"gemv_{batch}x{N}x{K}", y, ref, tol=4e-3)
    # patchify / unpatchify
    from einops import rearrange
    lat = torch.randn(16, 64, 96, device=dev).bfloat16()
    tok = torch.empty(32 * 48, 64, device=dev, dtype=torch.bfloat16)
    nat.patchify(lat, tok)
    rec("patchify", tok, rearrange(lat[None], "B C (H P) (W Q) -> B (H W) (C P Q)", P=2, Q=2)[0], exact=True)
    lat2 = torch.empty_like(lat)
    nat.unpatchify(tok, lat2)
    rec("unpatchify", lat2, lat, exact=True)
    # cfg + euler
    n = 16 * 128 * 128
    latv = torch.randn(n, device=dev).bfloat16(); posi = torch.randn(n, device=dev).bfloat16(); nega = torch.randn(n, device=dev).bfloat16()
    ds = torch.tensor(-0.0123456, dtype=torch.float32)
    ref = latv + (nega + 4.0 * (posi - nega)) * ds
    nat.cfg_euler_step(latv, posi, nega, 4.0, float(ds))
    rec("cfg_euler", latv, ref, exact=True)
    # timestep embedding: compare with the reference formula evaluated by torch on the GPU
    for tval in (1000.0, 989.7009, 500.0, 20.0):
        t = torch.tensor([tval], device=dev).bfloat16()
        out = torch.empty(256, device=dev, dtype=torch.bfloat16)
        nat.timestep_embedding(t, out)
        ts = t / 1000
        exponent = -math.log(10000) * torch.arange(0, 128, dtype=torch.float32, device=dev) / 128
        emb = torch.exp(exponent).to(ts.dtype)
        emb = 1000 * (ts[:, None].float() * emb[None, :])
        ref = torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1).to(torch.bfloat16)[0]
        rec(f"temb_{tval}", out, ref, exact=True)
    # special gather / blend-scatter
    T, C = 300, 3584
    pe = torch.randn(T, C, device=dev).bfloat16()
    mask = torch.zeros(T, dtype=torch.bool, device=dev); mask[T - 69:T - 5] = True
    dst = torch.empty(64, C, device=dev, dtype=torch.bfloat16); idx = torch.empty(65, device=dev, dtype=torch.int32)
    nat.special_gather(pe, mask.to(torch.uint8), dst, idx)
    rec("special_gather", dst, pe[mask], exact=True)
    res["sub"]["special_count"] = {"ok": bool(idx[64].item() == 64), "count": int(idx[64].item())}
    ok_all = ok_all and idx[64].item() == 64
    pd = torch.randn(64, C, device=dev).bfloat16(); pv = torch.randn(64, C, device=dev).bfloat16()
    t = torch.tensor([744.0], device=dev).bfloat16()
    t_min, t_max = 19.999980926513672, 1000.0
    alpha = ((t - t_min) / (t_max - t_min + 1e-6)).clamp(0.0, 1.0).view(-1, 1)
    ref_pe = pe.clone(); ref_pe[mask] = alpha * pd + (1 - alpha) * pv
    nat.special_blend_scatter(pe, idx, pd, pv, t, t_min, t_max)
    rec("special_blend_scatter", pe, ref_pe, exact=True)
    # small attention
    for (B, Hh, Sq, Skv, D) in ((2, 12, 261, 261, 64), (1, 8, 64, 1100, 64)):
        q = torch.randn(B * Sq, Hh * D, device=dev).bfloat16(); k = torch.randn(B * Skv, Hh * D, device=dev).bfloat16()
        v = torch.randn(B * Skv, Hh * D, device=dev).bfloat16(); o = torch.empty_like(q)
        nat.small_attention(q, k, v, o, B, Hh, Sq, Skv, D, D ** -0.5)
        ref = torch.nn.functional.scaled_dot_product_attention(q.view(B, Sq, Hh, D).transpose(1, 2).float(), k.view(B, Skv, Hh, D).transpose(1, 2).float(),
                                                               v.view(B, Skv, Hh, D).transpose(1, 2).float()).transpose(1, 2).reshape(B * Sq, Hh * D)
        rec(f"small_attn_{B}x{Hh}x{Sq}x{Skv}", o, ref, tol=5e-3)
    x = torch.randn(12, 64, device=dev).bfloat16(); add = torch.randn(4, 64, device=dev).bfloat16()
    ref = (x.float() + 1.0 * add.float().repeat(3, 1)).bfloat16()
    nat.add_rows(x, add, 4, 1.0)
    rec("add_rows", x, ref, exact=True)
    nat.check_async()
    res["ok"] = bool(ok_all)
    return res


CASES = {
    "rowwise": lambda: case_rowwise("rowwise"),
    "gemm_cg1_small": lambda: case_gemm("gemm_cg1_small", 256, 512, 128),
    "gemm_cg1_ragged": lambda: case_gemm("gemm_cg1_ragged", 300, 264, 192, M2=77),
    "gemm_cg1_big": lambda: case_gemm("gemm_cg1_big", 8192, 3072, 3072, M2=512, timing=True),
    "gemm_cg1_up": lambda: case_gemm("gemm_cg1_up", 8192, 12288, 3072, epi="gelu", M2=512, timing=True),
    "gemm_cg1_down": lambda: case_gemm("gemm_cg1_down", 8192, 3072, 12288, epi="gate", M2=512, timing=True),
    "gemm_cg1_qkv": lambda: case_gemm("gemm_cg1_qkv", 8192, 9216, 3072, epi="qkv", M2=512, timing=True),
    "gemm_cg1_qkv_small": lambda: case_gemm("gemm_cg1_qkv_small", 200, 768, 256, epi="qkv"),
    "gemm_cg1_gelu_erf": lambda: case_gemm("gemm_cg1_gelu_erf", 64, 10752, 3584, epi="gelu_erf", timing=True),
    "gemm_cg1_silu": lambda: case_gemm("gemm_cg1_silu", 64, 512, 256, epi="silu"),
    "gemm_cg1_k64_n64": lambda: case_gemm("gemm_cg1_k64_n64", 1000, 64, 3072),
    "gemm_cg2_small": lambda: case_gemm("gemm_cg2_small", 512, 512, 128, cg=2),
    "gemm_cg2_ragged": lambda: case_gemm("gemm_cg2_ragged", 300, 264, 192, cg=2, M2=77),
    "gemm_cg2_big": lambda: case_gemm("gemm_cg2_big", 8192, 3072, 3072, cg=2, M2=512, timing=True),
    "gemm_cg2_up": lambda: case_gemm("gemm_cg2_up", 8192, 12288, 3072, epi="gelu", cg=2, M2=512, timing=True),
    "gemm_cg2_down": lambda: case_gemm("gemm_cg2_down", 8192, 3072, 12288, epi="gate", cg=2, M2=512, timing=True),
    "gemm_cg2_qkv": lambda: case_gemm("gemm_cg2_qkv", 8192, 9216, 3072, epi="qkv", cg=2, M2=512, timing=True),
}
for _q, _qn in ((1, "q1"), (0, "q2")):
    for _p, _pn in ((2, "psmem"), (0, "ptmem")):
        for _s, _sn in ((0, ""),):
            _f = _q | _p | _s
            CASES[f"attn_{_qn}_{_pn}{_sn}_small"] = (lambda f=_f, n=f"attn_{_qn}_{_pn}{_sn}_small": case_attn(n, 256, 2, f))
            CASES[f"attn_{_qn}_{_pn}{_sn}_ragged"] = (lambda f=_f, n=f"attn_{_qn}_{_pn}{_sn}_ragged": case_attn(n, 1000, 3, f))
            CASES[f"attn_{_qn}_{_pn}{_sn}_big"] = (lambda f=_f, n=f"attn_{_qn}_{_pn}{_sn}_big": case_attn(n, 8704, 24, f, timing=True))


CASES["attn_v2_small"] = lambda: case_attn("attn_v2_small", 256, 2, 8)
CASES["attn_v2_ragged"] = lambda: case_attn("attn_v2_ragged", 1000, 3, 8)
CASES["attn_v2_big"] = lambda: case_attn("attn_v2_big", 8704, 24, 8, timing=True)
CASES["attn_v2_big_t288"] = lambda: case_attn("attn_v2_big_t288", 8480, 24, 8, timing=True)
CASES["attn_v1_big"] = lambda: case_attn("attn_v1_big", 8704, 24, 0, timing=True)
CASES["attn_v2_2048"] = lambda: case_attn("attn_v2_2048", 20992, 24, 8, timing=True)
CASES["attn_v1_2048"] = lambda: case_attn("attn_v1_2048", 20992, 24, 0, timing=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1] == "case":
        name = sys.argv[2]
        try:
            res = CASES[name]()
        except Exception as e:  # noqa: BLE001
            res = {"name": name, "ok": False, "error": f"{type(e).__name__}: {e}"[:600]}
        print("PROBE " + json.dumps(res), flush=True)
        return
    names = sys.argv[2:] if len(sys.argv) > 2 else list(CASES)
    if sys.argv[1] == "all":
        names = [n for n in names if n in CASES]
    log = open(os.path.join(OUT, "probe.jsonl"), "a")
    for name in names:
        t0 = time.time()
        try:
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), "case", name], capture_output=True, text=True, timeout=180)
            line = [l for l in cp.stdout.splitlines() if l.startswith("PROBE ")]
            if line:
                res = json.loads(line[-1][6:])
            else:
                res = {"name": name, "ok": False, "error": "no result", "rc": cp.returncode, "stderr": cp.stderr[-800:]}
        except subprocess.TimeoutExpired:
            res = {"name": name, "ok": False, "error": "timeout 180s"}
        res["wall_s"] = round(time.time() - t0, 1)
        log.write(json.dumps(res) + "\n"); log.flush()
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
