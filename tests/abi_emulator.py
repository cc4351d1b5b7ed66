"""TEST INFRASTRUCTURE: a torch-CPU emulation of the C-ABI entry points the VAE host code calls (include/pe_b200.h, last section).

It restates each entry point's documented contract (argument layout, padding rules, epilogues) with plain tensor ops so that the
HOST logic of physicedit_b200/vae.py -- weight repacking, tap order, space-to-depth mapping, buffer strides, the attention
chunking -- can be checked against the oracle without a GPU.  It is never imported by the product; the GPU tests run the same host
code on the real library.  Arithmetic is fp32 with bf16 rounding at the points the header documents."""
import torch
import torch.nn.functional as F

EPI_BIAS, EPI_GATE_RESIDUAL, EPI_F32 = 0, 4, 7


def _r(x):
    return x.to(torch.bfloat16).float()


class EmulatedNative:
    def __init__(self):
        self.tag = None
        self.launches = 0
        self.calls = []

    def _note(self, what):
        self.calls.append((what, self.tag))
        self.tag = None
        self.launches += 1

    # out = epi(A . W^T + bias), A [M, K] (row stride = a.stride(0)), W [N, K]
    def gemm(self, segs, N, K, epilogue=EPI_BIAS, flags=0):
        for s in segs:
            a, w, out = s["a"], s["w"], s["out"]
            assert a.shape[1] >= K and w.shape == (N, K) and w.is_contiguous() and N % 8 == 0 and K % 8 == 0
            assert a.stride(0) >= K and out.shape[0] == a.shape[0]
            acc = a[:, :K].float() @ w.float().t()
            if epilogue == EPI_F32:
                assert out.dtype == torch.float32
                out[:, :N] = acc
                continue
            if s.get("bias") is not None:
                acc = acc + s["bias"].float()[:N]
            y = _r(acc)
            if epilogue == 3:                                   # PE_EPI_BIAS_GELU_ERF
                y = 0.5 * y * (1.0 + torch.erf(y * 0.70710678118654752))
            if epilogue == EPI_GATE_RESIDUAL:
                y = out[:, :N].float() + _r(s["gate"].float()[:N] * y)
            out[:, :N] = y.to(torch.bfloat16)
        self._note("pe_gemm")

    def rmsnorm(self, x, out, w, eps=1e-6):
        """models/utils.py:250-257: bf16(bf16(x * rsqrt(mean(x^2) + eps)) * w)."""
        v = x.float()
        y = _r(v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps))
        out.copy_((y * w.float() if w is not None else y).to(torch.bfloat16))
        self._note("pe_rmsnorm")

    def add_rows(self, x, add, period, alpha=1.0):
        """x[r, :] = bf16(x[r, :] + alpha * add[r % period, :]) in place."""
        rows = torch.arange(x.shape[0]) % period
        x.copy_((x.float() + alpha * add.float()[rows]).to(torch.bfloat16))
        self._note("pe_add_rows")

    def conv2d(self, x, H, W, C, w, bias, out, N, kh, kw, pad, epilogue=EPI_BIAS, gate=None, flags=0):
        cpad = (C + 63) // 64 * 64
        assert x.shape[0] == H * W and x.stride(0) >= C and w.shape == (N, kh * kw * cpad) and N % 8 == 0 and C % 8 == 0
        xm = x[:, :C].float().reshape(H, W, C).permute(2, 0, 1).unsqueeze(0)
        wk = w.float().reshape(N, kh, kw, cpad)[..., :C].permute(0, 3, 1, 2)
        xm = F.pad(xm, (pad, kw - 1 - pad, pad, kh - 1 - pad))          # zeros outside the map on every side
        acc = F.conv2d(xm, wk)[0].permute(1, 2, 0).reshape(H * W, N)
        if bias is not None:
            acc = acc + bias.float()[:N]
        y = _r(acc)
        if epilogue == EPI_GATE_RESIDUAL:
            y = out[:, :N].float() + _r(gate.float()[:N] * y)
        out[:, :N] = y.to(torch.bfloat16)
        self._note("pe_conv2d")

    def channel_rmsnorm(self, x, out, C, gamma, act):
        v = x[:, :C].float()
        n = _r(v.pow(2).sum(-1, keepdim=True).sqrt()).clamp_min(1e-12)
        y = _r(_r(_r(v / n) * float(torch.tensor(C ** 0.5, dtype=torch.float32))) * gamma.float())
        if act:
            y = y / (1 + torch.exp(-y))
        out[:, :C] = y.to(torch.bfloat16)
        self._note("pe_channel_rmsnorm")

    def upsample2x(self, x, out, H, W, C):
        m = x.reshape(H, W, C)
        out.copy_(m.repeat_interleave(2, 0).repeat_interleave(2, 1).reshape(4 * H * W, C))
        self._note("pe_upsample2x")

    def space_to_depth(self, x, out, H, W, C):
        m = x.reshape(H // 2, 2, W // 2, 2, C).permute(0, 2, 1, 3, 4)          # [y, x, py, px, c]
        out.copy_(m.reshape(H * W // 4, 4 * C))
        self._note("pe_space_to_depth")

    @staticmethod
    def _affine(v, op, p0, p1):
        if op == 1:
            return _r(_r(v / p1.float()) + p0.float())
        if op == 2:
            return _r(_r(v - p0.float()) * p1.float())
        return v

    def nchw_to_nhwc(self, src, dst, C, op=0, p0=None, p1=None):
        v = src.reshape(C, -1).t().float()
        dst[:, :C] = self._affine(v, op, p0, p1).to(torch.bfloat16)
        self._note("pe_nchw_to_nhwc")

    def nhwc_to_nchw(self, src, dst, C, op=0, p0=None, p1=None):
        v = self._affine(src[:, :C].float(), op, p0, p1)
        dst.copy_(v.t().reshape(dst.shape).to(torch.bfloat16))
        self._note("pe_nhwc_to_nchw")

    def transpose(self, src, dst):
        dst.copy_(src.t())
        self._note("pe_transpose")

    def softmax_rows(self, scores, probs, n, scale, mask=None):
        sc = scores[:, :n].float() * scale
        if mask is not None:                                   # byte mask [period, >= n], 0 = hidden; score row r uses mask row r % period
            rows = torch.arange(scores.shape[0]) % mask.shape[0]
            sc = sc.masked_fill(mask[rows, :n] == 0, float("-inf"))
        p = torch.softmax(sc, dim=-1)
        probs.zero_()
        probs[:, :n] = p.to(torch.bfloat16)
        self._note("pe_softmax_rows")

    # ---- training path (include/pe_b200.h: pe_gemm_batched, pe_attention_fwd_lse, pe_attention_bwd_delta) ------------------------------
    def linear(self, x, w, bias, epilogue=EPI_BIAS, flags=0):
        out = torch.empty(x.shape[0], w.shape[0], dtype=torch.bfloat16)
        self.gemm([dict(a=x, w=w, bias=bias, out=out)], w.shape[0], w.shape[1], epilogue, flags)
        return out

    def attention_lse(self, q, k, v, o, lse, H, scale, flags=0):
        """o = softmax(scale q k^T) v per head (token-major [S, H * 128]); lse[h, s] = log2 sum_j exp2(scale log2(e) s_j)."""
        S = q.shape[0]
        hm = lambda t: t.float().view(S, H, -1).transpose(0, 1)
        sc = hm(q) @ hm(k).transpose(1, 2) * scale
        lse.copy_(torch.logsumexp(sc, dim=-1) * 1.4426950408889634)
        p = _r(torch.softmax(sc, dim=-1))                               # the kernel feeds bf16 probabilities to the P V product
        o.copy_((p @ hm(v)).transpose(0, 1).reshape(S, -1).to(torch.bfloat16))
        self._note("pe_attention_fwd_lse")

    def attention_bwd_delta(self, d_o, o, delta, H):
        S = o.shape[0]
        delta[:, :S] = (d_o.float() * o.float()).view(S, H, -1).sum(-1).t()
        self._note("pe_attention_bwd_delta")

    def gemm_batched(self, a, w, out, batch, M, N, K, a_batch_rows, w_batch_rows, out_batch_rows, epilogue=EPI_BIAS, vec=None, vec_batch_stride=0,
                     vec_per_column=False, alpha=1.0, flags=0):
        """Problem b: rows [b * a_batch_rows, + M) of a times rows [b * w_batch_rows, + N) of w -> rows [b * out_batch_rows, + M) of out; only those rows /
        the first N columns are written.  Epilogues: 0 plain, 7 fp32, 8 exp2(acc * alpha - vec[i]), 9 out * (acc - vec[i]) * alpha in place."""
        assert N % 8 == 0 and K % 8 == 0 and w.stride(0) == K and a.stride(0) >= K
        for b in range(batch):
            acc = a[b * a_batch_rows: b * a_batch_rows + M, :K].float() @ w[b * w_batch_rows: b * w_batch_rows + N, :K].float().t()
            dst = out[b * out_batch_rows: b * out_batch_rows + M]
            if epilogue in (8, 9):
                st = vec.reshape(-1)[b * vec_batch_stride:]
                st = st[:N][None, :] if vec_per_column else st[:M][:, None]
                val = torch.exp2(acc * alpha - st) if epilogue == 8 else dst[:, :N].float() * (acc - st) * alpha
                dst[:, :N] = val.to(torch.bfloat16)
            elif epilogue == EPI_F32:
                dst[:, :N] = acc
            else:
                dst[:, :N] = acc.to(torch.bfloat16)
        self._note("pe_gemm_batched")

    # ---- DiT block (include/pe_b200.h: pe_layernorm_modulate2, pe_gemv, pe_act, pe_attention_fwd, QKV / GELU epilogues of pe_gemm) -----
    def act(self, x, y, act):
        v = x.float()
        y.copy_((_r(v / (1 + torch.exp(-v))) if act == 1 else v).to(torch.bfloat16))
        self._note("pe_act")

    def gemv(self, x, w, bias, y, act_in=0, act_out=0, one_plus_mask=None):
        """y[b] = bf16(x[b] w^T + bias) (+ SiLU in / out); entries flagged in one_plus_mask come out as bf16(1 + y)."""
        v = x.float()
        if act_in == 1:
            v = _r(v / (1 + torch.exp(-v)))
        o = _r(v @ w.float().t() + (bias.float() if bias is not None else 0.0))
        if act_out == 1:
            o = _r(o / (1 + torch.exp(-o)))
        if one_plus_mask is not None:
            o = torch.where(one_plus_mask.bool()[None, :], _r(1.0 + o), o)
        y.copy_(o.to(torch.bfloat16))
        self._note("pe_gemv")

    def layernorm_modulate2(self, x, out, split_row, shift0, ops0, shift1, ops1):
        """rows < split_row: bf16(bf16(bf16(LN(x)) * ops0) + shift0); the others with (shift1, ops1); LN without affine, eps 1e-6, fp32 statistics."""
        v = x.float()
        n = _r((v - v.mean(-1, keepdim=True)) * torch.rsqrt(v.var(-1, unbiased=False, keepdim=True) + 1e-6))
        first = (torch.arange(x.shape[0]) < split_row)[:, None]
        ops = torch.where(first, ops0.float()[None], ops1.float()[None])
        sh = torch.where(first, shift0.float()[None], shift1.float()[None])
        out.copy_(_r(_r(n * ops) + sh).to(torch.bfloat16))
        self._note("pe_layernorm_modulate2")

    def attention(self, q, k, v, o, H, scale, flags=0):
        S = q.shape[0]
        hm = lambda t: t.float().view(S, H, -1).transpose(0, 1)
        p = _r(torch.softmax(hm(q) @ hm(k).transpose(1, 2) * scale, dim=-1))
        o.copy_((p @ hm(v)).transpose(0, 1).reshape(S, -1).to(torch.bfloat16))
        self._note("pe_attention_fwd")


def _qkv_epilogue(acc, seg, N):
    """PE_EPI_QKV_NORM_ROPE: acc [M, 3 * H * 128] = (q heads | k heads | v heads) + bias -> bf16; per-head RMSNorm * w and RoPE (fp32 complex multiply with
    the (cos, sin) table [M, 64, 2]) on q and k; three outputs."""
    y = _r(acc + seg["bias"].float()[:N])
    M, HD = y.shape[0], N // 3
    H = HD // 128
    cs = seg["rope"].float()[:M]
    c, s_ = cs[:, None, :, 0], cs[:, None, :, 1]
    for part, dst, wn in ((0, seg["out"], seg["norm_q_w"]), (1, seg["out_k"], seg["norm_k_w"]), (2, seg["out_v"], None)):
        t = y[:, part * HD:(part + 1) * HD].reshape(M, H, 128)
        if wn is not None:
            t = _r(_r(t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6)) * wn.float())
            pr = t.reshape(M, H, 64, 2)
            t = torch.stack((pr[..., 0] * c - pr[..., 1] * s_, pr[..., 0] * s_ + pr[..., 1] * c), dim=-1).reshape(M, H, 128)
        dst[:, :HD] = t.reshape(M, HD).to(torch.bfloat16)


_plain_gemm = EmulatedNative.gemm


def _gemm_with_block_epilogues(self, segs, N, K, epilogue=EPI_BIAS, flags=0):
    if epilogue == 5:                                          # PE_EPI_QKV_NORM_ROPE
        for s in segs:
            _qkv_epilogue(s["a"][:, :K].float() @ s["w"].float().t(), s, N)
        return self._note("pe_gemm")
    if epilogue == 2:                                          # PE_EPI_BIAS_GELU_SIGMOID: h = bf16(acc + bias); out = h * bf16(sigmoid(bf16(1.702 h)))
        for s in segs:
            h = _r(s["a"][:, :K].float() @ s["w"].float().t() + s["bias"].float()[:N])
            s["out"][:, :N] = (h * _r(torch.sigmoid(_r(1.702 * h)))).to(torch.bfloat16)
        return self._note("pe_gemm")
    return _plain_gemm(self, segs, N, K, epilogue, flags)


EmulatedNative.gemm = _gemm_with_block_epilogues


# ---- the rest of a whole DiT forward: pe_patchify / pe_unpatchify / pe_layernorm_modulate / pe_timestep_embedding -----------------------
def _patchify(self, latents, tokens):
    C, H2, W2 = latents.shape[-3:]
    x = latents.reshape(C, H2 // 2, 2, W2 // 2, 2)                 # C H P W Q -> (H W) (C P Q)
    tokens.copy_(x.permute(1, 3, 0, 2, 4).reshape((H2 // 2) * (W2 // 2), C * 4))
    self._note("pe_patchify")


def _unpatchify(self, tokens, latents):
    C, H2, W2 = latents.shape[-3:]
    x = tokens[:, :C * 4].reshape(H2 // 2, W2 // 2, C, 2, 2)       # H W C P Q -> C (H P) (W Q)
    latents.copy_(x.permute(2, 0, 3, 1, 4).reshape(latents.shape))
    self._note("pe_unpatchify")


def _layernorm_modulate(self, x, out, shift, one_plus_scale):
    v = x.float()
    n = _r((v - v.mean(-1, keepdim=True)) * torch.rsqrt(v.var(-1, unbiased=False, keepdim=True) + 1e-6))
    out.copy_(_r(_r(n * one_plus_scale.float()) + shift.float()).to(torch.bfloat16))
    self._note("pe_layernorm_modulate")


def _timestep_embedding(self, t_in, out, raw=True):
    """256-d sinusoid with the reference's bf16 quirks: ts = bf16(t * float(1/1000)) when raw, frequencies rounded to bf16, fp32 argument, cos half first."""
    import math
    t0 = t_in.float().reshape(-1)[:1]
    ts = _r(t0 * torch.tensor(1.0 / 1000.0, dtype=torch.float64).float()) if raw else t0
    freq = _r(torch.exp(-math.log(10000.0) * torch.arange(128, dtype=torch.float32) / 128.0))
    arg = 1000.0 * (ts * freq)
    out.copy_(torch.cat([torch.cos(arg), torch.sin(arg)]).to(torch.bfloat16))
    self._note("pe_timestep_embedding")


EmulatedNative.patchify, EmulatedNative.unpatchify = _patchify, _unpatchify
EmulatedNative.layernorm_modulate, EmulatedNative.timestep_embedding = _layernorm_modulate, _timestep_embedding


# ---- Qwen2.5-VL text-encoder path (include/pe_b200.h, last section) ----------------------------------------------------------------------
def _swiglu(self, x, out, I):
    g, u = x[:, :I].float(), x[:, I:2 * I].float()
    out[:, :I] = (_r(g / (1 + torch.exp(-g))) * u).to(torch.bfloat16)
    self._note("pe_swiglu")


def _rope_half(self, x, H, D, cos, sin, row0=0, row_ptr=None, mode=1):
    """rotate-half RoPE in place on x [T, >= H * D]; token t uses table row (row_ptr ? *row_ptr : row0) + t.  mode 0: fp32 arithmetic;
    mode 1: bf16 op order with bf16-rounded cos / sin (bf16(bf16(x c) + bf16(rot s)))."""
    T = x.shape[0]
    r0 = int(row_ptr.reshape(-1)[0]) if row_ptr is not None else row0
    c, s = cos[r0:r0 + T].float()[:, None, :], sin[r0:r0 + T].float()[:, None, :]
    v = x[:, :H * D].float().reshape(T, H, D)
    rot = torch.cat([-v[..., D // 2:], v[..., :D // 2]], dim=-1)
    y = (v * c + rot * s) if mode == 0 else _r(_r(v * _r(c)) + _r(rot * _r(s)))
    x[:, :H * D] = y.reshape(T, H * D).to(torch.bfloat16)
    self._note("pe_rope_half")


def _range_attention(self, q, k, v, o, H, Hkv, D, scale, kv_lo=None, kv_hi=None, kv_len_ptr=None, Skv=None):
    """query i of head h attends to keys [lo_i, hi_i) of KV head h // (H / Hkv); defaults lo = 0, hi = kv_len (or all rows)."""
    Sq = q.shape[0]
    n = k.shape[0] if Skv is None else Skv
    if kv_len_ptr is not None:
        n = min(n, int(kv_len_ptr.reshape(-1)[0]))
    lo = kv_lo.long() if kv_lo is not None else torch.zeros(Sq, dtype=torch.long)
    hi = kv_hi.long() if kv_hi is not None else torch.full((Sq,), n, dtype=torch.long)
    pos = torch.arange(n)
    allow = (pos[None, :] >= lo[:, None]) & (pos[None, :] < hi[:, None])
    qh = q[:, :H * D].float().reshape(Sq, H, D).transpose(0, 1)
    kh = k[:n, :Hkv * D].float().reshape(n, Hkv, D).transpose(0, 1).repeat_interleave(H // Hkv, dim=0)
    vh = v[:n, :Hkv * D].float().reshape(n, Hkv, D).transpose(0, 1).repeat_interleave(H // Hkv, dim=0)
    p = torch.softmax((qh @ kh.transpose(1, 2) * scale).masked_fill(~allow[None], float("-inf")), dim=-1)
    o[:, :H * D] = (p @ vh).transpose(0, 1).reshape(Sq, H * D).to(torch.bfloat16)
    self._note("pe_range_attention")


def _gather_rows(self, table, ids, out):
    keep = ids >= 0
    out[keep] = table[ids[keep]]
    self._note("pe_gather_rows")


def _check_async(self):
    return None


EmulatedNative.swiglu, EmulatedNative.rope_half, EmulatedNative.range_attention = _swiglu, _rope_half, _range_attention
EmulatedNative.gather_rows, EmulatedNative.check_async = _gather_rows, _check_async


# ---- decode step -----------------------------------------------------------------------------------------------------------------
def _gemv_fused(self, x, w, bias, y, act_in=0, norm_w=None, eps=1e-6, residual=None):
    """pe_gemv_fused: optional SwiGLU (act_in 2, x = gate | up) or RMSNorm prologue on the input, y = bf16(residual + bf16(x w^T + bias))."""
    v = x.float().reshape(-1, x.shape[-1])
    K = w.shape[1]
    if act_in == 2:
        g, u = v[:, :K], v[:, K:2 * K]
        v = _r(_r(g / (1 + torch.exp(-g))) * u)
    elif act_in == 1:
        v = _r(v / (1 + torch.exp(-v)))
    if norm_w is not None:
        v = _r(norm_w.float() * _r(v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)))
    o = _r(v @ w.float().t() + (bias.float() if bias is not None else 0.0))
    if residual is not None:
        o = residual.float().reshape(o.shape) + o
    y.copy_(o.reshape(y.shape).to(torch.bfloat16))
    self._note("pe_gemv_fused")


def _gemv_swiglu(self, x, w, bias, y, norm_w=None, eps=1e-6):
    v = x.float().reshape(-1, x.shape[-1])
    if norm_w is not None:
        v = _r(norm_w.float() * _r(v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)))
    gu = _r(v @ w.float().t() + (bias.float() if bias is not None else 0.0))
    I = w.shape[0] // 2
    g, u = gu[:, :I], gu[:, I:]
    y.copy_((_r(g / (1 + torch.exp(-g))) * u).reshape(y.shape).to(torch.bfloat16))
    self._note("pe_gemv_swiglu")


def _argmax(self, x, out, log=None, log_pos=None):
    i = int(torch.argmax(x.float().reshape(-1)))               # first maximal index
    out.reshape(-1)[0] = i
    if log is not None:
        log.reshape(-1)[int(log_pos.reshape(-1)[0])] = i
    self._note("pe_argmax")


def _kv_append(self, k_new, v_new, cache_k, cache_v, pos):
    r = int(pos.reshape(-1)[0])
    cache_k[r, :k_new.numel()] = k_new.reshape(-1)
    cache_v[r, :v_new.numel()] = v_new.reshape(-1)
    self._note("pe_kv_append")


def _rope_kv_append(self, qkv_row, Hq, Hkv, D, cos, sin, cache_k, cache_v, counters):
    row = qkv_row.reshape(1, -1)
    _rope_half(self, row[:, :(Hq + Hkv) * D], Hq + Hkv, D, cos, sin, row_ptr=counters.reshape(-1)[1:2], mode=1)
    r = int(counters.reshape(-1)[0])
    cache_k[r, :Hkv * D] = row[0, Hq * D:(Hq + Hkv) * D]
    cache_v[r, :Hkv * D] = row[0, (Hq + Hkv) * D:(Hq + 2 * Hkv) * D]
    self._note("pe_rope_kv_append")


def _decode_attention_fused(self, qkv_rows, caches, outs, counters, Hq, Hkv, D, cos, sin, scale):
    """rope of the new q / k heads, KV append at counters[0], attention of the new query over rows [0, counters[0]]; qkv is left untouched."""
    for qkv, (ck, cv), out, ctr in zip(qkv_rows, caches, outs, counters):
        row = qkv.clone().reshape(1, -1)
        _rope_kv_append(self, row[0], Hq, Hkv, D, cos, sin, ck, cv, ctr)
        n = int(ctr.reshape(-1)[0]) + 1
        _range_attention(self, row[:, :Hq * D], ck, cv, out.reshape(1, -1), Hq, Hkv, D, scale, kv_len_ptr=torch.tensor([n], dtype=torch.int32))
        self.calls = self.calls[:-2]
        self.launches -= 2
    self._note("pe_decode_attention_fused")


def _advance(self, counters, n):
    counters.reshape(-1)[:n] += 1
    self._note("pe_advance")


EmulatedNative.gemv_fused, EmulatedNative.gemv_swiglu, EmulatedNative.argmax = _gemv_fused, _gemv_swiglu, _argmax
EmulatedNative.kv_append, EmulatedNative.rope_kv_append, EmulatedNative.decode_attention_fused, EmulatedNative.advance = _kv_append, _rope_kv_append, _decode_attention_fused, _advance


# ---- adapter plumbing and the CFG + Euler update -----------------------------------------------------------------------------------------
def _special_gather(self, prompt_emb, mask_u8, dst, idx):
    rows = torch.nonzero(mask_u8.reshape(-1) != 0).reshape(-1)
    n = dst.shape[0]
    assert rows.numel() <= n, "more special rows than the caller made room for (the kernel raises PE_ERR_INVALID_ARGUMENT)"
    dst.zero_()
    idx.fill_(-1)
    dst[:rows.numel()] = prompt_emb[rows]
    idx[:rows.numel()] = rows.to(torch.int32)
    idx[n] = rows.numel()
    self._note("pe_special_gather")


def _special_blend_scatter(self, prompt_emb, idx, pred_dino, pred_vae, t_in, t_min, t_max):
    """helpers.py:142-164 on a bf16 timestep: alpha = clamp(bf16(bf16(t - t_min) * float(1 / range))), out = bf16(alpha d) + bf16(bf16(1 - alpha) v)."""
    inv = torch.tensor(1.0 / (t_max - t_min + 1e-6), dtype=torch.float64).float()
    alpha = _r(_r(t_in.float().reshape(-1)[0] - t_min) * inv).clamp(0.0, 1.0)
    oma = _r(1.0 - alpha)
    for i in range(pred_dino.shape[0]):
        t = int(idx[i])
        if t >= 0:
            prompt_emb[t] = (_r(alpha * pred_dino[i].float()) + _r(oma * pred_vae[i].float())).to(torch.bfloat16)
    self._note("pe_special_blend_scatter")


def _cfg_euler_step(self, latents, posi, nega, cfg_scale, dsigma):
    p = posi.float()
    if nega is not None:
        q = nega.float()
        p = _r(q + _r(cfg_scale * _r(p - q)))
    latents.copy_((latents.float() + _r(p * dsigma)).to(torch.bfloat16))
    self._note("pe_cfg_euler_step")


EmulatedNative.special_gather, EmulatedNative.special_blend_scatter, EmulatedNative.cfg_euler_step = _special_gather, _special_blend_scatter, _cfg_euler_step


# ---- training-path feature extractors: pe_layernorm (affine), pe_small_attention ---------------------------------------------------------
def _layernorm(self, x, out, w=None, b=None, eps=1e-5):
    v = x.float()
    y = (v - v.mean(-1, keepdim=True)) * torch.rsqrt(v.var(-1, unbiased=False, keepdim=True) + eps)
    if w is not None:
        y = y * w.float() + (b.float() if b is not None else 0.0)
    out.copy_(y.to(torch.bfloat16))
    self._note("pe_layernorm")


def _small_attention(self, q, k, v, o, B, H, Sq, Skv, D, scale):
    """q [B * Sq, >= H * D], k / v [B * Skv, >= H * D]: per (batch, head) softmax(scale q k^T) v."""
    qh = q[:, :H * D].float().reshape(B, Sq, H, D).permute(0, 2, 1, 3)
    kh = k[:, :H * D].float().reshape(B, Skv, H, D).permute(0, 2, 1, 3)
    vh = v[:, :H * D].float().reshape(B, Skv, H, D).permute(0, 2, 1, 3)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    o[:, :H * D] = (p @ vh).permute(0, 2, 1, 3).reshape(B * Sq, H * D).to(torch.bfloat16)
    self._note("pe_small_attention")


EmulatedNative.layernorm, EmulatedNative.small_attention = _layernorm, _small_attention
