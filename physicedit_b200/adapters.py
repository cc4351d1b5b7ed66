"""PhysicEdit's adapters and (training-path) feature extractors, executed by libpe_b200.

Mirrors DiffSynth-Studio/diffsynth/pipelines/helpers.py (FeedForward, PerceiverAttention,
PerceiverResampler, VisualThinkingAdapter, VisualThinkingDualAdapter) and pipelines/dinov2.py
(Dinov2withNorm around transformers' Dinov2WithRegistersModel) with identical parameter names, so the
`pipe.`-prefixed checkpoint keys of the reference (`pipe.visual_thinking_adapter.head_dino.0.weight`,
`pipe.dino_resampler.latents`, ...) load with `load_state_dict`.  Forward passes run on the C-ABI
kernels: the tcgen05 GEMM with bias / GELU(erf) / residual epilogues, `pe_layernorm`,
`pe_small_attention`.  bf16 on an sm_100 GPU only.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native as nv


def _nat(t: torch.Tensor) -> nv.Native:
    if not t.is_cuda or t.dtype != torch.bfloat16:
        raise nv.NativeUnavailable(f"native adapters run in bfloat16 on an sm_100 GPU only (got {t.dtype} on {t.device})")
    return nv.Native.get(t.device.index or 0)


def _ones(n, device):
    return torch.ones(n, dtype=torch.bfloat16, device=device)


def _mlp_gelu(nat: nv.Native, x2d: torch.Tensor, l0: nn.Linear, l2: nn.Linear, residual: torch.Tensor = None) -> torch.Tensor:
    """Linear -> GELU(erf) -> Linear [+ residual], two GEMM launches."""
    h = nat.linear(x2d, l0.weight, l0.bias, nv.EPI_BIAS_GELU_ERF)
    if residual is None:
        return nat.linear(h, l2.weight, l2.bias, nv.EPI_BIAS)
    nat.gemm([dict(a=h, w=l2.weight, bias=l2.bias, out=residual, gate=_ones(l2.weight.shape[0], x2d.device))],
             l2.weight.shape[0], l2.weight.shape[1], nv.EPI_GATE_RESIDUAL)
    return residual


class VisualThinkingAdapter(nn.Module):
    """helpers.py:112-121."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(in_dim, out_dim * 3), nn.GELU(), nn.Linear(out_dim * 3, out_dim))

    def forward(self, x):
        from . import autograd as ag
        if ag.module_trains(self, x):                                    # training: same GEMMs under autograd
            return ag.mlp_gelu(x, self.net[0], self.net[2])
        nat = _nat(x)
        y = _mlp_gelu(nat, x.reshape(-1, x.shape[-1]).contiguous(), self.net[0], self.net[2])
        return y.view(*x.shape[:-1], -1)


class VisualThinkingDualAdapter(nn.Module):
    """helpers.py:123-183.  forward(x, timestep) -> (mixed, pred_dino, pred_vae)."""

    def __init__(self, in_dim, out_dim, t_min, t_max):
        super().__init__()
        self.head_dino = nn.Sequential(nn.Linear(in_dim, out_dim * 3), nn.GELU(), nn.Linear(out_dim * 3, out_dim))
        self.head_vae = nn.Sequential(nn.Linear(in_dim, out_dim * 3), nn.GELU(), nn.Linear(out_dim * 3, out_dim))
        self.t_min, self.t_max = t_min, t_max

    def _get_alpha(self, timestep, device):
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], device=device, dtype=torch.float32)
        alpha = (timestep - self.t_min) / (self.t_max - self.t_min + 1e-6)
        return alpha.clamp(0.0, 1.0).view(-1, 1, 1)

    def heads(self, x2d: torch.Tensor):
        nat = _nat(x2d)
        return (_mlp_gelu(nat, x2d, self.head_dino[0], self.head_dino[2]), _mlp_gelu(nat, x2d, self.head_vae[0], self.head_vae[2]))

    def forward(self, x, timestep):
        from . import autograd as ag
        if ag.module_trains(self, x):
            return ag.dual_adapter_forward(self, x, timestep)
        nat = _nat(x)
        x2d = x.reshape(-1, x.shape[-1]).contiguous()
        pd, pv = self.heads(x2d)
        n = x2d.shape[0]
        mixed = torch.empty_like(pd)
        idx = torch.arange(n + 1, dtype=torch.int32, device=x.device)
        nat.special_blend_scatter(mixed, idx, pd, pv, timestep.to(torch.bfloat16).reshape(-1)[:1].contiguous(), self.t_min, self.t_max)
        shp = (*x.shape[:-1], pd.shape[-1])
        return mixed.view(shp), pd.view(shp), pv.view(shp)

    def get_loss(self, pred_dino, pred_vae, gt_dino, gt_vae, timestep, epsilon=0.1):
        """Scalar training loss (helpers.py:166-183); a handful of reductions, evaluated with torch on the device."""
        alpha = self._get_alpha(timestep, pred_dino.device).type_as(pred_dino)
        loss_dino = F.mse_loss(pred_dino, gt_dino, reduction="none").mean(dim=[1, 2])
        loss_vae = F.mse_loss(pred_vae, gt_vae, reduction="none").mean(dim=[1, 2])
        w = alpha.squeeze()
        wd, wv = w + epsilon, (1 - w) + epsilon
        tot = wd + wv
        return ((wd / tot) * loss_dino + (wv / tot) * loss_vae).mean()


class FeedForward(nn.Module):
    """helpers.py:8-19."""

    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, dim * mult), nn.GELU(), nn.Linear(dim * mult, dim))


class PerceiverAttention(nn.Module):
    """helpers.py:21-65 (flamingo-style cross attention: keys / values over cat(media, latents))."""

    def __init__(self, dim, dim_head=64, heads=8):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads, self.dim_head = heads, dim_head
        inner = dim_head * heads
        self.norm_media = nn.LayerNorm(dim)
        self.norm_latents = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)


class PerceiverResampler(nn.Module):
    """helpers.py:67-110.  x [1, N, dim] -> [1, num_latents, dim]."""

    def __init__(self, dim=1024, depth=2, dim_head=64, heads=8, num_latents=32, max_num_media_tokens=4096):
        super().__init__()
        self.latents = nn.Parameter(torch.randn(num_latents, dim) * 0.02)
        self.pos_emb = nn.Embedding(max_num_media_tokens, dim)
        self.layers = nn.ModuleList([nn.ModuleList([PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads), FeedForward(dim=dim)])
                                     for _ in range(depth)])
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        from . import autograd as ag
        assert x.shape[0] == 1, "the reference always calls the resampler with batch 1 (frames are flattened into the sequence)"
        if ag.module_trains(self, x):
            return ag.resampler_forward(self, x)
        nat = _nat(x)
        n, dim = x.shape[1], x.shape[2]
        m = self.latents.shape[0]
        xm = x[0].contiguous().clone()
        nat.add_rows(xm, self.pos_emb.weight[:n], n, 1.0)                    # x + pos_emb(arange(n))
        lat = self.latents.detach().clone()
        kv_in = torch.empty(n + m, dim, dtype=torch.bfloat16, device=x.device)
        for attn, ff in self.layers:
            inner = attn.heads * attn.dim_head
            nat.layernorm(xm, kv_in[:n], attn.norm_media.weight, attn.norm_media.bias, attn.norm_media.eps)
            nat.layernorm(lat, kv_in[n:], attn.norm_latents.weight, attn.norm_latents.bias, attn.norm_latents.eps)
            q = nat.linear(kv_in[n:], attn.to_q.weight, None)
            kv = nat.linear(kv_in, attn.to_kv.weight, None)
            o = torch.empty(m, inner, dtype=torch.bfloat16, device=x.device)
            nat.small_attention(q, kv[:, :inner], kv[:, inner:], o, 1, attn.heads, m, n + m, attn.dim_head, attn.scale)
            nat.gemm([dict(a=o, w=attn.to_out.weight, bias=None, out=lat, gate=_ones(dim, x.device))], dim, inner, nv.EPI_GATE_RESIDUAL)
            h = torch.empty_like(lat)
            nat.layernorm(lat, h, ff.net[0].weight, ff.net[0].bias, ff.net[0].eps)
            _mlp_gelu(nat, h, ff.net[1], ff.net[3], residual=lat)
        out = torch.empty_like(lat)
        nat.layernorm(lat, out, self.norm.weight, self.norm.bias, self.norm.eps)
        return out.unsqueeze(0)


# ------------------------------------------------------------------------------------------------
# DINOv2-with-registers ViT (transformers modeling_dinov2_with_registers.py), parameter names as HF
# ------------------------------------------------------------------------------------------------
class _PatchEmb(nn.Module):
    def __init__(self, hidden, patch, channels=3):
        super().__init__()
        self.projection = nn.Conv2d(channels, hidden, kernel_size=patch, stride=patch)


class _Embeddings(nn.Module):
    def __init__(self, hidden, patch, image_size, n_reg):
        super().__init__()
        self.cls_token = nn.Parameter(torch.randn(1, 1, hidden))
        self.mask_token = nn.Parameter(torch.zeros(1, hidden))
        self.register_tokens = nn.Parameter(torch.zeros(1, n_reg, hidden))
        self.patch_embeddings = _PatchEmb(hidden, patch)
        self.position_embeddings = nn.Parameter(torch.randn(1, (image_size // patch) ** 2 + 1, hidden))


class _SelfAttn(nn.Module):
    def __init__(self, hidden):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(hidden, hidden), nn.Linear(hidden, hidden), nn.Linear(hidden, hidden)


class _SelfOut(nn.Module):
    def __init__(self, hidden):
        super().__init__()
        self.dense = nn.Linear(hidden, hidden)


class _Attention(nn.Module):
    def __init__(self, hidden):
        super().__init__()
        self.attention = _SelfAttn(hidden)
        self.output = _SelfOut(hidden)


class _LayerScale(nn.Module):
    def __init__(self, hidden, value=1.0):
        super().__init__()
        self.lambda1 = nn.Parameter(value * torch.ones(hidden))


class _MLP(nn.Module):
    def __init__(self, hidden, ratio=4):
        super().__init__()
        self.fc1 = nn.Linear(hidden, hidden * ratio)
        self.fc2 = nn.Linear(hidden * ratio, hidden)


class _Layer(nn.Module):
    def __init__(self, hidden, eps, mlp_ratio=4):
        super().__init__()
        self.norm1 = nn.LayerNorm(hidden, eps=eps)
        self.attention = _Attention(hidden)
        self.layer_scale1 = _LayerScale(hidden)
        self.norm2 = nn.LayerNorm(hidden, eps=eps)
        self.mlp = _MLP(hidden, mlp_ratio)
        self.layer_scale2 = _LayerScale(hidden)


class _Encoder(nn.Module):
    def __init__(self, hidden, layers, eps, mlp_ratio=4):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(hidden, eps, mlp_ratio) for _ in range(layers)])


class Dinov2Encoder(nn.Module):
    """Same state_dict keys as transformers' Dinov2WithRegistersModel (embeddings.*, encoder.layer.N.*, layernorm.*)."""

    def __init__(self, hidden=768, layers=12, heads=12, patch=14, image_size=518, n_reg=4, eps=1e-6, mlp_ratio=4):
        super().__init__()
        self.hidden, self.heads, self.patch, self.n_reg, self.eps = hidden, heads, patch, n_reg, eps
        self.embeddings = _Embeddings(hidden, patch, image_size, n_reg)
        self.encoder = _Encoder(hidden, layers, eps, mlp_ratio)
        self.layernorm = nn.LayerNorm(hidden, eps=eps)
        self._packed = None
        self._pos_cache = {}

    # packed copies (patch-embed weight, fused QKV) and the interpolated position table are derived from the parameters:
    # any weight load or device / dtype move drops them (an in-place load keeps data_ptr, so pointer checks are not enough)
    def load_state_dict(self, *args, **kwargs):
        self._packed, self._pos_cache = None, {}
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._packed, self._pos_cache = None, {}
        return super()._apply(fn, *args, **kwargs)

    def _pos(self, gh: int, gw: int) -> torch.Tensor:
        """interpolate_pos_encoding (:90-140): bicubic antialias in fp32, depends on the weights and grid only -> cached."""
        key = (gh, gw, self.embeddings.position_embeddings.data_ptr())
        if key not in self._pos_cache:
            pe = self.embeddings.position_embeddings
            npos = pe.shape[1] - 1
            s = int(npos ** 0.5)
            if gh * gw == npos and gh == gw:
                out = pe[0]
            else:
                patch = pe[:, 1:].reshape(1, s, s, -1).permute(0, 3, 1, 2)
                patch = F.interpolate(patch.to(torch.float32), size=(gh, gw), mode="bicubic", align_corners=False, antialias=True).to(pe.dtype)
                out = torch.cat((pe[:, 0], patch.permute(0, 2, 3, 1).reshape(-1, pe.shape[-1])), dim=0)
            self._pos_cache = {key: out.contiguous()}
        return self._pos_cache[key]

    def _pack(self):
        dev = self.layernorm.weight.device if self.layernorm.weight is not None else self.embeddings.cls_token.device
        w = self.embeddings.patch_embeddings.projection.weight
        k = w.shape[1] * w.shape[2] * w.shape[3]
        kp = (k + 7) // 8 * 8                                             # K = 588 -> 592 (zero padded) for 16-byte TMA rows
        wp = torch.zeros(w.shape[0], kp, dtype=torch.bfloat16, device=dev)
        wp[:, :k] = w.reshape(w.shape[0], k)
        qkv = []
        for l in self.encoder.layer:
            a = l.attention.attention
            qkv.append((torch.cat([a.query.weight, a.key.weight, a.value.weight], 0).contiguous(),
                        torch.cat([a.query.bias, a.key.bias, a.value.bias], 0).contiguous()))
        self._packed = (wp, kp, qkv)

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """pixel_values [F,3,H,W] bf16 -> last_hidden_state [F, 1 + n_reg + (H/14)(W/14), hidden] after the final LayerNorm."""
        nat = _nat(pixel_values)
        if self._packed is None:
            self._pack()
        wp, kp, qkv = self._packed
        Fn, C, H, W = pixel_values.shape
        p, hid = self.patch, self.hidden
        gh, gw = H // p, W // p
        # im2col of the stride-14 conv (pure data movement) -> GEMM with K = 3*14*14 (+4 zero columns)
        cols = pixel_values[:, :, :gh * p, :gw * p].reshape(Fn, C, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(Fn * gh * gw, C * p * p)
        a = torch.zeros(Fn * gh * gw, kp, dtype=torch.bfloat16, device=pixel_values.device)
        a[:, :C * p * p] = cols
        proj = self.embeddings.patch_embeddings.projection
        patches = nat.linear(a, wp, proj.bias)
        n_tok = 1 + self.n_reg + gh * gw
        x = torch.empty(Fn, n_tok, hid, dtype=torch.bfloat16, device=pixel_values.device)
        pos = self._pos(gh, gw)
        nat.add_rows(patches, pos[1:], gh * gw, 1.0)                          # patch tokens + interpolated position embedding
        x[:, 0] = self.embeddings.cls_token[0, 0] + pos[0]
        x[:, 1:1 + self.n_reg] = self.embeddings.register_tokens[0]           # registers carry no position embedding
        x[:, 1 + self.n_reg:] = patches.view(Fn, gh * gw, hid)
        x = x.view(Fn * n_tok, hid)
        h = torch.empty_like(x)
        o = torch.empty_like(x)
        d = hid // self.heads
        for l, (wq, bq) in zip(self.encoder.layer, qkv):
            nat.layernorm(x, h, l.norm1.weight, l.norm1.bias, l.norm1.eps)
            qkv_o = nat.linear(h, wq, bq)
            nat.small_attention(qkv_o[:, :hid], qkv_o[:, hid:2 * hid], qkv_o[:, 2 * hid:], o, Fn, self.heads, n_tok, n_tok, d, d ** -0.5)
            dn = l.attention.output.dense
            nat.gemm([dict(a=o, w=dn.weight, bias=dn.bias, out=x, gate=l.layer_scale1.lambda1)], hid, hid, nv.EPI_GATE_RESIDUAL)
            nat.layernorm(x, h, l.norm2.weight, l.norm2.bias, l.norm2.eps)
            m1 = nat.linear(h, l.mlp.fc1.weight, l.mlp.fc1.bias, nv.EPI_BIAS_GELU_ERF)
            nat.gemm([dict(a=m1, w=l.mlp.fc2.weight, bias=l.mlp.fc2.bias, out=x, gate=l.layer_scale2.lambda1)], hid, 4 * hid, nv.EPI_GATE_RESIDUAL)
        ln = self.layernorm
        nat.layernorm(x, h, ln.weight if ln.elementwise_affine else None, ln.bias if ln.elementwise_affine else None, ln.eps)
        return h.view(Fn, n_tok, hid)


class Dinov2withNorm(nn.Module):
    """pipelines/dinov2.py:8-35: final LayerNorm made non-affine, CLS + 4 register tokens dropped."""

    def __init__(self, dinov2_path: str = None, normalize: bool = True, config: dict = None):
        super().__init__()
        cfg = dict(hidden=768, layers=12, heads=12, patch=14, image_size=518, n_reg=4, eps=1e-6)
        sd = None
        if dinov2_path is not None:
            import json
            import os
            with open(os.path.join(dinov2_path, "config.json")) as f:
                hf = json.load(f)
            cfg = dict(hidden=hf["hidden_size"], layers=hf["num_hidden_layers"], heads=hf["num_attention_heads"], patch=hf["patch_size"],
                       image_size=hf["image_size"], n_reg=hf.get("num_register_tokens", 4), eps=hf.get("layer_norm_eps", 1e-6),
                       mlp_ratio=hf.get("mlp_ratio", 4))
            st = os.path.join(dinov2_path, "model.safetensors")
            if os.path.exists(st):
                from safetensors.torch import load_file
                sd = load_file(st)
            else:
                sd = torch.load(os.path.join(dinov2_path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
        if config:
            cfg.update(config)
        self.encoder = Dinov2Encoder(**cfg)
        if sd is not None:
            self.encoder.load_state_dict(sd, strict=True)
        self.encoder.requires_grad_(False)
        if normalize:
            self.encoder.layernorm.elementwise_affine = False
            self.encoder.layernorm.weight = None
            self.encoder.layernorm.bias = None
        self.patch_size = self.encoder.patch
        self.hidden_size = self.encoder.hidden

    def dinov2_forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.encoder(x)[:, 1 + self.encoder.n_reg:]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.dinov2_forward(x)
