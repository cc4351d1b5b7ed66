"""ORACLE (test infrastructure, not product code) -- CPU restatement of PhysicEdit's training-path feature extractors.

  * perceiver resampler / flamingo-style cross attention: DiffSynth-Studio/diffsynth/pipelines/helpers.py:8-110
  * VisualThinkingAdapter: helpers.py:112-121
  * DINOv2-with-registers ViT-B/14: NOT in /root/reference -- it is `transformers` (unpinned in
    DiffSynth-Studio/requirements.txt:3; installed 5.5.0) modeling_dinov2_with_registers.py:42-171 (embeddings),
    :174-254 (attention), :364-405 (layer), :459-509 (model); call sites pipelines/dinov2.py:17-31 and
    qwen_image_physical.py:1071,1082.  Dinov2withNorm makes the final LayerNorm non-affine and drops CLS + 4 registers.
  * QwenImageUnit_PhysicalVisualEmbedder.process: qwen_image_physical.py:1057-1118 (from pre-processed tensors).

Pinned by tests/test_oracle_golden.py::test_aux_* against tests/golden/aux.pt (oracle/make_golden.py runs the reference's
PerceiverResampler / VisualThinkingAdapter and transformers' Dinov2WithRegistersModel on seeded synthetic weights).
Weights are flat {state_dict key: tensor} dicts with the reference's / HF's key names.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]


def _lin(x, W, p, bias=True):
    return F.linear(x, W[p + ".weight"], W.get(p + ".bias") if bias else None)


def _ln(x, W, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), W[p + ".weight"], W[p + ".bias"], eps)


# ---- perceiver resampler -----------------------------------------------------------------------
def perceiver_attention(W: Weights, p: str, x, latents, heads=8):
    x = _ln(x, W, p + ".norm_media")
    latents = _ln(latents, W, p + ".norm_latents")
    b, m = latents.shape[0], latents.shape[1]
    q = _lin(latents, W, p + ".to_q", bias=False)
    k, v = _lin(torch.cat((x, latents), dim=1), W, p + ".to_kv", bias=False).chunk(2, dim=-1)
    d = q.shape[-1] // heads
    split = lambda t: t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)
    q, k, v = split(q), split(k), split(v)
    dots = torch.einsum("bhid,bhjd->bhij", q, k) * d ** -0.5
    dots = dots - dots.amax(dim=-1, keepdim=True)
    attn = dots.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(b, m, heads * d)
    return _lin(out, W, p + ".to_out", bias=False)


def perceiver_resampler(W: Weights, x, depth=2, heads=8):
    b, n = x.shape[:2]
    latents = W["latents"].unsqueeze(0).expand(b, -1, -1)
    x = x + W["pos_emb.weight"][:n]
    for l in range(depth):
        latents = latents + perceiver_attention(W, f"layers.{l}.0", x, latents, heads)
        h = _ln(latents, W, f"layers.{l}.1.net.0")
        h = _lin(F.gelu(_lin(h, W, f"layers.{l}.1.net.1")), W, f"layers.{l}.1.net.3")
        latents = latents + h
    return _ln(latents, W, "norm")


def visual_thinking_adapter(W: Weights, x):
    return _lin(F.gelu(_lin(x, W, "net.0")), W, "net.2")


def resampler_param_shapes(dim: int, max_tokens: int, num_latents=64, depth=2, heads=8, dim_head=64):
    inner = heads * dim_head
    s = {"latents": (num_latents, dim), "pos_emb.weight": (max_tokens, dim), "norm.weight": (dim,), "norm.bias": (dim,)}
    for l in range(depth):
        a, f = f"layers.{l}.0", f"layers.{l}.1.net"
        for n in ("norm_media", "norm_latents"):
            s[f"{a}.{n}.weight"], s[f"{a}.{n}.bias"] = (dim,), (dim,)
        s[f"{a}.to_q.weight"], s[f"{a}.to_kv.weight"], s[f"{a}.to_out.weight"] = (inner, dim), (2 * inner, dim), (dim, inner)
        s[f"{f}.0.weight"], s[f"{f}.0.bias"] = (dim,), (dim,)
        s[f"{f}.1.weight"], s[f"{f}.1.bias"] = (4 * dim, dim), (4 * dim,)
        s[f"{f}.3.weight"], s[f"{f}.3.bias"] = (dim, 4 * dim), (dim,)
    return s


def vt_adapter_param_shapes(in_dim: int, out_dim: int = 3584):
    return {"net.0.weight": (3 * out_dim, in_dim), "net.0.bias": (3 * out_dim,), "net.2.weight": (out_dim, 3 * out_dim), "net.2.bias": (out_dim,)}


# ---- DINOv2 with registers -----------------------------------------------------------------------
def dinov2_param_shapes(hidden=768, layers=12, patch=14, image_size=518, n_reg=4, ratio=4):
    s = {"embeddings.cls_token": (1, 1, hidden), "embeddings.mask_token": (1, hidden), "embeddings.register_tokens": (1, n_reg, hidden),
         "embeddings.position_embeddings": (1, (image_size // patch) ** 2 + 1, hidden),
         "embeddings.patch_embeddings.projection.weight": (hidden, 3, patch, patch), "embeddings.patch_embeddings.projection.bias": (hidden,),
         "layernorm.weight": (hidden,), "layernorm.bias": (hidden,)}
    for i in range(layers):
        p = f"encoder.layer.{i}"
        for n in ("norm1", "norm2"):
            s[f"{p}.{n}.weight"], s[f"{p}.{n}.bias"] = (hidden,), (hidden,)
        for n in ("query", "key", "value"):
            s[f"{p}.attention.attention.{n}.weight"], s[f"{p}.attention.attention.{n}.bias"] = (hidden, hidden), (hidden,)
        s[f"{p}.attention.output.dense.weight"], s[f"{p}.attention.output.dense.bias"] = (hidden, hidden), (hidden,)
        s[f"{p}.layer_scale1.lambda1"], s[f"{p}.layer_scale2.lambda1"] = (hidden,), (hidden,)
        s[f"{p}.mlp.fc1.weight"], s[f"{p}.mlp.fc1.bias"] = (ratio * hidden, hidden), (ratio * hidden,)
        s[f"{p}.mlp.fc2.weight"], s[f"{p}.mlp.fc2.bias"] = (hidden, ratio * hidden), (hidden,)
    return s


def dinov2_synth_weights(seed: int, dtype=torch.float32, **cfg) -> Weights:
    shapes = dinov2_param_shapes(**cfg)
    out = {}
    for n, key in enumerate(sorted(shapes)):
        shp = shapes[key]
        g = torch.Generator("cpu").manual_seed(seed * 7919 + n)
        if key.endswith("projection.weight"):
            t = torch.randn(shp, generator=g) / math.sqrt(shp[1] * shp[2] * shp[3])
        elif len(shp) == 2 and key.endswith(".weight"):
            t = torch.randn(shp, generator=g) / math.sqrt(shp[1])
        elif "lambda1" in key:
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif key.endswith("norm1.weight") or key.endswith("norm2.weight") or key == "layernorm.weight":
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            t = 0.02 * torch.randn(shp, generator=g)
        out[key] = t.to(dtype)
    return out


def dinov2_pos(W: Weights, gh: int, gw: int):
    pe = W["embeddings.position_embeddings"]
    npos = pe.shape[1] - 1
    s = int(npos ** 0.5)
    if gh * gw == npos and gh == gw:
        return pe
    patch = pe[:, 1:].reshape(1, s, s, -1).permute(0, 3, 1, 2)
    patch = F.interpolate(patch.to(torch.float32), size=(gh, gw), mode="bicubic", align_corners=False, antialias=True).to(pe.dtype)
    return torch.cat((pe[:, 0].unsqueeze(0), patch.permute(0, 2, 3, 1).reshape(1, -1, pe.shape[-1])), dim=1)


def dinov2_with_norm(W: Weights, pixel_values, heads=12, patch=14, n_reg=4, eps=1e-6):
    """Dinov2withNorm.forward: HF model -> last_hidden_state with a NON-affine final LayerNorm -> drop CLS + registers."""
    B, _, H, Wd = pixel_values.shape
    x = F.conv2d(pixel_values, W["embeddings.patch_embeddings.projection.weight"], W["embeddings.patch_embeddings.projection.bias"], stride=patch)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((W["embeddings.cls_token"].expand(B, -1, -1), x), dim=1)
    x = x + dinov2_pos(W, H // patch, Wd // patch)
    x = torch.cat((x[:, :1], W["embeddings.register_tokens"].expand(B, -1, -1), x[:, 1:]), dim=1)
    hid = x.shape[-1]
    d = hid // heads
    n_layers = 1 + max(int(k.split(".")[2]) for k in W if k.startswith("encoder.layer."))
    for i in range(n_layers):
        p = f"encoder.layer.{i}"
        h = F.layer_norm(x, (hid,), W[p + ".norm1.weight"], W[p + ".norm1.bias"], eps)
        a = p + ".attention.attention"
        split = lambda t: t.reshape(B, -1, heads, d).transpose(1, 2)
        q, k, v = split(_lin(h, W, a + ".query")), split(_lin(h, W, a + ".key")), split(_lin(h, W, a + ".value"))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, -1, hid)
        o = _lin(o, W, p + ".attention.output.dense")
        x = o * W[p + ".layer_scale1.lambda1"] + x
        h = F.layer_norm(x, (hid,), W[p + ".norm2.weight"], W[p + ".norm2.bias"], eps)
        h = _lin(F.gelu(_lin(h, W, p + ".mlp.fc1")), W, p + ".mlp.fc2")
        x = h * W[p + ".layer_scale2.lambda1"] + x
    x = F.layer_norm(x, (hid,), None, None, eps)
    return x[:, 1 + n_reg:]


# ---- QwenImageUnit_PhysicalVisualEmbedder.process (qwen_image_physical.py:1057-1118) from tensors -----------------
def physical_visual_embeddings(P: Dict[str, Weights], dino_middle, dino_source, vae_middle_latents, vae_source_latents):
    """P: weights of dinov2, dino_resampler, dino_time_embed, dino_resampler_adapter, vae_resampler, vae_time_embed,
    vae_resampler_adapter.  Returns (pseudo_special_emb_dino, pseudo_special_emb_vae), each [1, 64, 3584]."""
    from .dit_oracle import patchify

    def dino_branch(px, with_time):
        hs = dinov2_with_norm(P["dinov2"], px)
        if with_time:
            hs = hs + P["dino_time_embed"]["weight"][: hs.shape[0]].unsqueeze(1)
        hs = hs.reshape(1, -1, hs.shape[-1])
        return visual_thinking_adapter(P["dino_resampler_adapter"], perceiver_resampler(P["dino_resampler"], hs))

    def vae_branch(lat, with_time):
        tok = patchify(lat)
        if with_time:
            tok = tok + P["vae_time_embed"]["weight"][: tok.shape[0]].unsqueeze(1)
        tok = tok.reshape(1, -1, tok.shape[-1])
        return visual_thinking_adapter(P["vae_resampler_adapter"], perceiver_resampler(P["vae_resampler"], tok))

    return (dino_branch(dino_middle, True) - dino_branch(dino_source, False), vae_branch(vae_middle_latents, True) - vae_branch(vae_source_latents, False))


def aux_synth(seed: int, dtype=torch.float32) -> Dict[str, Weights]:
    """Seeded weights for every training-path module (shared by make_golden.py and the tests)."""
    from .dit_oracle import synth_weights
    P = {"dinov2": dinov2_synth_weights(seed, dtype)}
    P["dino_resampler"] = synth_weights(resampler_param_shapes(768, 4096), seed + 1, dtype)
    P["vae_resampler"] = synth_weights(resampler_param_shapes(64, 10240), seed + 2, dtype)
    for k in ("dino_resampler", "vae_resampler"):       # embeddings / latents: small normal values like the reference init
        g = torch.Generator("cpu").manual_seed(seed + 50)
        P[k]["latents"] = (0.02 * torch.randn(P[k]["latents"].shape, generator=g)).to(dtype)
        P[k]["pos_emb.weight"] = (0.5 * torch.randn(P[k]["pos_emb.weight"].shape, generator=g)).to(dtype)
    P["dino_resampler_adapter"] = synth_weights(vt_adapter_param_shapes(768), seed + 3, dtype)
    P["vae_resampler_adapter"] = synth_weights(vt_adapter_param_shapes(64), seed + 4, dtype)
    g = torch.Generator("cpu").manual_seed(seed + 5)
    P["dino_time_embed"] = {"weight": torch.randn(6, 768, generator=g).to(dtype)}
    P["vae_time_embed"] = {"weight": torch.randn(6, 64, generator=g).to(dtype)}
    return P


def aux_inputs(seed: int, n_mid: int = 3, lat_hw: Tuple[int, int] = (32, 32), dtype=torch.float32):
    g = torch.Generator("cpu").manual_seed(seed)
    return dict(dino_middle=torch.randn(n_mid, 3, 224, 224, generator=g).to(dtype), dino_source=torch.randn(1, 3, 224, 224, generator=g).to(dtype),
                vae_middle_latents=torch.randn(n_mid, 16, *lat_hw, generator=g).to(dtype), vae_source_latents=torch.randn(1, 16, *lat_hw, generator=g).to(dtype))
