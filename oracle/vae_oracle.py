"""ORACLE (test infrastructure, not product code) -- CPU restatement of QwenImageVAE.encode / .decode for single images.

Follows DiffSynth-Studio/diffsynth/models/qwen_image_vae.py (call sites pipelines/qwen_image_physical.py:665,
1273, 1298, 1092, 1106).  The pipeline only ever passes 4-D image tensors, so T = 1 and no feature cache is used
(`encode` / `decode` call the encoder / decoder without `feat_cache`, :706-735).  Under those conditions

  * QwenImageCausalConv3d (:8-51) pads 2*p frames of zeros IN FRONT of the single frame, so only the LAST temporal slice of a
    [Cout, Cin, 3, kh, kw] kernel ever meets data: the layer is a 2-D convolution with weight[:, :, -1];
  * the `time_conv` of the 3-D resample layers is skipped (they only run when a feature cache exists, :259-301);
  * QwenImageRMS_norm (:54-78) is F.normalize over channels * sqrt(C) * gamma;
  * QwenImageAttentionBlock (:156-199) is single-head attention with head dim = C (384) over the H*W positions.

Weights are flat {state_dict key: tensor} dicts with the reference's key names and shapes (conv kernels stay 5-D).
Pinned by tests/test_oracle_golden.py::test_vae_* against tests/golden/vae.pt, which oracle/make_golden_vae.py produced by
running the reference's own QwenImageVAE on the same seeded synthetic weights.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]

# qwen_image_vae.py:667-704
LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497,
               0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251,
              1.9160]


def _conv(x: torch.Tensor, W: Weights, p: str, padding=0, stride=1) -> torch.Tensor:
    w = W[p + ".weight"]
    if w.dim() == 5:                      # causal 3-D kernel at T = 1: only the last temporal slice sees the frame (:39-51)
        w = w[:, :, -1]
    return F.conv2d(x, w, W[p + ".bias"], stride=stride, padding=padding)


def rms_norm(x: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    """QwenImageRMS_norm.forward (:76-78): F.normalize(x, dim=1) * sqrt(C) * gamma (+ 0.0)."""
    c = x.shape[1]
    return F.normalize(x, dim=1) * (c ** 0.5) * gamma.reshape(1, c, 1, 1) + 0.0


def residual_block(W: Weights, p: str, x: torch.Tensor) -> torch.Tensor:
    """QwenImageResidualBlock.forward (:112-152), feat_cache None."""
    h = _conv(x, W, p + ".conv_shortcut") if (p + ".conv_shortcut.weight") in W else x
    x = F.silu(rms_norm(x, W[p + ".norm1.gamma"]))
    x = _conv(x, W, p + ".conv1", padding=1)
    x = F.silu(rms_norm(x, W[p + ".norm2.gamma"]))
    x = _conv(x, W, p + ".conv2", padding=1)
    return x + h


def attention_block(W: Weights, p: str, x: torch.Tensor) -> torch.Tensor:
    """QwenImageAttentionBlock.forward (:173-199): one head of dim C over the H*W positions."""
    identity = x
    b, c, hh, ww = x.shape
    x = rms_norm(x, W[p + ".norm.gamma"])
    qkv = _conv(x, W, p + ".to_qkv")
    qkv = qkv.reshape(b, 1, c * 3, -1).permute(0, 1, 3, 2).contiguous()
    q, k, v = qkv.chunk(3, dim=-1)
    x = F.scaled_dot_product_attention(q, k, v)
    x = x.squeeze(1).permute(0, 2, 1).reshape(b, c, hh, ww)
    return _conv(x, W, p + ".proj") + identity


def resample(W: Weights, p: str, x: torch.Tensor, mode: str) -> torch.Tensor:
    """QwenImageResample.forward (:257-301) at T = 1 without a feature cache (time_conv never runs)."""
    if mode.startswith("upsample"):
        x = F.interpolate(x.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(x)     # QwenImageUpsample (:213-215)
        return _conv(x, W, p + ".resample.1", padding=1)
    x = F.pad(x, (0, 1, 0, 1))                                                                      # nn.ZeroPad2d((0,1,0,1)) (:246-249)
    return _conv(x, W, p + ".resample.1", stride=2)


def mid_block(W: Weights, p: str, x: torch.Tensor) -> torch.Tensor:
    x = residual_block(W, p + ".resnets.0", x)
    x = attention_block(W, p + ".attentions.0", x)
    return residual_block(W, p + ".resnets.1", x)


def encoder_layout(dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_downsample=(False, True, True)) -> List[Tuple[str, int]]:
    """[(kind, index in encoder.down_blocks)] in module order (:379-398)."""
    out, i = [], 0
    for lvl in range(len(dim_mult)):
        for _ in range(num_res_blocks):
            out.append(("res", i)); i += 1
        if lvl != len(dim_mult) - 1:
            out.append(("downsample3d" if temporal_downsample[lvl] else "downsample2d", i)); i += 1
    return out


def encode(W: Weights, image: torch.Tensor) -> torch.Tensor:
    """QwenImageVAE.encode (:706-717): image [B,3,H,W] in [-1,1] -> normalised latents [B,16,H/8,W/8]."""
    x = _conv(image, W, "encoder.conv_in", padding=1)
    for kind, i in encoder_layout():
        p = f"encoder.down_blocks.{i}"
        x = residual_block(W, p, x) if kind == "res" else resample(W, p, x, kind)
    x = mid_block(W, "encoder.mid_block", x)
    x = F.silu(rms_norm(x, W["encoder.norm_out.gamma"]))
    x = _conv(x, W, "encoder.conv_out", padding=1)
    x = _conv(x, W, "quant_conv")
    x = x[:, :16]
    mean = torch.tensor(LATENT_MEAN).view(1, 16, 1, 1).to(x.dtype)
    std = (1 / torch.tensor(LATENT_STD).view(1, 16, 1, 1)).to(x.dtype)
    return (x - mean) * std


def decode(W: Weights, latents: torch.Tensor) -> torch.Tensor:
    """QwenImageVAE.decode (:719-731): latents [B,16,h,w] -> image [B,3,8h,8w]."""
    mean = torch.tensor(LATENT_MEAN).view(1, 16, 1, 1).to(latents.dtype)
    std = (1 / torch.tensor(LATENT_STD).view(1, 16, 1, 1)).to(latents.dtype)
    x = latents / std + mean
    x = _conv(x, W, "post_quant_conv")
    x = _conv(x, W, "decoder.conv_in", padding=1)
    x = mid_block(W, "decoder.mid_block", x)
    for b in range(4):
        for r in range(3):
            x = residual_block(W, f"decoder.up_blocks.{b}.resnets.{r}", x)
        if b != 3:
            x = resample(W, f"decoder.up_blocks.{b}.upsamplers.0", x, "upsample")
    x = F.silu(rms_norm(x, W["decoder.norm_out.gamma"]))
    return _conv(x, W, "decoder.conv_out", padding=1)


# ---- parameter shapes (state_dict order / names of the reference class) + synthetic weights -----------------------
def vae_param_shapes(base_dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2,
                     temporal_downsample=(False, True, True)) -> Dict[str, Tuple[int, ...]]:
    S: Dict[str, Tuple[int, ...]] = {}

    def conv3(p, cin, cout, k=3):
        S[p + ".weight"] = (cout, cin, k, k, k)
        S[p + ".bias"] = (cout,)

    def conv2(p, cin, cout, k):
        S[p + ".weight"] = (cout, cin, k, k)
        S[p + ".bias"] = (cout,)

    def res(p, cin, cout):
        S[p + ".norm1.gamma"] = (cin, 1, 1, 1)
        conv3(p + ".conv1", cin, cout)
        S[p + ".norm2.gamma"] = (cout, 1, 1, 1)
        conv3(p + ".conv2", cout, cout)
        if cin != cout:
            conv3(p + ".conv_shortcut", cin, cout, 1)

    def attn(p, c):
        S[p + ".norm.gamma"] = (c, 1, 1)
        conv2(p + ".to_qkv", c, 3 * c, 1)
        conv2(p + ".proj", c, c, 1)

    def mid(p, c):
        attn(p + ".attentions.0", c)
        res(p + ".resnets.0", c, c)
        res(p + ".resnets.1", c, c)

    # encoder (:359-409)
    dims = [base_dim * u for u in (1,) + tuple(dim_mult)]
    conv3("encoder.conv_in", 3, dims[0])
    i = 0
    for lvl, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            res(f"encoder.down_blocks.{i}", cin, cout); i += 1
            cin = cout
        if lvl != len(dim_mult) - 1:
            conv2(f"encoder.down_blocks.{i}.resample.1", cout, cout, 3)
            if temporal_downsample[lvl]:
                S[f"encoder.down_blocks.{i}.time_conv.weight"] = (cout, cout, 3, 1, 1)
                S[f"encoder.down_blocks.{i}.time_conv.bias"] = (cout,)
            i += 1
    mid("encoder.mid_block", dims[-1])
    S["encoder.norm_out.gamma"] = (dims[-1], 1, 1, 1)
    conv3("encoder.conv_out", dims[-1], z_dim * 2)
    conv3("quant_conv", z_dim * 2, z_dim * 2, 1)
    conv3("post_quant_conv", z_dim, z_dim, 1)
    # decoder (:537-599)
    ddims = [base_dim * u for u in (dim_mult[-1],) + tuple(dim_mult[::-1])]
    temporal_upsample = tuple(temporal_downsample[::-1])
    conv3("decoder.conv_in", z_dim, ddims[0])
    mid("decoder.mid_block", ddims[0])
    for b, (cin, cout) in enumerate(zip(ddims[:-1], ddims[1:])):
        if b > 0:
            cin = cin // 2
        for r in range(num_res_blocks + 1):
            res(f"decoder.up_blocks.{b}.resnets.{r}", cin, cout)
            cin = cout
        if b != len(dim_mult) - 1:
            conv2(f"decoder.up_blocks.{b}.upsamplers.0.resample.1", cout, cout // 2, 3)
            if temporal_upsample[b]:
                S[f"decoder.up_blocks.{b}.upsamplers.0.time_conv.weight"] = (cout * 2, cout, 3, 1, 1)
                S[f"decoder.up_blocks.{b}.upsamplers.0.time_conv.bias"] = (cout * 2,)
    S["decoder.norm_out.gamma"] = (ddims[-1], 1, 1, 1)
    conv3("decoder.conv_out", ddims[-1], 3)
    return S


def vae_synth_weights(seed: int, dtype=torch.float32, shapes=None, gain: float = 1.7) -> Weights:
    """Deterministic weights, one generator per tensor (seed, position in the sorted key list).  Conv kernels ~ U(-b, b),
    b = gain/sqrt(fan_in) with fan_in = Cin*kh*kw of the 2-D slice that is live at T = 1 (gain > 1 keeps activations O(1)
    through the RMS-norm/SiLU/conv chain); biases ~ U(-1, 1)/sqrt(fan_in); gammas = 1 + 0.1 N(0,1)."""
    shapes = shapes or vae_param_shapes()
    out = {}
    for n, key in enumerate(sorted(shapes)):
        shp = shapes[key]
        g = torch.Generator("cpu").manual_seed(seed * 1000003 + n)
        if key.endswith(".weight"):
            fan_in = shp[1] * shp[-1] * shp[-2]
            t = (torch.rand(shp, generator=g) * 2 - 1) * (gain / math.sqrt(fan_in))
        elif key.endswith(".bias"):
            w = shapes[key[:-5] + ".weight"]
            t = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(w[1] * w[-1] * w[-2])
        else:
            t = 1 + 0.1 * torch.randn(shp, generator=g)
        out[key] = t.to(dtype)
    return out


def vae_inputs(h8: int, w8: int, seed: int, dtype=torch.float32):
    """A synthetic image in [-1, 1] ([1,3,8*h8,8*w8]: smooth gradients + noise, like preprocess_image's range) and latents
    ~ N(0,1) ([1,16,h8,w8], the scale of the denoised, normalised latents)."""
    g = torch.Generator("cpu").manual_seed(seed)
    H, Wd = 8 * h8, 8 * w8
    yy = torch.linspace(-1, 1, H).view(1, 1, H, 1)
    xx = torch.linspace(-1, 1, Wd).view(1, 1, 1, Wd)
    ph = torch.rand(3, generator=g).view(1, 3, 1, 1) * 6.28
    img = 0.6 * torch.sin(3.1 * yy + ph) * torch.cos(2.3 * xx - ph) + 0.25 * torch.randn(1, 3, H, Wd, generator=g)
    lat = torch.randn(1, 16, h8, w8, generator=g)
    return dict(image=img.clamp(-1, 1).to(dtype), latents=lat.to(dtype))
