"""QwenImageVAE (single-image encode / decode) executed by libpe_b200 -- SURVEY.md 8f1, the step either side of the denoise loop.

Mirrors DiffSynth-Studio/diffsynth/models/qwen_image_vae.py: same module tree, parameter names and shapes (194 tensors,
conv kernels stay 5-D `[Cout, Cin, 3, kh, kw]`), so the reference's checkpoint loads with `load_state_dict`, and the same
`encode(x, **kwargs)` / `decode(x, **kwargs)` signatures the pipeline calls (pipelines/qwen_image_physical.py:665, 1273, 1298,
1092, 1106; `tiled=`, `tile_size=`, `tile_stride=`, `device=` are accepted and ignored exactly as the reference ignores them).

How it runs (bf16 on an sm_100 GPU only, no fallback):
  * activation maps are channels-last `[H*W, C]`, so a pixel is a row of the implicit-GEMM convolution `pe_conv2d` -- the DiT's
    tcgen05 GEMM kernel whose TMA producer reads one shifted pixel patch per tap (no im2col buffer, zero padding = TMA
    out-of-bounds fill).  At T = 1 without a feature cache a QwenImageCausalConv3d is a 2-D conv with the LAST temporal slice
    of its kernel (the causal padding puts zeros in front of the only frame, :39-51), and `time_conv` never runs (:259-301);
  * conv weights are repacked once (`prepare()`) to `[N, taps * round_up(C, 64)]`, tap-major; the ZeroPad2d((0,1,0,1)) +
    stride-2 conv of the downsample layers (:246-249) becomes `pe_space_to_depth` + a 2x2 stride-1 conv over 4C channels;
  * RMS norm + SiLU, 2x nearest upsample, the latent (de)normalisation and the NCHW<->NHWC turns are single HBM passes
    (vae_kernels.cu); the residual add is the conv's epilogue; the single-head (d = 384) attention of the mid block is
    scores = q.k^T (GEMM, fp32 out) -> row softmax -> P.V (GEMM), chunked over query rows.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import native as nv

LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497,
               0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251,
              1.9160]

_PAIR_MIN_PIXELS = 16384      # >= 64 CTA-pair tiles of 256 pixels
_ATTN_QUERY_CHUNK = 8192      # query rows per score block: 8192 x S fp32 scores (512 MB at S = 16384)


# ---- module tree (parameters only; the forward passes below never call nn.Conv*.forward) ---------------------------------------
class _RMSNorm(nn.Module):
    def __init__(self, dim: int, images: bool):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones((dim, 1, 1) if images else (dim, 1, 1, 1)))


class _ResidualBlock(nn.Module):
    """qwen_image_vae.py:81-152."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm1 = _RMSNorm(cin, images=False)
        self.conv1 = nn.Conv3d(cin, cout, 3)
        self.norm2 = _RMSNorm(cout, images=False)
        self.conv2 = nn.Conv3d(cout, cout, 3)
        self.conv_shortcut = nn.Conv3d(cin, cout, 1) if cin != cout else nn.Identity()


class _AttentionBlock(nn.Module):
    """:156-199."""

    def __init__(self, dim: int):
        super().__init__()
        self.norm = _RMSNorm(dim, images=True)
        self.to_qkv = nn.Conv2d(dim, dim * 3, 1)
        self.proj = nn.Conv2d(dim, dim, 1)


class _Resample(nn.Module):
    """:218-301.  `resample.1` is the 2-D conv; `time_conv` exists for checkpoint compatibility (unused for single images)."""

    def __init__(self, dim: int, mode: str):
        super().__init__()
        self.mode = mode
        if mode.startswith("upsample"):
            self.resample = nn.Sequential(nn.Identity(), nn.Conv2d(dim, dim // 2, 3, padding=1))
            if mode == "upsample3d":
                self.time_conv = nn.Conv3d(dim, dim * 2, (3, 1, 1))
        else:
            self.resample = nn.Sequential(nn.Identity(), nn.Conv2d(dim, dim, 3, stride=(2, 2)))
            if mode == "downsample3d":
                self.time_conv = nn.Conv3d(dim, dim, (3, 1, 1))


class _MidBlock(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.attentions = nn.ModuleList([_AttentionBlock(dim)])
        self.resnets = nn.ModuleList([_ResidualBlock(dim, dim), _ResidualBlock(dim, dim)])


class _Encoder(nn.Module):
    """:344-448."""

    def __init__(self, dim, z_dim, dim_mult, num_res_blocks, temporal_downsample):
        super().__init__()
        dims = [dim * u for u in [1] + list(dim_mult)]
        self.conv_in = nn.Conv3d(3, dims[0], 3)
        blocks: List[nn.Module] = []
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(num_res_blocks):
                blocks.append(_ResidualBlock(cin, cout))
                cin = cout
            if i != len(dim_mult) - 1:
                blocks.append(_Resample(cout, "downsample3d" if temporal_downsample[i] else "downsample2d"))
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = _MidBlock(dims[-1])
        self.norm_out = _RMSNorm(dims[-1], images=False)
        self.conv_out = nn.Conv3d(dims[-1], z_dim, 3)


class _UpBlock(nn.Module):
    def __init__(self, cin, cout, num_res_blocks, upsample_mode: Optional[str]):
        super().__init__()
        res = []
        for _ in range(num_res_blocks + 1):
            res.append(_ResidualBlock(cin, cout))
            cin = cout
        self.resnets = nn.ModuleList(res)
        self.upsamplers = nn.ModuleList([_Resample(cout, upsample_mode)]) if upsample_mode is not None else None


class _Decoder(nn.Module):
    """:522-637."""

    def __init__(self, dim, z_dim, dim_mult, num_res_blocks, temporal_upsample):
        super().__init__()
        dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
        self.conv_in = nn.Conv3d(z_dim, dims[0], 3)
        self.mid_block = _MidBlock(dims[0])
        ups = []
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            if i > 0:
                cin = cin // 2
            mode = None
            if i != len(dim_mult) - 1:
                mode = "upsample3d" if temporal_upsample[i] else "upsample2d"
            ups.append(_UpBlock(cin, cout, num_res_blocks, mode))
        self.up_blocks = nn.ModuleList(ups)
        self.norm_out = _RMSNorm(dims[-1], images=False)
        self.conv_out = nn.Conv3d(dims[-1], 3, 3)


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class QwenImageVAE(nn.Module):
    """Drop-in for diffsynth.models.qwen_image_vae.QwenImageVAE (:640-735) for 4-D image tensors."""

    def __init__(self, base_dim: int = 96, z_dim: int = 16, dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2, attn_scales=(),
                 temperal_downsample=(False, True, True), dropout: float = 0.0):
        super().__init__()
        if z_dim != 16 or list(attn_scales):
            raise ValueError("native QwenImageVAE: z_dim must be 16 and attn_scales empty (the released configuration)")
        self.z_dim = z_dim
        self.temperal_downsample = list(temperal_downsample)
        self.temperal_upsample = self.temperal_downsample[::-1]
        self.encoder = _Encoder(base_dim, z_dim * 2, list(dim_mult), num_res_blocks, self.temperal_downsample)
        self.quant_conv = nn.Conv3d(z_dim * 2, z_dim * 2, 1)
        self.post_quant_conv = nn.Conv3d(z_dim, z_dim, 1)
        self.decoder = _Decoder(base_dim, z_dim, list(dim_mult), num_res_blocks, self.temperal_upsample)
        self.mean = torch.tensor(LATENT_MEAN, device="cpu").view(1, 16, 1, 1, 1)        # plain attributes, as in the reference (:703-704)
        self.std = 1 / torch.tensor(LATENT_STD, device="cpu").view(1, 16, 1, 1, 1)
        self._packed: Optional[Dict[str, Tuple]] = None

    @staticmethod
    def state_dict_converter():
        return QwenImageVAEStateDictConverter()

    # ---- weight repacking ---------------------------------------------------------------------------------------------------
    def load_state_dict(self, *args, **kwargs):
        self._packed = None
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    @torch.no_grad()
    def prepare(self) -> None:
        """Repack every convolution for pe_conv2d / pe_gemm (layout work only; done once per weight load)."""
        dev = self.post_quant_conv.weight.device          # (device / dtype are enforced where kernels are launched: _ctx)
        P: Dict[str, Tuple] = {}

        def pad_rows(w2d: torch.Tensor, bias: torch.Tensor):
            n = w2d.shape[0]
            n8 = _round_up(n, 8)
            if n8 != n:
                w2d = torch.cat([w2d, w2d.new_zeros(n8 - n, w2d.shape[1])])
                bias = torch.cat([bias, bias.new_zeros(n8 - n)])
            return w2d.contiguous(), bias.contiguous(), n8

        def pack_conv(name: str, conv: nn.Module, cin_store: Optional[int] = None, rows: Optional[int] = None):
            """3x3 (or 1x1) stride-1 conv -> [N8, kh*kw*cpad]; `cin_store`: channels of the (zero padded) input map."""
            w = conv.weight
            if w.dim() == 5:
                w = w[:, :, -1]                      # the only temporal slice that meets the frame
            if rows is not None:
                w, b = w[:rows], conv.bias[:rows]
            else:
                b = conv.bias
            n, c, kh, kw = w.shape
            # 3x3: every tap starts on a 64-channel block; 1x1 is a plain GEMM whose K is the stored channel count
            cpad = _round_up(cin_store or c, 64) if kh * kw > 1 else (cin_store or c)
            wp = w.new_zeros(n, kh, kw, cpad)
            wp[..., :c] = w.permute(0, 2, 3, 1)
            w2d, b, n8 = pad_rows(wp.reshape(n, kh * kw * cpad), b)
            P[name] = (w2d, b, n8, kh, kw, cin_store or c)

        def pack_down(name: str, conv: nn.Conv2d):
            """pad (0,1,0,1) + 3x3 stride 2  ==  2x2 stride 1 over the space-to-depth map (4C channels, phase-major)."""
            w = conv.weight                          # [N, C, 3, 3]
            n, c = w.shape[:2]
            wp = w.new_zeros(n, 2, 2, 4, c)          # [n, a, b, phase = py*2+px, c]
            for ky in range(3):
                for kx in range(3):
                    wp[:, ky // 2, kx // 2, (ky % 2) * 2 + (kx % 2)] = w[:, :, ky, kx]
            w2d, b, n8 = pad_rows(wp.reshape(n, 2 * 2 * 4 * c), conv.bias)
            P[name] = (w2d, b, n8, 2, 2, 4 * c)

        def pack_res(name: str, blk: _ResidualBlock):
            pack_conv(name + ".conv1", blk.conv1)
            pack_conv(name + ".conv2", blk.conv2)
            P[name + ".norm1"] = blk.norm1.gamma.reshape(-1).contiguous()
            P[name + ".norm2"] = blk.norm2.gamma.reshape(-1).contiguous()
            if not isinstance(blk.conv_shortcut, nn.Identity):
                pack_conv(name + ".conv_shortcut", blk.conv_shortcut)

        def pack_mid(name: str, mid: _MidBlock):
            pack_res(name + ".resnets.0", mid.resnets[0])
            pack_res(name + ".resnets.1", mid.resnets[1])
            a = mid.attentions[0]
            P[name + ".attn.norm"] = a.norm.gamma.reshape(-1).contiguous()
            c = a.proj.weight.shape[0]
            P[name + ".attn.qkv_w"] = a.to_qkv.weight.reshape(3 * c, c).contiguous()
            P[name + ".attn.qkv_b"] = a.to_qkv.bias.contiguous()
            P[name + ".attn.proj_w"] = a.proj.weight.reshape(c, c).contiguous()
            P[name + ".attn.proj_b"] = a.proj.bias.contiguous()

        enc, dec = self.encoder, self.decoder
        pack_conv("encoder.conv_in", enc.conv_in, cin_store=64)
        for i, blk in enumerate(enc.down_blocks):
            if isinstance(blk, _ResidualBlock):
                pack_res(f"encoder.down_blocks.{i}", blk)
            else:
                pack_down(f"encoder.down_blocks.{i}", blk.resample[1])
        pack_mid("encoder.mid_block", enc.mid_block)
        P["encoder.norm_out"] = enc.norm_out.gamma.reshape(-1).contiguous()
        pack_conv("encoder.conv_out", enc.conv_out)
        pack_conv("quant_conv", self.quant_conv, cin_store=64, rows=16)         # encode keeps x[:, :16] only (:711)
        pack_conv("post_quant_conv", self.post_quant_conv, cin_store=64)
        pack_conv("decoder.conv_in", dec.conv_in, cin_store=64)
        pack_mid("decoder.mid_block", dec.mid_block)
        for b, ub in enumerate(dec.up_blocks):
            for r, blk in enumerate(ub.resnets):
                pack_res(f"decoder.up_blocks.{b}.resnets.{r}", blk)
            if ub.upsamplers is not None:
                pack_conv(f"decoder.up_blocks.{b}.upsample", ub.upsamplers[0].resample[1])
        P["decoder.norm_out"] = dec.norm_out.gamma.reshape(-1).contiguous()
        pack_conv("decoder.conv_out", dec.conv_out)
        P["mean"] = self.mean.reshape(-1).to(device=dev, dtype=torch.bfloat16)          # self.mean.to(dtype=x.dtype) (:712, :723)
        P["stdinv"] = self.std.reshape(-1).to(device=dev, dtype=torch.bfloat16)
        P["ones"] = torch.ones(512, dtype=torch.bfloat16, device=dev)
        self._packed = P

    # ---- layer runners --------------------------------------------------------------------------------------------------------
    def _ctx(self, x: torch.Tensor):
        if not x.is_cuda or x.dtype != torch.bfloat16:
            raise nv.NativeUnavailable(f"native QwenImageVAE runs in bfloat16 on an sm_100 GPU only (got {x.dtype} on {x.device})")
        if self._packed is None:
            self.prepare()
        return nv.Native.get(x.device.index or 0), self._packed

    @staticmethod
    def _conv(nat, P, name, x, H, W, out=None, residual=False, ldo=None):
        w2d, b, n8, kh, kw, cin = P[name]
        if out is None and ldo is None:
            out = torch.empty((H * W, n8), dtype=torch.bfloat16, device=x.device)
        elif out is None:                       # narrow output inside a zero-padded `ldo`-wide map (feeds a K = ldo layer)
            out = torch.zeros((H * W, ldo), dtype=torch.bfloat16, device=x.device)
        nat.tag = f"vae_conv{kh}x{kw}_{H}x{W}_{cin}to{n8}"
        # 192 / 384-channel layers on maps with enough tiles to fill the GPU run on CTA pairs (cta_group::2): each CTA loads half of the
        # weight rows, which halves their L2 -> SM weight traffic; layers of <= 128 channels pair two pixel patches inside one CTA instead
        # (pe_conv2d's default), which doubles the MMA work per k-block iteration of the producer / issuer instruction chains
        pair = H * W >= _PAIR_MIN_PIXELS
        if kh == 1 and kw == 1:
            nat.gemm([dict(a=x, w=w2d, bias=b, out=out, gate=P["ones"] if residual else None)], n8, w2d.shape[1],
                     nv.EPI_GATE_RESIDUAL if residual else nv.EPI_BIAS, nv.GEMM_FLAG_TRIM_N | (nv.GEMM_FLAG_CTA_PAIR if pair else 0))
        else:
            nat.conv2d(x, H, W, cin, w2d, b, out, n8, kh, kw, 1 if kh == 3 else 0, nv.EPI_GATE_RESIDUAL if residual else nv.EPI_BIAS,
                       gate=P["ones"] if residual else None, flags=nv.CONV_FLAG_CTA_PAIR if (pair and n8 > 128) else 0)
        return out

    @staticmethod
    def _norm(nat, gamma, x, act=True):
        out = torch.empty_like(x)
        nat.tag = f"vae_rmsnorm_{x.shape[0]}x{gamma.numel()}"
        nat.channel_rmsnorm(x, out, gamma.numel(), gamma, act)
        return out

    def _res(self, nat, P, name, x, H, W):
        """QwenImageResidualBlock.forward (:112-152): the residual add is conv2's epilogue, in place on the shortcut."""
        h = self._conv(nat, P, name + ".conv_shortcut", x, H, W) if (name + ".conv_shortcut") in P else x
        t = self._norm(nat, P[name + ".norm1"], x)
        t = self._conv(nat, P, name + ".conv1", t, H, W)
        t = self._norm(nat, P[name + ".norm2"], t)
        return self._conv(nat, P, name + ".conv2", t, H, W, out=h, residual=True)

    def _attn(self, nat, P, name, x):
        """QwenImageAttentionBlock.forward (:173-199): one head of dim C over all S = H*W positions."""
        S, C = x.shape
        S8 = _round_up(S, 8)
        dev = x.device
        xn = self._norm(nat, P[name + ".norm"], x, act=False)
        qw, qb = P[name + ".qkv_w"], P[name + ".qkv_b"]
        q = torch.empty((S, C), dtype=torch.bfloat16, device=dev)
        k = torch.zeros((S8, C), dtype=torch.bfloat16, device=dev) if S8 != S else torch.empty((S, C), dtype=torch.bfloat16, device=dev)
        v = torch.empty((S, C), dtype=torch.bfloat16, device=dev)
        for i, dst in enumerate((q, k, v)):
            nat.tag = "vae_attn_qkv"
            nat.gemm([dict(a=xn, w=qw[i * C:(i + 1) * C], bias=qb[i * C:(i + 1) * C], out=dst[:S])], C, C, nv.EPI_BIAS, nv.GEMM_FLAG_TRIM_N)
        vt = torch.zeros((C, S8), dtype=torch.bfloat16, device=dev) if S8 != S else torch.empty((C, S), dtype=torch.bfloat16, device=dev)
        nat.tag = "vae_attn_transpose"
        nat.transpose(v, vt[:, :S])
        o = torch.empty((S, C), dtype=torch.bfloat16, device=dev)
        rq = min(S, _ATTN_QUERY_CHUNK)
        scores = torch.empty((rq, S8), dtype=torch.float32, device=dev)
        probs = torch.empty((rq, S8), dtype=torch.bfloat16, device=dev)
        for r0 in range(0, S, rq):
            r1 = min(S, r0 + rq)
            nat.tag = "vae_attn_scores"
            nat.gemm([dict(a=q[r0:r1], w=k, bias=None, out=scores[:r1 - r0])], S8, C, nv.EPI_F32, nv.GEMM_FLAG_TRIM_N)
            nat.tag = "vae_attn_softmax"
            nat.softmax_rows(scores[:r1 - r0], probs[:r1 - r0], S, C ** -0.5)
            nat.tag = "vae_attn_pv"
            nat.gemm([dict(a=probs[:r1 - r0], w=vt, bias=None, out=o[r0:r1])], C, S8, nv.EPI_BIAS, nv.GEMM_FLAG_TRIM_N)
        nat.tag = "vae_attn_proj"
        nat.gemm([dict(a=o, w=P[name + ".proj_w"], bias=P[name + ".proj_b"], out=x, gate=P["ones"])], C, C, nv.EPI_GATE_RESIDUAL,
                 nv.GEMM_FLAG_TRIM_N)
        return x

    def _mid(self, nat, P, name, x, H, W):
        x = self._res(nat, P, name + ".resnets.0", x, H, W)
        x = self._attn(nat, P, name + ".attn", x)
        return self._res(nat, P, name + ".resnets.1", x, H, W)

    # ---- public API -------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        """latents [B,16,h,w] (or [B,16,1,h,w]) -> image [B,3,8h,8w] (:719-731)."""
        five_d = x.dim() == 5
        if five_d:
            if x.shape[2] != 1:
                raise NotImplementedError("native QwenImageVAE handles single frames (T = 1) only")
            x = x[:, :, 0]
        nat, P = self._ctx(x)
        out = torch.empty((x.shape[0], 3, x.shape[2] * 8, x.shape[3] * 8), dtype=torch.bfloat16, device=x.device)
        for b in range(x.shape[0]):
            self._decode_one(nat, P, x[b].contiguous(), out[b])
        if kwargs.get("check_async", True):
            nat.check_async()          # the image goes to the host next: a timed-out kernel must raise here, not return garbage
        return out.unsqueeze(2) if five_d else out

    def _decode_one(self, nat, P, lat, out):
        H, W = lat.shape[1], lat.shape[2]
        z = torch.zeros((H * W, 64), dtype=torch.bfloat16, device=lat.device)
        nat.tag = "vae_layout"
        nat.nchw_to_nhwc(lat, z, 16, op=1, p0=P["mean"], p1=P["stdinv"])               # x / std + mean (:724-725)
        x = self._conv(nat, P, "post_quant_conv", z, H, W, ldo=64)
        x = self._conv(nat, P, "decoder.conv_in", x, H, W)
        x = self._mid(nat, P, "decoder.mid_block", x, H, W)
        for b, ub in enumerate(self.decoder.up_blocks):
            for r in range(len(ub.resnets)):
                x = self._res(nat, P, f"decoder.up_blocks.{b}.resnets.{r}", x, H, W)
            if ub.upsamplers is not None:
                C = x.shape[1]
                up = torch.empty((4 * H * W, C), dtype=torch.bfloat16, device=x.device)
                nat.tag = "vae_upsample"
                nat.upsample2x(x, up, H, W, C)
                H, W = 2 * H, 2 * W
                x = self._conv(nat, P, f"decoder.up_blocks.{b}.upsample", up, H, W)
        x = self._norm(nat, P["decoder.norm_out"], x)
        x = self._conv(nat, P, "decoder.conv_out", x, H, W)                               # [H*W, 8], channels 0..2 live
        nat.tag = "vae_layout"
        nat.nhwc_to_nchw(x, out, 3)

    @torch.no_grad()
    def encode(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        """image [B,3,H,W] in [-1,1] (H, W multiples of 8) -> normalised latents [B,16,H/8,W/8] (:706-717)."""
        five_d = x.dim() == 5
        if five_d:
            if x.shape[2] != 1:
                raise NotImplementedError("native QwenImageVAE handles single frames (T = 1) only")
            x = x[:, :, 0]
        if x.shape[2] % 8 or x.shape[3] % 8:
            raise ValueError(f"QwenImageVAE.encode: H and W must be multiples of 8, got {tuple(x.shape)}")
        nat, P = self._ctx(x)
        out = torch.empty((x.shape[0], 16, x.shape[2] // 8, x.shape[3] // 8), dtype=torch.bfloat16, device=x.device)
        for b in range(x.shape[0]):
            self._encode_one(nat, P, x[b].contiguous(), out[b])
        if kwargs.get("check_async", True):
            nat.check_async()          # once per image, before the latents enter a 50-step loop
        return out.unsqueeze(2) if five_d else out

    def _encode_one(self, nat, P, img, out):
        H, W = img.shape[1], img.shape[2]
        z = torch.zeros((H * W, 64), dtype=torch.bfloat16, device=img.device)
        nat.tag = "vae_layout"
        nat.nchw_to_nhwc(img, z, 3)
        x = self._conv(nat, P, "encoder.conv_in", z, H, W)
        for i, blk in enumerate(self.encoder.down_blocks):
            name = f"encoder.down_blocks.{i}"
            if isinstance(blk, _ResidualBlock):
                x = self._res(nat, P, name, x, H, W)
            else:
                C = x.shape[1]
                s2d = torch.empty((H * W // 4, 4 * C), dtype=torch.bfloat16, device=x.device)
                nat.tag = "vae_space_to_depth"
                nat.space_to_depth(x, s2d, H, W, C)
                H, W = H // 2, W // 2
                x = self._conv(nat, P, name, s2d, H, W)
        x = self._mid(nat, P, "encoder.mid_block", x, H, W)
        x = self._norm(nat, P["encoder.norm_out"], x)
        x = self._conv(nat, P, "encoder.conv_out", x, H, W, ldo=64)                         # 32 channels in a zero-padded 64-wide map
        x = self._conv(nat, P, "quant_conv", x, H, W)                                       # rows 0..15 of quant_conv only
        nat.tag = "vae_layout"
        nat.nhwc_to_nchw(x, out, 16, op=2, p0=P["mean"], p1=P["stdinv"])                 # (x - mean) * std (:712-714)


class QwenImageVAEStateDictConverter:
    """:737-742 (identity)."""

    def from_diffusers(self, state_dict):
        return state_dict

    def from_civitai(self, state_dict):
        return state_dict
