"""Blockwise controlnet of Qwen-Image: after every DiT block the NOISE-image tokens receive a correction computed from
themselves and a control image's tokens.

Reference: DiffSynth-Studio/diffsynth/models/qwen_image_controlnet.py:6-74 (`BlockWiseControlBlock`: RMSNorm(x) + RMSNorm(y) -> Linear ->
GELU -> Linear; `QwenImageBlockWiseControlNet`: `img_in` on the patchified control latents + one block per DiT layer),
pipelines/qwen_image_physical.py:157-180 (`QwenImageBlockwiseMultiControlNet`: several controlnets, each active inside its [end, start]
window of the denoising progress, outputs scaled and summed) and :1389-1396 (the call site inside the block loop).
Same parameter names / shapes (checkpoints load unchanged); the arithmetic runs on the native kernels: `pe_rmsnorm`, `pe_add_rows`,
`pe_gemm` with the GELU(erf) epilogue and the gate-residual epilogue (`scale` is the gate).  No PhysicEdit script enables it.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from . import native as nv
from .dit import RMSNorm

DIM = 3072


class BlockWiseControlBlock(nn.Module):
    """qwen_image_controlnet.py:6-27."""

    def __init__(self, dim: int = DIM):
        super().__init__()
        self.x_rms = RMSNorm(dim, eps=1e-6)
        self.y_rms = RMSNorm(dim, eps=1e-6)
        self.input_proj = nn.Linear(dim, dim)
        self.act = nn.GELU()
        self.output_proj = nn.Linear(dim, dim)

    def hidden(self, x2d: torch.Tensor, y2d: torch.Tensor) -> torch.Tensor:
        """GELU(input_proj(rms(x) + rms(y))) for [n, dim] inputs."""
        nat = nv.Native.get(x2d.device.index or 0)
        xr, yr = torch.empty_like(x2d), torch.empty_like(y2d)
        nat.rmsnorm(x2d, xr, self.x_rms.weight, self.x_rms.eps)
        nat.rmsnorm(y2d, yr, self.y_rms.weight, self.y_rms.eps)
        nat.add_rows(xr, yr, xr.shape[0], 1.0)
        return nat.linear(xr, self.input_proj.weight, self.input_proj.bias, nv.EPI_BIAS_GELU_ERF)

    def forward(self, x, y):
        h = self.hidden(x.reshape(-1, x.shape[-1]).contiguous(), y.reshape(-1, y.shape[-1]).contiguous())
        nat = nv.Native.get(h.device.index or 0)
        return nat.linear(h, self.output_proj.weight, self.output_proj.bias).view(x.shape)

    def init_weights(self):
        nn.init.zeros_(self.output_proj.weight)
        nn.init.zeros_(self.output_proj.bias)


class QwenImageBlockWiseControlNet(nn.Module):
    """qwen_image_controlnet.py:30-61."""

    def __init__(self, num_layers: int = 60, in_dim: int = 64, additional_in_dim: int = 0, dim: int = DIM):
        super().__init__()
        self.img_in = nn.Linear(in_dim + additional_in_dim, dim)
        self.controlnet_blocks = nn.ModuleList([BlockWiseControlBlock(dim) for _ in range(num_layers)])

    def init_weight(self):
        nn.init.zeros_(self.img_in.weight)
        nn.init.zeros_(self.img_in.bias)
        for block in self.controlnet_blocks:
            block.init_weights()

    def process_controlnet_conditioning(self, controlnet_conditioning):
        x = controlnet_conditioning
        nat = nv.Native.get(x.device.index or 0)
        x2 = x.reshape(-1, x.shape[-1])
        k = x2.shape[1]
        w = self.img_in.weight
        if k % 8:                                                          # the inpaint variant has 64 + 4 input channels: zero-pad K to a multiple of 8
            pad = 8 - k % 8
            x2, w = torch.nn.functional.pad(x2, (0, pad)), torch.nn.functional.pad(w, (0, pad))
        return nat.linear(x2.contiguous(), w.contiguous(), self.img_in.bias).view(*x.shape[:-1], -1)

    def blockwise_forward(self, img, controlnet_conditioning, block_id):
        return self.controlnet_blocks[block_id](img, controlnet_conditioning)

    @staticmethod
    def state_dict_converter():
        return QwenImageBlockWiseControlNetStateDictConverter()


class QwenImageBlockWiseControlNetStateDictConverter:
    """:64-74: the inpaint controlnet (key hash a9e54e48...) has 4 extra input channels."""

    def from_civitai(self, state_dict):
        extra = {"additional_in_dim": state_dict["img_in.weight"].shape[1] - 64} if state_dict["img_in.weight"].shape[1] != 64 else {}
        return state_dict, extra


class QwenImageBlockwiseMultiControlNet(nn.Module):
    """qwen_image_physical.py:157-180."""

    def __init__(self, models):
        super().__init__()
        self.models = nn.ModuleList(models if isinstance(models, (list, tuple)) else [models])

    def preprocess(self, controlnet_inputs, conditionings: Sequence[torch.Tensor], **kwargs) -> List[torch.Tensor]:
        """Control latents [1, C, h8, w8] -> tokens (2 x 2 patches, channel-major like the DiT's) -> `img_in` of their controlnet."""
        out = []
        for ci, cond in zip(controlnet_inputs, conditionings):
            B, C, Hh, Ww = cond.shape
            tok = cond.view(B, C, Hh // 2, 2, Ww // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (Hh // 2) * (Ww // 2), C * 4)
            out.append(self.models[ci.controlnet_id].process_controlnet_conditioning(tok.contiguous()))
        return out

    @staticmethod
    def active(ci, progress_id, num_inference_steps) -> bool:
        progress = (num_inference_steps - 1 - progress_id) / max(num_inference_steps - 1, 1)
        return not (progress > ci.start + 1e-4 or progress < ci.end - 1e-4)

    def blockwise_forward(self, image, conditionings, controlnet_inputs, progress_id, num_inference_steps, block_id, **kwargs):
        """The reference's call (returns the summed, scaled correction; 0 when no controlnet is active)."""
        res = 0
        for ci, cond in zip(controlnet_inputs, conditionings):
            if self.active(ci, progress_id, num_inference_steps):
                res = res + self.models[ci.controlnet_id].blockwise_forward(image, cond, block_id) * ci.scale
        return res

    def apply_(self, image2d: torch.Tensor, conditionings, controlnet_inputs, progress_id, num_inference_steps, block_id) -> None:
        """image2d [n0, 3072] += sum_i scale_i * block_i(image2d, cond_i), in the reference's rounding order (:1390-1396: every controlnet sees
        the SAME pre-update tokens; res = res + out * scale in bf16; image = slice + res), the last GEMM's epilogue doing the scale-and-add."""
        live = [(ci, c) for ci, c in zip(controlnet_inputs, conditionings) if self.active(ci, progress_id, num_inference_steps)]
        if not live:
            return
        nat = nv.Native.get(image2d.device.index or 0)
        n, dim = image2d.shape
        hidden = [self.models[ci.controlnet_id].controlnet_blocks[block_id].hidden(image2d, c.reshape(n, dim)) for ci, c in live]
        target = image2d if len(live) == 1 else torch.zeros_like(image2d)
        for (ci, _), h in zip(live, hidden):
            blk = self.models[ci.controlnet_id].controlnet_blocks[block_id]
            gate = torch.full((dim,), float(ci.scale), dtype=torch.bfloat16, device=image2d.device)
            nat.gemm([dict(a=h, w=blk.output_proj.weight, bias=blk.output_proj.bias, out=target, gate=gate)], dim, dim, nv.EPI_GATE_RESIDUAL)
        if target is not image2d:
            nat.add_rows(image2d, target, n, 1.0)
