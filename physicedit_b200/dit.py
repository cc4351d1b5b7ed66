"""Qwen-Image-Edit DiT with the reference's exact module / parameter layout, executed by libpe_b200.

Mirrors DiffSynth-Studio/diffsynth/models/qwen_image_dit.py (QwenImageDiT, QwenImageTransformerBlock,
QwenDoubleStreamAttention, QwenFeedForward, ApproximateGELU, QwenEmbedRope) and models/utils.py
(RMSNorm, AdaLayerNorm, TimestepEmbeddings): identical attribute names and parameter shapes, so
`state_dict()` hashes to the registry value 0319a1cb19835fb510907dd3367c95ff
(configs/model_config.py:21) and the reference's loader, LoRA loader, PEFT injection and checkpoints
work unchanged.  The arithmetic never runs through torch ops: `DiTEngine` drives the C-ABI kernels
(9 launches per block instead of ~70 ATen launches).  bf16 on an sm_100 GPU only -- no fallback.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import native as nv

DIM = 3072
NUM_HEADS = 24
HEAD_DIM = 128


# ------------------------------------------------------------------------------------------------
# parameter containers (names / shapes as the reference)
# ------------------------------------------------------------------------------------------------
class RMSNorm(nn.Module):
    """models/utils.py:241-257."""

    def __init__(self, dim, eps, elementwise_affine=True):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones((dim,))) if elementwise_affine else None

    def forward(self, hidden_states):
        nat = nv.Native.get(hidden_states.device.index or 0)
        x = hidden_states.reshape(-1, hidden_states.shape[-1]).contiguous()
        out = torch.empty_like(x)
        nat.rmsnorm(x, out, self.weight, self.eps)
        return out.view(hidden_states.shape)


class _TimestepProj(nn.Module):
    """_DiffusersCompatibleTimestepProj (models/utils.py:259-271)."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.linear_1 = nn.Linear(dim_in, dim_out)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(dim_out, dim_out)


class TimestepEmbeddings(nn.Module):
    """models/utils.py:274-293 with flip_sin_to_cos, scale=1000, align_dtype_to_timestep (qwen_image_dit.py:413)."""

    def __init__(self, dim_in=256, dim_out=DIM):
        super().__init__()
        self.timestep_embedder = _TimestepProj(dim_in, dim_out)

    def forward(self, timestep: torch.Tensor, dtype=torch.bfloat16, raw: bool = False) -> torch.Tensor:
        """timestep: bf16 [1] holding t/1000 as in the reference (raw=False), or the loop's bf16(t) with the
        division fused into the kernel (raw=True).  Returns temb [1, 3072]."""
        nat = nv.Native.get(timestep.device.index or 0)
        dev = timestep.device
        sinus = torch.empty(256, dtype=torch.bfloat16, device=dev)
        nat.timestep_embedding(timestep, sinus, raw)
        h = torch.empty(1, DIM, dtype=torch.bfloat16, device=dev)
        te = self.timestep_embedder
        nat.gemv(sinus.view(1, 256), te.linear_1.weight, te.linear_1.bias, h, 0, 1)
        out = torch.empty(1, DIM, dtype=torch.bfloat16, device=dev)
        nat.gemv(h, te.linear_2.weight, te.linear_2.bias, out, 0, 0)
        return out


class AdaLayerNorm(nn.Module):
    """models/utils.py:296-309, single=True: linear(silu(emb)) -> (scale, shift)."""

    def __init__(self, dim, single=True):
        super().__init__()
        assert single
        self.single = True
        self.linear = nn.Linear(dim, dim * 2)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)


class ApproximateGELU(nn.Module):
    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)


class QwenFeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, dropout=0.0):
        super().__init__()
        inner = int(dim * 4)
        self.net = nn.ModuleList([ApproximateGELU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out)])


class QwenDoubleStreamAttention(nn.Module):
    def __init__(self, dim_a, dim_b, num_heads, head_dim):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, head_dim
        self.to_q = nn.Linear(dim_a, dim_a)
        self.to_k = nn.Linear(dim_a, dim_a)
        self.to_v = nn.Linear(dim_a, dim_a)
        self.norm_q = RMSNorm(head_dim, eps=1e-6)
        self.norm_k = RMSNorm(head_dim, eps=1e-6)
        self.add_q_proj = nn.Linear(dim_b, dim_b)
        self.add_k_proj = nn.Linear(dim_b, dim_b)
        self.add_v_proj = nn.Linear(dim_b, dim_b)
        self.norm_added_q = RMSNorm(head_dim, eps=1e-6)
        self.norm_added_k = RMSNorm(head_dim, eps=1e-6)
        self.to_out = nn.Sequential(nn.Linear(dim_a, dim_a))
        self.to_add_out = nn.Linear(dim_b, dim_b)


class QwenImageTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, eps=1e-6):
        super().__init__()
        self.dim, self.num_attention_heads, self.attention_head_dim = dim, num_attention_heads, attention_head_dim
        self.img_mod = nn.Sequential(nn.SiLU(), nn.Linear(dim, 6 * dim))
        self.img_norm1 = nn.LayerNorm(dim, elementwise_affine=False, eps=eps)
        self.attn = QwenDoubleStreamAttention(dim, dim, num_attention_heads, attention_head_dim)
        self.img_norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=eps)
        self.img_mlp = QwenFeedForward(dim=dim, dim_out=dim)
        self.txt_mod = nn.Sequential(nn.SiLU(), nn.Linear(dim, 6 * dim, bias=True))
        self.txt_norm1 = nn.LayerNorm(dim, elementwise_affine=False, eps=eps)
        self.txt_norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=eps)
        self.txt_mlp = QwenFeedForward(dim=dim, dim_out=dim)
        self._owner = None      # set by QwenImageDiT: (dit, index) for the engine

    def forward(self, image, text, temb, image_rotary_emb=None, attention_mask=None, enable_fp8_attention=False):
        """Same signature / return order as the reference (qwen_image_dit.py:359-401): returns (text, image).
        image [1,S_img,3072], text [1,T,3072], temb [1,3072]; image_rotary_emb = (vid complex [S_img,64], txt complex [T,64])."""
        dit, idx = self._owner
        eng = dit.engine()
        T, S_img = text.shape[1], image.shape[1]
        x = torch.cat([text[0], image[0]], dim=0).contiguous()
        vid, txt = image_rotary_emb
        rope = torch.view_as_real(torch.cat([txt, vid], dim=0).to(torch.complex64)).contiguous().to(x.device)
        ws = eng.workspace(S_img, T)
        mods = eng.block_mods(temb.reshape(1, DIM), [idx])
        mask_u8 = None
        if attention_mask is not None:                   # the reference's additive 0 / -inf mask [1, 1, S, S] (EliGen)
            mask_u8 = (attention_mask.reshape(attention_mask.shape[-2], attention_mask.shape[-1]) == 0).to(torch.uint8).contiguous()
        eng.run_block(idx, x, T, mods[0, 0], rope, ws, mask_u8, bool(enable_fp8_attention) and mask_u8 is None)
        return x[:T].unsqueeze(0), x[T:].unsqueeze(0)


class QwenEmbedRope(nn.Module):
    """qwen_image_dit.py:60-165: 3-axis RoPE tables (complex64 on the host), scale_rope=True."""

    def __init__(self, theta: int, axes_dim: Sequence[int], scale_rope=False):
        super().__init__()
        self.theta, self.axes_dim, self.scale_rope = theta, list(axes_dim), scale_rope
        self._build(4096)
        self.rope_cache: Dict[str, torch.Tensor] = {}

    def rope_params(self, index, dim, theta=10000):
        assert dim % 2 == 0
        freqs = torch.outer(index, 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float32).div(dim)))
        return torch.polar(torch.ones_like(freqs), freqs)

    def _build(self, n):
        pos_index = torch.arange(n)
        neg_index = torch.arange(n).flip(0) * -1 - 1
        self.pos_freqs = torch.cat([self.rope_params(pos_index, d, self.theta) for d in self.axes_dim], dim=1)
        self.neg_freqs = torch.cat([self.rope_params(neg_index, d, self.theta) for d in self.axes_dim], dim=1)

    def _expand_pos_freqs_if_needed(self, video_fhw, txt_seq_lens):
        if isinstance(video_fhw, list):
            video_fhw = tuple(max([i[j] for i in video_fhw]) for j in range(3))
        _, height, width = video_fhw
        max_vid_index = max(height // 2, width // 2) if self.scale_rope else max(height, width)
        required = max_vid_index + max(txt_seq_lens)
        if required > self.pos_freqs.shape[0]:
            self._build(math.ceil(required / 512) * 512)

    def _axis_table(self, idx, frame, height, width):
        """[frame * height * width, 64] complex: frame axis at index idx.., centred height / width axes when scale_rope (:131-150)."""
        split = [x // 2 for x in self.axes_dim]
        fpos, fneg = self.pos_freqs.split(split, dim=1), self.neg_freqs.split(split, dim=1)
        f_frame = fpos[0][idx: idx + frame].view(frame, 1, 1, -1).expand(frame, height, width, -1)
        if self.scale_rope:
            f_h = torch.cat([fneg[1][-(height - height // 2):], fpos[1][: height // 2]], dim=0)
            f_w = torch.cat([fneg[2][-(width - width // 2):], fpos[2][: width // 2]], dim=0)
        else:
            f_h, f_w = fpos[1][:height], fpos[2][:width]
        f_h = f_h.view(1, height, 1, -1).expand(frame, height, width, -1)
        f_w = f_w.view(1, 1, width, -1).expand(frame, height, width, -1)
        return torch.cat([f_frame, f_h, f_w], dim=-1).reshape(frame * height * width, -1).clone().contiguous()

    def _tables(self, video_fhw, txt_seq_lens, device, sampling):
        self._expand_pos_freqs_if_needed(video_fhw, txt_seq_lens)
        vid_freqs = []
        max_vid_index = 0
        for idx, (frame, height, width) in enumerate(video_fhw):
            key = f"{idx}_{height}_{width}"
            if sampling and idx > 0 and f"0_{height}_{width}" not in self.rope_cache:
                # edit_rope_interpolation (:179-194): an edit image of another size takes the positions of the NOISE image's grid sampled at
                # linspace(0, n0 - 1, n).long() rows / columns, with its own frame index.  (Like the reference, the entry lands in the cache
                # shared with forward(): whichever call creates a key first decides it.)
                f0, h0, w0 = video_fhw[0]
                grid0 = self.rope_cache[f"0_{h0}_{w0}"].reshape(f0, h0, w0, -1)
                hi = torch.linspace(0, h0 - 1, height).long()
                wi = torch.linspace(0, w0 - 1, width).long()
                hg, wg = torch.meshgrid(hi, wi, indexing="ij")
                sampled = grid0[:, hg, wg, :]
                n_frame = self.axes_dim[0] // 2
                sampled[:, :, :, :n_frame] = self.pos_freqs[idx: idx + frame, :n_frame].view(frame, 1, 1, -1).expand(frame, height, width, -1)
                self.rope_cache[key] = sampled.reshape(frame * height * width, -1).clone()
            if key not in self.rope_cache:
                self.rope_cache[key] = self._axis_table(idx, frame, height, width)
            vid_freqs.append(self.rope_cache[key].contiguous())
            max_vid_index = max(height // 2, width // 2, max_vid_index) if self.scale_rope else max(height, width, max_vid_index)
        max_len = max(txt_seq_lens)
        txt_freqs = self.pos_freqs[max_vid_index: max_vid_index + max_len, ...]
        vid_freqs = torch.cat(vid_freqs, dim=0)
        if device is not None:
            vid_freqs, txt_freqs = vid_freqs.to(device), txt_freqs.to(device)
        return vid_freqs, txt_freqs

    def forward(self, video_fhw, txt_seq_lens, device=None):
        """:122-165."""
        return self._tables(video_fhw, txt_seq_lens, device, sampling=False)

    def forward_sampling(self, video_fhw, txt_seq_lens, device=None):
        """:168-225 (`edit_rope_interpolation=True`, qwen_image_physical.py:1367-1368)."""
        return self._tables(video_fhw, txt_seq_lens, device, sampling=True)


class QwenImageDiTStateDictConverter:
    def from_civitai(self, state_dict):
        return state_dict

    def from_diffusers(self, state_dict):
        return state_dict


class QwenImageDiT(nn.Module):
    """qwen_image_dit.py:404-546.  `forward` is the stock (non-edit) entry; PhysicEdit calls
    physicedit_b200.model_fn.model_fn_qwen_image, which uses `engine()` directly."""

    def __init__(self, num_layers: int = 60):
        super().__init__()
        self.pos_embed = QwenEmbedRope(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
        self.time_text_embed = TimestepEmbeddings(256, DIM)
        self.txt_norm = RMSNorm(3584, eps=1e-6)
        self.img_in = nn.Linear(64, DIM)
        self.txt_in = nn.Linear(3584, DIM)
        self.transformer_blocks = nn.ModuleList(
            [QwenImageTransformerBlock(dim=DIM, num_attention_heads=NUM_HEADS, attention_head_dim=HEAD_DIM) for _ in range(num_layers)])
        self.norm_out = AdaLayerNorm(DIM, single=True)
        self.proj_out = nn.Linear(DIM, 64)
        self._engine: Optional["DiTEngine"] = None
        for i, b in enumerate(self.transformer_blocks):
            object.__setattr__(b, "_owner", (self, i))

    # Timestep-keyed caches of the engine (temb, modulation vectors, norm_out scale/shift) are derived from the weights.  An
    # in-place load (`load_state_dict(ckpt, strict=False)` without assign, as INTEGRATION.md shows) keeps every data_ptr, so the
    # engine's pointer check cannot see it: drop the caches on every load and on every device / dtype move.
    def load_state_dict(self, *args, **kwargs):
        res = super().load_state_dict(*args, **kwargs)
        if self._engine is not None:
            self._engine.invalidate()
        return res

    def _apply(self, fn, *args, **kwargs):
        res = super()._apply(fn, *args, **kwargs)
        if getattr(self, "_engine", None) is not None:
            object.__setattr__(self, "_engine", None)          # storage may have moved: re-pack lazily
        return res

    def engine(self) -> "DiTEngine":
        if self._engine is None:
            object.__setattr__(self, "_engine", DiTEngine(self))
        self._engine.refresh_if_stale()
        return self._engine

    def forward(self, latents=None, timestep=None, prompt_emb=None, prompt_emb_mask=None, height=None, width=None):
        from .model_fn import model_fn_qwen_image
        out, _ = model_fn_qwen_image(dit=self, latents=latents, timestep=timestep * 1000 if timestep is not None else None,
                                     prompt_emb=prompt_emb, prompt_emb_mask=prompt_emb_mask, special_token_mask=None,
                                     height=height, width=width, is_train=False)
        return out

    @staticmethod
    def state_dict_converter():
        return QwenImageDiTStateDictConverter()


# ------------------------------------------------------------------------------------------------
# the engine: packed weights, workspaces, kernel sequencing
# ------------------------------------------------------------------------------------------------
class Workspace:
    def __init__(self, S_img: int, T: int, device):
        S = S_img + T
        bf = dict(dtype=torch.bfloat16, device=device)
        self.S_img, self.T, self.S = S_img, T, S
        self.x = torch.empty(S, DIM, **bf)          # residual stream, [text; image]
        self.xhat = torch.empty(S, DIM, **bf)       # LN + modulate output
        self.q = torch.empty(S, DIM, **bf)
        self.k = torch.empty(S, DIM, **bf)
        self.v = torch.empty(S, DIM, **bf)
        self.att = torch.empty(S, DIM, **bf)
        self.h = torch.empty(S, 4 * DIM, **bf)      # MLP hidden
        self.tok = torch.empty(S_img, 64, **bf)     # patchified latents
        self.txt_n = torch.empty(T, 3584, **bf)     # txt_norm output
        self.out_tok = torch.empty(S_img, 64, **bf)


class DiTEngine:
    """Holds fused QKV weights (the module's to_q/to_k/to_v parameters are re-pointed at views of the
    fused buffer, so in-place LoRA folds and load_state_dict keep working) and runs the kernel sequence."""

    def __init__(self, dit: QwenImageDiT):
        self.dit = dit
        p = next(dit.parameters())
        if not p.is_cuda or p.dtype != torch.bfloat16:
            raise nv.NativeUnavailable(f"the native DiT runs in bfloat16 on an sm_100 GPU only (got {p.dtype} on {p.device}); "
                                       "there is no CPU / fp32 fallback")
        self.device = p.device
        self.nat = nv.Native.get(self.device.index or 0)
        self.use_cta_pair = True
        self.attn_flags = 0
        self._ws: Dict[Tuple[int, int, int], Workspace] = {}
        self._rope: Dict[tuple, torch.Tensor] = {}
        self._mods_cache: Optional[Tuple[float, torch.Tensor, torch.Tensor, torch.Tensor]] = None
        self.sp = None          # ulysses.UlyssesContext: ONE image on the ranks of a group (sequence-parallel rows, head-parallel attention)
        self._pack()

    # -- weights ---------------------------------------------------------------------------------
    def _pack(self):
        self.qkv_w: List[Tuple[torch.Tensor, torch.Tensor]] = []
        self.qkv_b: List[Tuple[torch.Tensor, torch.Tensor]] = []
        with torch.no_grad():
            for blk in self.dit.transformer_blocks:
                a = blk.attn
                packs = []
                for mods in ((a.to_q, a.to_k, a.to_v), (a.add_q_proj, a.add_k_proj, a.add_v_proj)):
                    w = torch.cat([m.weight.data for m in mods], dim=0).contiguous()
                    b = torch.cat([m.bias.data for m in mods], dim=0).contiguous()
                    for j, m in enumerate(mods):                      # re-point: parameters become views of the fused buffers
                        m.weight.data = w[j * DIM:(j + 1) * DIM]
                        m.bias.data = b[j * DIM:(j + 1) * DIM]
                    packs.append((w, b))
                self.qkv_w.append((packs[0][0], packs[1][0]))
                self.qkv_b.append((packs[0][1], packs[1][1]))
        # which output features of the modulation linears are "scale" (stored as 1 + scale)
        m6 = torch.zeros(6 * DIM, dtype=torch.uint8, device=self.device)
        m6[DIM:2 * DIM] = 1
        m6[4 * DIM:5 * DIM] = 1
        self.mask6 = m6
        m2 = torch.zeros(2 * DIM, dtype=torch.uint8, device=self.device)
        m2[:DIM] = 1                                                   # AdaLayerNorm(single): (scale, shift)
        self.mask2 = m2
        self.invalidate()

    def refresh_if_stale(self):
        blocks = self.dit.transformer_blocks
        for i in (0, len(blocks) - 1):
            a = blocks[i].attn
            if a.to_q.weight.data_ptr() != self.qkv_w[i][0].data_ptr() or a.add_v_proj.weight.data_ptr() != self.qkv_w[i][1][2 * DIM:].data_ptr():
                self._pack()      # parameters were replaced (e.g. load_state_dict(assign=True))
                return

    def invalidate(self):
        """Call after modifying weights in place (LoRA fold): drops cached modulation vectors and the training path's cached weight transposes
        (a write through `.data`, as the LoRA loader does, does not bump the version the transpose cache is keyed on)."""
        self._mods_cache = None
        self._cond_table = None
        import sys
        ag = sys.modules.get("physicedit_b200.autograd")
        if ag is not None:
            ag.weight_transposes.clear()

    # -- cached per-shape state ---------------------------------------------------------------------
    def workspace(self, S_img: int, T: int, branch: int = 0) -> Workspace:
        """Activation buffers for one forward.  `branch` separates the two CFG branches when they run concurrently on two
        streams (pipeline.denoise_step, cfg_streams=2): equal shapes must not share buffers then."""
        key = (S_img, T, branch)
        if key not in self._ws:
            if len(self._ws) >= 6:
                self._ws.pop(next(iter(self._ws)))
            self._ws[key] = Workspace(S_img, T, self.device)
        return self._ws[key]

    def rope(self, img_shapes: Sequence[Tuple[int, int, int]], T: int, sampling: bool = False) -> torch.Tensor:
        """float2 (cos, sin) table [T + S_img, 64] in joint [text; image] order; `sampling` = edit_rope_interpolation (forward_sampling)."""
        key = (tuple(tuple(s) for s in img_shapes), T, bool(sampling))
        if key not in self._rope:
            pe = self.dit.pos_embed
            vid, txt = (pe.forward_sampling if sampling else pe.forward)(list(img_shapes), [T], device=None)
            joint = torch.cat([txt, vid], dim=0).to(torch.complex64)
            self._rope[key] = torch.view_as_real(joint).contiguous().to(self.device)
        return self._rope[key]

    # -- pieces ---------------------------------------------------------------------------------------
    def block_mods(self, temb: torch.Tensor, indices: Optional[Sequence[int]] = None, temb_act: Optional[torch.Tensor] = None) -> torch.Tensor:
        """img_mod / txt_mod of the given blocks for a batch of B <= 8 conditioning vectors temb [B, 3072]:
        returns [B, n, 2 (img, txt), 18432] with bf16(1+scale) in the scale slots.  The GEMV streams each weight matrix once
        for the whole batch; SiLU(temb) (the nn.SiLU in front of every modulation linear) is materialised once."""
        blocks = self.dit.transformer_blocks
        indices = list(range(len(blocks))) if indices is None else list(indices)
        B = temb.shape[0]
        if temb_act is None:
            temb_act = torch.empty_like(temb)
            self.nat.tag = "gemv_mod"
            self.nat.act(temb.contiguous(), temb_act, 1)
        out = torch.empty(len(indices), 2, B, 6 * DIM, dtype=torch.bfloat16, device=self.device)
        for n, i in enumerate(indices):
            b = blocks[i]
            self.nat.tag = "gemv_mod"
            self.nat.gemv(temb_act, b.img_mod[1].weight, b.img_mod[1].bias, out[n, 0], 0, 0, self.mask6)
            self.nat.tag = "gemv_mod"
            self.nat.gemv(temb_act, b.txt_mod[1].weight, b.txt_mod[1].bias, out[n, 1], 0, 0, self.mask6)
        return out.permute(2, 0, 1, 3)          # [B, n, 2, 18432] view; [b] is contiguous per (n, stream) row

    def _conditioning_batch(self, timesteps_bf16: torch.Tensor):
        """temb, block modulation table and norm_out (scale, shift) for B <= 8 timesteps at once."""
        B = timesteps_bf16.shape[0]
        temb = torch.cat([self.dit.time_text_embed(timesteps_bf16[b:b + 1], raw=True) for b in range(B)], dim=0)
        temb_act = torch.empty_like(temb)
        self.nat.tag = "gemv_mod"
        self.nat.act(temb, temb_act, 1)
        mods = self.block_mods(temb, temb_act=temb_act)
        no = self.dit.norm_out.linear
        out_mod = torch.empty(B, 2 * DIM, dtype=torch.bfloat16, device=self.device)
        self.nat.tag = "gemv_mod"
        self.nat.gemv(temb_act, no.weight, no.bias, out_mod, 0, 0, self.mask2)
        return temb, mods, out_mod

    def precompute_conditioning(self, timesteps_bf16: torch.Tensor, keys: Sequence[float]) -> None:
        """Everything that depends on the timestep only (temb, the 120 modulation vectors per forward, norm_out's scale/shift)
        for a whole schedule: the M=1 linears are run as batch-8 GEMVs, so the 13.6 GB of modulation weights stream once per 8
        denoise steps instead of once per forward.  Values are identical to the per-forward computation of the reference
        (qwen_image_dit.py:373-374 recomputes them in every block of every forward from the same temb)."""
        table = {}
        for c0 in range(0, len(keys), 8):
            ts = timesteps_bf16[c0:c0 + 8].contiguous()
            temb, mods, out_mod = self._conditioning_batch(ts)
            for b in range(ts.shape[0]):
                table[float(keys[c0 + b])] = (temb[b:b + 1], mods[b], out_mod[b:b + 1])
        self._cond_table = table

    def conditioning(self, timestep_bf16: torch.Tensor, t_key: Optional[float]):
        """temb, all block modulation vectors and the norm_out (scale, shift) for this timestep.  They depend on
        the timestep only, so the two CFG branches of a denoise step share them (cache keyed by the host value)."""
        tab = getattr(self, "_cond_table", None)
        if t_key is not None and tab is not None and t_key in tab:
            return tab[t_key]
        if t_key is not None and self._mods_cache is not None and self._mods_cache[0] == t_key:
            return self._mods_cache[1:]
        temb, mods, out_mod = self._conditioning_batch(timestep_bf16.reshape(-1)[:1])
        res = (temb, mods[0], out_mod)
        if t_key is not None:
            self._mods_cache = (t_key,) + res
        return res

    def run_block(self, i: int, x: torch.Tensor, T: int, mods: torch.Tensor, rope: torch.Tensor, ws: Workspace, attn_mask: Optional[torch.Tensor] = None,
                  fp8_attention: bool = False):
        """One double-stream block in place on the joint residual stream x [T + S_img, 3072].
        mods: [2, 18432] = (img, txt) x (shift_a, 1+scale_a, gate_a, shift_m, 1+scale_m, gate_m).
        attn_mask: uint8 [S, S], 0 = hidden (EliGen) -> the masked attention path; fp8_attention: q / k / v quantised to e4m3."""
        nat, blk = self.nat, self.dit.transformer_blocks[i]
        a = blk.attn
        flags = nv.GEMM_FLAG_CTA_PAIR if self.use_cta_pair else 0
        mi, mt = mods[0], mods[1]
        D = DIM
        xi, xt = x[T:], x[:T]
        hi, ht = ws.xhat[T:], ws.xhat[:T]
        # LN + modulate (attention branch), both streams in one launch
        nat.tag = "ln_mod"
        nat.layernorm_modulate2(x, ws.xhat, T, mt[0:D], mt[D:2 * D], mi[0:D], mi[D:2 * D])
        # fused QKV projection + per-head RMSNorm + RoPE, written straight into the joint q/k/v buffers
        (wq_i, wq_t), (bq_i, bq_t) = self.qkv_w[i], self.qkv_b[i]
        nat.tag = "gemm_qkv"
        nat.gemm([dict(a=hi, w=wq_i, bias=bq_i, out=ws.q[T:], out_k=ws.k[T:], out_v=ws.v[T:], norm_q_w=a.norm_q.weight,
                       norm_k_w=a.norm_k.weight, rope=rope[T:]),
                  dict(a=ht, w=wq_t, bias=bq_t, out=ws.q[:T], out_k=ws.k[:T], out_v=ws.v[:T], norm_q_w=a.norm_added_q.weight,
                       norm_k_w=a.norm_added_k.weight, rope=rope[:T])], 3 * D, D, nv.EPI_QKV_NORM_ROPE, flags)
        nat.tag = "attention"
        if attn_mask is not None:
            self.masked_attention(ws.q, ws.k, ws.v, ws.att, attn_mask)
        elif fp8_attention:
            self.fp8_attention(ws.q, ws.k, ws.v, ws.att)
        else:
            nat.attention(ws.q, ws.k, ws.v, ws.att, NUM_HEADS, 1.0 / math.sqrt(HEAD_DIM), self.attn_flags)
        # output projections + gate * o + residual (in place on x)
        nat.tag = "gemm_out"
        nat.gemm([dict(a=ws.att[T:], w=a.to_out[0].weight, bias=a.to_out[0].bias, out=xi, gate=mi[2 * D:3 * D]),
                  dict(a=ws.att[:T], w=a.to_add_out.weight, bias=a.to_add_out.bias, out=xt, gate=mt[2 * D:3 * D])],
                 D, D, nv.EPI_GATE_RESIDUAL, flags)
        # MLP branch
        nat.tag = "ln_mod"
        nat.layernorm_modulate2(x, ws.xhat, T, mt[3 * D:4 * D], mt[4 * D:5 * D], mi[3 * D:4 * D], mi[4 * D:5 * D])
        im, tm = blk.img_mlp.net, blk.txt_mlp.net
        nat.tag = "gemm_up"
        nat.gemm([dict(a=hi, w=im[0].proj.weight, bias=im[0].proj.bias, out=ws.h[T:]),
                  dict(a=ht, w=tm[0].proj.weight, bias=tm[0].proj.bias, out=ws.h[:T])], 4 * D, D, nv.EPI_BIAS_GELU_SIGMOID, flags)
        nat.tag = "gemm_down"
        nat.gemm([dict(a=ws.h[T:], w=im[2].weight, bias=im[2].bias, out=xi, gate=mi[5 * D:6 * D]),
                  dict(a=ws.h[:T], w=tm[2].weight, bias=tm[2].bias, out=xt, gate=mt[5 * D:6 * D])], D, 4 * D, nv.EPI_GATE_RESIDUAL, flags)

    MASKED_ATTN_SCRATCH_BYTES = 12 << 30

    def masked_attention(self, q, k, v, o, mask_u8) -> None:
        """Joint attention under an arbitrary key mask (EliGen: `F.scaled_dot_product_attention(q, k, v, attn_mask=...)`, qwen_image_dit.py:36
        with the 0 / -inf mask of process_entity_masks :433-498).  The flash kernel has no mask input, so this option takes the materialised
        route, batched over heads: scores = Q K^T (tcgen05 GEMM, fp32 epilogue) -> masked row softmax -> O = P V (GEMM).  HBM-bound
        (12 bytes per score element) -- EliGen is not on the PhysicEdit path; correctness first."""
        nat = self.nat
        S, H, D = q.shape[0], NUM_HEADS, HEAD_DIM
        Sp = (S + 7) // 8 * 8
        dev = q.device

        def head_major(t):
            out = torch.zeros(H, Sp, D, dtype=torch.bfloat16, device=dev) if Sp != S else torch.empty(H, S, D, dtype=torch.bfloat16, device=dev)
            out[:, :S].copy_(t.view(S, H, D).transpose(0, 1))
            return out
        qh, kh = head_major(q), head_major(k)
        vt = head_major(v).transpose(1, 2).contiguous()                       # [H, 128, Sp]
        oh = torch.empty(H, Sp, D, dtype=torch.bfloat16, device=dev)
        hc = max(1, min(H, self.MASKED_ATTN_SCRATCH_BYTES // (6 * Sp * Sp)))
        scores = torch.empty(hc * Sp, Sp, dtype=torch.float32, device=dev)
        probs = torch.empty(hc * Sp, Sp, dtype=torch.bfloat16, device=dev)
        if mask_u8.shape[0] != Sp:                                            # pad rows (never stored) so that row r -> mask row r % Sp
            m = torch.ones(Sp, mask_u8.shape[1], dtype=torch.uint8, device=dev)
            m[:S] = mask_u8
            mask_u8 = m
        flat = lambda t, h0, n: t[h0:h0 + n].reshape(n * t.shape[1], t.shape[2])
        for h0 in range(0, H, hc):
            n = min(hc, H - h0)
            kw = dict(batch=n, M=S, a_batch_rows=Sp, out_batch_rows=Sp)
            nat.gemm_batched(flat(qh, h0, n), flat(kh, h0, n), scores, N=Sp, K=D, w_batch_rows=Sp, epilogue=nv.EPI_F32, **kw)
            nat.softmax_rows(scores[:n * Sp], probs[:n * Sp], S, 1.0 / math.sqrt(D), mask=mask_u8)
            nat.gemm_batched(probs, flat(vt, h0, n), flat(oh, h0, n), N=D, K=Sp, w_batch_rows=D, **kw)
        o.view(S, H, D).copy_(oh[:, :S].transpose(0, 1))

    def fp8_attention(self, q, k, v, o) -> None:
        """`enable_fp8_attention` (qwen_image_dit.py:24-35; only taken when FlashAttention-3 is installed, otherwise the reference silently runs
        the bf16 SDPA): q / k / v divided by their standard deviation and rounded to float8_e4m3fn, softmax scale q_std k_std / sqrt(d), output
        times v_std.  Here the e4m3 VALUES are fed to the bf16 tensor-core kernel (every e4m3 number is a bf16 number), so Q K^T and the softmax
        see exactly the quantised operands; P stays bf16 where FA3 rounds it to e4m3 as well.  The quantisation glue is torch ops (a non-default
        option); FA3 is absent here, so this branch is "parity unpinned"."""
        qs, ks, vs = q.std(), k.std(), v.std()
        f8 = lambda t, s_: (t / s_).to(torch.float8_e4m3fn).to(torch.bfloat16)
        scale = float(qs * ks) / math.sqrt(HEAD_DIM)
        self.nat.attention(f8(q, qs), f8(k, ks), f8(v, vs), o, NUM_HEADS, scale, self.attn_flags)
        o.mul_(vs)

    def rope_segments(self, img_shapes: Sequence[Tuple[int, int, int]], seg_lens: Sequence[int]) -> torch.Tensor:
        """(cos, sin) table for a text stream made of several prompts (EliGen, :441-446): every prompt's positions start at the same index."""
        key = (tuple(tuple(s) for s in img_shapes), tuple(int(x) for x in seg_lens), "segments")
        if key not in self._rope:
            pe = self.dit.pos_embed
            vid, _ = pe.forward(list(img_shapes), [max(seg_lens)], device=None)
            txt = [pe.forward(list(img_shapes), [int(n)], device=None)[1] for n in seg_lens]
            joint = torch.cat(txt + [vid], dim=0).to(torch.complex64)
            self._rope[key] = torch.view_as_real(joint).contiguous().to(self.device)
        return self._rope[key]

    def forward(self, latents_list: Sequence[torch.Tensor], timestep_bf16: torch.Tensor, prompt_emb: torch.Tensor,
                out_latents: torch.Tensor, t_key: Optional[float] = None, branch: int = 0, rope_sampling: bool = False,
                after_block=None, text_segments: Optional[Sequence[int]] = None, attn_mask: Optional[torch.Tensor] = None,
                fp8_attention: bool = False) -> torch.Tensor:
        """latents_list: [noise latents, edit/context latents ...] each [1,16,h8,w8] bf16; prompt_emb [T,3584] bf16
        (already updated by the adapter).  Writes the velocity for the first entry into out_latents [1,16,h8,w8].
        `after_block(i, noise_tokens [n0, 3072])` may update the noise-image tokens in place after block i (blockwise controlnet, :1389-1396)."""
        if self.sp is not None:
            if after_block is not None or attn_mask is not None or fp8_attention:
                raise NotImplementedError("blockwise controlnet / EliGen masks / fp8 attention are not wired into the sequence-parallel mode")
            from .ulysses import forward_sp
            return forward_sp(self, self.sp, latents_list, timestep_bf16, prompt_emb, out_latents, t_key, rope_sampling=rope_sampling)
        nat, dit = self.nat, self.dit
        T = prompt_emb.shape[0]
        shapes = [(1, l.shape[-2] // 2, l.shape[-1] // 2) for l in latents_list]
        S_img = sum(h * w for _, h, w in shapes)
        ws = self.workspace(S_img, T, branch)
        # EliGen: prompt_emb holds [entity prompts ..., global prompt] back to back (text_segments = their lengths), each with its own positions
        rope = self.rope(shapes, T, rope_sampling) if text_segments is None else self.rope_segments(shapes, text_segments)
        temb, mods, out_mod = self.conditioning(timestep_bf16, t_key)
        # patchify + img_in
        off = 0
        for l, (_, h, w) in zip(latents_list, shapes):
            nat.patchify(l.reshape(16, l.shape[-2], l.shape[-1]), ws.tok[off:off + h * w])
            off += h * w
        nat.gemm([dict(a=ws.tok, w=dit.img_in.weight, bias=dit.img_in.bias, out=ws.x[T:])], DIM, 64, nv.EPI_BIAS)
        # txt_norm + txt_in
        nat.rmsnorm(prompt_emb, ws.txt_n, dit.txt_norm.weight, dit.txt_norm.eps)
        nat.gemm([dict(a=ws.txt_n, w=dit.txt_in.weight, bias=dit.txt_in.bias, out=ws.x[:T])], DIM, 3584, nv.EPI_BIAS)
        n0 = shapes[0][1] * shapes[0][2]
        for i in range(len(dit.transformer_blocks)):
            self.run_block(i, ws.x, T, mods[i], rope, ws, attn_mask, fp8_attention)
            if after_block is not None:
                after_block(i, ws.x[T:T + n0])
        # norm_out + proj_out on the noise tokens only (the reference slices after computing all image rows)
        nat.layernorm_modulate(ws.x[T:T + n0], ws.xhat[:n0], out_mod[0, DIM:], out_mod[0, :DIM])
        nat.gemm([dict(a=ws.xhat[:n0], w=dit.proj_out.weight, bias=dit.proj_out.bias, out=ws.out_tok[:n0])], 64, DIM, nv.EPI_BIAS)
        nat.unpatchify(ws.out_tok[:n0], out_latents.reshape(16, out_latents.shape[-2], out_latents.shape[-1]))
        return out_latents
