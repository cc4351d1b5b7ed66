"""Qwen2.5-VL text encoder of Qwen-Image-Edit / PhysicEdit on the native path.

The pipeline uses the VL model twice per CFG branch before the denoise loop (pipelines/qwen_image_physical.py):
  * `generate(**processor_outputs, max_new_tokens=1000)`  -- the "physical thinking" text (:859-873, 943-967), greedy,
  * `edit_forward(input_ids, attention_mask, pixel_values, image_grid_thw)[-1]` -- the final hidden states that become `prompt_emb`
    (:774-800; wrapper models/qwen_image_text_encoder_withdecode.py:188-275).
The arithmetic is transformers' `modeling_qwen2_5_vl.py` (an un-pinned third-party dependency of the reference: requirements.txt:3;
the wrapper's config says 4.54.0, this image has 5.5.0): vision tower (patch embed, 2-D rotary, windowed + full attention blocks,
SwiGLU MLP, 2x2 patch merger), 28-layer GQA decoder with multimodal rotary embedding, lm_head.  Here every linear runs on
`pe_gemm` (prefill) / `pe_gemv` (one-token decode), the rest on the kernels of csrc/llm_kernels.cu; one decode step has no host-side
state (positions and the KV length live in device counters), so it is captured ONCE per generate call in a CUDA graph and replayed
per token -- B=1 decode is HBM-bound (15.2 GB of weights per token), not launch-bound.

Module / parameter names are the wrapper's (`model.language_model.*`, `model.visual.*`, `lm_head.weight`; checkpoints in the
"diffusers" layout go through `state_dict_converter().from_diffusers`, :283-297), so the reference's files load.

Position ids -- a version-dependent corner of the dependency, made explicit (`rope_mode`):
  "mrope"        3-D positions derived from input_ids + image_grid_thw whenever an image is present: what transformers 4.54 (the version
                 the reference was written against) does in BOTH entry points.  Default.
  "mrope_hf55"   the same with transformers 5.5's vision temporal index (= 2 x start position, its tokens_per_second handling), which is
                 what `generate` does in this image because the 5.x processor hands over `mm_token_type_ids`.
  "sequential"   plain 1-D positions: what `edit_forward` silently degrades to under transformers 5.5 (no `mm_token_type_ids` passed).
bf16 on an sm_100 GPU only; batch 1 (the pipeline is strictly B=1, :688,821).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import native as nv

TEXT_ENCODER_KEY_HASH = "8004730443f55db63092006dd9f7110e"      # configs/model_config.py:23 (source-layout keys + shapes)


@dataclass
class VLConfig:
    """Sizes of Qwen2.5-VL-7B as hard-coded by the wrapper (qwen_image_text_encoder_withdecode.py:9-143)."""
    hidden: int = 3584
    layers: int = 28
    heads: int = 28
    kv_heads: int = 4
    head_dim: int = 128
    intermediate: int = 18944
    vocab: int = 152064
    rms_eps: float = 1e-6
    rope_theta: float = 1000000.0
    mrope_section: Tuple[int, int, int] = (16, 24, 24)
    v_hidden: int = 1280
    v_depth: int = 32
    v_heads: int = 16
    v_intermediate: int = 3420
    v_out: int = 3584
    patch: int = 14
    temporal_patch: int = 2
    merge: int = 2
    window: int = 112
    fullatt: Tuple[int, ...] = (7, 15, 23, 31)
    in_ch: int = 3
    image_token_id: int = 151655
    eos_token_id: int = 151645
    tokens_per_second: int = 2


def _round_up(x, m):
    return (x + m - 1) // m * m


# ---------------------------------------------------------------------------------------------------------------------------
# parameter containers (names as transformers / the wrapper)
# ---------------------------------------------------------------------------------------------------------------------------
class _Norm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class _TextAttn(nn.Module):
    def __init__(self, c: VLConfig):
        super().__init__()
        self.q_proj = nn.Linear(c.hidden, c.heads * c.head_dim, bias=True)
        self.k_proj = nn.Linear(c.hidden, c.kv_heads * c.head_dim, bias=True)
        self.v_proj = nn.Linear(c.hidden, c.kv_heads * c.head_dim, bias=True)
        self.o_proj = nn.Linear(c.heads * c.head_dim, c.hidden, bias=False)


class _MLP(nn.Module):
    def __init__(self, dim, inter, bias):
        super().__init__()
        self.gate_proj = nn.Linear(dim, inter, bias=bias)
        self.up_proj = nn.Linear(dim, inter, bias=bias)
        self.down_proj = nn.Linear(inter, dim, bias=bias)


class _TextLayer(nn.Module):
    def __init__(self, c: VLConfig):
        super().__init__()
        self.self_attn = _TextAttn(c)
        self.mlp = _MLP(c.hidden, c.intermediate, False)
        self.input_layernorm = _Norm(c.hidden)
        self.post_attention_layernorm = _Norm(c.hidden)


class _TextModel(nn.Module):
    def __init__(self, c: VLConfig):
        super().__init__()
        self.embed_tokens = nn.Embedding(c.vocab, c.hidden)
        self.layers = nn.ModuleList([_TextLayer(c) for _ in range(c.layers)])
        self.norm = _Norm(c.hidden)


class _PatchEmbed(nn.Module):
    def __init__(self, c: VLConfig):
        super().__init__()
        self.proj = nn.Conv3d(c.in_ch, c.v_hidden, kernel_size=(c.temporal_patch, c.patch, c.patch), stride=(c.temporal_patch, c.patch, c.patch), bias=False)


class _VisionAttn(nn.Module):
    def __init__(self, c: VLConfig):
        super().__init__()
        self.qkv = nn.Linear(c.v_hidden, 3 * c.v_hidden, bias=True)
        self.proj = nn.Linear(c.v_hidden, c.v_hidden)


class _VisionBlock(nn.Module):
    def __init__(self, c: VLConfig):
        super().__init__()
        self.norm1 = _Norm(c.v_hidden)
        self.norm2 = _Norm(c.v_hidden)
        self.attn = _VisionAttn(c)
        self.mlp = _MLP(c.v_hidden, c.v_intermediate, True)


class _Merger(nn.Module):
    def __init__(self, c: VLConfig):
        super().__init__()
        d = c.v_hidden * c.merge * c.merge
        self.ln_q = _Norm(c.v_hidden)
        self.mlp = nn.Sequential(nn.Linear(d, d), nn.GELU(), nn.Linear(d, c.v_out))


class _Vision(nn.Module):
    def __init__(self, c: VLConfig):
        super().__init__()
        self.patch_embed = _PatchEmbed(c)
        self.blocks = nn.ModuleList([_VisionBlock(c) for _ in range(c.v_depth)])
        self.merger = _Merger(c)


class _VLModel(nn.Module):
    def __init__(self, c: VLConfig):
        super().__init__()
        self.visual = _Vision(c)
        self.language_model = _TextModel(c)


class QwenImageTextEncoderStateDictConverter:
    """qwen_image_text_encoder_withdecode.py:283-297: `visual.*` -> `model.visual.*`, `model.*` -> `model.language_model.*`."""

    def from_diffusers(self, state_dict):
        out = {}
        for k, v in state_dict.items():
            if k.startswith("visual."):
                k = "model." + k
            elif k.startswith("model.") and not k.startswith("model.language_model.") and not k.startswith("model.visual."):
                k = k.replace("model.", "model.language_model.", 1)
            out[k] = v
        return out

    def from_civitai(self, state_dict):
        return state_dict


# ---------------------------------------------------------------------------------------------------------------------------
# host-side bookkeeping (position ids, rotary tables, window index): restated from modeling_qwen2_5_vl.py
# ---------------------------------------------------------------------------------------------------------------------------
def vision_rot_table(cfg: VLConfig, grid_thw: Sequence[Sequence[int]]) -> torch.Tensor:
    """rot_pos_emb (:382-409): per patch (in the processor's merge-block order) the angles [h * inv_freq | w * inv_freq], fp32 [N, head_dim/2]."""
    m = cfg.merge
    dim = (cfg.v_hidden // cfg.v_heads) // 2
    inv = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))
    out = []
    for t, h, w in grid_thw:
        hp = torch.arange(h).unsqueeze(1).expand(-1, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        wp = torch.arange(w).unsqueeze(0).expand(h, -1).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        out.append(torch.stack([hp, wp], dim=-1).repeat(t, 1))
    pos = torch.cat(out, dim=0)
    freqs = torch.outer(torch.arange(int(max(max(g[1], g[2]) for g in grid_thw)), dtype=torch.float), inv)
    return freqs[pos].flatten(1)


def vision_window_index(cfg: VLConfig, grid_thw: Sequence[Sequence[int]]):
    """get_window_index (:411-452): permutation of the merged tokens into window-major order + cumulative window lengths (in patches)."""
    import torch.nn.functional as F
    ws = cfg.window // cfg.merge // cfg.patch
    unit = cfg.merge * cfg.merge
    index, cu, base = [], [0], 0
    for t, h, w in grid_thw:
        gh, gw = h // cfg.merge, w // cfg.merge
        idx = torch.arange(t * gh * gw).reshape(t, gh, gw)
        ph, pw = ws - gh % ws, ws - gw % ws
        nh, nw = (gh + ph) // ws, (gw + pw) // ws
        pad = F.pad(idx, (0, pw, 0, ph), "constant", -100).reshape(t, nh, ws, nw, ws).permute(0, 1, 3, 2, 4).reshape(t, nh * nw, ws, ws)
        lens = (pad != -100).sum([2, 3]).reshape(-1)
        flat = pad.reshape(-1)
        index.append(flat[flat != -100] + base)
        cu.extend((lens.cumsum(0) * unit + cu[-1]).tolist())
        base += t * gh * gw
    cu_t = torch.unique_consecutive(torch.tensor(cu, dtype=torch.int32))
    return torch.cat(index, dim=0), cu_t


def mrope_position_ids(cfg: VLConfig, input_ids: torch.Tensor, grid_thw: Optional[Sequence[Sequence[int]]], mode: str) -> Tuple[torch.Tensor, int]:
    """get_rope_index for one unpadded sequence (:1024-1135): text runs count 1, 2, 3, ... on all three axes; an image run of
    (t, h, w) patches gets (temporal, row, column) indices offset by the running position, which then advances by max(h, w) / merge.
    Returns (positions int64 [3, T], delta) with delta = max position + 1 - T (the offset decode steps continue from)."""
    ids = input_ids.reshape(-1).tolist()
    T = len(ids)
    if mode == "sequential" or not grid_thw or cfg.image_token_id not in ids:
        return torch.arange(T).view(1, -1).expand(3, -1).contiguous(), 0
    grids = iter(grid_thw)
    chunks, cur, i = [], 0, 0
    while i < T:
        j = i
        is_img = ids[i] == cfg.image_token_id
        while j < T and (ids[j] == cfg.image_token_id) == is_img:
            j += 1
        if not is_img:
            chunks.append(torch.arange(j - i).view(1, -1).expand(3, -1) + cur)
            cur += j - i
            i = j
            continue
        k = i
        while k < j:                                   # consecutive images without text in between: one grid each
            t, h, w = next(grids)
            gh, gw = h // cfg.merge, w // cfg.merge
            n = t * gh * gw
            temporal = torch.full((n,), cur * (cfg.tokens_per_second if mode == "mrope_hf55" else 1), dtype=torch.long)
            rows = torch.arange(cur, cur + gh).repeat_interleave(gw * t)
            cols = torch.arange(cur, cur + gw).repeat(gh * t)
            chunks.append(torch.stack([temporal, rows, cols], dim=0))
            cur += max(h, w) // cfg.merge
            k += n
        i = j
    pos = torch.cat(chunks, dim=1)
    return pos.contiguous(), int(pos.max()) + 1 - T


def text_rope_tables(cfg: VLConfig, positions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Qwen2_5_VLRotaryEmbedding.forward (:595-608) + the section interleave of apply_multimodal_rotary_pos_emb (:657-662):
    fp32 cos / sin [T, head_dim]; channel chunk i of (16,24,24,16,24,24) takes axis i % 3.  (The bf16 rounding of the reference's
    `cos.to(x.dtype)` happens inside pe_rope_half, mode 1.)"""
    D = cfg.head_dim
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, D, 2, dtype=torch.int64).to(torch.float) / D))
    freqs = positions.to(torch.float)[:, :, None] * inv[None, None, :]          # [3, T, D/2]
    emb = torch.cat([freqs, freqs], dim=-1)                                      # [3, T, D]
    sel = torch.cat([torch.full((s,), i % 3) for i, s in enumerate(list(cfg.mrope_section) * 2)])
    emb = emb.gather(0, sel.view(1, 1, D).expand(1, emb.shape[1], D))[0]          # [T, D]
    return emb.cos().contiguous(), emb.sin().contiguous()


# ---------------------------------------------------------------------------------------------------------------------------
class QwenImageTextEncoder(nn.Module):
    """Native stand-in for QwenImageTextEncoderWithDecode: same state dict, `edit_forward` and `generate`."""

    def __init__(self, config: Optional[VLConfig] = None, rope_mode: str = "mrope"):
        super().__init__()
        self.cfg = config or VLConfig()
        self.rope_mode = rope_mode
        self.model = _VLModel(self.cfg)
        self.lm_head = nn.Linear(self.cfg.hidden, self.cfg.vocab, bias=False)
        self._packed = None
        self.use_cuda_graph = True
        self.fused_decode = True          # decode step with the norms / SwiGLU / skip adds fused into the GEMVs (False: one kernel per op)
        self.last_generate_stats: dict = {}

    @staticmethod
    def state_dict_converter():
        return QwenImageTextEncoderStateDictConverter()

    def load_state_dict(self, *args, **kwargs):
        self._packed = None
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    # ---- packing ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _pack(self):
        c = self.cfg
        p = self.lm_head.weight
        if p.dtype != torch.bfloat16:
            raise nv.NativeUnavailable(f"the native text encoder runs in bfloat16 on an sm_100 GPU only (got {p.dtype} on {p.device}); no fallback")
        dev = p.device            # the device itself is checked where the handle is taken (_ctx: Native.get raises without libpe_b200.so / an sm_100 GPU)
        P = {"ones_h": torch.ones(c.hidden, dtype=torch.bfloat16, device=dev), "ones_v": torch.ones(c.v_hidden, dtype=torch.bfloat16, device=dev)}
        P["text"] = []
        for l in self.model.language_model.layers:
            a, m = l.self_attn, l.mlp
            P["text"].append(dict(
                wqkv=torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0).contiguous(),
                bqkv=torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0).contiguous(),
                wgu=torch.cat([m.gate_proj.weight, m.up_proj.weight], 0).contiguous()))
        Ip = _round_up(c.v_intermediate, 8)            # 3420 -> 3424: 16-byte rows for TMA (zero weights / biases in the pad columns)
        P["v_inter_pad"] = Ip
        P["vision"] = []
        for b in self.model.visual.blocks:
            m = b.mlp
            wgu = torch.zeros(2 * Ip, c.v_hidden, dtype=torch.bfloat16, device=dev)
            bgu = torch.zeros(2 * Ip, dtype=torch.bfloat16, device=dev)
            wgu[:c.v_intermediate], wgu[Ip:Ip + c.v_intermediate] = m.gate_proj.weight, m.up_proj.weight
            bgu[:c.v_intermediate], bgu[Ip:Ip + c.v_intermediate] = m.gate_proj.bias, m.up_proj.bias
            wd = torch.zeros(c.v_hidden, Ip, dtype=torch.bfloat16, device=dev)
            wd[:, :c.v_intermediate] = m.down_proj.weight
            P["vision"].append(dict(wgu=wgu, bgu=bgu, wd=wd))
        P["w_patch"] = self.model.visual.patch_embed.proj.weight.reshape(c.v_hidden, -1).contiguous()
        self._packed = P
        return P

    def _ctx(self):
        nat = nv.Native.get(self.lm_head.weight.device.index or 0)         # NativeUnavailable when the library or the GPU is missing: no fallback
        return nat, (self._packed or self._pack())

    # ---- vision tower (Qwen2_5_VisionTransformerPretrainedModel.forward :455-523) ---------------------------------------------
    @torch.no_grad()
    def vision(self, pixel_values: torch.Tensor, grid_thw) -> torch.Tensor:
        """pixel_values [N, 3*2*14*14] (the processor's flattened patches) -> merged image embeddings [N / 4, 3584]."""
        c = self.cfg
        nat, P = self._ctx()
        dev = self.lm_head.weight.device
        grid = [tuple(int(v) for v in g) for g in (grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw)]
        bf = dict(dtype=torch.bfloat16, device=dev)
        x_in = pixel_values.to(**bf).contiguous()
        N, unit, D = x_in.shape[0], c.merge * c.merge, c.v_hidden // c.v_heads
        x = torch.empty(N, c.v_hidden, **bf)
        nat.gemm([dict(a=x_in, w=P["w_patch"], bias=None, out=x)], c.v_hidden, x_in.shape[1], nv.EPI_BIAS)
        win, cu_win = vision_window_index(c, grid)
        rot = vision_rot_table(c, grid)
        order = (win.view(-1, 1) * unit + torch.arange(unit).view(1, -1)).reshape(-1)            # patch order = window-major merged tokens x 4
        x = x[order.to(dev)].contiguous()
        rot = rot[order]
        emb = torch.cat([rot, rot], dim=-1)
        cos, sin = emb.cos().contiguous().to(dev), emb.sin().contiguous().to(dev)
        # KV ranges per patch: its window (windowed blocks) or its whole image / frame (full-attention blocks)
        def ranges(cu):
            cu = [int(v) for v in cu]
            lo = torch.empty(N, dtype=torch.int32)
            hi = torch.empty(N, dtype=torch.int32)
            for a, b in zip(cu[:-1], cu[1:]):
                lo[a:b], hi[a:b] = a, b
            return lo.to(dev), hi.to(dev)
        cu_full = [0]
        for t, h, w in grid:
            for _ in range(t):
                cu_full.append(cu_full[-1] + h * w)
        r_win, r_full = ranges(cu_win.tolist()), ranges(cu_full)
        xn = torch.empty_like(x)
        qkv = torch.empty(N, 3 * c.v_hidden, **bf)
        att = torch.empty(N, c.v_hidden, **bf)
        Ip = P["v_inter_pad"]
        gu = torch.empty(N, 2 * Ip, **bf)
        hm = torch.empty(N, Ip, **bf)
        H, C = c.v_heads, c.v_hidden
        for i, blk in enumerate(self.model.visual.blocks):
            pk = P["vision"][i]
            nat.rmsnorm(x, xn, blk.norm1.weight, 1e-6)
            nat.gemm([dict(a=xn, w=blk.attn.qkv.weight, bias=blk.attn.qkv.bias, out=qkv)], 3 * C, C, nv.EPI_BIAS)
            nat.rope_half(qkv[:, :C], H, D, cos, sin, mode=0)
            nat.rope_half(qkv[:, C:2 * C], H, D, cos, sin, mode=0)
            lo, hi = r_full if i in c.fullatt else r_win
            nat.range_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], att, H, H, D, D ** -0.5, kv_lo=lo, kv_hi=hi)
            nat.gemm([dict(a=att, w=blk.attn.proj.weight, bias=blk.attn.proj.bias, out=x, gate=P["ones_v"])], C, C, nv.EPI_GATE_RESIDUAL)
            nat.rmsnorm(x, xn, blk.norm2.weight, 1e-6)
            nat.gemm([dict(a=xn, w=pk["wgu"], bias=pk["bgu"], out=gu)], 2 * Ip, C, nv.EPI_BIAS)
            nat.swiglu(gu, hm, Ip)
            nat.gemm([dict(a=hm, w=pk["wd"], bias=blk.mlp.down_proj.bias, out=x, gate=P["ones_v"])], C, Ip, nv.EPI_GATE_RESIDUAL)
        mg = self.model.visual.merger
        nat.rmsnorm(x, xn, mg.ln_q.weight, 1e-6)
        y = xn.view(N // unit, C * unit)
        h1 = torch.empty(N // unit, C * unit, **bf)
        nat.gemm([dict(a=y, w=mg.mlp[0].weight, bias=mg.mlp[0].bias, out=h1)], C * unit, C * unit, nv.EPI_BIAS_GELU_ERF)
        out = torch.empty(N // unit, c.v_out, **bf)
        nat.gemm([dict(a=h1, w=mg.mlp[2].weight, bias=mg.mlp[2].bias, out=out)], c.v_out, C * unit, nv.EPI_BIAS)
        return out[torch.argsort(win).to(dev)].contiguous()                                      # back to raster order (:518-519)

    # ---- language model -------------------------------------------------------------------------------------------------------
    def _embed(self, nat, input_ids: torch.Tensor, pixel_values, image_grid_thw) -> torch.Tensor:
        """embed_tokens + masked_scatter of the image embeddings at the <|image_pad|> positions (Qwen2_5_VLModel.forward :1305-1316)."""
        c = self.cfg
        ids = input_ids.reshape(-1).to(self.lm_head.weight.device, torch.int64).contiguous()
        x = torch.empty(ids.numel(), c.hidden, dtype=torch.bfloat16, device=ids.device)
        nat.gather_rows(self.model.language_model.embed_tokens.weight, ids, x)
        if pixel_values is not None:
            img = self.vision(pixel_values, image_grid_thw)
            is_img = ids == c.image_token_id
            n_img = int(is_img.sum().item())
            if n_img != img.shape[0]:
                raise ValueError(f"Image features and image tokens do not match, tokens: {n_img}, features: {img.shape[0]}")
            src = torch.where(is_img, torch.cumsum(is_img.to(torch.int64), 0) - 1, torch.full_like(ids, -1))
            nat.gather_rows(img, src.contiguous(), x)
        return x

    def _prefill(self, nat, P, x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, cache=None) -> torch.Tensor:
        """All decoder layers on T tokens (Qwen2_5_VLDecoderLayer.forward :778-827), causal; fills `cache[layer] = (k, v)` rows [0, T)."""
        c = self.cfg
        T, dev = x.shape[0], x.device
        bf = dict(dtype=torch.bfloat16, device=dev)
        HD, KD = c.heads * c.head_dim, c.kv_heads * c.head_dim
        xn = torch.empty_like(x)
        qkv = torch.empty(T, HD + 2 * KD, **bf)
        att = torch.empty(T, HD, **bf)
        gu = torch.empty(T, 2 * c.intermediate, **bf)
        hm = torch.empty(T, c.intermediate, **bf)
        causal_hi = torch.arange(1, T + 1, dtype=torch.int32, device=dev)
        for i, l in enumerate(self.model.language_model.layers):
            pk = P["text"][i]
            nat.rmsnorm(x, xn, l.input_layernorm.weight, c.rms_eps)
            nat.gemm([dict(a=xn, w=pk["wqkv"], bias=pk["bqkv"], out=qkv)], HD + 2 * KD, c.hidden, nv.EPI_BIAS)
            nat.rope_half(qkv[:, :HD + KD], c.heads + c.kv_heads, c.head_dim, cos, sin, mode=1)        # q and k heads are adjacent columns
            nat.range_attention(qkv[:, :HD], qkv[:, HD:HD + KD], qkv[:, HD + KD:], att, c.heads, c.kv_heads, c.head_dim, c.head_dim ** -0.5, kv_hi=causal_hi)
            if cache is not None:
                cache[i][0][:T].copy_(qkv[:, HD:HD + KD])
                cache[i][1][:T].copy_(qkv[:, HD + KD:])
            nat.gemm([dict(a=att, w=l.self_attn.o_proj.weight, bias=None, out=x, gate=P["ones_h"])], c.hidden, HD, nv.EPI_GATE_RESIDUAL)
            nat.rmsnorm(x, xn, l.post_attention_layernorm.weight, c.rms_eps)
            nat.gemm([dict(a=xn, w=pk["wgu"], bias=None, out=gu)], 2 * c.intermediate, c.hidden, nv.EPI_BIAS)
            nat.swiglu(gu, hm, c.intermediate)
            nat.gemm([dict(a=hm, w=l.mlp.down_proj.weight, bias=None, out=x, gate=P["ones_h"])], c.hidden, c.intermediate, nv.EPI_GATE_RESIDUAL)
        out = torch.empty_like(x)
        nat.rmsnorm(x, out, self.model.language_model.norm.weight, c.rms_eps)
        return out

    def _positions(self, input_ids, image_grid_thw):
        grid = None if image_grid_thw is None else [tuple(int(v) for v in g) for g in (image_grid_thw.tolist() if torch.is_tensor(image_grid_thw) else image_grid_thw)]
        return mrope_position_ids(self.cfg, input_ids.cpu(), grid, self.rope_mode)

    def _check_b1(self, input_ids, attention_mask):
        if input_ids.dim() == 2 and input_ids.shape[0] != 1:
            raise ValueError("the native text encoder is batch 1 (the pipeline encodes one prompt at a time, qwen_image_physical.py:817)")
        if attention_mask is not None and not bool(attention_mask.bool().all()):
            raise ValueError("padded inputs are not produced by the B=1 pipeline")

    @torch.no_grad()
    def edit_forward(self, input_ids=None, attention_mask=None, pixel_values=None, image_grid_thw=None, output_hidden_states=True, **kwargs):
        """-> (final hidden states [1, T, hidden],): the element the pipeline takes with `[-1]` (the last entry of the reference's
        hidden-state tuple is the output of the final norm)."""
        self._check_b1(input_ids, attention_mask)
        nat, P = self._ctx()
        x = self._embed(nat, input_ids, pixel_values, image_grid_thw)
        pos, _ = self._positions(input_ids, image_grid_thw)
        cos, sin = (t.to(x.device) for t in text_rope_tables(self.cfg, pos))
        h = self._prefill(nat, P, x, cos, sin)
        nat.check_async()
        return (h.unsqueeze(0),)

    # ---- greedy generation (GenerationMixin.generate with the wrapper's generation config: do_sample False, eos 151645) ----------
    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, pixel_values=None, image_grid_thw=None, max_new_tokens: int = 1000, **kwargs):
        """Returns [1, T + n] = the prompt followed by the generated ids (ending with the EOS token if one was produced)."""
        return self.generate_batch([dict(input_ids=input_ids, attention_mask=attention_mask, pixel_values=pixel_values, image_grid_thw=image_grid_thw)],
                                   max_new_tokens=max_new_tokens)[0]

    @torch.no_grad()
    def generate_batch(self, requests: Sequence[dict], max_new_tokens: int = 1000) -> List[torch.Tensor]:
        """Greedy generation for up to 8 independent requests decoded TOGETHER: the pipeline generates once per CFG branch
        (qwen_image_physical.py:859-873 runs for the positive and for the negative prompt), and a B=1 decode step streams all 15 GB of
        weights for one token -- decoding both branches in one batch-2 step streams them once for two tokens.  Each request is
        prefilled on its own (different lengths), keeps its own KV cache and device-side counters, and is cut at its own EOS; rows of a
        batched GEMV are computed independently, so the token ids are those of the one-by-one runs."""
        import time
        c = self.cfg
        nat, P = self._ctx()
        dev = self.lm_head.weight.device
        bf = dict(dtype=torch.bfloat16, device=dev)
        nb = len(requests)
        if not 1 <= nb <= 8:
            raise ValueError("generate_batch takes 1..8 requests")
        t0 = time.time()
        KD, HD = c.kv_heads * c.head_dim, c.heads * c.head_dim
        Ts, caches, h_last, deltas = [], [], [], []
        for r in requests:
            self._check_b1(r["input_ids"], r.get("attention_mask"))
            T = r["input_ids"].numel()
            cache = [(torch.empty(T + max_new_tokens, KD, **bf), torch.empty(T + max_new_tokens, KD, **bf)) for _ in range(c.layers)]
            x = self._embed(nat, r["input_ids"], r.get("pixel_values"), r.get("image_grid_thw"))
            pos, delta = self._positions(r["input_ids"], r.get("image_grid_thw"))
            cos, sin = (t.to(dev) for t in text_rope_tables(c, pos))
            h = self._prefill(nat, P, x, cos, sin, cache)
            Ts.append(T); caches.append(cache); h_last.append(h[T - 1:T]); deltas.append(delta)
        # decode-time rotary table: all three axes share the position, row r = position r
        n_rows = max(T + d for T, d in zip(Ts, deltas)) + max_new_tokens + 1
        cos_d, sin_d = (t.to(dev) for t in text_rope_tables(c, torch.arange(n_rows).view(1, -1).expand(3, -1)))
        # device-side state of a decode step, per request: [kv rows cached, rope row, step, kv rows after this step's append]
        ctr = torch.tensor([[T, T + d, 0, T + 1] for T, d in zip(Ts, deltas)], dtype=torch.int32, device=dev)
        log = torch.full((nb, max_new_tokens + 1), -1, dtype=torch.int64, device=dev)
        tok = torch.zeros(nb, dtype=torch.int64, device=dev)
        B = dict(x=torch.empty(nb, c.hidden, **bf), xn=torch.empty(nb, c.hidden, **bf), qkv=torch.empty(nb, HD + 2 * KD, **bf), att=torch.empty(nb, HD, **bf),
                 o=torch.empty(nb, c.hidden, **bf), gu=torch.empty(nb, 2 * c.intermediate, **bf), hm=torch.empty(nb, c.intermediate, **bf),
                 logits=torch.empty(nb, c.vocab, **bf))
        emb_w = self.model.language_model.embed_tokens.weight

        def head(hidden_rows, norm_w=None):
            nat.tag = "te_lm_head"
            if norm_w is not None:
                nat.gemv_fused(hidden_rows, self.lm_head.weight, None, B["logits"], norm_w=norm_w, eps=c.rms_eps)      # final norm fused in
            else:
                nat.gemv(hidden_rows, self.lm_head.weight, None, B["logits"])
            for b in range(nb):
                nat.tag = "te_argmax"
                nat.argmax(B["logits"][b], tok[b:b + 1], log[b], ctr[b, 2:3])

        def step():
            """one token per request: embed(tok) -> 28 layers with the KV caches -> norm -> lm_head -> argmax; every position comes from `ctr`."""
            nat.tag = "te_embed"
            nat.gather_rows(emb_w, tok, B["x"])
            for i, l in enumerate(self.model.language_model.layers):
                pk = P["text"][i]
                if self.fused_decode:
                    # 4 + 2 x requests launches per layer: the norms, SwiGLU and skip adds ride in the GEMVs' prologues / epilogues,
                    # rope + KV append are one kernel (same arithmetic and rounding points as the unfused sequence below)
                    nat.tag = "te_gemv_qkv"
                    nat.gemv_fused(B["x"], pk["wqkv"], pk["bqkv"], B["qkv"], norm_w=l.input_layernorm.weight, eps=c.rms_eps)
                    nat.tag = "te_attention"                     # rope + KV append + attention of every request: one launch
                    nat.decode_attention_fused([B["qkv"][b] for b in range(nb)], [caches[b][i] for b in range(nb)], [B["att"][b] for b in range(nb)],
                                               [ctr[b] for b in range(nb)], c.heads, c.kv_heads, c.head_dim, cos_d, sin_d, c.head_dim ** -0.5)
                    nat.tag = "te_gemv_o"
                    nat.gemv_fused(B["att"], l.self_attn.o_proj.weight, None, B["x"], residual=B["x"])
                    nat.tag = "te_gemv_gate_up"
                    nat.gemv_swiglu(B["x"], pk["wgu"], None, B["hm"], norm_w=l.post_attention_layernorm.weight, eps=c.rms_eps)   # act_fn(gate) * up in the epilogue
                    nat.tag = "te_gemv_down"
                    nat.gemv_fused(B["hm"], l.mlp.down_proj.weight, None, B["x"], residual=B["x"])                                # long K: cluster split-K kernel
                    continue
                nat.tag = "te_rmsnorm"
                nat.rmsnorm(B["x"], B["xn"], l.input_layernorm.weight, c.rms_eps)
                nat.tag = "te_gemv_qkv"
                nat.gemv(B["xn"], pk["wqkv"], pk["bqkv"], B["qkv"])
                for b in range(nb):
                    nat.tag = "te_rope"
                    nat.rope_half(B["qkv"][b:b + 1, :HD + KD], c.heads + c.kv_heads, c.head_dim, cos_d, sin_d, row_ptr=ctr[b, 1:2], mode=1)   # q and k heads are adjacent
                    nat.tag = "te_kv_append"
                    nat.kv_append(B["qkv"][b, HD:HD + KD], B["qkv"][b, HD + KD:], caches[b][i][0], caches[b][i][1], ctr[b, 0:1])
                    nat.tag = "te_attention"
                    nat.range_attention(B["qkv"][b:b + 1, :HD], caches[b][i][0], caches[b][i][1], B["att"][b:b + 1], c.heads, c.kv_heads, c.head_dim,
                                        c.head_dim ** -0.5, kv_len_ptr=ctr[b, 3:4])
                nat.tag = "te_gemv_o"
                nat.gemv(B["att"], l.self_attn.o_proj.weight, None, B["o"])
                nat.tag = "te_residual"
                nat.add_rows(B["x"], B["o"], nb, 1.0)
                nat.tag = "te_rmsnorm"
                nat.rmsnorm(B["x"], B["xn"], l.post_attention_layernorm.weight, c.rms_eps)
                nat.tag = "te_gemv_gate_up"
                nat.gemv(B["xn"], pk["wgu"], None, B["gu"])
                nat.tag = "te_swiglu"
                nat.swiglu(B["gu"], B["hm"], c.intermediate)
                nat.tag = "te_gemv_down"
                nat.gemv(B["hm"], l.mlp.down_proj.weight, None, B["o"])
                nat.tag = "te_residual"
                nat.add_rows(B["x"], B["o"], nb, 1.0)
            nat.tag = "te_advance"
            nat.advance(ctr, 4 * nb)
            if self.fused_decode:
                head(B["x"], self.model.language_model.norm.weight)
            else:
                nat.tag = "te_rmsnorm"
                nat.rmsnorm(B["x"], B["xn"], self.model.language_model.norm.weight, c.rms_eps)
                head(B["xn"])

        head(torch.cat(h_last, dim=0).contiguous())     # token 1 of every request from its prefill's last position (log slot 0)
        done = 1                                        # tokens generated so far (per request)
        t_prefill = time.time()
        graph = None
        if self.use_cuda_graph and max_new_tokens > 2:
            step()                                      # token 2 eagerly (warms every kernel up), then capture ONE step and replay it
            done = 2
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):      # records the launches, does not run them: ctr / log / caches stay as they are
                    step()
            torch.cuda.current_stream(dev).wait_stream(side)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        t_loop, done_at_loop = time.time(), done       # capture / warm-up excluded from the per-token figure
        n_out = [None] * nb
        while any(n is None for n in n_out):
            upto = min(max_new_tokens, done + 32)
            while done < upto:
                if graph is not None:
                    graph.replay()
                else:
                    step()
                done += 1
            got = log[:, :done].tolist()                # one small D2H per 32 tokens: has every request reached its EOS?
            for b in range(nb):
                if n_out[b] is None and c.eos_token_id in got[b]:
                    n_out[b] = got[b].index(c.eos_token_id) + 1
                elif n_out[b] is None and done >= max_new_tokens:
                    n_out[b] = max_new_tokens
        nat.check_async()
        t1 = time.time()
        self.last_generate_stats = {"requests": nb, "prompt_tokens": Ts[0] if nb == 1 else Ts, "new_tokens": int(n_out[0]) if nb == 1 else [int(n) for n in n_out],
                                    "tokens_computed": int(done), "prefill_s": t_prefill - t0, "decode_s": t1 - t_prefill, "graph_capture_s": t_loop - t_prefill,
                                    "ms_per_token": (t1 - t_loop) * 1e3 / max(done - done_at_loop, 1), "cuda_graph": graph is not None}
        return [torch.cat([r["input_ids"].reshape(1, -1).to(dev), log[b, :n_out[b]].view(1, -1)], dim=1) for b, r in enumerate(requests)]


def load_text_encoder(state_dict, torch_dtype=torch.bfloat16, device="cuda", config: Optional[VLConfig] = None):
    """ModelManager.load_model for the `qwen_image_text_encoder` registry row (configs/model_config.py:23; converter "diffusers"):
    recognised by the md5 of the source keys + shapes, converted, meta-init + assign.  Returns None for anything else."""
    from .pipeline import hash_state_dict_keys
    if config is None and hash_state_dict_keys(state_dict) != TEXT_ENCODER_KEY_HASH:
        return None
    sd = QwenImageTextEncoderStateDictConverter().from_diffusers(state_dict)
    with torch.device("meta"):
        te = QwenImageTextEncoder(config)
    te.load_state_dict(sd, assign=True)
    return te.to(dtype=torch_dtype, device=device).eval()
