"""ORACLE (test infrastructure, not product code) -- CPU restatement of PhysicEdit's per-step hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product path (physicedit_b200/) never does and has no CPU fallback.

The reference is pure PyTorch, so the restatement is functional torch on CPU tensors: weights come in
as a flat {state_dict key: tensor} dict (no nn.Module, no einops), every function names the
reference lines it follows (paths relative to /root/reference/DiffSynth-Studio/diffsynth/).

Parity pin: tests/test_oracle_golden.py checks every function here against tests/golden/*.pt, which
oracle/make_golden.py produced by importing and running the reference itself in the authoring
container (seeded synthetic weights; the reference ships no weights, tests or golden vectors).

`dtype` selects the arithmetic: torch.float32 is the oracle proper; torch.bfloat16 replays the
reference's bf16 path op by op (each torch op rounds to bf16 exactly where the reference
materialises a bf16 tensor) and is what the bf16 goldens are checked with.
`cuda_scalar_div=True` replays ATen's CUDA behaviour for `tensor / python_scalar`
(multiply by float(1/scalar), BinaryDivTrueKernel.cu) which the reference hits on the GPU for
`timestep / 1000` and the adapter's alpha.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]

NUM_HEADS = 24
HEAD_DIM = 128
DIM = 3072
AXES_DIM = (16, 56, 56)
ROPE_THETA = 10000


# -------------------------------------------------------------------------------------------------
# small helpers
# -------------------------------------------------------------------------------------------------
def _div_scalar(x: torch.Tensor, s: float, cuda_scalar_div: bool) -> torch.Tensor:
    """tensor / python scalar.  CPU: true division in fp32 opmath.  CUDA: x * float(1/double(s))."""
    if not cuda_scalar_div:
        return x / s
    inv = torch.tensor(1.0 / float(s), dtype=torch.float64).to(torch.float32)
    return (x.float() * inv).to(x.dtype)


def linear(x: torch.Tensor, W: Weights, prefix: str) -> torch.Tensor:
    return F.linear(x, W[prefix + ".weight"], W.get(prefix + ".bias"))


def patchify(latents: torch.Tensor) -> torch.Tensor:
    """rearrange "B C (H P) (W Q) -> B (H W) (C P Q)", P=Q=2  (pipelines/qwen_image_physical.py:1344,1354)."""
    B, C, H2, W2 = latents.shape
    x = latents.reshape(B, C, H2 // 2, 2, W2 // 2, 2)          # B C H P W Q
    return x.permute(0, 2, 4, 1, 3, 5).reshape(B, (H2 // 2) * (W2 // 2), C * 4)


def unpatchify(tokens: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """rearrange "B (H W) (C P Q) -> B C (H P) (W Q)"  (qwen_image_physical.py:1402); H, W in patches."""
    B = tokens.shape[0]
    x = tokens.reshape(B, H, W, 16, 2, 2)                      # B H W C P Q
    return x.permute(0, 3, 1, 4, 2, 5).reshape(B, 16, H * 2, W * 2)


# -------------------------------------------------------------------------------------------------
# timestep embedding (models/utils.py:189-216, 274-293; ctor flags models/qwen_image_dit.py:413)
# -------------------------------------------------------------------------------------------------
def timestep_sinusoid(ts: torch.Tensor) -> torch.Tensor:
    """get_timestep_embedding(ts, 256, flip_sin_to_cos=True, downscale_freq_shift=0, scale=1000,
    align_dtype_to_timestep=True) -> fp32 [B, 256] (cos half first).  ts already holds timestep/1000."""
    half = 128
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=ts.device)
    exponent = exponent / (half - 0)
    emb = torch.exp(exponent)
    emb = emb.to(ts.dtype)                                     # align_dtype_to_timestep: bf16 freqs on the bf16 path
    emb = ts[:, None].float() * emb[None, :]
    emb = 1000 * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)   # flip_sin_to_cos


def time_text_embed(W: Weights, ts: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """TimestepEmbeddings.forward (models/utils.py:289-293) with the diffusers-compatible MLP (:259-270)."""
    e = timestep_sinusoid(ts).to(dtype)
    e = linear(e, W, "time_text_embed.timestep_embedder.linear_1")
    e = F.silu(e)
    return linear(e, W, "time_text_embed.timestep_embedder.linear_2")


# -------------------------------------------------------------------------------------------------
# RoPE (models/qwen_image_dit.py:51-165)
# -------------------------------------------------------------------------------------------------
def _rope_params(index: torch.Tensor, dim: int) -> torch.Tensor:
    freqs = torch.outer(index, 1.0 / torch.pow(ROPE_THETA, torch.arange(0, dim, 2).to(torch.float32).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


def rope_tables(img_shapes: Sequence[Tuple[int, int, int]], txt_len: int, sampling: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """QwenEmbedRope(theta=1e4, axes_dim=[16,56,56], scale_rope=True).forward -> complex64 (vid [S_img,64], txt [T,64]).
    Tables auto-extend in 512 steps beyond 4096 positions (:94-120).  sampling=True: forward_sampling (:168-225,
    `edit_rope_interpolation`) from an EMPTY cache: an image idx > 0 whose (h, w) differs from image 0's takes image 0's table
    sampled at linspace(0, n0 - 1, n).long() rows / columns, with its own frame-axis entries."""
    max_hw = max(max(h // 2, w // 2) for _, h, w in img_shapes)
    n = 4096
    if max_hw + txt_len > n:
        n = math.ceil((max_hw + txt_len) / 512) * 512
    pos_index = torch.arange(n)
    neg_index = torch.arange(n).flip(0) * -1 - 1
    pos = torch.cat([_rope_params(pos_index, d) for d in AXES_DIM], dim=1)
    neg = torch.cat([_rope_params(neg_index, d) for d in AXES_DIM], dim=1)
    split = [d // 2 for d in AXES_DIM]
    fpos = pos.split(split, dim=1)
    fneg = neg.split(split, dim=1)
    vid = []
    max_vid_index = 0
    for idx, (frame, height, width) in enumerate(img_shapes):
        f_frame = fpos[0][idx: idx + frame].view(frame, 1, 1, -1).expand(frame, height, width, -1)
        f_h = torch.cat([fneg[1][-(height - height // 2):], fpos[1][: height // 2]], dim=0)
        f_h = f_h.view(1, height, 1, -1).expand(frame, height, width, -1)
        f_w = torch.cat([fneg[2][-(width - width // 2):], fpos[2][: width // 2]], dim=0)
        f_w = f_w.view(1, 1, width, -1).expand(frame, height, width, -1)
        table = torch.cat([f_frame, f_h, f_w], dim=-1).reshape(frame * height * width, -1)
        f0, h0, w0 = img_shapes[0]
        if sampling and idx > 0 and (height, width) != (h0, w0):
            grid0 = vid[0].reshape(f0, h0, w0, -1)
            hg, wg = torch.meshgrid(torch.linspace(0, h0 - 1, height).long(), torch.linspace(0, w0 - 1, width).long(), indexing="ij")
            table = grid0[:, hg, wg, :].clone()
            table[..., :split[0]] = f_frame
            table = table.reshape(frame * height * width, -1)
        vid.append(table)
        max_vid_index = max(height // 2, width // 2, max_vid_index)
    txt = pos[max_vid_index: max_vid_index + txt_len]
    return torch.cat(vid, dim=0), txt


def apply_rope(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """apply_rotary_emb_qwen (:51-57): adjacent pairs as complex, fp32 multiply, cast back.  x [B,h,S,128]."""
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    return torch.view_as_real(xc * freqs.to(x.device)).flatten(3).type_as(x)


# -------------------------------------------------------------------------------------------------
# norms (models/utils.py:241-257)
# -------------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    var = x.to(torch.float32).square().mean(-1, keepdim=True)
    y = (x * torch.rsqrt(var + eps)).to(x.dtype)
    return y * weight


def layernorm(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), eps=eps)


# -------------------------------------------------------------------------------------------------
# one double-stream block (models/qwen_image_dit.py:247-401)
# -------------------------------------------------------------------------------------------------
def _heads(x: torch.Tensor) -> torch.Tensor:
    B, S, _ = x.shape
    return x.reshape(B, S, NUM_HEADS, HEAD_DIM).permute(0, 2, 1, 3)          # 'b s (h d) -> b h s d'


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qwen_image_flash_attention default branch (:37-38): SDPA, scale 1/sqrt(128); 'b n s d -> b s (n d)'.  attention_mask: the additive
    0 / -inf mask [B, 1, S, S] of process_entity_masks (EliGen)."""
    if attention_mask is not None:
        x = torch.empty_like(q)
        for h in range(q.shape[1]):                      # one head at a time (exact in fp32, bounded memory)
            sc = (q[:, h].float() @ k[:, h].float().transpose(-1, -2)) * (q.shape[-1] ** -0.5) + attention_mask[:, 0].float()
            x[:, h] = (torch.softmax(sc, dim=-1).to(q.dtype) @ v[:, h]) if q.dtype != torch.float32 else torch.softmax(sc, dim=-1) @ v[:, h]
    elif q.is_cuda and q.dtype == torch.float32:
        # The oracle proper on a GPU (tests at the benchmark's sequence lengths): SDPA's fused fp32 kernels may use TF32 tensor
        # cores, and its math backend materialises all heads' S x S scores at once (42 GB at S = 20992).  Same formula, exact
        # fp32, one head at a time.
        x = torch.empty_like(q)
        for h in range(q.shape[1]):
            p = torch.softmax((q[:, h] @ k[:, h].transpose(-1, -2)) * (q.shape[-1] ** -0.5), dim=-1)
            x[:, h] = p @ v[:, h]
    else:
        x = F.scaled_dot_product_attention(q, k, v)
    B, n, S, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, S, n * d)


def _modulate(x: torch.Tensor, mod: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    shift, scale, gate = mod.chunk(3, dim=-1)                                  # (:355-357)
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1), gate.unsqueeze(1)


def _mlp(W: Weights, pre: str, x: torch.Tensor) -> torch.Tensor:
    h = linear(x, W, pre + ".net.0.proj")                                      # ApproximateGELU (:42-49)
    h = h * torch.sigmoid(1.702 * h)
    return linear(h, W, pre + ".net.2")                                        # Dropout(0) is the identity


def block_forward(W: Weights, i: int, image: torch.Tensor, text: torch.Tensor, temb: torch.Tensor,
                  rope: Tuple[torch.Tensor, torch.Tensor], attention_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """QwenImageTransformerBlock.forward (:359-401) + QwenDoubleStreamAttention.forward (:274-316)."""
    p = f"transformer_blocks.{i}"
    img_mod_attn, img_mod_mlp = linear(F.silu(temb), W, p + ".img_mod.1").chunk(2, dim=-1)
    txt_mod_attn, txt_mod_mlp = linear(F.silu(temb), W, p + ".txt_mod.1").chunk(2, dim=-1)
    img_m, img_gate = _modulate(layernorm(image), img_mod_attn)
    txt_m, txt_gate = _modulate(layernorm(text), txt_mod_attn)

    a = p + ".attn"
    img_q, img_k, img_v = (_heads(linear(img_m, W, a + n)) for n in (".to_q", ".to_k", ".to_v"))
    txt_q, txt_k, txt_v = (_heads(linear(txt_m, W, a + n)) for n in (".add_q_proj", ".add_k_proj", ".add_v_proj"))
    seq_txt = txt_q.shape[2]
    img_q, img_k = rmsnorm(img_q, W[a + ".norm_q.weight"]), rmsnorm(img_k, W[a + ".norm_k.weight"])
    txt_q, txt_k = rmsnorm(txt_q, W[a + ".norm_added_q.weight"]), rmsnorm(txt_k, W[a + ".norm_added_k.weight"])
    img_f, txt_f = rope
    img_q, img_k = apply_rope(img_q, img_f), apply_rope(img_k, img_f)
    txt_q, txt_k = apply_rope(txt_q, txt_f), apply_rope(txt_k, txt_f)
    joint = attention(torch.cat([txt_q, img_q], dim=2), torch.cat([txt_k, img_k], dim=2),
                      torch.cat([txt_v, img_v], dim=2), attention_mask).to(img_q.dtype)
    txt_o = linear(joint[:, :seq_txt], W, a + ".to_add_out")
    img_o = linear(joint[:, seq_txt:], W, a + ".to_out.0")

    image = image + img_gate * img_o
    text = text + txt_gate * txt_o
    img_m2, img_gate2 = _modulate(layernorm(image), img_mod_mlp)
    txt_m2, txt_gate2 = _modulate(layernorm(text), txt_mod_mlp)
    image = image + img_gate2 * _mlp(W, p + ".img_mlp", img_m2)
    text = text + txt_gate2 * _mlp(W, p + ".txt_mlp", txt_m2)
    return text, image


# -------------------------------------------------------------------------------------------------
# the PhysicEdit adapter (pipelines/helpers.py:123-164)
# -------------------------------------------------------------------------------------------------
def adapter_alpha(timestep: torch.Tensor, t_min: float, t_max: float, cuda_scalar_div: bool = False) -> torch.Tensor:
    alpha = _div_scalar(timestep - t_min, t_max - t_min + 1e-6, cuda_scalar_div)
    return alpha.clamp(0.0, 1.0).view(-1, 1, 1)


def dual_adapter(A: Weights, x: torch.Tensor, timestep: torch.Tensor, t_min: float, t_max: float,
                 cuda_scalar_div: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """VisualThinkingDualAdapter.forward: two Linear-GELU(erf)-Linear heads blended by alpha(t)."""
    def head(n):
        return linear(F.gelu(linear(x, A, n + ".0")), A, n + ".2")
    pred_dino, pred_vae = head("head_dino"), head("head_vae")
    alpha = adapter_alpha(timestep, t_min, t_max, cuda_scalar_div).type_as(pred_dino)
    return alpha * pred_dino + (1 - alpha) * pred_vae, pred_dino, pred_vae


def adapter_loss(pred_dino, pred_vae, gt_dino, gt_vae, timestep, t_min, t_max, epsilon=0.1):
    """VisualThinkingDualAdapter.get_loss (helpers.py:166-183)."""
    alpha = adapter_alpha(timestep, t_min, t_max).type_as(pred_dino)
    loss_dino = F.mse_loss(pred_dino, gt_dino, reduction="none").mean(dim=[1, 2])
    loss_vae = F.mse_loss(pred_vae, gt_vae, reduction="none").mean(dim=[1, 2])
    w = alpha.squeeze()
    wd, wv = w + epsilon, (1 - w) + epsilon
    tot = wd + wv
    return ((wd / tot) * loss_dino + (wv / tot) * loss_vae).mean()


# -------------------------------------------------------------------------------------------------
# model_fn_qwen_image (pipelines/qwen_image_physical.py:1302-1403), inference branches used by PhysicEdit
# -------------------------------------------------------------------------------------------------
def model_fn(W: Weights, A: Optional[Weights], latents: torch.Tensor, timestep: torch.Tensor, prompt_emb: torch.Tensor,
             prompt_emb_mask: torch.Tensor, special_token_mask: Optional[torch.Tensor], height: int, width: int,
             edit_latents=None, num_layers: Optional[int] = None, t_min: float = 19.999980926513672, t_max: float = 1000.0,
             cuda_scalar_div: bool = False, collect: Optional[dict] = None, edit_rope_interpolation: bool = False,
             controlnet: Optional[list] = None, progress_id: int = 0, num_inference_steps: int = 1, entity: Optional[dict] = None) -> torch.Tensor:
    """Returns the predicted velocity [B,16,H/8,W/8].  MUTATES prompt_emb in place like the reference (:1336)."""
    dtype = latents.dtype
    if special_token_mask is not None:
        special = prompt_emb[special_token_mask].view(prompt_emb.shape[0], -1, prompt_emb.size(-1))
        special, dino_pred, vae_pred = dual_adapter(A, special, timestep, t_min, t_max, cuda_scalar_div)
        prompt_emb[special_token_mask] = special.reshape(-1, special.shape[-1])
        if collect is not None:
            collect["dino_pred"], collect["vae_pred"] = dino_pred, vae_pred
    img_shapes = [(latents.shape[0], latents.shape[2] // 2, latents.shape[3] // 2)]
    txt_len = int(prompt_emb_mask.sum(dim=1).max().item())
    ts = _div_scalar(timestep, 1000, cuda_scalar_div)
    image = patchify(latents)
    image_seq_len = image.shape[1]
    if edit_latents is not None:
        edits = edit_latents if isinstance(edit_latents, list) else [edit_latents]
        img_shapes += [(e.shape[0], e.shape[2] // 2, e.shape[3] // 2) for e in edits]
        image = torch.cat([image] + [patchify(e) for e in edits], dim=1)
    image = linear(image, W, "img_in")
    temb = time_text_embed(W, ts, dtype)
    attention_mask = None
    if entity is not None:                                                                      # :1360-1364, qwen_image_dit.py:433-498
        text, (vid_f, txt_f), attention_mask = process_entity_masks(W, latents, prompt_emb, txt_len, entity["prompt_emb"], entity["masks"], height, width,
                                                                    image.shape[1], img_shapes)
    else:
        text = linear(rmsnorm(prompt_emb, W["txt_norm.weight"]), W, "txt_in")
        vid_f, txt_f = rope_tables(img_shapes, txt_len, sampling=edit_rope_interpolation)      # :1367-1370
    if num_layers is None:
        num_layers = 1 + max(int(k.split(".")[1]) for k in W if k.startswith("transformer_blocks."))
    conds = [controlnet_img_in(c["weights"], patchify(c["latents"])) for c in controlnet] if controlnet else None      # :1372-1374, :164-170
    for i in range(num_layers):
        text, image = block_forward(W, i, image, text, temb, (vid_f, txt_f), attention_mask)
        if conds is not None:                                                                                          # :1389-1396
            image_slice = image[:, :image_seq_len].clone()
            image = image.clone()
            image[:, :image_seq_len] = image_slice + controlnet_sum(controlnet, conds, image_slice, i, progress_id, num_inference_steps)
        if collect is not None:
            collect[f"block{i}"] = (text, image)
    # AdaLayerNorm(single=True): (scale, shift) order (models/utils.py:304-308)
    emb = linear(F.silu(temb), W, "norm_out.linear")
    scale, shift = emb.unsqueeze(1).chunk(2, dim=2)
    image = layernorm(image) * (1 + scale) + shift
    image = linear(image, W, "proj_out")[:, :image_seq_len]
    return unpatchify(image, height // 16, width // 16)


# -------------------------------------------------------------------------------------------------
# EliGen entity control (models/qwen_image_dit.py:433-498)
# -------------------------------------------------------------------------------------------------
def process_entity_masks(W: Weights, latents, prompt_emb, txt_len: int, entity_prompt_emb, entity_masks, height: int, width: int, n_image_tokens: int,
                         img_shapes):
    """entity_prompt_emb: list of [1, L_i, 3584] (unpadded, batch 1); entity_masks [1, N, 1, H/8, W/8] (non-negative).  Returns the joint text
    stream [entity prompts ..., global prompt] after txt_norm / txt_in, (image table, concatenated text tables -- every prompt starts at the same
    position) and the additive attention mask [1, 1, S, S]: prompt i <-> the image tokens whose 2 x 2 latent patch touches mask i (the global
    prompt sees all), repeated over every image of the sequence; different prompts never see each other."""
    text = torch.cat([linear(rmsnorm(e, W["txt_norm.weight"]), W, "txt_in") for e in list(entity_prompt_emb) + [prompt_emb]], dim=1)
    seq_lens = [int(e.shape[1]) for e in entity_prompt_emb] + [txt_len]
    vid_f, _ = rope_tables(img_shapes, txt_len)
    txt_f = torch.cat([rope_tables(img_shapes, n)[1] for n in seq_lens], dim=0)
    N = entity_masks.shape[1] + 1
    patched = [F.max_pool2d(entity_masks[:, i].float(), 2).flatten(1) > 0 for i in range(N - 1)]            # sum over (C P Q) of the patch > 0
    patched.append(torch.ones_like(patched[0]))
    total = sum(seq_lens) + n_image_tokens
    allow = torch.ones(1, total, total, dtype=torch.bool, device=latents.device)
    cum = [0]
    for n in seq_lens:
        cum.append(cum[-1] + n)
    i0 = cum[-1]
    for i in range(N):
        im = patched[i].unsqueeze(1).repeat(1, seq_lens[i], n_image_tokens // patched[i].shape[-1])
        allow[:, cum[i]:cum[i + 1], i0:] = im
        allow[:, i0:, cum[i]:cum[i + 1]] = im.transpose(1, 2)
        for j in range(N):
            if j != i:
                allow[:, cum[i]:cum[i + 1], cum[j]:cum[j + 1]] = False
    mask = torch.zeros(allow.shape, dtype=torch.float32, device=latents.device).masked_fill(~allow, float("-inf"))
    return text, (vid_f, txt_f), mask.to(latents.dtype).unsqueeze(1)


# -------------------------------------------------------------------------------------------------
# blockwise controlnet (models/qwen_image_controlnet.py:6-61, pipelines/qwen_image_physical.py:157-180)
# -------------------------------------------------------------------------------------------------
def controlnet_img_in(C: Weights, tokens: torch.Tensor) -> torch.Tensor:
    return linear(tokens, C, "img_in")


def controlnet_block(C: Weights, i: int, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """BlockWiseControlBlock.forward (:17-22): output_proj(GELU(input_proj(rms(x) + rms(y))))."""
    p = f"controlnet_blocks.{i}."
    h = rmsnorm(x, C[p + "x_rms.weight"]) + rmsnorm(y, C[p + "y_rms.weight"])
    return linear(F.gelu(linear(h, C, p + "input_proj")), C, p + "output_proj")


def controlnet_sum(controlnet, conds, image_slice, block_id, progress_id, num_inference_steps):
    """QwenImageBlockwiseMultiControlNet.blockwise_forward (:172-180); each entry of `controlnet`: dict(weights, latents, scale, start, end)."""
    res = 0
    for c, cond in zip(controlnet, conds):
        progress = (num_inference_steps - 1 - progress_id) / max(num_inference_steps - 1, 1)
        if progress > c.get("start", 1.0) + 1e-4 or progress < c.get("end", 0.0) - 1e-4:
            continue
        res = res + controlnet_block(c["weights"], block_id, image_slice, cond) * c.get("scale", 1.0)
    return res


def controlnet_param_shapes(num_layers: int, in_dim: int = 64, dim: int = 3072) -> Dict[str, Tuple[int, ...]]:
    s = {"img_in.weight": (dim, in_dim), "img_in.bias": (dim,)}
    for i in range(num_layers):
        p = f"controlnet_blocks.{i}."
        s.update({p + "x_rms.weight": (dim,), p + "y_rms.weight": (dim,), p + "input_proj.weight": (dim, dim), p + "input_proj.bias": (dim,),
                  p + "output_proj.weight": (dim, dim), p + "output_proj.bias": (dim,)})
    return s


# -------------------------------------------------------------------------------------------------
# scheduler (schedulers/flow_match.py) and the denoise loop (qwen_image_physical.py:646-661)
# -------------------------------------------------------------------------------------------------
class FlowMatchOracle:
    """FlowMatchScheduler(sigma_min=0, sigma_max=1, extra_one_step=True, exponential_shift=True,
    exponential_shift_mu=0.8, shift_terminal=0.02) -- the ctor arguments at qwen_image_physical.py:192."""

    def __init__(self):
        self.num_train_timesteps = 1000
        self.sigma_min, self.sigma_max = 0.0, 1.0
        self.exponential_shift_mu = 0.8
        self.shift_terminal = 0.02
        self.set_timesteps(100)

    @staticmethod
    def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=8192, base_shift=0.5, max_shift=0.9):
        m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
        b = base_shift - m * base_seq_len
        return image_seq_len * m + b

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, training=False, dynamic_shift_len=None):
        sigma_start = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        sigmas = torch.linspace(sigma_start, self.sigma_min, num_inference_steps + 1)[:-1]
        mu = self.calculate_shift(dynamic_shift_len) if dynamic_shift_len is not None else self.exponential_shift_mu
        sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1))
        one_minus_z = 1 - sigmas
        scale_factor = one_minus_z[-1] / (1 - self.shift_terminal)
        self.sigmas = 1 - (one_minus_z / scale_factor)
        self.timesteps = self.sigmas * self.num_train_timesteps
        if training:
            x = self.timesteps
            y = torch.exp(-2 * ((x - num_inference_steps / 2) / num_inference_steps) ** 2)
            y_shifted = y - y.min()
            self.linear_timesteps_weights = y_shifted * (num_inference_steps / y_shifted.sum())

    def dsigma(self, progress_id: int):
        """(sigma_next - sigma) exactly as FlowMatchScheduler.step computes it (flow_match.py:72-82)."""
        timestep = self.timesteps[progress_id]
        tid = torch.argmin((self.timesteps - timestep).abs())
        sigma = self.sigmas[tid]
        sigma_ = 0 if tid + 1 >= len(self.timesteps) else self.sigmas[tid + 1]
        return sigma_ - sigma

    def step(self, model_output, progress_id, sample):
        return sample + model_output * self.dsigma(progress_id)

    def add_noise(self, original, noise, timestep):
        tid = torch.argmin((self.timesteps - timestep).abs())
        sigma = self.sigmas[tid]
        return (1 - sigma) * original + sigma * noise

    def training_weight(self, timestep):
        tid = torch.argmin((self.timesteps - timestep).abs())
        return self.linear_timesteps_weights[tid]


def denoise_loop(W, A, latents, posi, nega, edit_latents, height, width, num_inference_steps, cfg_scale=4.0,
                 num_layers=None, cuda_scalar_div=False, timestep_dtype=None):
    """QwenImagePhysicPipeline.__call__ lines 600, 646-661: CFG (2 forwards / step) + Euler update.
    posi / nega: dicts with prompt_emb, prompt_emb_mask, special_token_mask (prompt_emb is mutated in place)."""
    sch = FlowMatchOracle()
    sch.set_timesteps(num_inference_steps, dynamic_shift_len=(height // 16) * (width // 16))
    dtype = timestep_dtype or latents.dtype          # timestep_dtype=bf16 with fp32 latents: fp32 arithmetic on the bf16 path's timestep bookkeeping
    for pid, t in enumerate(sch.timesteps):
        t = t.unsqueeze(0).to(dtype).to(latents.device)
        kw = dict(height=height, width=width, edit_latents=edit_latents, num_layers=num_layers, cuda_scalar_div=cuda_scalar_div)
        vp = model_fn(W, A, latents, t, posi["prompt_emb"], posi["prompt_emb_mask"], posi["special_token_mask"], **kw)
        if cfg_scale != 1.0:
            vn = model_fn(W, A, latents, t, nega["prompt_emb"], nega["prompt_emb_mask"], nega["special_token_mask"], **kw)
            v = vn + cfg_scale * (vp - vn)
        else:
            v = vp
        latents = sch.step(v, pid, latents)
    return latents


# -------------------------------------------------------------------------------------------------
# LoRA fold (lora/__init__.py:5-45) and loader key hashing (models/utils.py:148-182)
# -------------------------------------------------------------------------------------------------
def lora_name_dict(lora_state_dict) -> Dict[str, Tuple[str, str]]:
    out = {}
    for key in lora_state_dict:
        if ".lora_B." not in key:
            continue
        keys = key.split(".")
        if len(keys) > keys.index("lora_B") + 2:
            keys.pop(keys.index("lora_B") + 1)
        keys.pop(keys.index("lora_B"))
        if keys[0] == "diffusion_model":
            keys.pop(0)
        keys.pop(-1)
        out[".".join(keys)] = (key, key.replace(".lora_B.", ".lora_A."))
    return out


def lora_fold(W: Weights, lora_sd, alpha: float = 1.0, dtype=torch.float32) -> int:
    """W[name.weight] <- W + alpha * (B @ A), computed in `dtype` (the pipe dtype, bf16 in the scripts)."""
    n = 0
    for name, (kb, ka) in lora_name_dict(lora_sd).items():
        key = name + ".weight"
        if key not in W:
            continue
        up, down = lora_sd[kb].to(dtype), lora_sd[ka].to(dtype)
        W[key] = W[key].to(dtype) + alpha * torch.mm(up, down)
        n += 1
    return n


def state_dict_key_hash(shapes: Dict[str, Sequence[int]]) -> str:
    """hash_state_dict_keys(with_shape=True): md5 of sorted 'key:shape' and 'key' entries joined by ','."""
    import hashlib
    keys: List[str] = []
    for k, shp in shapes.items():
        keys.append(k + ":" + "_".join(map(str, list(shp))))
        keys.append(k)
    keys.sort()
    return hashlib.md5(",".join(keys).encode("UTF-8")).hexdigest()


# -------------------------------------------------------------------------------------------------
# seeded synthetic weights (the reference ships none): shared by the goldens, the tests and bench.py
# -------------------------------------------------------------------------------------------------
def dit_param_shapes(num_layers: int) -> Dict[str, Tuple[int, ...]]:
    """Exact parameter names / shapes of QwenImageDiT (models/qwen_image_dit.py:404-430)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(name, out_f, in_f):
        s[name + ".weight"] = (out_f, in_f)
        s[name + ".bias"] = (out_f,)
    lin("time_text_embed.timestep_embedder.linear_1", DIM, 256)
    lin("time_text_embed.timestep_embedder.linear_2", DIM, DIM)
    s["txt_norm.weight"] = (3584,)
    lin("img_in", DIM, 64)
    lin("txt_in", DIM, 3584)
    for i in range(num_layers):
        p = f"transformer_blocks.{i}"
        lin(p + ".img_mod.1", 6 * DIM, DIM)
        for n in ("to_q", "to_k", "to_v"):
            lin(f"{p}.attn.{n}", DIM, DIM)
        s[p + ".attn.norm_q.weight"] = (HEAD_DIM,)
        s[p + ".attn.norm_k.weight"] = (HEAD_DIM,)
        for n in ("add_q_proj", "add_k_proj", "add_v_proj"):
            lin(f"{p}.attn.{n}", DIM, DIM)
        s[p + ".attn.norm_added_q.weight"] = (HEAD_DIM,)
        s[p + ".attn.norm_added_k.weight"] = (HEAD_DIM,)
        lin(p + ".attn.to_out.0", DIM, DIM)
        lin(p + ".attn.to_add_out", DIM, DIM)
        lin(p + ".img_mlp.net.0.proj", 4 * DIM, DIM)
        lin(p + ".img_mlp.net.2", DIM, 4 * DIM)
        lin(p + ".txt_mod.1", 6 * DIM, DIM)
        lin(p + ".txt_mlp.net.0.proj", 4 * DIM, DIM)
        lin(p + ".txt_mlp.net.2", DIM, 4 * DIM)
    lin("norm_out.linear", 2 * DIM, DIM)
    lin("proj_out", 64, DIM)
    return s


def adapter_param_shapes(dim: int = 3584) -> Dict[str, Tuple[int, ...]]:
    s = {}
    for h in ("head_dino", "head_vae"):
        s[f"{h}.0.weight"], s[f"{h}.0.bias"] = (3 * dim, dim), (3 * dim,)
        s[f"{h}.2.weight"], s[f"{h}.2.bias"] = (dim, 3 * dim), (dim,)
    return s


def synth_weights(shapes: Dict[str, Tuple[int, ...]], seed: int, dtype=torch.float32, device="cpu", weight_gain: float = 1.0) -> Weights:
    """Deterministic weights: one torch.Generator per tensor seeded by (seed, position in the sorted key list),
    so any subset of layers can be rebuilt without generating the others.  Linear weights ~ U(-b, b) with
    b = gain/sqrt(fan_in) (PyTorch's default Linear init bound), biases ~ U(-b, b), norm weights 1 + 0.1 N(0,1)."""
    out = {}
    for n, key in enumerate(sorted(shapes)):
        shp = shapes[key]
        g = torch.Generator("cpu").manual_seed(seed * 1000003 + n)
        if len(shp) == 2:
            b = weight_gain / math.sqrt(shp[1])
            t = (torch.rand(shp, generator=g, dtype=torch.float32) * 2 - 1) * b
        elif key.endswith(".bias"):
            wshape = shapes.get(key[:-5] + ".weight", ())
            fan_in = wshape[1] if len(wshape) == 2 else 2500          # LayerNorm biases: small values
            t = (torch.rand(shp, generator=g, dtype=torch.float32) * 2 - 1) / math.sqrt(fan_in)
        else:
            t = 1 + 0.1 * torch.randn(shp, generator=g, dtype=torch.float32)
        out[key] = t.to(dtype).to(device)
    return out


def synth_inputs(height: int, width: int, T: int, seed: int, dtype=torch.float32, edit_hw: Optional[Tuple[int, int]] = None, n_special: int = 64):
    """Synthetic step inputs in the shapes of SURVEY 8d: latents via the reference's generate_noise recipe
    (utils/__init__.py:119-124: CPU generator, fp32 randn, cast), prompt_emb ~ 3 N(0,1), all-ones mask,
    special mask = n_special consecutive rows ending 5 before the end."""
    g = torch.Generator("cpu").manual_seed(seed)
    latents = torch.randn((1, 16, height // 8, width // 8), generator=g, dtype=torch.float32).to(dtype)
    eh, ew = edit_hw or (height, width)
    edit = torch.randn((1, 16, eh // 8, ew // 8), generator=g, dtype=torch.float32).to(dtype)
    prompt = (3 * torch.randn((1, T, 3584), generator=g, dtype=torch.float32)).to(dtype)
    mask = torch.ones((1, T), dtype=torch.int64)
    special = torch.zeros((1, T), dtype=torch.bool)
    special[0, T - 5 - n_special: T - 5] = True
    return dict(latents=latents, edit_latents=edit, prompt_emb=prompt, prompt_emb_mask=mask, special_token_mask=special)
