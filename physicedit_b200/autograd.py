"""Training path: the differentiable forward of the DiT + adapters and its backward on the native kernels.

What the reference does: `pipe.training_loss` (DiffSynth-Studio/diffsynth/pipelines/qwen_image_physical.py:313-329) runs `model_fn_qwen_image`
(:1302-1403) under autograd with PEFT LoRA r = 128 injected un-merged into 12 linears of each of the 60 blocks
(diffsynth/trainers/utils.py:799-808, scripts/train/train_multigpu.sh:30-31), the adapter stack trainable, gradient checkpointing per block
(:1379-1389), `accelerator.backward(loss)` (scripts/train/train_physicedit.py:648-652).

What runs here -- first slice of the row, stated honestly:
  * every matrix product of the forward AND the backward is the tcgen05 GEMM (`pe_gemm`): Y = X W^T (+ b), dX = dY W (W^T materialised by
    `pe_transpose`, cached for frozen weights), dW | db = dY^T [X | 1] (one GEMM: the bias gradient is the extra column), LoRA down / up
    projections and their gradients (N or K = 128);
  * attention forward is the flash kernel, which also returns the row statistics (`pe_attention_fwd_lse`); its backward is seven BATCHED GEMM
    launches per block (`pe_gemm_batched`, one problem per head) whose epilogues rebuild P / P^T from the scores and turn dP into dS / dS^T in
    place (PE_EPI_ATTN_P / PE_EPI_ATTN_DS) -- the bf16 S x S matrices still pass through HBM (18 bytes per score element), so it is HBM-bound;
    a fully fused tcgen05 flash backward is the next step;
  * the row-wise glue between them (LayerNorm + modulation, per-head RMSNorm + RoPE, x sigmoid(1.702 x), gate * branch + residual, GELU(erf),
    the blend, the losses) is written with torch ops in the reference's own op order and differentiated by autograd: ATen CUDA kernels, not
    CPU code and not the oracle, < 3 % of the FLOPs; native fused backward passes for them are also "next".
Inference never comes here (model_fn routes here only when a gradient is required or un-merged LoRA is present).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

from . import native as nv
from .lora import HotLoRALinear, LoRALinear

HEAD_DIM = 128
BF16 = torch.bfloat16


def _nat(t: torch.Tensor) -> nv.Native:
    if not t.is_cuda or t.dtype != BF16:
        raise nv.NativeUnavailable(f"the native training path runs in bfloat16 on an sm_100 GPU only (got {t.dtype} on {t.device}); no fallback")
    return nv.Native.get(t.device.index or 0)


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _rows(x: torch.Tensor) -> torch.Tensor:
    """[..., K] -> a [M, K] view / copy the GEMM can read (innermost stride 1, row stride a multiple of 8)."""
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1 or x2.stride(0) % 8 or x2.stride(0) < x2.shape[1] or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    return x2


def _transposed(nat: nv.Native, x2: torch.Tensor, extra_rows: int = 0, ones_row: bool = False) -> torch.Tensor:
    """x2 [M, C] -> [C + extra_rows, pad8(M)] holding x2^T, zero in the padding; row C = ones over the M valid columns if asked."""
    M, C = x2.shape
    out = torch.zeros(C + extra_rows, _pad8(M), dtype=BF16, device=x2.device) if (extra_rows or M % 8) else \
        torch.empty(C, M, dtype=BF16, device=x2.device)
    nat.transpose(x2, out[:C, :M])
    if ones_row:
        out[C, :M] = 1
    return out


class _WeightTransposes:
    """W^T of FROZEN weights for the dX GEMMs, keyed by storage address + version (an in-place LoRA fold bumps the version).  An address alone
    can be recycled by the allocator for a different tensor, so the forward registers the live weight object: when that object dies its entry
    is dropped by a weakref finalizer before any new tensor can appear at the address."""

    def __init__(self, cap_bytes: int = 48 << 30):
        self.cap, self.used, self.d, self.owners = cap_bytes, 0, {}, {}

    @staticmethod
    def _key(w: torch.Tensor):
        return (w.data_ptr(), w._version, tuple(w.shape))

    def register(self, w: torch.Tensor) -> None:
        """Called in the forward with the weight object itself (a module parameter)."""
        if w.requires_grad:
            return
        key = self._key(w)
        if key not in self.owners:
            import weakref
            self.owners[key] = weakref.ref(w, lambda _ref, key=key: self._drop(key))

    def _drop(self, key) -> None:
        self.owners.pop(key, None)
        t = self.d.pop(key, None)
        if t is not None:
            self.used -= t.numel() * 2

    def get(self, nat: nv.Native, w: torch.Tensor) -> torch.Tensor:
        key = self._key(w)
        if w.requires_grad or key not in self.owners:          # trainable, or an unregistered temporary: never cached
            return _transposed(nat, w)
        t = self.d.get(key)
        if t is None:
            t = _transposed(nat, w)
            nbytes = t.numel() * 2
            if self.used + nbytes <= self.cap:
                self.d[key], self.used = t, self.used + nbytes
        return t

    def clear(self):
        self.d.clear()
        self.owners.clear()
        self.used = 0


weight_transposes = _WeightTransposes()


def _gemm(nat: nv.Native, a2: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a2 w^T (+ b) with the tile shape the inference engine uses for the big layers: cta_group::2 (256-row tiles on an SM pair)."""
    pair = w.shape[0] % 256 == 0 and a2.shape[0] >= 512
    return nat.linear(a2, w, b, nv.EPI_BIAS, nv.GEMM_FLAG_CTA_PAIR if pair else 0)


class _LinearFn(torch.autograd.Function):
    """y = x w^T (+ b) on pe_gemm; either operand may require grad (activations x activations products use it too)."""

    @staticmethod
    def forward(ctx, x, w, b):
        nat = _nat(x)
        x2 = _rows(x)
        wc = w if w.is_contiguous() else w.contiguous()
        if wc is w and isinstance(w, torch.nn.Parameter):
            weight_transposes.register(w)
        y = _gemm(nat, x2, wc, b)
        ctx.save_for_backward(x2, wc)
        ctx.has_bias, ctx.in_shape = b is not None, x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        nat = _nat(x2)
        dy2 = _rows(dy)
        N, K = w.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _gemm(nat, dy2, weight_transposes.get(nat, w)).view(ctx.in_shape)
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1] or need_b:
            dyt = _transposed(nat, dy2)                                    # [N, Mp]
            xt = _transposed(nat, x2, extra_rows=8 if need_b else 0, ones_row=need_b)     # [K (+8), Mp]: the ones row makes column K the bias gradient
            g = _gemm(nat, dyt, xt)                                        # [N, K (+8)]
            if ctx.needs_input_grad[1]:
                dw = g[:, :K].contiguous() if need_b else g
            if need_b:
                db = g[:, K].contiguous()
        return dx, dw, db


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _LinearFn.apply(x, w, b)


def lora_linear(x: torch.Tensor, m: LoRALinear) -> torch.Tensor:
    """peft/tuners/lora/layer.py Linear.forward at dropout 0: base(x) + lora_B(lora_A(x)) * scaling, every product a native GEMM."""
    y = linear(x, m.base_layer.weight, m.base_layer.bias)
    z = linear(linear(x, m.lora_A["default"].weight), m.lora_B["default"].weight)
    return y + z * m.scaling


def module_linear(m: torch.nn.Module, x: torch.Tensor) -> torch.Tensor:
    if isinstance(m, LoRALinear):
        return lora_linear(x, m)
    y = linear(x, m.weight, m.bias)
    if isinstance(m, HotLoRALinear):                      # AutoWrappedLinear.forward (vram_management/layers.py:177-179): out + x A^T B^T per attached LoRA
        for a, b in zip(m.lora_A_weights, m.lora_B_weights):
            y = y + linear(linear(x, a), b)
    return y


ATTN_BWD_SCRATCH_BYTES = 16 << 30        # S x S scratch of the attention backward: heads are processed in chunks that fit


class _AttentionFn(torch.autograd.Function):
    """Joint non-causal attention, head dim 128, q / k / v / o token-major [S, H * 128].

    Backward, per chunk of heads, SEVEN batched tcgen05 GEMM launches (`pe_gemm_batched`, one problem per head) and no separate softmax /
    transpose pass over the S x S matrices: with L the forward's log2-domain row statistic and delta = rowsum(dO * O),
        P    = exp2(c Q K^T - L[row])   P^T  = exp2(c K Q^T - L[col])            (epilogue PE_EPI_ATTN_P)
        dS   = P (dO V^T - delta[row]) s   dS^T = P^T (V dO^T - delta[col]) s    (epilogue PE_EPI_ATTN_DS, in place over P / P^T)
        dV = P^T dO      dQ = dS K      dK = dS^T Q
    Recomputing the transposed matrices on the tensor cores (2 extra products of 2 S^2 128 FLOPs) is cheaper than transposing them through
    HBM.  18 bytes of HBM traffic per score element in all (bf16 matrices only)."""

    @staticmethod
    def forward(ctx, q, k, v, H):
        nat = _nat(q)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        o = torch.empty_like(q)
        lse = torch.empty(H, q.shape[0], dtype=torch.float32, device=q.device)
        nat.attention_lse(q, k, v, o, lse, H, 1.0 / math.sqrt(HEAD_DIM))
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.H = H
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        nat, H = _nat(q), ctx.H
        S, Sp, D = q.shape[0], (q.shape[0] + 15) // 16 * 16, HEAD_DIM          # 16: row strides that allow the epilogues' 32-byte stores
        dev, scale = q.device, 1.0 / math.sqrt(HEAD_DIM)
        do = do.contiguous()

        def head_major(t):                                   # [H, Sp, 128], rows >= S zero (they only ever meet zeros of the other operand)
            out = torch.zeros(H, Sp, D, dtype=BF16, device=dev) if Sp != S else torch.empty(H, S, D, dtype=BF16, device=dev)
            out[:, :S].copy_(t.view(S, H, D).transpose(0, 1))
            return out
        qh, kh, vh, doh = (head_major(t) for t in (q, k, v, do))
        qt, kt, dot = (t.transpose(1, 2).contiguous() for t in (qh, kh, doh))          # [H, 128, Sp]: the K-major operands of dK, dQ, dV
        stat = torch.zeros(2, H, Sp, dtype=torch.float32, device=dev)                  # [0] = L, [1] = delta; padding 0 (finite)
        stat[0, :, :S] = lse
        nat.attention_bwd_delta(do, o, stat[1], H)
        dq, dk, dv = (torch.empty(H, Sp, D, dtype=BF16, device=dev) for _ in range(3))
        hc = max(1, min(H, ATTN_BWD_SCRATCH_BYTES // (4 * Sp * Sp)))                    # two bf16 S x S matrices per head
        P, Pt = (torch.empty(hc * Sp, Sp, dtype=BF16, device=dev) for _ in range(2))
        c = scale * 1.4426950408889634
        flat = lambda t, h0, n: t[h0:h0 + n].reshape(n * t.shape[1], t.shape[2])
        for h0 in range(0, H, hc):
            n = min(hc, H - h0)
            L, dl = stat[0, h0:h0 + n], stat[1, h0:h0 + n]
            Q, K, V, DO = (flat(t, h0, n) for t in (qh, kh, vh, doh))
            kw = dict(batch=n, M=S, a_batch_rows=Sp, out_batch_rows=Sp, vec_batch_stride=Sp)
            nat.gemm_batched(Q, K, P, N=Sp, K=D, w_batch_rows=Sp, epilogue=nv.EPI_ATTN_P, vec=L, alpha=c, **kw)                          # P
            nat.gemm_batched(K, Q, Pt, N=Sp, K=D, w_batch_rows=Sp, epilogue=nv.EPI_ATTN_P, vec=L, vec_per_column=True, alpha=c, **kw)     # P^T
            nat.gemm_batched(Pt, flat(dot, h0, n), flat(dv, h0, n), N=D, K=Sp, w_batch_rows=D, **kw)                                     # dV = P^T dO
            nat.gemm_batched(DO, V, P, N=Sp, K=D, w_batch_rows=Sp, epilogue=nv.EPI_ATTN_DS, vec=dl, alpha=scale, **kw)                    # dS over P
            nat.gemm_batched(V, DO, Pt, N=Sp, K=D, w_batch_rows=Sp, epilogue=nv.EPI_ATTN_DS, vec=dl, vec_per_column=True, alpha=scale, **kw)   # dS^T over P^T
            nat.gemm_batched(P, flat(kt, h0, n), flat(dq, h0, n), N=D, K=Sp, w_batch_rows=D, **kw)                                       # dQ = dS K
            nat.gemm_batched(Pt, flat(qt, h0, n), flat(dk, h0, n), N=D, K=Sp, w_batch_rows=D, **kw)                                      # dK = dS^T Q

        def token_major(t):
            return t[:, :S].transpose(0, 1).reshape(S, H * D)
        return token_major(dq), token_major(dk), token_major(dv), None


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, H: int) -> torch.Tensor:
    return _AttentionFn.apply(q, k, v, H)


# ------------------------------------------------------------------------------------------------
# row-wise glue in the reference's op order (torch ops; autograd differentiates them)
# ------------------------------------------------------------------------------------------------
def _rmsnorm(x: torch.Tensor, weight: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """models/utils.py:241-257."""
    var = x.to(torch.float32).square().mean(-1, keepdim=True)
    y = (x * torch.rsqrt(var + eps)).to(x.dtype)
    return y * weight if weight is not None else y


def _rope(x: torch.Tensor, cs: torch.Tensor) -> torch.Tensor:
    """apply_rotary_emb_qwen (qwen_image_dit.py:51-57) with the complex table as (cos, sin): x [S, H, 128], cs fp32 [S, 64, 2]."""
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    c, s = cs[:, None, :, 0], cs[:, None, :, 1]
    out = torch.stack((xf[..., 0] * c - xf[..., 1] * s, xf[..., 0] * s + xf[..., 1] * c), dim=-1)
    return out.flatten(-2).type_as(x)


def _modulate(x: torch.Tensor, mod: torch.Tensor):
    """QwenImageTransformerBlock._modulate (:345-347) on 2-D streams: mod [1, 3 * dim] -> (x (1 + scale) + shift, gate)."""
    shift, scale, gate = mod.chunk(3, dim=-1)
    return x * (1 + scale) + shift, gate


def block_forward(blk, image: torch.Tensor, text: torch.Tensor, temb: torch.Tensor, rope_img: torch.Tensor, rope_txt: torch.Tensor):
    """QwenImageTransformerBlock.forward + QwenDoubleStreamAttention.forward (qwen_image_dit.py:247-316, 359-401) on [S, 3072] streams."""
    D, H = image.shape[-1], blk.num_attention_heads
    st = F.silu(temb)                                                    # the nn.SiLU in front of img_mod / txt_mod
    img_mod_attn, img_mod_mlp = module_linear(blk.img_mod[1], st).chunk(2, dim=-1)
    txt_mod_attn, txt_mod_mlp = module_linear(blk.txt_mod[1], st).chunk(2, dim=-1)
    ln = lambda x: F.layer_norm(x, (D,), eps=1e-6)
    img_m, img_gate = _modulate(ln(image), img_mod_attn)
    txt_m, txt_gate = _modulate(ln(text), txt_mod_attn)
    a = blk.attn

    def heads(x, proj, norm, rope):
        y = module_linear(proj, x).view(x.shape[0], H, HEAD_DIM)
        if norm is not None:
            y = _rope(_rmsnorm(y, norm.weight, norm.eps), rope)
        return y.reshape(x.shape[0], D)
    q = torch.cat([heads(txt_m, a.add_q_proj, a.norm_added_q, rope_txt), heads(img_m, a.to_q, a.norm_q, rope_img)], dim=0)
    k = torch.cat([heads(txt_m, a.add_k_proj, a.norm_added_k, rope_txt), heads(img_m, a.to_k, a.norm_k, rope_img)], dim=0)
    v = torch.cat([heads(txt_m, a.add_v_proj, None, None), heads(img_m, a.to_v, None, None)], dim=0)
    o = attention(q, k, v, H)
    T = text.shape[0]
    image = image + img_gate * module_linear(a.to_out[0], o[T:])
    text = text + txt_gate * module_linear(a.to_add_out, o[:T])
    img_m2, img_gate2 = _modulate(ln(image), img_mod_mlp)
    txt_m2, txt_gate2 = _modulate(ln(text), txt_mod_mlp)

    def mlp(net, x):
        h = module_linear(net[0].proj, x)
        return module_linear(net[2], h * torch.sigmoid(1.702 * h))       # ApproximateGELU (:42-49); Dropout(0) is the identity
    image = image + img_gate2 * mlp(blk.img_mlp.net, img_m2)
    text = text + txt_gate2 * mlp(blk.txt_mlp.net, txt_m2)
    return text, image


def controlnet_conditionings(multi, controlnet_inputs, conditionings):
    """QwenImageBlockwiseMultiControlNet.preprocess (qwen_image_physical.py:164-170) under autograd: control latents -> 2 x 2 patch tokens -> `img_in`
    of their controlnet (input width padded to a multiple of 8 for the inpaint variant's 64 + 4 channels)."""
    out = []
    for ci, cond in zip(controlnet_inputs, conditionings):
        B, C, Hh, Ww = cond.shape
        tok = cond.view(B, C, Hh // 2, 2, Ww // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape((Hh // 2) * (Ww // 2), C * 4).to(BF16)
        lin = multi.models[ci.controlnet_id].img_in
        w, pad = lin.weight, (-tok.shape[1]) % 8
        if pad:
            tok, w = F.pad(tok, (0, pad)), F.pad(w, (0, pad))
        out.append(linear(tok.contiguous(), w, lin.bias))
    return out


def controlnet_residual(multi, image: torch.Tensor, conds, controlnet_inputs, progress_id, num_inference_steps, block_id):
    """QwenImageBlockwiseMultiControlNet.blockwise_forward (:172-180) + BlockWiseControlBlock.forward (qwen_image_controlnet.py:15-20) on [n0, 3072]
    tokens: sum over the controlnets active at this progress of scale * output_proj(GELU(input_proj(rms(x) + rms(y))))."""
    res = 0
    for ci, cond in zip(controlnet_inputs, conds):
        if not multi.active(ci, progress_id, num_inference_steps):
            continue
        blk = multi.models[ci.controlnet_id].controlnet_blocks[block_id]
        h = _rmsnorm(image, blk.x_rms.weight, blk.x_rms.eps) + _rmsnorm(cond, blk.y_rms.weight, blk.y_rms.eps)
        res = res + module_linear(blk.output_proj, F.gelu(module_linear(blk.input_proj, h))) * ci.scale
    return res


def dit_forward(dit, latents_list: Sequence[torch.Tensor], timestep_bf16: torch.Tensor, prompt_emb: torch.Tensor,
                use_gradient_checkpointing: bool = True, rope_sampling: bool = False, after_block=None) -> torch.Tensor:
    """model_fn_qwen_image :1340-1403 under autograd.  latents_list = [noise latents, (context), edit...] each [1, 16, h8, w8] bf16 (no gradient
    flows into them: the path trains LoRA and adapters only); timestep_bf16 [1] = the loop's bf16(t); prompt_emb [1, T, 3584] (may carry the
    adapter's graph).  Returns the velocity [1, 16, h8, w8]."""
    eng = dit.engine()
    nat = eng.nat
    T = prompt_emb.shape[1]
    shapes = [(1, l.shape[-2] // 2, l.shape[-1] // 2) for l in latents_list]
    S_img = sum(h * w for _, h, w in shapes)
    with torch.no_grad():                                                 # frozen entry stages whose inputs need no gradient
        tok = torch.empty(S_img, 64, dtype=BF16, device=prompt_emb.device)
        off = 0
        for l, (_, h, w) in zip(latents_list, shapes):
            nat.patchify(l.reshape(16, l.shape[-2], l.shape[-1]).contiguous(), tok[off:off + h * w])
            off += h * w
        temb = dit.time_text_embed(timestep_bf16, raw=True)               # [1, 3072]
        rope = eng.rope(shapes, T, rope_sampling)                         # fp32 (cos, sin) [T + S_img, 64, 2], text rows first
    image = module_linear(dit.img_in, tok)
    text = module_linear(dit.txt_in, _rmsnorm(prompt_emb[0], dit.txt_norm.weight, dit.txt_norm.eps))
    rope_txt, rope_img = rope[:T], rope[T:]
    for block_id, blk in enumerate(dit.transformer_blocks):
        if use_gradient_checkpointing and torch.is_grad_enabled():
            text, image = checkpoint(block_forward, blk, image, text, temb, rope_img, rope_txt, use_reentrant=False)
        else:
            text, image = block_forward(blk, image, text, temb, rope_img, rope_txt)
        if after_block is not None:                                       # blockwise controlnet (:1389-1396): the noise tokens get block_id's correction
            n_noise = shapes[0][1] * shapes[0][2]
            res = after_block(block_id, image[:n_noise])
            if torch.is_tensor(res):
                image = torch.cat([image[:n_noise] + res, image[n_noise:]], dim=0)
    _, h0, w0 = shapes[0]
    n0 = h0 * w0
    emb = module_linear(dit.norm_out.linear, F.silu(temb))                # AdaLayerNorm(single) (models/utils.py:296-309): (scale, shift)
    scale, shift = emb.chunk(2, dim=-1)
    x = F.layer_norm(image[:n0], (image.shape[-1],), eps=1e-6) * (1 + scale) + shift
    out = module_linear(dit.proj_out, x)                                  # [n0, 64] = (h w) (c p q)
    return out.view(h0, w0, 16, 2, 2).permute(2, 0, 3, 1, 4).reshape(1, 16, 2 * h0, 2 * w0)


# ------------------------------------------------------------------------------------------------
# adapter stack under autograd
# ------------------------------------------------------------------------------------------------
def mlp_gelu(x: torch.Tensor, l0: torch.nn.Linear, l2: torch.nn.Linear) -> torch.Tensor:
    """Linear -> nn.GELU (erf) -> Linear (helpers.py:112-121, 127-137, 8-19)."""
    return module_linear(l2, F.gelu(module_linear(l0, x)))


def dual_adapter_forward(ad, x: torch.Tensor, timestep: torch.Tensor):
    """VisualThinkingDualAdapter.forward (helpers.py:150-164): (mixed, pred_dino, pred_vae)."""
    pd = mlp_gelu(x, ad.head_dino[0], ad.head_dino[2])
    pv = mlp_gelu(x, ad.head_vae[0], ad.head_vae[2])
    alpha = ad._get_alpha(timestep, x.device).type_as(pd)
    return alpha * pd + (1 - alpha) * pv, pd, pv


def perceiver_attention(attn, x: torch.Tensor, latents: torch.Tensor) -> torch.Tensor:
    """PerceiverAttention.forward (helpers.py:34-65) on 2-D inputs x [n, dim], latents [m, dim]; products on the native GEMM."""
    x = F.layer_norm(x, (x.shape[-1],), attn.norm_media.weight, attn.norm_media.bias, attn.norm_media.eps)
    latents = F.layer_norm(latents, (latents.shape[-1],), attn.norm_latents.weight, attn.norm_latents.bias, attn.norm_latents.eps)
    h, d = attn.heads, attn.dim_head
    q = module_linear(attn.to_q, latents)
    k, v = module_linear(attn.to_kv, torch.cat((x, latents), dim=0)).chunk(2, dim=-1)
    n = k.shape[0]
    if n % 8:                                                             # GEMM output width / reduction length: multiples of 8 (zero rows, sliced off below)
        k, v = F.pad(k, (0, 0, 0, 8 - n % 8)), F.pad(v, (0, 0, 0, 8 - n % 8))
    outs = []
    for i in range(h):
        qi, ki, vi = (t[:, i * d:(i + 1) * d] for t in (q, k, v))
        dots = linear(qi, ki.contiguous())[:, :n] * attn.scale           # einsum('i d, j d -> i j')
        dots = dots - dots.amax(dim=-1, keepdim=True).detach()
        p = dots.softmax(dim=-1)
        if n % 8:
            p = F.pad(p, (0, 8 - n % 8))
        outs.append(linear(p, vi.t().contiguous()))                       # einsum('i j, j d -> i d')
    return module_linear(attn.to_out, torch.cat(outs, dim=-1))


def resampler_forward(rs, x: torch.Tensor) -> torch.Tensor:
    """PerceiverResampler.forward (helpers.py:95-110): x [1, n, dim] -> [1, num_latents, dim]."""
    n = x.shape[1]
    xm = x[0] + rs.pos_emb.weight[:n]
    lat = rs.latents
    for attn, ff in rs.layers:
        lat = lat + perceiver_attention(attn, xm, lat)
        h = F.layer_norm(lat, (lat.shape[-1],), ff.net[0].weight, ff.net[0].bias, ff.net[0].eps)
        lat = lat + mlp_gelu(h, ff.net[1], ff.net[3])
    return F.layer_norm(lat, (lat.shape[-1],), rs.norm.weight, rs.norm.bias, rs.norm.eps).unsqueeze(0)


def needs_grad(*modules) -> bool:
    """True when autograd is on and some parameter of the given modules is trainable."""
    return torch.is_grad_enabled() and any(p.requires_grad for m in modules if m is not None for p in m.parameters())


def module_trains(m: torch.nn.Module, x: torch.Tensor) -> bool:
    """Should module `m` run under autograd for input x?  Yes when grad mode is on and either the input already carries a graph or the module
    is in training mode with trainable parameters (`pipe.freeze_except` puts exactly the trainable models in train mode, :254-259)."""
    return torch.is_grad_enabled() and (x.requires_grad or (m.training and any(p.requires_grad for p in m.parameters())))
