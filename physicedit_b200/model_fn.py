"""The per-forward step function of PhysicEdit, native.

Drop-in for `model_fn_qwen_image` (DiffSynth-Studio/diffsynth/pipelines/qwen_image_physical.py:1302-1403):
same keyword signature, same return `(latents, special_token_loss)`, same in-place mutation of
`prompt_emb` (the adapter output is written back into the caller's tensor, :1336, so step k sees the
output of step k-1), same bf16 timestep bookkeeping (`t -> bf16 -> /1000 -> bf16`, bf16 frequencies).
The pipeline installs it as `pipe.model_fn`.  Everything numeric runs in libpe_b200 kernels.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import native as nv

def _txt_len(mask: Optional[torch.Tensor], T: int) -> int:
    """`prompt_emb_mask.sum(dim=1).tolist()` (:1341): a device->host sync per forward in the reference.  The pipeline's loops
    read it once per request and pass `txt_len=`; a direct call without it pays the same sync the reference pays (no cache:
    a cache keyed on the tensor's address can be hit by a later request's mask at a recycled address)."""
    if mask is None:
        return T
    return int(mask.sum(dim=1).max().item())


def prompt_lengths(prompt_emb_mask: Optional[torch.Tensor], special_token_mask: Optional[torch.Tensor], T: int) -> dict:
    """The two data-dependent sizes of a request -- text length (:1341) and number of special tokens (:1334) -- read from the
    device ONCE; pass the result as `**prompt_lengths(...)` (keys `txt_len`, `n_special`) to keep the denoise loop sync-free."""
    n_sp = 0 if special_token_mask is None else int(special_token_mask.sum().item())
    return dict(txt_len=_txt_len(prompt_emb_mask, T), n_special=n_sp)


def model_fn_qwen_image(
    dit=None,
    blockwise_controlnet=None,
    visual_thinking_adapter=None,
    latents=None,
    timestep=None,
    prompt_emb=None,
    prompt_emb_mask=None,
    special_token_mask=None,
    height=None,
    width=None,
    blockwise_controlnet_conditioning=None,
    blockwise_controlnet_inputs=None,
    progress_id=0,
    num_inference_steps=1,
    entity_prompt_emb=None,
    entity_prompt_emb_mask=None,
    entity_masks=None,
    edit_latents=None,
    context_latents=None,
    enable_fp8_attention=False,
    use_gradient_checkpointing=False,
    use_gradient_checkpointing_offload=False,
    edit_rope_interpolation=False,
    is_train=True,
    pseudo_special_emb_dino=None,
    pseudo_special_emb_vae=None,
    timestep_host: Optional[float] = None,
    out: Optional[torch.Tensor] = None,
    cfg_branch: int = 0,
    txt_len: Optional[int] = None,
    n_special: Optional[int] = None,
    **kwargs,
):
    if latents.shape[0] != 1 or prompt_emb.shape[0] != 1:
        raise ValueError("the pipeline is strictly batch 1 (qwen_image_physical.py:688,821); batch edits shard one image per GPU")
    if not (latents.dtype == torch.bfloat16 and prompt_emb.dtype == torch.bfloat16):
        raise nv.NativeUnavailable(f"native model_fn needs CUDA bfloat16 tensors (latents {latents.dtype} on {latents.device}); no fallback")
    eng = dit.engine()          # raises NativeUnavailable for weights that are not bf16 on an sm_100 GPU / without libpe_b200.so: there is no CPU path
    nat = eng.nat
    t_bf16 = timestep.to(device=latents.device, dtype=torch.bfloat16).reshape(-1)[:1].contiguous()
    if timestep_host is None:
        # The reference loop hands over a CUDA bf16 timestep (:649).  Its value keys the conditioning cache (temb, the 120 modulation
        # vectors, norm_out's scale/shift depend on nothing else), so the two CFG branches of a step share one pass over the 13.6 GB of
        # modulation weights.  Reading it back is one 2-byte D2H per forward -- the reference itself syncs every forward at :1341.
        timestep_host = float(t_bf16[0].item()) if timestep.is_cuda else float(timestep.reshape(-1)[0].to(torch.bfloat16))
    T = prompt_emb.shape[1]
    if txt_len is None:
        txt_len = _txt_len(prompt_emb_mask, T)
    if txt_len != T:
        raise ValueError(f"padded prompts ({txt_len} valid of {T} tokens) are not produced by the B=1 pipeline; trim prompt_emb to its mask")
    pe2d = prompt_emb[0]
    if not pe2d.is_contiguous():
        raise ValueError("prompt_emb must be contiguous: it is updated in place")
    from . import autograd as ag
    if getattr(dit, "_lora_injected", False) or (torch.is_grad_enabled() and is_train and
                                                  (prompt_emb.requires_grad or ag.needs_grad(dit, visual_thinking_adapter,
                                                                                             blockwise_controlnet if blockwise_controlnet_conditioning is not None else None))):
        # training: un-merged LoRA, or a training call (is_train, grad mode on) with trainable parameters on the path -> the
        # differentiable path on the same GEMM / attention kernels.  Inference calls (is_train=False, or under no_grad) stay on the engine.
        if entity_prompt_emb is not None or enable_fp8_attention:
            raise NotImplementedError("EliGen / fp8 attention under autograd are not part of the PhysicEdit training path")
        after_block = None
        if blockwise_controlnet_conditioning is not None:               # a trainable (or merely present) blockwise controlnet on the training path
            conds = ag.controlnet_conditionings(blockwise_controlnet, blockwise_controlnet_inputs, blockwise_controlnet_conditioning)
            after_block = lambda block_id, noise_tokens: ag.controlnet_residual(blockwise_controlnet, noise_tokens, conds, blockwise_controlnet_inputs,
                                                                                progress_id, num_inference_steps, block_id)
        pred, special_token_loss = _model_fn_autograd(dit, visual_thinking_adapter, latents, timestep, t_bf16, prompt_emb, special_token_mask, edit_latents,
                                                      context_latents, use_gradient_checkpointing, is_train, pseudo_special_emb_dino, pseudo_special_emb_vae,
                                                      bool(edit_rope_interpolation), after_block)
        if out is not None:
            # the pipeline's denoise loop reads the prediction from the buffer it passed (run_cfg_branches): an evaluation in the middle of a
            # training run (un-merged LoRA still injected, scripts/train/train_physicedit.py:39-169) comes through here
            with torch.no_grad():
                out.copy_(pred)
        return pred, special_token_loss

    special_token_loss = 0
    if special_token_mask is not None:
        ad = visual_thinking_adapter
        # any number of special tokens, like the reference's boolean gather (:1334); a count the caller did not pass is read from the
        # device (the gather kernel raises the handle's async error if the mask holds MORE rows than `n_special`: never a silent drop)
        n_sp = int(special_token_mask.sum().item()) if n_special is None else int(n_special)
        if n_sp > 0:
            gathered = torch.empty(n_sp, pe2d.shape[1], dtype=torch.bfloat16, device=pe2d.device)
            idx = torch.empty(n_sp + 1, dtype=torch.int32, device=pe2d.device)
            nat.special_gather(pe2d, special_token_mask[0].view(torch.uint8), gathered, idx)
            pred_dino, pred_vae = ad.heads(gathered)
            nat.special_blend_scatter(pe2d, idx, pred_dino, pred_vae, t_bf16, ad.t_min, ad.t_max)       # in place (:1336)
            if is_train:
                special_token_loss = ad.get_loss(pred_dino.unsqueeze(0), pred_vae.unsqueeze(0), pseudo_special_emb_dino, pseudo_special_emb_vae, timestep)

    lat_list = [latents]
    if context_latents is not None:
        lat_list.append(context_latents)
    if edit_latents is not None:
        lat_list += list(edit_latents) if isinstance(edit_latents, list) else [edit_latents]
    lat_list = [l.contiguous() for l in lat_list]
    if out is None:
        out = torch.empty_like(latents)
    text_segments, attn_mask = None, None
    if entity_prompt_emb is not None:
        # EliGen (:1360-1364; QwenImageDiT.process_entity_masks, qwen_image_dit.py:433-498): the text stream becomes [entity prompts ..., global
        # prompt], every prompt with its own positions, under a mask that ties each entity prompt to its image region
        ents = [e[0] for e in entity_prompt_emb]
        if entity_prompt_emb_mask is not None and any(int(m.shape[1]) != e.shape[0] for m, e in zip(entity_prompt_emb_mask, ents)):
            raise ValueError("entity prompt masks must match their embeddings (batch 1: unpadded)")
        text_segments = [int(e.shape[0]) for e in ents] + [T]
        attn_mask = entity_attention_mask(entity_masks, text_segments, lat_list)
        pe2d = torch.cat(ents + [pe2d], dim=0).contiguous()
    after_block = None
    if blockwise_controlnet_conditioning is not None:
        # :1372-1374, :1389-1396: control latents -> tokens -> img_in once per forward; after every block the noise tokens get the scaled sum
        conds = blockwise_controlnet.preprocess(blockwise_controlnet_inputs, blockwise_controlnet_conditioning)

        def after_block(block_id, noise_tokens):
            blockwise_controlnet.apply_(noise_tokens, conds, blockwise_controlnet_inputs, progress_id, num_inference_steps, block_id)
    eng.forward(lat_list, t_bf16, pe2d, out, t_key=timestep_host, branch=cfg_branch, rope_sampling=bool(edit_rope_interpolation), after_block=after_block,
                text_segments=text_segments, attn_mask=attn_mask, fp8_attention=bool(enable_fp8_attention) and attn_mask is None)
    return out, special_token_loss


def entity_attention_mask(entity_masks: torch.Tensor, text_segments, lat_list) -> torch.Tensor:
    """The boolean form (uint8, 1 = visible) of process_entity_masks' attention mask (qwen_image_dit.py:447-496) in the joint [text; image] order:
    entity prompt i sees, and is seen by, the image tokens whose 2 x 2 latent patch touches entity mask i -- in EVERY image of the sequence (the
    mask is repeated over the edit images, which therefore must have the noise image's token count); the global prompt sees everything; different
    prompts never see each other.  entity_masks [1, N, 1, h8, w8] (what QwenImageUnit_EntityControl produces)."""
    dev = entity_masks.device
    n0 = (lat_list[0].shape[-2] // 2) * (lat_list[0].shape[-1] // 2)
    S_img = sum((l.shape[-2] // 2) * (l.shape[-1] // 2) for l in lat_list)
    if S_img % n0:
        raise ValueError("EliGen needs edit / context images with the noise image's latent size (the reference repeats the entity mask over them)")
    N = entity_masks.shape[1] + 1
    if len(text_segments) != N:
        raise ValueError(f"{N - 1} entity masks for {len(text_segments) - 1} entity prompts")
    patched = torch.nn.functional.max_pool2d(entity_masks[0, :, 0].float().unsqueeze(0), 2)[0].flatten(1) > 0       # [N - 1, n0]
    patched = torch.cat([patched, torch.ones(1, n0, dtype=torch.bool, device=dev)], dim=0).repeat(1, S_img // n0)     # + the global prompt
    T_all = sum(text_segments)
    S = T_all + S_img
    seg_of_row = torch.repeat_interleave(torch.arange(N, device=dev), torch.tensor(text_segments, device=dev))        # [T_all]
    mask = torch.ones(S, S, dtype=torch.uint8, device=dev)
    txt_img = patched[seg_of_row].to(torch.uint8)                                                                     # [T_all, S_img]
    mask[:T_all, T_all:] = txt_img
    mask[T_all:, :T_all] = txt_img.t()
    mask[:T_all, :T_all] = (seg_of_row[:, None] == seg_of_row[None, :]).to(torch.uint8)
    return mask


def _model_fn_autograd(dit, ad, latents, timestep, t_bf16, prompt_emb, special_token_mask, edit_latents, context_latents, use_gradient_checkpointing,
                       is_train, pseudo_special_emb_dino, pseudo_special_emb_vae, rope_sampling=False, after_block=None):
    """:1331-1403 under autograd (physicedit_b200/autograd.py): same in-place write of the adapter output into `prompt_emb` (:1336), same
    `(latents, special_token_loss)` return; gradients reach the LoRA factors, the adapter heads and whatever produced the pseudo targets."""
    from . import autograd as ag
    special_token_loss = 0
    if special_token_mask is not None and bool(special_token_mask.any()):
        special = prompt_emb[special_token_mask].view(prompt_emb.shape[0], -1, prompt_emb.shape[-1])
        mixed, pred_dino, pred_vae = ag.dual_adapter_forward(ad, special, timestep)
        prompt_emb[special_token_mask] = mixed.reshape(-1, prompt_emb.shape[-1])
        if is_train:
            special_token_loss = ad.get_loss(pred_dino, pred_vae, pseudo_special_emb_dino, pseudo_special_emb_vae, timestep)
    lat_list = [latents]
    if context_latents is not None:
        lat_list.append(context_latents)
    if edit_latents is not None:
        lat_list += list(edit_latents) if isinstance(edit_latents, list) else [edit_latents]
    out = ag.dit_forward(dit, [l.contiguous() for l in lat_list], t_bf16, prompt_emb, use_gradient_checkpointing=use_gradient_checkpointing,
                         rope_sampling=rope_sampling, after_block=after_block)
    return out, special_token_loss
