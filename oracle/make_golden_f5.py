#!/usr/bin/env python
"""Generates tests/golden/f5.pt: outputs of the REFERENCE's own modules for the SURVEY 8f5 options (authoring container only, like
oracle/make_golden.py):   PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_f5.py
  rope_sampling : QwenEmbedRope.forward_sampling (models/qwen_image_dit.py:168-225), fresh cache per case
  eligen        : QwenImageDiT.process_entity_masks (models/qwen_image_dit.py:433-498) + one block with the mask, 1-block model, fp32
  controlnet    : QwenImageBlockWiseControlNet (models/qwen_image_controlnet.py) + QwenImageBlockwiseMultiControlNet
                  (pipelines/qwen_image_physical.py:157-180) on seeded synthetic weights (oracle.dit_oracle.synth_weights), fp32 and bf16
"""
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

out = {}
with ref_import.ReferenceModules() as ref:
    rs = {}
    for shapes, T in (([(1, 8, 6), (1, 12, 10)], 40), ([(1, 8, 8), (1, 8, 8)], 33), ([(1, 10, 12), (1, 5, 7), (1, 16, 16)], 21), ([(1, 64, 64), (1, 52, 80)], 64)):
        rope = ref.dit.QwenEmbedRope(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
        vf, tf = rope.forward_sampling(shapes, [T], device="cpu")
        key = "_".join(f"{a}x{b}x{c}" for a, b, c in shapes) + f"_T{T}"
        idx = torch.arange(0, vf.shape[0], 37 if vf.shape[0] > 512 else 1)
        rs[key] = dict(shapes=shapes, T=T, n_vid=vf.shape[0], vid_idx=idx, vid=vf[idx].clone(), txt=tf.clone())
    out["rope_sampling"] = rs

    # ---- blockwise controlnet ----
    from oracle import dit_oracle as O
    import importlib
    cn_mod = importlib.import_module("diffsynth.models.qwen_image_controlnet")
    flux = importlib.import_module("diffsynth.pipelines.flux_image_new")
    L, n = 2, 16
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, n, 3072, generator=g)
    lat = [torch.randn(1, 16, 8, 8, generator=g), torch.randn(1, 16, 8, 8, generator=g)]
    cn = {}
    for dtype in (torch.float32, torch.bfloat16):
        nets = []
        for seed in (61, 62):
            m = cn_mod.QwenImageBlockWiseControlNet(num_layers=L)
            m.load_state_dict(O.synth_weights(O.controlnet_param_shapes(L), seed=seed))
            nets.append(m.to(dtype).eval())
        multi = ref.phys.QwenImageBlockwiseMultiControlNet(nets)
        inputs = [flux.ControlNetInput(controlnet_id=0, scale=1.0, start=1.0, end=0.0), flux.ControlNetInput(controlnet_id=1, scale=0.5, start=0.8, end=0.3)]
        with torch.no_grad():
            conds = multi.preprocess(inputs, [l.to(dtype) for l in lat])
            res = {}
            for pid in (0, 2, 4):                       # of 5 steps: progress 1.0 (only net 0), 0.5 (both), 0.0 (only net 0)
                for blk in range(L):
                    res[(pid, blk)] = multi.blockwise_forward(image=x.to(dtype), conditionings=conds, controlnet_inputs=inputs, progress_id=pid,
                                                              num_inference_steps=5, block_id=blk)[:, :, ::8].clone()
        cn[str(dtype)] = dict(cond0=conds[0][:, :, ::8].clone(), res=res)
    cn["meta"] = dict(L=L, n=n, seeds=(61, 62), x=x, latents=lat, stride=8)
    out["controlnet"] = cn

    # ---- EliGen: the reference's model_fn with entity prompts / masks, 1 block, 64 x 64 image + edit image, fp32 ----
    H = Wd = 64
    Wsd = O.synth_weights(O.dit_param_shapes(1), seed=71)
    rdit = ref_import.build_reference_dit(ref, Wsd, 1, torch.float32, "cpu")
    inp = O.synth_inputs(H, Wd, 24, seed=72, dtype=torch.float32, n_special=8)
    g = torch.Generator().manual_seed(73)
    ent_emb = [3 * torch.randn(1, n, 3584, generator=g) for n in (9, 14)]
    ent_mask = [torch.ones(1, n, dtype=torch.long) for n in (9, 14)]
    masks = torch.zeros(1, 2, 1, H // 8, Wd // 8)
    masks[0, 0, 0, :4, :5] = 1
    masks[0, 1, 0, 3:, 2:] = 1
    t = torch.tensor([500.0])
    with torch.no_grad():
        y, _ = ref.phys.model_fn_qwen_image(dit=rdit, latents=inp["latents"], timestep=t, prompt_emb=inp["prompt_emb"].clone(), prompt_emb_mask=inp["prompt_emb_mask"],
                                            special_token_mask=None, height=H, width=Wd, edit_latents=inp["edit_latents"], entity_prompt_emb=ent_emb,
                                            entity_prompt_emb_mask=ent_mask, entity_masks=masks, is_train=False)
        y_plain, _ = ref.phys.model_fn_qwen_image(dit=rdit, latents=inp["latents"], timestep=t, prompt_emb=inp["prompt_emb"].clone(),
                                                  prompt_emb_mask=inp["prompt_emb_mask"], special_token_mask=None, height=H, width=Wd,
                                                  edit_latents=inp["edit_latents"], is_train=False)
    out["eligen"] = dict(meta=dict(H=H, W=Wd, T=24, w_seed=71, in_seed=72, n_special=8, t=500.0), entity_prompt_emb=ent_emb, entity_masks=masks, y=y.clone(),
                         differs_from_plain=float((y - y_plain).norm() / y_plain.norm()))
torch.save(out, os.path.join(ROOT, "tests", "golden", "f5.pt"))
print({k: list(v) for k, v in out.items()}, os.path.getsize(os.path.join(ROOT, "tests", "golden", "f5.pt")))
