#!/usr/bin/env python
"""Generates tests/golden/*.pt by importing and running the REFERENCE itself (authoring container only).

The reference ships no tests, golden vectors or weights, so the oracle is pinned
against outputs of the reference's own modules on seeded synthetic weights:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

`import diffsynth` fails here (imageio / modelscope missing), so the package __init__ is skipped by
registering an empty namespace module whose __path__ points at the read-only tree, plus a stub
`modelscope`.  Nothing in tests/, bench.py or smoke() reads /root/reference at run time:
they read only the fixtures this script wrote.
"""
import importlib
import math
import os
import sys
import types

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dit_oracle as O  # noqa: E402

REF = "/root/reference/DiffSynth-Studio/diffsynth"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    pkg = types.ModuleType("diffsynth")
    pkg.__path__ = [REF]
    sys.modules["diffsynth"] = pkg
    ms = types.ModuleType("modelscope")
    ms.snapshot_download = lambda *a, **k: None
    sys.modules["modelscope"] = ms
    phys = importlib.import_module("diffsynth.pipelines.qwen_image_physical")
    dit = importlib.import_module("diffsynth.models.qwen_image_dit")
    utils = importlib.import_module("diffsynth.models.utils")
    fm = importlib.import_module("diffsynth.schedulers.flow_match")
    lora = importlib.import_module("diffsynth.lora")
    helpers = importlib.import_module("diffsynth.pipelines.helpers")
    cfg = importlib.import_module("diffsynth.configs.model_config")
    return phys, dit, utils, fm, lora, helpers, cfg


def build_ref_dit(dit_mod, num_layers, W, dtype):
    with torch.device("meta"):
        m = dit_mod.QwenImageDiT(num_layers=num_layers)
    m.load_state_dict({k: v.to(dtype).clone() for k, v in W.items()}, assign=True)   # clone: the LoRA loader writes in place
    # pos_embed tables are plain attributes built in the ctor (on meta here): rebuild them for real
    m.pos_embed = dit_mod.QwenEmbedRope(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    return m.eval()


def build_ref_adapter(helpers, A, dtype, t_min, t_max):
    ad = helpers.VisualThinkingDualAdapter(3584, 3584, t_min, t_max)
    ad.load_state_dict({k: v.to(dtype) for k, v in A.items()})
    return ad.to(dtype).eval()


@torch.no_grad()
def main():
    os.makedirs(GOLD, exist_ok=True)
    phys, dit_mod, utils, fm, lora, helpers, cfg = import_reference()
    out = {}

    # ---- scheduler (flow_match.py; ctor args qwen_image_physical.py:192) ----
    def sched():
        return fm.FlowMatchScheduler(sigma_min=0.0, sigma_max=1.0, extra_one_step=True, exponential_shift=True,
                                     exponential_shift_mu=0.8, shift_terminal=0.02)
    s = sched()
    g = {"default_t_min": s.timesteps.min().item(), "default_t_max": s.timesteps.max().item(), "cases": {}}
    for (h, w, n) in ((256, 256, 4), (1024, 1024, 50), (2048, 2048, 30), (512, 512, 40), (1536, 1536, 40), (480, 832, 40), (1024, 1024, 40)):
        s = sched()
        s.set_timesteps(n, dynamic_shift_len=(h // 16) * (w // 16))
        lat = torch.linspace(-1, 1, 64).bfloat16()
        vel = torch.linspace(2, -2, 64).bfloat16()
        steps = torch.stack([s.step(vel, s.timesteps[i], lat) for i in range(n)])
        g["cases"][f"{h}x{w}_{n}"] = dict(sigmas=s.sigmas.clone(), timesteps=s.timesteps.clone(), mu=s.calculate_shift((h // 16) * (w // 16)),
                                           bf16_timesteps=s.timesteps.to(torch.bfloat16), step_out=steps)
    s = sched()
    s.set_timesteps(1000, training=True)
    g["training"] = dict(timesteps=s.timesteps.clone(), weights=s.linear_timesteps_weights.clone(),
                         add_noise=s.add_noise(torch.ones(4), torch.full((4,), 3.0), s.timesteps[123]))
    out["scheduler"] = g

    # ---- timestep embedding incl. bf16 quirks (models/utils.py:189-216) ----
    te = utils.TimestepEmbeddings(256, 3072, diffusers_compatible_format=True, scale=1000, align_dtype_to_timestep=True)
    tcases = {}
    for t in (1000.0, 989.7009, 979.1915, 744.611382484436, 500.0, 426.6734719276428, 20.0):
        tb = torch.tensor([t]).to(torch.bfloat16)
        tcases[t] = dict(bf16_t=tb.clone(), ts_bf16=(tb / 1000).clone(), sinus_bf16=te.time_proj(tb / 1000).clone(),
                         sinus_fp32=te.time_proj(torch.tensor([t]) / 1000).clone())
    out["timestep"] = tcases

    # ---- RoPE tables ----
    rope = dit_mod.QwenEmbedRope(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    rc = {}
    for shapes, T in (([(1, 8, 8), (1, 8, 8)], 80), ([(1, 6, 10), (1, 4, 4)], 33), ([(1, 64, 64), (1, 64, 64)], 512)):
        rope.rope_cache = {}
        vf, tf = rope(shapes, [T], device="cpu")
        key = "_".join(f"{a}x{b}x{c}" for a, b, c in shapes) + f"_T{T}"
        if vf.shape[0] > 256:
            idx = torch.arange(0, vf.shape[0], 37)
            rc[key] = dict(shapes=shapes, T=T, vid_idx=idx, vid=vf[idx].clone(), txt=tf[::7].clone(), txt_stride=7, n_vid=vf.shape[0])
        else:
            rc[key] = dict(shapes=shapes, T=T, vid=vf.clone(), txt=tf.clone(), n_vid=vf.shape[0])
    out["rope"] = rc

    # ---- registry hash (configs/model_config.py:21) + hash function ----
    with torch.device("meta"):
        full = dit_mod.QwenImageDiT()
    out["dit_hash"] = dict(hash=utils.hash_state_dict_keys(full.state_dict(), with_shape=True),
                           registry=[r[1] for r in cfg.model_loader_configs if "qwen_image_dit" in r[2]],
                           n_tensors=len(full.state_dict()), n_params=sum(v.numel() for v in full.state_dict().values()))

    # ---- LoRA key mapping + fold (lora/__init__.py) ----
    gen = torch.Generator().manual_seed(5)
    lsd = {}
    for name, (o, i) in (("transformer_blocks.0.attn.to_q", (3072, 3072)), ("transformer_blocks.0.img_mlp.net.2", (3072, 12288)),
                         ("transformer_blocks.0.img_mod.1", (18432, 3072))):
        lsd[f"{name}.lora_A.default.weight"] = torch.randn(16, i, generator=gen) * 0.02
        lsd[f"{name}.lora_B.default.weight"] = torch.randn(o, 16, generator=gen) * 0.02
    lsd["diffusion_model.transformer_blocks.0.attn.to_k.lora_A.weight"] = torch.randn(16, 3072, generator=gen) * 0.02
    lsd["diffusion_model.transformer_blocks.0.attn.to_k.lora_B.weight"] = torch.randn(3072, 16, generator=gen) * 0.02
    loader = lora.GeneralLoRALoader(device="cpu", torch_dtype=torch.bfloat16)
    out["lora"] = dict(name_dict=loader.get_name_dict(lsd))

    # ---- DiT forward goldens: 2 blocks, 128x128 + 128x128 edit, T=80 ----
    NL, H, Wd, T = 2, 128, 128, 80
    Wts = O.synth_weights(O.dit_param_shapes(NL), seed=1)
    Ats = O.synth_weights(O.adapter_param_shapes(), seed=2)
    t_min, t_max = out["scheduler"]["default_t_min"], out["scheduler"]["default_t_max"]
    fw = {}
    for dtype, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        # weights are bf16-representable in both runs so fp32-vs-bf16 isolates the arithmetic
        Wq = {k: v.to(torch.bfloat16).to(dtype) for k, v in Wts.items()}
        Aq = {k: v.to(torch.bfloat16).to(dtype) for k, v in Ats.items()}
        m = build_ref_dit(dit_mod, NL, Wq, dtype)
        ad = build_ref_adapter(helpers, Aq, dtype, t_min, t_max)
        if tag == "bf16":
            lw = {k: v for k, v in lsd.items() if "diffusion_model" not in k}
            m_l = build_ref_dit(dit_mod, NL, Wq, dtype)
            loader.load(m_l, lw, alpha=1.0)
            fw["lora_folded_to_q_bf16"] = m_l.transformer_blocks[0].attn.to_q.weight.detach()[:64, :64].clone()
            fw["lora_folded_mlp2_bf16"] = m_l.transformer_blocks[0].img_mlp.net[2].weight.detach()[:64, :64].clone()
        inp = O.synth_inputs(H, Wd, T, seed=3, dtype=torch.bfloat16)
        inp = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in inp.items()}
        pe = inp["prompt_emb"].clone()
        outs = []
        for tval in (744.611382484436, 426.6734719276428):       # two consecutive calls: in-place compounding
            t = torch.tensor([tval]).to(torch.bfloat16).to(dtype)  # the loop casts the fp32 table entry to pipe dtype (:649)
            y, loss = phys.model_fn_qwen_image(dit=m, visual_thinking_adapter=ad, latents=inp["latents"], timestep=t, prompt_emb=pe,
                                               prompt_emb_mask=inp["prompt_emb_mask"], special_token_mask=inp["special_token_mask"],
                                               height=H, width=Wd, edit_latents=inp["edit_latents"], is_train=False)
            outs.append(y.clone())
        fw[tag] = dict(out=outs, prompt_emb_after=pe[:, ::4, ::8].clone(), special_after=pe[inp["special_token_mask"]][:, ::8].clone(), special_abs_mean=pe[inp["special_token_mask"]].float().abs().mean().item())
        if tag == "bf16":
            # per-op goldens of one block on the bf16 path
            t = torch.tensor([500.0]).to(dtype)
            temb = m.time_text_embed(t / 1000, dtype)
            image = m.img_in(torch.cat([O.patchify(inp["latents"]), O.patchify(inp["edit_latents"])], dim=1))
            text = m.txt_in(m.txt_norm(inp["prompt_emb"]))
            ropes = m.pos_embed([(1, 8, 8), (1, 8, 8)], [T], device="cpu")
            t1, i1 = m.transformer_blocks[0](image=image, text=text, temb=temb, image_rotary_emb=ropes)
            fw["block0_bf16"] = dict(temb=temb.clone(), image_in=image[..., ::4].clone(), text_in=text[..., ::4].clone(), text_out=t1[..., ::4].clone(), image_out=i1[..., ::4].clone())
            xa = inp["prompt_emb"][inp["special_token_mask"]].view(1, -1, 3584)
            mixed, pd, pv = ad(xa, torch.tensor([744.611382484436]).to(dtype))
            fw["adapter_bf16"] = dict(mixed=mixed[..., ::4].clone(), pred_dino=pd[..., ::4].clone(), pred_vae=pv[..., ::4].clone(),
                                      loss=ad.get_loss(pd, pv, pd * 0.5, pv * 0.25, torch.tensor([744.611382484436]).to(dtype)).item())
    fw["meta"] = dict(num_layers=NL, height=H, width=Wd, T=T, w_seed=1, a_seed=2, in_seed=3, t_min=t_min, t_max=t_max)
    out["forward"] = fw

    # ---- tiny end-to-end loop: 4 steps, CFG 4.0, 1 block, 64x64 + edit 64x64 (config #1 plumbing at toy size) ----
    NL2, H2, T2 = 1, 64, 72
    W1 = {k: v.to(torch.bfloat16).float() for k, v in O.synth_weights(O.dit_param_shapes(NL2), seed=4).items()}
    A1 = {k: v.to(torch.bfloat16).float() for k, v in Ats.items()}
    m = build_ref_dit(dit_mod, NL2, W1, torch.float32)
    ad = build_ref_adapter(helpers, A1, torch.float32, t_min, t_max)
    posi = O.synth_inputs(H2, H2, T2, seed=6)
    nega = O.synth_inputs(H2, H2, T2 - 3, seed=7)
    s = sched()
    s.set_timesteps(4, dynamic_shift_len=(H2 // 16) ** 2)
    lat = posi["latents"].clone()
    pe_p, pe_n = posi["prompt_emb"].clone(), nega["prompt_emb"].clone()
    for pid, t in enumerate(s.timesteps):
        tt = t.unsqueeze(0).to(torch.float32)
        kw = dict(dit=m, visual_thinking_adapter=ad, latents=lat, timestep=tt, height=H2, width=H2, edit_latents=posi["edit_latents"], is_train=False)
        vp, _ = phys.model_fn_qwen_image(prompt_emb=pe_p, prompt_emb_mask=posi["prompt_emb_mask"], special_token_mask=posi["special_token_mask"], **kw)
        vn, _ = phys.model_fn_qwen_image(prompt_emb=pe_n, prompt_emb_mask=nega["prompt_emb_mask"], special_token_mask=nega["special_token_mask"], **kw)
        v = vn + 4.0 * (vp - vn)
        lat = s.step(v, s.timesteps[pid], lat)
    out["loop"] = dict(meta=dict(num_layers=NL2, height=H2, T_posi=T2, T_nega=T2 - 3, steps=4, w_seed=4, a_seed=2, posi_seed=6, nega_seed=7),
                       latents=lat.clone(), prompt_emb_posi_after=pe_p[posi["special_token_mask"]][:, ::8].clone())

    # ---- training-path feature extractors: reference PerceiverResampler / VisualThinkingAdapter and HF DINOv2 ----
    from oracle import aux_oracle as AO
    import tempfile
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    dino_mod = importlib.import_module("diffsynth.pipelines.dinov2")
    P = AO.aux_synth(seed=11)
    ain = AO.aux_inputs(seed=12)
    cfgd = Dinov2WithRegistersConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, mlp_ratio=4, patch_size=14, image_size=518,
                                     num_register_tokens=4)
    with tempfile.TemporaryDirectory() as td:
        hf = Dinov2WithRegistersModel(cfgd)
        missing = hf.load_state_dict(P["dinov2"], strict=True)
        hf.save_pretrained(td)
        dn = dino_mod.Dinov2withNorm(dinov2_path=td).eval()          # the reference wrapper around the HF model
    dino_src = dn(ain["dino_source"])
    dino_mid = dn(ain["dino_middle"])
    def ref_resampler(dim, max_tok, W):
        m = helpers.PerceiverResampler(dim=dim, num_latents=64, depth=2, max_num_media_tokens=max_tok)
        m.load_state_dict(W)
        return m.eval()
    dr, vr = ref_resampler(768, 4096, P["dino_resampler"]), ref_resampler(64, 10240, P["vae_resampler"])
    da = helpers.VisualThinkingAdapter(768, 3584); da.load_state_dict(P["dino_resampler_adapter"])
    va = helpers.VisualThinkingAdapter(64, 3584); va.load_state_dict(P["vae_resampler_adapter"])
    hs_mid = dino_mid + P["dino_time_embed"]["weight"][:3].unsqueeze(1)
    r_mid = dr(hs_mid.reshape(1, -1, 768))
    emb_dino = da(r_mid) - da(dr(dino_src.reshape(1, -1, 768)))
    tok_mid = O.patchify(ain["vae_middle_latents"]) + P["vae_time_embed"]["weight"][:3].unsqueeze(1)
    emb_vae = va(vr(tok_mid.reshape(1, -1, 64))) - va(vr(O.patchify(ain["vae_source_latents"]).reshape(1, -1, 64)))
    out["aux"] = dict(seed=11, in_seed=12, dino_source=dino_src[:, ::4, ::8].clone(), dino_middle=dino_mid[:, ::4, ::8].clone(),
                      resampler_dino_mid=r_mid[..., ::4].clone(), pseudo_special_emb_dino=emb_dino[..., ::8].clone(),
                      pseudo_special_emb_vae=emb_vae[..., ::8].clone(), transformers=__import__("transformers").__version__)

    for k, v in out.items():
        torch.save(v, os.path.join(GOLD, f"{k}.pt"))
        print(k, os.path.getsize(os.path.join(GOLD, f"{k}.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
