"""Pins the oracle (oracle/dit_oracle.py) against fixtures produced by the reference itself
(oracle/make_golden.py, run in the authoring container).  CPU only."""
import math

import pytest
import torch

from oracle import dit_oracle as O


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item()


def test_scheduler_tables_bit_exact(golden):
    g = golden("scheduler")
    s = O.FlowMatchOracle()
    assert s.timesteps.min().item() == g["default_t_min"] == 19.999980926513672
    assert s.timesteps.max().item() == g["default_t_max"] == 1000.0
    for key, c in g["cases"].items():
        hw, n = key.split("_")
        h, w = map(int, hw.split("x"))
        n = int(n)
        s.set_timesteps(n, dynamic_shift_len=(h // 16) * (w // 16))
        assert torch.equal(s.sigmas, c["sigmas"]), key
        assert torch.equal(s.timesteps, c["timesteps"]), key
        assert s.calculate_shift((h // 16) * (w // 16)) == c["mu"]
        assert torch.equal(s.timesteps.to(torch.bfloat16), c["bf16_timesteps"])
        # bf16-rounded timesteps stay unique, so argmin bookkeeping is unambiguous
        assert len(set(c["bf16_timesteps"].tolist())) == n
        lat = torch.linspace(-1, 1, 64).bfloat16()
        vel = torch.linspace(2, -2, 64).bfloat16()
        steps = torch.stack([s.step(vel, i, lat) for i in range(n)])
        assert torch.equal(steps, c["step_out"]), key
    s.set_timesteps(1000, training=True)
    t = g["training"]
    assert torch.equal(s.timesteps, t["timesteps"]) and torch.equal(s.linear_timesteps_weights, t["weights"])
    assert torch.equal(s.add_noise(torch.ones(4), torch.full((4,), 3.0), s.timesteps[123]), t["add_noise"])
    assert abs(t["weights"].sum().item() - 1000) < 1e-2


def test_known_scheduler_values(golden):
    """SURVEY 8c known answers."""
    s = O.FlowMatchOracle()
    s.set_timesteps(4, dynamic_shift_len=256)
    assert s.sigmas.tolist() == [1.0, 0.744611382484436, 0.4266734719276428, 0.019999980926513672]
    assert s.timesteps.to(torch.bfloat16).tolist() == [1000.0, 744.0, 426.0, 20.0]
    for hw, mu in ((256, 0.5), (512, 0.538710), (1024, 0.693548), (1536, 0.951613), (2048, 1.312903)):
        assert abs(s.calculate_shift((hw // 16) ** 2) - mu) < 1e-6


def test_timestep_embedding_bit_exact(golden):
    for t, c in golden("timestep").items():
        tb = torch.tensor([t]).to(torch.bfloat16)
        assert torch.equal(tb, c["bf16_t"])
        ts = tb / 1000
        assert torch.equal(ts, c["ts_bf16"])
        assert torch.equal(O.timestep_sinusoid(ts), c["sinus_bf16"])
        assert torch.equal(O.timestep_sinusoid(torch.tensor([t]) / 1000), c["sinus_fp32"])


def test_cuda_scalar_division_matches_cpu_on_every_schedule_timestep(golden):
    """ATen's CUDA `tensor / scalar` multiplies by float(1/scalar); for every bf16 timestep the schedules
    produce, that agrees bit-for-bit with the CPU's true division (so CPU goldens are valid for the GPU path)."""
    g = golden("scheduler")
    vals = torch.cat([c["bf16_timesteps"] for c in g["cases"].values()] + [g["training"]["timesteps"].to(torch.bfloat16)]).unique()
    a = O._div_scalar(vals, 1000, False)
    b = O._div_scalar(vals, 1000, True)
    assert torch.equal(a, b)
    al_a = O.adapter_alpha(vals, g["default_t_min"], g["default_t_max"], False)
    al_b = O.adapter_alpha(vals, g["default_t_min"], g["default_t_max"], True)
    assert torch.equal(al_a, al_b)


def test_rope_tables_bit_exact(golden):
    for key, c in golden("rope").items():
        vid, txt = O.rope_tables([tuple(s) for s in c["shapes"]], c["T"])
        assert vid.shape[0] == c["n_vid"]
        if "vid_idx" in c:
            assert torch.equal(vid[c["vid_idx"]], c["vid"])
            assert torch.equal(txt[:: c["txt_stride"]], c["txt"])
        else:
            assert torch.equal(vid, c["vid"]) and torch.equal(txt, c["txt"])


def test_state_dict_hash_matches_registry(golden):
    g = golden("dit_hash")
    shapes = O.dit_param_shapes(60)
    assert len(shapes) == g["n_tensors"] == 1933
    assert sum(math.prod(s) for s in shapes.values()) == g["n_params"] == 20430401088
    assert O.state_dict_key_hash(shapes) == g["hash"] == "0319a1cb19835fb510907dd3367c95ff"
    assert g["hash"] in g["registry"]


def test_lora_name_dict(golden):
    g = golden("lora")["name_dict"]
    fake = {}
    for tgt, (kb, ka) in g.items():
        fake[kb] = fake[ka] = None
    assert O.lora_name_dict(fake) == g
    assert "transformer_blocks.0.attn.to_k" in g          # 'diffusion_model.' prefix stripped, no adapter-name segment


@pytest.fixture(scope="module")
def fwd(golden):
    return golden("forward")


def _weights(meta, dtype):
    W = {k: v.to(torch.bfloat16).to(dtype) for k, v in O.synth_weights(O.dit_param_shapes(meta["num_layers"]), seed=meta["w_seed"]).items()}
    A = {k: v.to(torch.bfloat16).to(dtype) for k, v in O.synth_weights(O.adapter_param_shapes(), seed=meta["a_seed"]).items()}
    return W, A


@pytest.mark.parametrize("tag,dtype", [("fp32", torch.float32), ("bf16", torch.bfloat16)])
def test_model_fn_matches_reference(fwd, tag, dtype):
    meta = fwd["meta"]
    W, A = _weights(meta, dtype)
    inp = O.synth_inputs(meta["height"], meta["width"], meta["T"], seed=meta["in_seed"], dtype=torch.bfloat16)
    inp = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in inp.items()}
    pe = inp["prompt_emb"].clone()
    for call, tval in enumerate((744.611382484436, 426.6734719276428)):
        t = torch.tensor([tval]).to(torch.bfloat16).to(dtype)
        y = O.model_fn(W, A, inp["latents"], t, pe, inp["prompt_emb_mask"], inp["special_token_mask"], meta["height"], meta["width"],
                       edit_latents=inp["edit_latents"], t_min=meta["t_min"], t_max=meta["t_max"])
        ref = fwd[tag]["out"][call]
        if dtype == torch.float32:
            assert rel_l2(y, ref) < 2e-6, (tag, call)
        else:
            assert torch.equal(y, ref), (tag, call)      # same torch ops in the same order: bit-identical on the same CPU
    # in-place compounding of the special tokens: mutated rows match, others untouched
    assert rel_l2(pe[:, ::4, ::8], fwd[tag]["prompt_emb_after"]) < 2e-6
    assert rel_l2(pe[inp["special_token_mask"]][:, ::8], fwd[tag]["special_after"]) < 2e-6
    assert torch.equal(pe[~inp["special_token_mask"]], inp["prompt_emb"][~inp["special_token_mask"]])
    assert not torch.equal(pe[inp["special_token_mask"]], inp["prompt_emb"][inp["special_token_mask"]])


def test_reference_bf16_noise_floor(fwd):
    """The reference's own bf16 forward differs from its fp32 forward at the 1e-2 level: this is the
    noise floor the CUDA path's parity tolerance is defined against."""
    e = rel_l2(fwd["bf16"]["out"][0], fwd["fp32"]["out"][0])
    assert 1e-3 < e < 3e-2


def test_block_and_adapter_goldens_bf16(fwd):
    meta = fwd["meta"]
    W, A = _weights(meta, torch.bfloat16)
    inp = O.synth_inputs(meta["height"], meta["width"], meta["T"], seed=meta["in_seed"], dtype=torch.bfloat16)
    g = fwd["block0_bf16"]
    t = torch.tensor([500.0]).to(torch.bfloat16)
    temb = O.time_text_embed(W, t / 1000, torch.bfloat16)
    assert torch.equal(temb, g["temb"])
    image = O.linear(torch.cat([O.patchify(inp["latents"]), O.patchify(inp["edit_latents"])], dim=1), W, "img_in")
    text = O.linear(O.rmsnorm(inp["prompt_emb"], W["txt_norm.weight"]), W, "txt_in")
    assert torch.equal(image[..., ::4], g["image_in"]) and torch.equal(text[..., ::4], g["text_in"])
    rope = O.rope_tables([(1, 8, 8), (1, 8, 8)], meta["T"])
    t1, i1 = O.block_forward(W, 0, image, text, temb, rope)
    assert torch.equal(t1[..., ::4], g["text_out"]) and torch.equal(i1[..., ::4], g["image_out"])
    ga = fwd["adapter_bf16"]
    xa = inp["prompt_emb"][inp["special_token_mask"]].view(1, -1, 3584)
    tt = torch.tensor([744.611382484436]).to(torch.bfloat16)
    mixed, pd, pv = O.dual_adapter(A, xa, tt, meta["t_min"], meta["t_max"])
    assert torch.equal(mixed[..., ::4], ga["mixed"]) and torch.equal(pd[..., ::4], ga["pred_dino"]) and torch.equal(pv[..., ::4], ga["pred_vae"])
    loss = O.adapter_loss(pd, pv, pd * 0.5, pv * 0.25, tt, meta["t_min"], meta["t_max"]).item()
    assert abs(loss - ga["loss"]) <= 1e-6 * abs(ga["loss"])


def test_lora_fold_bf16(fwd, golden):
    meta = fwd["meta"]
    W, _ = _weights(meta, torch.bfloat16)
    gen = torch.Generator().manual_seed(5)
    lsd = {}
    for name, (o, i) in (("transformer_blocks.0.attn.to_q", (3072, 3072)), ("transformer_blocks.0.img_mlp.net.2", (3072, 12288)),
                         ("transformer_blocks.0.img_mod.1", (18432, 3072))):
        lsd[f"{name}.lora_A.default.weight"] = torch.randn(16, i, generator=gen) * 0.02
        lsd[f"{name}.lora_B.default.weight"] = torch.randn(o, 16, generator=gen) * 0.02
    assert O.lora_fold(W, lsd, 1.0, torch.bfloat16) == 3
    assert torch.equal(W["transformer_blocks.0.attn.to_q.weight"][:64, :64], fwd["lora_folded_to_q_bf16"])
    assert torch.equal(W["transformer_blocks.0.img_mlp.net.2.weight"][:64, :64], fwd["lora_folded_mlp2_bf16"])


def test_denoise_loop_matches_reference(golden):
    g = golden("loop")
    meta = g["meta"]
    W = {k: v.to(torch.bfloat16).float() for k, v in O.synth_weights(O.dit_param_shapes(meta["num_layers"]), seed=meta["w_seed"]).items()}
    A = {k: v.to(torch.bfloat16).float() for k, v in O.synth_weights(O.adapter_param_shapes(), seed=meta["a_seed"]).items()}
    posi = O.synth_inputs(meta["height"], meta["height"], meta["T_posi"], seed=meta["posi_seed"])
    nega = O.synth_inputs(meta["height"], meta["height"], meta["T_nega"], seed=meta["nega_seed"])
    lat = O.denoise_loop(W, A, posi["latents"].clone(), posi, nega, posi["edit_latents"], meta["height"], meta["height"], meta["steps"])
    assert rel_l2(lat, g["latents"]) < 5e-6
    assert rel_l2(posi["prompt_emb"][posi["special_token_mask"]][:, ::8], g["prompt_emb_posi_after"]) < 5e-6


def test_aux_training_path_oracle_matches_reference(golden):
    """DINOv2-with-registers (transformers 5.5.0 through the reference's Dinov2withNorm), perceiver resamplers and
    VisualThinkingAdapters: oracle/aux_oracle.py vs outputs of the reference modules (fp32, seeded weights)."""
    from oracle import aux_oracle as AO
    g = golden("aux")
    P = AO.aux_synth(seed=g["seed"])
    ain = AO.aux_inputs(seed=g["in_seed"])
    with torch.no_grad():
        src = AO.dinov2_with_norm(P["dinov2"], ain["dino_source"])
        mid = AO.dinov2_with_norm(P["dinov2"], ain["dino_middle"])
        assert src.shape == (1, 256, 768) and mid.shape == (3, 256, 768)
        assert rel_l2(src[:, ::4, ::8], g["dino_source"]) < 2e-5
        assert rel_l2(mid[:, ::4, ::8], g["dino_middle"]) < 2e-5
        # non-affine final LayerNorm: unit variance per token
        assert abs(src.var(dim=-1, unbiased=False).mean().item() - 1.0) < 1e-3
        hs = (mid + P["dino_time_embed"]["weight"][:3].unsqueeze(1)).reshape(1, -1, 768)
        assert rel_l2(AO.perceiver_resampler(P["dino_resampler"], hs)[..., ::4], g["resampler_dino_mid"]) < 2e-5
        ed, ev = AO.physical_visual_embeddings(P, **ain)
        assert ed.shape == ev.shape == (1, 64, 3584)
        assert rel_l2(ed[..., ::8], g["pseudo_special_emb_dino"]) < 5e-5
        assert rel_l2(ev[..., ::8], g["pseudo_special_emb_vae"]) < 5e-5


def test_rope_sampling_tables_bit_exact(golden):
    """edit_rope_interpolation (QwenEmbedRope.forward_sampling, qwen_image_dit.py:168-225): oracle and the product's host code vs the reference."""
    from physicedit_b200.dit import QwenEmbedRope
    for key, c in golden("f5")["rope_sampling"].items():
        shapes = [tuple(s) for s in c["shapes"]]
        vid, txt = O.rope_tables(shapes, c["T"], sampling=True)
        assert vid.shape[0] == c["n_vid"] and torch.equal(vid[c["vid_idx"]], c["vid"]) and torch.equal(txt, c["txt"]), key
        pe = QwenEmbedRope(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
        v2, t2 = pe.forward_sampling(shapes, [c["T"]])
        assert torch.equal(v2[c["vid_idx"]], c["vid"]) and torch.equal(t2, c["txt"]), key
        # the plain tables differ whenever an edit image has another size than the noise image
        v3, _ = QwenEmbedRope(theta=10000, axes_dim=[16, 56, 56], scale_rope=True).forward(shapes, [c["T"]])
        assert torch.equal(v3, v2) == all(s[1:] == shapes[0][1:] for s in shapes[1:]), key


def test_controlnet_oracle_matches_reference(golden):
    """oracle.controlnet_* vs the reference's QwenImageBlockWiseControlNet / QwenImageBlockwiseMultiControlNet (fp32 near-exact, bf16 bit-exact)."""
    g = golden("f5")["controlnet"]
    meta = g["meta"]
    for dtype in (torch.float32, torch.bfloat16):
        nets = [dict(weights={k: v.to(dtype) for k, v in O.synth_weights(O.controlnet_param_shapes(meta["L"]), seed=s).items()}, latents=l.to(dtype), **kw)
                for s, l, kw in zip(meta["seeds"], meta["latents"], (dict(scale=1.0, start=1.0, end=0.0), dict(scale=0.5, start=0.8, end=0.3)))]
        conds = [O.controlnet_img_in(c["weights"], O.patchify(c["latents"])) for c in nets]
        ref = g[str(dtype)]
        tol = dict(rtol=0, atol=0) if dtype == torch.bfloat16 else dict(rtol=1e-5, atol=1e-5)
        assert torch.allclose(conds[0][:, :, ::meta["stride"]], ref["cond0"], **tol)
        for (pid, blk), want in ref["res"].items():
            got = O.controlnet_sum(nets, conds, meta["x"].to(dtype), blk, pid, 5)
            assert torch.allclose(got[:, :, ::meta["stride"]], want, **tol), (dtype, pid, blk)


def test_eligen_oracle_matches_reference(golden):
    """oracle.process_entity_masks + masked attention inside model_fn vs the reference's model_fn_qwen_image with entity prompts / masks (fp32)."""
    g = golden("f5")["eligen"]
    m = g["meta"]
    W = O.synth_weights(O.dit_param_shapes(1), seed=m["w_seed"])
    inp = O.synth_inputs(m["H"], m["W"], m["T"], seed=m["in_seed"], dtype=torch.float32, n_special=m["n_special"])
    y = O.model_fn(W, None, inp["latents"], torch.tensor([m["t"]]), inp["prompt_emb"].clone(), inp["prompt_emb_mask"], None, m["H"], m["W"],
                   edit_latents=inp["edit_latents"], entity=dict(prompt_emb=g["entity_prompt_emb"], masks=g["entity_masks"]))
    assert g["differs_from_plain"] > 1e-3
    assert ((y - g["y"]).norm() / g["y"].norm()).item() < 2e-5


def test_entity_attention_mask_host_logic_matches_the_oracle():
    """physicedit_b200.model_fn.entity_attention_mask (pure torch: runs on the CPU) == the boolean form of the oracle's process_entity_masks mask
    (itself pinned to the reference by test_eligen_oracle_matches_reference), with an edit image and three entities of different lengths."""
    from physicedit_b200.model_fn import entity_attention_mask
    H = W = 96
    g = torch.Generator().manual_seed(1)
    masks = (torch.rand(1, 3, 1, H // 8, W // 8, generator=g) > 0.6).float()
    lat = [torch.zeros(1, 16, H // 8, W // 8), torch.zeros(1, 16, H // 8, W // 8)]
    seg = [5, 11, 3, 20]
    ents = [torch.zeros(1, n, 3584) for n in seg[:-1]]
    Wt = {"txt_norm.weight": torch.ones(3584), "txt_in.weight": torch.zeros(8, 3584), "txt_in.bias": torch.zeros(8)}
    n_img = 2 * (H // 16) * (W // 16)
    _, _, add = O.process_entity_masks(Wt, lat[0], torch.zeros(1, seg[-1], 3584), seg[-1], ents, masks, H, W, n_img, [(1, H // 16, W // 16)] * 2)
    got = entity_attention_mask(masks, seg, lat)
    assert got.dtype == torch.uint8 and got.shape == (sum(seg) + n_img,) * 2
    assert torch.equal(got.bool(), add[0, 0] == 0)
    assert got.diagonal().all()                                   # every token sees itself: no empty softmax row
    with pytest.raises(ValueError, match="entity"):
        entity_attention_mask(masks, seg[1:], lat)
    with pytest.raises(ValueError, match="latent size"):
        entity_attention_mask(masks, seg, [lat[0], torch.zeros(1, 16, 10, 14)])
