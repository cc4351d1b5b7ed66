"""Parity of the CUDA path (through the C ABI, physicedit_b200/lib/libpe_b200.so) against the oracle and the
reference-generated goldens.  Run with `-m gpu` on a B200.

Tolerance rule: the reference's own bf16 forward differs from its fp32 forward by
~1e-2 (measured noise floor, `test_reference_bf16_noise_floor`), so the bar is
    err(native, fp32 oracle)  <=  err(reference bf16, fp32 oracle) + 1e-3        (relative L2 on the latents)
and bit-exact for integer / bookkeeping outputs (patchify, gather, timestep sinusoid input, scheduler).
"""
import math

import pytest
import torch

from oracle import dit_oracle as O

gpu = pytest.mark.gpu
TOL_EXTRA = 1e-3


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item()


@pytest.fixture(scope="module")
def nat():
    from physicedit_b200 import native as nv
    return nv.Native.get(0)


def _build_dit(num_layers, seed):
    from physicedit_b200.dit import QwenImageDiT
    W = {k: v.to(torch.bfloat16) for k, v in O.synth_weights(O.dit_param_shapes(num_layers), seed=seed).items()}
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=num_layers)
    dit.load_state_dict({k: v.clone() for k, v in W.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    return dit.to("cuda").eval(), W


def _build_adapter(seed, t_min, t_max):
    from physicedit_b200.adapters import VisualThinkingDualAdapter
    A = {k: v.to(torch.bfloat16) for k, v in O.synth_weights(O.adapter_param_shapes(), seed=seed).items()}
    ad = VisualThinkingDualAdapter(3584, 3584, t_min, t_max)
    ad.load_state_dict(A)
    return ad.to(device="cuda", dtype=torch.bfloat16).eval(), A


@gpu
def test_library_is_loaded_and_device_is_sm100(nat):
    assert nat.sm_count >= 100
    assert nat.lib.pe_abi_version() == 2


@gpu
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("M,N,K,M2", [(256, 512, 128, 0), (300, 264, 192, 77), (1, 64, 64, 0), (1000, 3072, 3072, 33)])
def test_gemm_bias(nat, cg, M, N, K, M2):
    from physicedit_b200 import native as nv
    torch.manual_seed(M + N)
    segs, refs = [], []
    for m in [M] + ([M2] if M2 else []):
        a = torch.randn(m, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        out = torch.zeros(m, N, device="cuda", dtype=torch.bfloat16)
        segs.append(dict(a=a, w=w, bias=b, out=out))
        refs.append((a.float() @ w.float().t() + b.float()))
    nat.gemm(segs, N, K, nv.EPI_BIAS, nv.GEMM_FLAG_CTA_PAIR if cg == 2 else 0)
    nat.check_async()
    for s, r in zip(segs, refs):
        assert rel_l2(s["out"], r) < 3e-3


@gpu
def test_gemm_rejects_bad_arguments(nat):
    from physicedit_b200 import native as nv
    a = torch.zeros(8, 12, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(16, 12, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(nv.NativeError):
        nat.gemm([dict(a=a, w=w, bias=None, out=torch.zeros(8, 16, device="cuda", dtype=torch.bfloat16))], 16, 12)   # K % 8 != 0
    with pytest.raises(nv.NativeError):
        nat.gemm([dict(a=a.float(), w=w, bias=None, out=a)], 16, 12)


@gpu
@pytest.mark.parametrize("flags", [0, 32, 16, 8, 1, 2, 3])
@pytest.mark.parametrize("S,H", [(256, 2), (1000, 3), (130, 1), (60, 1), (64, 1), (8480, 2)])
def test_attention_vs_fp32_softmax(nat, flags, S, H):
    torch.manual_seed(S)
    d = H * 128
    q, k, v = (torch.randn(S, d, device="cuda").bfloat16() for _ in range(3))
    o = torch.zeros(S, d, device="cuda", dtype=torch.bfloat16)
    nat.attention(q, k, v, o, H, 1 / math.sqrt(128), flags)
    nat.check_async()
    qh, kh, vh = (t.view(S, H, 128).transpose(0, 1).float() for t in (q, k, v))
    ref = (torch.softmax(qh @ kh.transpose(1, 2) / math.sqrt(128), dim=-1) @ vh).transpose(0, 1).reshape(S, d)
    assert rel_l2(o, ref) < 5e-3


@gpu
def test_attention_peaked_rows_trigger_lazy_rescale(nat):
    """Rows whose maximum moves by far more than 2^8 between KV tiles exercise the O-rescale path."""
    torch.manual_seed(1)
    S, H = 640, 1
    q = torch.randn(S, 128, device="cuda").bfloat16()
    k = torch.randn(S, 128, device="cuda").bfloat16()
    k[300:] *= 6.0          # later tiles carry much larger logits
    k[600:] *= 3.0
    v = torch.randn(S, 128, device="cuda").bfloat16()
    o = torch.zeros_like(q)
    ref = torch.softmax(q.float() @ k.float().t() / math.sqrt(128), dim=-1) @ v.float()
    for flags in (0, 8, 16, 32):
        nat.attention(q, k, v, o, H, 1 / math.sqrt(128), flags)
        nat.check_async()
        assert rel_l2(o, ref) < 5e-3
    # a jump of more than 2^100 between consecutive KV tiles takes the overflow-guard (redo) path
    k2 = torch.randn(S, 128, device="cuda").bfloat16()
    k2[384:] *= 60.0
    for flags in (0, 8, 3, 16, 32):        # (32: the trailing-reference kernel flags the item and the exact kernel redoes it)
        nat.attention(q, k2, v, o, H, 1 / math.sqrt(128), flags)
        nat.check_async()
        ref = torch.softmax(q.float() @ k2.float().t() / math.sqrt(128), dim=-1) @ v.float()
        assert torch.isfinite(o.float()).all()
        assert rel_l2(o, ref) < 1e-2


@gpu
def test_rowwise_kernels_bit_exact_bookkeeping(nat, golden):
    # timestep sinusoid: bit-exact with the reference's CPU result for every golden timestep
    for t, c in golden("timestep").items():
        tb = torch.tensor([t]).to(torch.bfloat16).cuda()
        out = torch.empty(256, device="cuda", dtype=torch.bfloat16)
        nat.timestep_embedding(tb, out, True)
        ref = c["sinus_bf16"].to(torch.bfloat16)[0]
        mism = (out.cpu() != ref)
        # sin/cos of arguments up to 1000 rad: CUDA and CPU libm may differ by 1 ulp before bf16 rounding
        assert mism.float().mean().item() <= 2 / 256, (t, mism.sum().item())
        assert (out.cpu().float() - ref.float()).abs().max().item() <= 2 ** -7
        out2 = torch.empty_like(out)
        nat.timestep_embedding(c["ts_bf16"].cuda(), out2, False)
        assert torch.equal(out, out2)
    # patchify round trip and layout
    lat = torch.randn(1, 16, 32, 48).bfloat16()
    tok = torch.empty(16 * 24, 64, device="cuda", dtype=torch.bfloat16)
    nat.patchify(lat[0].cuda(), tok)
    assert torch.equal(tok.cpu(), O.patchify(lat)[0])
    back = torch.empty(16, 32, 48, device="cuda", dtype=torch.bfloat16)
    nat.unpatchify(tok, back)
    assert torch.equal(back.cpu(), lat[0])
    # CFG + Euler with the scheduler's own dsigma, bit-exact against torch bf16 arithmetic
    s = O.FlowMatchOracle()
    s.set_timesteps(50, dynamic_shift_len=4096)
    x, p, n = (torch.randn(16 * 64 * 64).bfloat16() for _ in range(3))
    for pid in (0, 17, 49):
        ref = s.step(n + 4.0 * (p - n), pid, x)
        xd = x.cuda().clone()
        nat.cfg_euler_step(xd, p.cuda(), n.cuda(), 4.0, float(s.dsigma(pid)))
        assert torch.equal(xd.cpu(), ref)


@gpu
def test_layernorm_modulate_tail_is_bit_exact(nat):
    """LN + AdaLN modulate (qwen_image_dit.py:319-401 `_modulate`): the bf16 rounding points of `norm(x) * (1 + scale) + shift` are
    replayed bit for bit.  The normalised tensor n is obtained from the kernel itself (1+scale = 1, shift = 0 is exact), then
    bf16(bf16(n * ops) + shift) computed by ATen's bf16 ops must equal the kernel's fused output -- including huge / tiny / mixed-sign
    modulation values that stress the packed bf16x2 multiply and add."""
    torch.manual_seed(3)
    rows, C, split = 333, 3072, 100
    x = (torch.randn(rows, C, device="cuda") * 3 + 0.5).bfloat16()
    one = torch.ones(C, device="cuda", dtype=torch.bfloat16)
    zero = torch.zeros(C, device="cuda", dtype=torch.bfloat16)
    n = torch.empty_like(x)
    nat.layernorm_modulate2(x, n, split, zero, one, zero, one)
    ref_n = torch.nn.functional.layer_norm(x.float(), (C,), eps=1e-6)
    assert (n.float() - ref_n).abs().max().item() <= 2 ** -6          # bf16 rounding of values up to ~4
    mags = torch.tensor([1e-3, 1.0, 37.0, 3e4], device="cuda")
    for seed in range(3):
        g = torch.Generator(device="cuda").manual_seed(seed)
        ops = [(torch.randn(C, device="cuda", generator=g) * mags[torch.randint(0, 4, (C,), device="cuda", generator=g)]).bfloat16() for _ in range(2)]
        sh = [(torch.randn(C, device="cuda", generator=g) * mags[torch.randint(0, 4, (C,), device="cuda", generator=g)]).bfloat16() for _ in range(2)]
        out = torch.empty_like(x)
        nat.layernorm_modulate2(x, out, split, sh[0], ops[0], sh[1], ops[1])
        ref = torch.cat([n[:split] * ops[0] + sh[0], n[split:] * ops[1] + sh[1]], dim=0)      # ATen bf16 mul, then bf16 add
        assert torch.equal(out, ref)
        out1 = torch.empty_like(x[:split])
        nat.layernorm_modulate(x[:split].contiguous(), out1, sh[0], ops[0])
        assert torch.equal(out1, ref[:split])
    nat.check_async()


@gpu
def test_scheduler_matches_golden(golden):
    from physicedit_b200.scheduler import FlowMatchScheduler
    g = golden("scheduler")
    s = FlowMatchScheduler(sigma_min=0, sigma_max=1, extra_one_step=True, exponential_shift=True, exponential_shift_mu=0.8, shift_terminal=0.02)
    for key, c in g["cases"].items():
        hw, n = key.split("_")
        h, w = map(int, hw.split("x"))
        s.set_timesteps(int(n), dynamic_shift_len=(h // 16) * (w // 16))
        assert torch.equal(s.sigmas, c["sigmas"]) and torch.equal(s.timesteps, c["timesteps"])
        lat = torch.linspace(-1, 1, 64).bfloat16().cuda()
        vel = torch.linspace(2, -2, 64).bfloat16().cuda()
        steps = torch.stack([s.step(vel, s.timesteps[i], lat).cpu() for i in range(int(n))])
        assert torch.equal(steps, c["step_out"])


@gpu
def test_model_fn_parity_two_blocks(golden):
    """2 blocks, 128x128 + 128x128 edit image, T=80 (64 special tokens), two consecutive calls."""
    from physicedit_b200.model_fn import model_fn_qwen_image
    fwd = golden("forward")
    meta = fwd["meta"]
    dit, W = _build_dit(meta["num_layers"], meta["w_seed"])
    ad, A = _build_adapter(meta["a_seed"], meta["t_min"], meta["t_max"])
    inp = O.synth_inputs(meta["height"], meta["width"], meta["T"], seed=meta["in_seed"], dtype=torch.bfloat16)
    dev = {k: v.cuda() for k, v in inp.items()}
    pe = dev["prompt_emb"].clone()
    # fp32 oracle on the same (bf16-representable) weights
    Wf = {k: v.float() for k, v in W.items()}
    Af = {k: v.float() for k, v in A.items()}
    pe_o = inp["prompt_emb"].float().clone()
    for call, tval in enumerate((744.611382484436, 426.6734719276428)):
        t = torch.tensor([tval]).to(torch.bfloat16)
        y, loss = model_fn_qwen_image(dit=dit, visual_thinking_adapter=ad, latents=dev["latents"], timestep=t.cuda(), prompt_emb=pe,
                                      prompt_emb_mask=dev["prompt_emb_mask"], special_token_mask=dev["special_token_mask"],
                                      height=meta["height"], width=meta["width"], edit_latents=dev["edit_latents"], is_train=False)
        dit.engine().nat.check_async()
        assert loss == 0
        y32 = O.model_fn(Wf, Af, inp["latents"].float(), t.float(), pe_o, inp["prompt_emb_mask"], inp["special_token_mask"],
                         meta["height"], meta["width"], edit_latents=inp["edit_latents"].float(), t_min=meta["t_min"], t_max=meta["t_max"])
        assert rel_l2(y32, fwd["fp32"]["out"][call]) < 1e-5            # the oracle reproduces the reference's fp32 output
        floor = rel_l2(fwd["bf16"]["out"][call], y32)                  # reference bf16 vs fp32: the noise floor
        err = rel_l2(y, y32)
        assert err <= floor + TOL_EXTRA, (call, err, floor)
        assert rel_l2(y, fwd["bf16"]["out"][call]) <= 2 * floor + TOL_EXTRA
    # in-place compounding of the special tokens, other rows untouched
    sm = inp["special_token_mask"][0]
    assert torch.equal(pe[0].cpu()[~sm], inp["prompt_emb"][0][~sm])
    assert rel_l2(pe[0].cpu()[sm][:, ::8], fwd["bf16"]["special_after"]) < 2e-2
    assert abs(pe[0].cpu()[sm].float().abs().mean().item() - fwd["bf16"]["special_abs_mean"]) < 0.05 * fwd["bf16"]["special_abs_mean"]


@gpu
def test_block_module_api_matches_golden(golden):
    """`block(image, text, temb, image_rotary_emb) -> (text, image)` keeps the reference's module contract."""
    fwd = golden("forward")
    meta = fwd["meta"]
    g = fwd["block0_bf16"]
    dit, W = _build_dit(meta["num_layers"], meta["w_seed"])
    inp = O.synth_inputs(meta["height"], meta["width"], meta["T"], seed=meta["in_seed"], dtype=torch.bfloat16)
    t = torch.tensor([500.0]).to(torch.bfloat16)
    temb = dit.time_text_embed((t / 1000).cuda(), torch.bfloat16)
    assert rel_l2(temb, g["temb"]) < 4e-3
    image = O.linear(torch.cat([O.patchify(inp["latents"]), O.patchify(inp["edit_latents"])], dim=1), W, "img_in")
    text = O.linear(O.rmsnorm(inp["prompt_emb"], W["txt_norm.weight"]), W, "txt_in")
    rope = dit.pos_embed([(1, 8, 8), (1, 8, 8)], [meta["T"]], device="cuda")
    t1, i1 = dit.transformer_blocks[0](image=image.cuda(), text=text.cuda(), temb=g["temb"].cuda(), image_rotary_emb=rope)
    assert rel_l2(t1[..., ::4], g["text_out"]) < 1e-2 and rel_l2(i1[..., ::4], g["image_out"]) < 1e-2


@gpu
def test_lora_fold_and_denoise_loop(golden):
    """LoRA fold in bf16 on device (bit-exact with the reference loader) + the 4-step CFG loop of config #1 at toy size."""
    from physicedit_b200.lora import GeneralLoRALoader
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    fwd = golden("forward")
    dit, _ = _build_dit(2, fwd["meta"]["w_seed"])
    dit.engine()                                      # pack first: the fold must land in the fused QKV buffer through the views
    gen = torch.Generator().manual_seed(5)
    lsd = {}
    for name, (o, i) in (("transformer_blocks.0.attn.to_q", (3072, 3072)), ("transformer_blocks.0.img_mlp.net.2", (3072, 12288)),
                         ("transformer_blocks.0.img_mod.1", (18432, 3072))):
        lsd[f"{name}.lora_A.default.weight"] = torch.randn(16, i, generator=gen) * 0.02
        lsd[f"{name}.lora_B.default.weight"] = torch.randn(o, 16, generator=gen) * 0.02
    assert GeneralLoRALoader(device="cuda", torch_dtype=torch.bfloat16).load(dit, lsd, alpha=1.0) == 3
    eng = dit.engine()
    assert rel_l2(eng.qkv_w[0][0][:64, :64], fwd["lora_folded_to_q_bf16"]) < 4e-3          # fused buffer sees the fold
    assert rel_l2(dit.transformer_blocks[0].img_mlp.net[2].weight[:64, :64], fwd["lora_folded_mlp2_bf16"]) < 4e-3
    del dit, eng
    torch.cuda.empty_cache()

    g = golden("loop")
    meta = g["meta"]
    dit, W = _build_dit(meta["num_layers"], meta["w_seed"])
    pipe = QwenImagePhysicPipeline(device="cuda", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit
    A = {k: v.to(torch.bfloat16) for k, v in O.synth_weights(O.adapter_param_shapes(), seed=meta["a_seed"]).items()}
    pipe.visual_thinking_adapter.load_state_dict(A)
    pipe.visual_thinking_adapter.to(device="cuda", dtype=torch.bfloat16)
    assert pipe.visual_thinking_adapter.t_min == 19.999980926513672 and pipe.visual_thinking_adapter.t_max == 1000.0
    posi = O.synth_inputs(meta["height"], meta["height"], meta["T_posi"], seed=meta["posi_seed"], dtype=torch.bfloat16)
    nega = O.synth_inputs(meta["height"], meta["height"], meta["T_nega"], seed=meta["nega_seed"], dtype=torch.bfloat16)
    keys = ("prompt_emb", "prompt_emb_mask", "special_token_mask")
    ip = {k: posi[k].cuda() for k in keys}
    in_ = {k: nega[k].cuda() for k in keys}
    lat = pipe.denoise(posi["latents"].cuda(), ip, in_, posi["edit_latents"].cuda(), height=meta["height"], width=meta["height"],
                       num_inference_steps=meta["steps"], cfg_scale=4.0)
    dit.engine().nat.check_async()
    # oracle loops: fp32 (truth) and bf16 (what the reference's arithmetic gives) on the same weights
    Wf = {k: v.float() for k, v in W.items()}
    Af = {k: v.float() for k, v in A.items()}
    pf = {k: (v.float() if v.is_floating_point() else v) for k, v in posi.items()}
    nf = {k: (v.float() if v.is_floating_point() else v) for k, v in nega.items()}
    lat32 = O.denoise_loop(Wf, Af, pf["latents"].clone(), pf, nf, pf["edit_latents"], meta["height"], meta["height"], meta["steps"])
    pb = {k: v.clone() for k, v in posi.items()}
    nb = {k: v.clone() for k, v in nega.items()}
    lat16 = O.denoise_loop(W, A, pb["latents"].clone(), pb, nb, pb["edit_latents"], meta["height"], meta["height"], meta["steps"])
    floor = rel_l2(lat16, lat32)
    err = rel_l2(lat, lat32)
    assert err <= floor + TOL_EXTRA, (err, floor)
    assert torch.isfinite(lat).all()
    # the two CFG branches on two streams (pipe.cfg_streams = 2): same kernels, same inputs -> bit-identical latents; run it with
    # equal text lengths too, where the branches would share a workspace without the per-branch key
    for t_nega in (meta["T_nega"], meta["T_posi"]):
        nega2 = O.synth_inputs(meta["height"], meta["height"], t_nega, seed=meta["nega_seed"], dtype=torch.bfloat16)
        res = []
        for streams in (1, 2):
            pipe.cfg_streams = streams
            ip = {k: posi[k].cuda() for k in keys}            # prompt_emb is mutated in place by the adapter: fresh copies per run
            in_ = {k: nega2[k].cuda() for k in keys}
            res.append(pipe.denoise(posi["latents"].cuda(), ip, in_, posi["edit_latents"].cuda(), height=meta["height"], width=meta["height"],
                                    num_inference_steps=meta["steps"], cfg_scale=4.0))
            dit.engine().nat.check_async()
        assert torch.equal(res[0], res[1])
    pipe.cfg_streams = 1


@gpu
def test_training_path_feature_extractors(golden):
    """DINOv2-with-registers, perceiver resamplers, VisualThinkingAdapters and the
    pseudo_special_emb targets on the CUDA path vs the fp32 oracle (same seeded, bf16-representable weights)."""
    from oracle import aux_oracle as AO
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    g = golden("aux")
    P = AO.aux_synth(seed=g["seed"], dtype=torch.bfloat16)
    ain = AO.aux_inputs(seed=g["in_seed"], dtype=torch.bfloat16)
    pipe = QwenImagePhysicPipeline(device="cuda", torch_dtype=torch.bfloat16, dinov2_config=dict(hidden=768, layers=12, heads=12))
    pipe.dinov2.encoder.load_state_dict({k: v for k, v in P["dinov2"].items() if not k.startswith("layernorm.")}, strict=True)
    pipe.dino_resampler.load_state_dict(P["dino_resampler"])
    pipe.vae_resampler.load_state_dict(P["vae_resampler"])
    pipe.dino_resampler_adapter.load_state_dict(P["dino_resampler_adapter"])
    pipe.vae_resampler_adapter.load_state_dict(P["vae_resampler_adapter"])
    pipe.dino_time_embed.load_state_dict(P["dino_time_embed"])
    pipe.vae_time_embed.load_state_dict(P["vae_time_embed"])
    pipe.to("cuda")
    Pf = {k: {n: t.float() for n, t in v.items()} for k, v in P.items()}
    af = {k: v.float() for k, v in ain.items()}
    with torch.no_grad():
        d_nat = pipe.dinov2(ain["dino_middle"].cuda())
        d_ref = AO.dinov2_with_norm(Pf["dinov2"], af["dino_middle"])
        d_b16 = AO.dinov2_with_norm(P["dinov2"], ain["dino_middle"])
        floor = rel_l2(d_b16, d_ref)
        assert d_nat.shape == (3, 256, 768)
        assert rel_l2(d_nat, d_ref) <= floor + 2e-3, (rel_l2(d_nat, d_ref), floor)
        hs = torch.randn(1, 700, 768, generator=torch.Generator().manual_seed(3)).bfloat16()
        r_nat = pipe.dino_resampler(hs.cuda())
        r_ref = AO.perceiver_resampler(Pf["dino_resampler"], hs.float())
        assert rel_l2(r_nat, r_ref) <= rel_l2(AO.perceiver_resampler(P["dino_resampler"], hs), r_ref) + 2e-3
        out = pipe.physical_visual_embeddings(**{k: v.cuda() for k, v in ain.items()})
        ed, ev = AO.physical_visual_embeddings(Pf, **af)
        ed16, ev16 = AO.physical_visual_embeddings(P, **ain)
        assert out["pseudo_special_emb_dino"].shape == (1, 64, 3584)
        assert rel_l2(out["pseudo_special_emb_dino"], ed) <= rel_l2(ed16, ed) + 5e-3
        assert rel_l2(out["pseudo_special_emb_vae"], ev) <= rel_l2(ev16, ev) + 5e-3
    from physicedit_b200 import native as nv
    nv.Native.get(0).check_async()


@gpu
def test_training_loss_forward_value(golden):
    """training_loss (qwen_image_physical.py:313-329): flow-matching MSE x weight + adapter loss, forward value."""
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    fwd = golden("forward")
    meta = fwd["meta"]
    dit, W = _build_dit(1, meta["w_seed"])
    pipe = QwenImagePhysicPipeline(device="cuda", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit
    A = {k: v.to(torch.bfloat16) for k, v in O.synth_weights(O.adapter_param_shapes(), seed=meta["a_seed"]).items()}
    pipe.visual_thinking_adapter.load_state_dict(A)
    pipe.visual_thinking_adapter.to(device="cuda", dtype=torch.bfloat16)
    pipe.scheduler.set_timesteps(1000, training=True)
    inp = O.synth_inputs(64, 64, 72, seed=8, dtype=torch.bfloat16)
    torch.manual_seed(0)
    gt_d = torch.randn(1, 64, 3584).bfloat16().cuda()
    gt_v = torch.randn(1, 64, 3584).bfloat16().cuda()
    pe = inp["prompt_emb"].cuda().clone()
    loss = pipe.training_loss(input_latents=inp["latents"].cuda(), prompt_emb=pe, prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                              special_token_mask=inp["special_token_mask"].cuda(), height=64, width=64, edit_latents=inp["edit_latents"].cuda(),
                              pseudo_special_emb_dino=gt_d, pseudo_special_emb_vae=gt_v, is_train=True)
    assert torch.isfinite(loss).all() and loss.item() > 0
    assert pipe.special_token_loss > 0
    # the adapter loss matches the oracle's formula on the predictions the native heads produced
    x = inp["prompt_emb"][inp["special_token_mask"]].view(1, 64, 3584).cuda()
    _, pd, pv = pipe.visual_thinking_adapter(x, torch.tensor([500.0]).bfloat16().cuda())
    l_nat = pipe.visual_thinking_adapter.get_loss(pd, pv, gt_d, gt_v, torch.tensor([500.0]).bfloat16().cuda()).item()
    l_ora = O.adapter_loss(pd.cpu().float(), pv.cpu().float(), gt_d.cpu().float(), gt_v.cpu().float(), torch.tensor([500.0]).bfloat16(),
                           meta["t_min"], meta["t_max"]).item()
    assert abs(l_nat - l_ora) < 2e-2 * abs(l_ora)


# ---------------------------------------------------------------------------------------------------------------------
# Full-size (BASELINE.json config #2 shapes) checks through size-independent properties: the oracle cannot run these
# sizes in seconds, the properties can.
# ---------------------------------------------------------------------------------------------------------------------
@gpu
@pytest.mark.parametrize("flags", [0, 8, 16, 32])
def test_full_size_attention_properties(nat, flags):
    S, H = 8192 + 512, 24
    d = H * 128
    g = torch.Generator(device="cuda").manual_seed(7)
    q, k, v1, v2 = (torch.randn(S, d, device="cuda", generator=g).bfloat16() for _ in range(4))
    scale = 1 / math.sqrt(128)
    o = torch.empty_like(q)
    # (1) softmax rows sum to one: with V == 1 the output is exactly 1 up to the bf16 rounding of P
    ones = torch.ones_like(q)
    nat.attention(q, k, ones, o, H, scale, flags)
    assert (o.float() - 1).abs().max().item() <= 2 ** -7
    # (2) linear in V
    o1, o2, o12 = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    nat.attention(q, k, v1, o1, H, scale, flags)
    nat.attention(q, k, v2, o2, H, scale, flags)
    nat.attention(q, k, (v1.float() + v2.float()).bfloat16(), o12, H, scale, flags)
    assert rel_l2(o12, o1.float() + o2.float()) < 8e-3
    # (3) invariant under a joint permutation of the keys / values (order of the KV tiles and of the online softmax)
    perm = torch.randperm(S, device="cuda", generator=g)
    op = torch.empty_like(q)
    nat.attention(q, k[perm].contiguous(), v1[perm].contiguous(), op, H, scale, flags)
    assert rel_l2(op, o1) < 4e-3
    # (4) deterministic: no atomics, fixed schedule
    o1b = torch.empty_like(q)
    nat.attention(q, k, v1, o1b, H, scale, flags)
    assert torch.equal(o1, o1b)
    # (5) a spot check of 64 rows of one head against exact fp32 softmax
    rows = torch.arange(0, S, S // 64, device="cuda")[:64]
    ref = torch.softmax(q[rows, :128].float() @ k[:, :128].float().t() * scale, dim=-1) @ v1[:, :128].float()
    assert rel_l2(o1[rows, :128], ref) < 5e-3
    nat.check_async()


@gpu
def test_full_size_gemm_properties(nat):
    from physicedit_b200 import native as nv
    M, T, N, K = 8192, 512, 12288, 3072
    g = torch.Generator(device="cuda").manual_seed(3)
    a1 = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    a2 = torch.randn(T, K, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    w2 = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    o1 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o2 = torch.empty(T, N, device="cuda", dtype=torch.bfloat16)
    for flags in (0, nv.GEMM_FLAG_CTA_PAIR):
        # zero activations -> exactly the bias in every row (both segments, every tile, every epilogue warp)
        nat.gemm([dict(a=torch.zeros_like(a1), w=w1, bias=b, out=o1), dict(a=torch.zeros_like(a2), w=w2, bias=b, out=o2)], N, K, nv.EPI_BIAS, flags)
        assert torch.equal(o1, b.expand(M, N)) and torch.equal(o2, b.expand(T, N))
        # two segments with different weights in one launch == two separate launches, bit for bit
        nat.gemm([dict(a=a1, w=w1, bias=b, out=o1), dict(a=a2, w=w2, bias=b, out=o2)], N, K, nv.EPI_BIAS, flags)
        s1 = torch.empty_like(o1)
        s2 = torch.empty_like(o2)
        nat.gemm([dict(a=a1, w=w1, bias=b, out=s1)], N, K, nv.EPI_BIAS, flags)
        nat.gemm([dict(a=a2, w=w2, bias=b, out=s2)], N, K, nv.EPI_BIAS, flags)
        assert torch.equal(o1, s1) and torch.equal(o2, s2)
        # sampled rows against fp32 matmul
        rows = torch.arange(0, M, 257, device="cuda")
        ref = a1[rows].float() @ w1.float().t() + b.float()
        assert rel_l2(o1[rows], ref) < 3e-3
    # CTA-pair and single-CTA kernels accumulate in the same k order -> identical results
    p1 = torch.empty_like(o1)
    nat.gemm([dict(a=a1, w=w1, bias=b, out=p1)], N, K, nv.EPI_BIAS, 0)
    nat.gemm([dict(a=a1, w=w1, bias=b, out=s1)], N, K, nv.EPI_BIAS, nv.GEMM_FLAG_CTA_PAIR)
    assert torch.equal(p1, s1)
    nat.check_async()


@gpu
def test_full_sequence_block_is_deterministic_and_finite():
    """One block at the full 1024^2 sequence (8192 image + 512 text tokens): two runs are bit-identical, outputs finite,
    the text rows depend on the image rows (joint attention) and the result matches the oracle block on sampled rows."""
    dit, W = _build_dit(1, 21)
    eng = dit.engine()
    T, S_img = 512, 8192
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(T + S_img, 3072, generator=g).bfloat16()
    temb = torch.randn(1, 3072, generator=g).bfloat16()
    rope = eng.rope([(1, 64, 64), (1, 64, 64)], T)
    ws = eng.workspace(S_img, T)
    mods = eng.block_mods(temb.cuda(), [0])[0]
    outs = []
    for _ in range(2):
        x = x0.cuda().clone()
        eng.run_block(0, x, T, mods[0], rope, ws)
        outs.append(x)
    eng.nat.check_async()
    assert torch.equal(outs[0], outs[1])
    assert torch.isfinite(outs[0].float()).all()
    x2 = x0.clone()
    x2[T + 100:T + 200] += 1.0                                   # perturb image rows only
    xp = x2.cuda()
    eng.run_block(0, xp, T, mods[0], rope, ws)
    assert not torch.equal(xp[:T], outs[0][:T])                  # text stream sees the image through the joint attention
    # oracle on the same block (bf16 on CPU, ~10 s): sampled rows
    vid, txt = O.rope_tables([(1, 64, 64), (1, 64, 64)], T)
    t_o, i_o = O.block_forward(W, 0, x0[T:].unsqueeze(0), x0[:T].unsqueeze(0), temb, (vid, txt))
    ref = torch.cat([t_o[0], i_o[0]], dim=0)
    rows = torch.arange(0, T + S_img, 97)
    assert rel_l2(outs[0].cpu()[rows], ref[rows]) < 1.5e-2


@gpu
def test_model_fn_ragged_shapes_context_and_edit_list():
    """Non-square latents, an edit-image LIST plus context latents (qwen_image_physical.py:1347-1355), an odd text length and
    no special tokens: every M / S tail path of the kernels in one forward, against the fp32 oracle."""
    from physicedit_b200.model_fn import model_fn_qwen_image
    dit, W = _build_dit(1, 33)
    g = torch.Generator().manual_seed(9)
    H, Wd, T = 96, 160, 77
    lat = torch.randn(1, 16, H // 8, Wd // 8, generator=g).bfloat16()
    ctx = torch.randn(1, 16, 8, 12, generator=g).bfloat16()
    e1 = torch.randn(1, 16, 8, 16, generator=g).bfloat16()
    e2 = torch.randn(1, 16, 6, 10, generator=g).bfloat16()
    pe = (3 * torch.randn(1, T, 3584, generator=g)).bfloat16()
    mask = torch.ones(1, T, dtype=torch.int64)
    t = torch.tensor([426.6734719276428]).to(torch.bfloat16)
    y, loss = model_fn_qwen_image(dit=dit, latents=lat.cuda(), timestep=t.cuda(), prompt_emb=pe.cuda(), prompt_emb_mask=mask.cuda(),
                                  special_token_mask=None, height=H, width=Wd, edit_latents=[e1.cuda(), e2.cuda()], context_latents=ctx.cuda(),
                                  is_train=False)
    dit.engine().nat.check_async()
    assert loss == 0 and y.shape == lat.shape
    Wf = {k: v.float() for k, v in W.items()}

    def oracle(dtype):
        Wd_ = Wf if dtype == torch.float32 else W
        c = lambda x: x.to(dtype)
        # the oracle takes the image list in the reference's order: latents, context, edits
        img = [c(lat), c(ctx), c(e1), c(e2)]
        shapes = [(1, x.shape[2] // 2, x.shape[3] // 2) for x in img]
        image = O.linear(torch.cat([O.patchify(x) for x in img], dim=1), Wd_, "img_in")
        ts = c(t) / 1000
        temb = O.time_text_embed(Wd_, ts, dtype)
        text = O.linear(O.rmsnorm(c(pe), Wd_["txt_norm.weight"]), Wd_, "txt_in")
        rope = O.rope_tables(shapes, T)
        text, image = O.block_forward(Wd_, 0, image, text, temb, rope)
        emb = O.linear(torch.nn.functional.silu(temb), Wd_, "norm_out.linear")
        scale, shift = emb.unsqueeze(1).chunk(2, dim=2)
        image = O.layernorm(image) * (1 + scale) + shift
        n0 = shapes[0][1] * shapes[0][2]
        return O.unpatchify(O.linear(image, Wd_, "proj_out")[:, :n0], H // 16, Wd // 16)
    y32, y16 = oracle(torch.float32), oracle(torch.bfloat16)
    floor = rel_l2(y16, y32)
    assert rel_l2(y, y32) <= floor + TOL_EXTRA, (rel_l2(y, y32), floor)
