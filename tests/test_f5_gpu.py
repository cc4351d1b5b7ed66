"""SURVEY.md 8f5 options of model_fn_qwen_image (none is enabled by a PhysicEdit script; the reference supports them, so does the drop-in):
edit_rope_interpolation, blockwise controlnet, EliGen entity masks, fp8 attention.  Same protocol as the other parity tests: error of the native
path vs the fp32 oracle within the reference-bf16 floor + 1e-3."""
import math

import pytest
import torch

from oracle import dit_oracle as O

gpu = pytest.mark.gpu
TOL_EXTRA = 1e-3


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _oracle(sd, ad, inp, t, H, W, dtype, layers, **kw):
    from test_parity_depth_gpu import Upcast
    Wd = Upcast(sd) if dtype == torch.float32 else sd
    Ad = Upcast(ad) if dtype == torch.float32 else ad
    c = lambda x: x.cuda().to(dtype) if x.is_floating_point() else x.cuda()
    with torch.no_grad():
        return O.model_fn(Wd, Ad, c(inp["latents"]), t.cuda().to(torch.bfloat16), c(inp["prompt_emb"]).clone(), c(inp["prompt_emb_mask"]),
                          c(inp["special_token_mask"]), H, W, edit_latents=c(inp["edit_latents"]), num_layers=layers, cuda_scalar_div=True, **kw)


@gpu
def test_edit_rope_interpolation_forward():
    """qwen_image_physical.py:1367-1368: an edit image of another size than the target takes the target grid's positions, sampled."""
    from test_parity_depth_gpu import device_model
    from physicedit_b200.model_fn import model_fn_qwen_image
    L, H, W, T = 2, 128, 96, 80
    pipe, sd, ad = device_model(L, seed=3)
    inp = O.synth_inputs(H, W, T, seed=51, dtype=torch.bfloat16, edit_hw=(192, 160))
    t = torch.tensor([311.0]).to(torch.bfloat16)

    def native(flag):
        with torch.no_grad():
            return model_fn_qwen_image(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"].cuda(), timestep=t.cuda(),
                                       prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                                       special_token_mask=inp["special_token_mask"].cuda(), height=H, width=W, edit_latents=inp["edit_latents"].cuda(),
                                       is_train=False, edit_rope_interpolation=flag)[0]
    y_plain = native(False)          # first: like the reference, forward() would otherwise find the SAMPLED entry in the shared cache
    y = native(True)
    y32 = _oracle(sd, ad, inp, t, H, W, torch.float32, L, edit_rope_interpolation=True)
    y16 = _oracle(sd, ad, inp, t, H, W, torch.bfloat16, L, edit_rope_interpolation=True)
    err, floor = rel_l2(y, y32), rel_l2(y16, y32)
    print(f"\nedit_rope_interpolation: native vs fp32 {err:.3e}, floor {floor:.3e}; vs the plain tables {rel_l2(y, y_plain):.3e}")
    assert err <= floor + TOL_EXTRA
    assert not torch.equal(y, y_plain)                                      # the option really changes the positions (bit-exact tables: CPU test)


def _controlnets(L, seeds, device="cuda"):
    from physicedit_b200.controlnet import QwenImageBlockWiseControlNet, QwenImageBlockwiseMultiControlNet
    nets, sds = [], []
    for s in seeds:
        sd = {k: v.to(torch.bfloat16) for k, v in O.synth_weights(O.controlnet_param_shapes(L), seed=s).items()}
        m = QwenImageBlockWiseControlNet(num_layers=L)
        m.load_state_dict(sd)
        nets.append(m.to(device=device, dtype=torch.bfloat16).eval())
        sds.append({k: v.to(device) for k, v in sd.items()})
    return QwenImageBlockwiseMultiControlNet(nets), sds


@gpu
@pytest.mark.parametrize("n_nets", [1, 2])
def test_blockwise_controlnet_forward(n_nets):
    """model_fn with blockwise_controlnet_conditioning (:1372-1374, :1389-1396): one controlnet (the fused scale-and-add epilogue) and two
    with different scales / activity windows, vs the fp32 oracle; and the same parameter names / shapes as the reference module."""
    from test_parity_depth_gpu import device_model
    from physicedit_b200.compat import ControlNetInput
    from physicedit_b200.model_fn import model_fn_qwen_image
    L, H, W, T = 2, 128, 128, 72
    pipe, sd, ad = device_model(L, seed=4)
    multi, csd = _controlnets(L, seeds=(61, 62)[:n_nets])
    assert set(multi.models[0].state_dict()) == set(O.controlnet_param_shapes(L))
    inp = O.synth_inputs(H, W, T, seed=52, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(9)
    ctrl = [torch.randn(1, 16, H // 8, W // 8, generator=g).bfloat16() for _ in range(n_nets)]
    spec = [dict(scale=1.0, start=1.0, end=0.0), dict(scale=0.5, start=0.8, end=0.3)][:n_nets]
    inputs = [ControlNetInput(controlnet_id=i, **s) for i, s in enumerate(spec)]
    t = torch.tensor([500.0]).to(torch.bfloat16)
    pid, steps = 2, 5                                                     # progress 0.5: both windows active
    with torch.no_grad():
        y = model_fn_qwen_image(dit=pipe.dit, blockwise_controlnet=multi, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"].cuda(),
                                timestep=t.cuda(), prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                                special_token_mask=inp["special_token_mask"].cuda(), height=H, width=W, edit_latents=inp["edit_latents"].cuda(), is_train=False,
                                blockwise_controlnet_conditioning=[c.cuda() for c in ctrl], blockwise_controlnet_inputs=inputs, progress_id=pid,
                                num_inference_steps=steps)[0]
        y_plain = model_fn_qwen_image(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"].cuda(), timestep=t.cuda(),
                                      prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                                      special_token_mask=inp["special_token_mask"].cuda(), height=H, width=W, edit_latents=inp["edit_latents"].cuda(), is_train=False)[0]

    def oracle(dtype):
        cn = [dict(weights={k: v.to(dtype) for k, v in w.items()}, latents=c.cuda().to(dtype), **s) for w, c, s in zip(csd, ctrl, spec)]
        return _oracle(sd, ad, inp, t, H, W, dtype, L, controlnet=cn, progress_id=pid, num_inference_steps=steps)
    y32, y16 = oracle(torch.float32), oracle(torch.bfloat16)
    err, floor = rel_l2(y, y32), rel_l2(y16, y32)
    print(f"\nblockwise controlnet x{n_nets}: native vs fp32 {err:.3e}, floor {floor:.3e}; vs no controlnet {rel_l2(y, y_plain):.3e}")
    assert err <= floor + TOL_EXTRA
    assert not torch.equal(y, y_plain)
    # a controlnet outside its window does nothing: bit-identical to the plain forward
    off = [ControlNetInput(controlnet_id=0, scale=1.0, start=0.4, end=0.0)]
    with torch.no_grad():
        y_off = model_fn_qwen_image(dit=pipe.dit, blockwise_controlnet=multi, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"].cuda(),
                                    timestep=t.cuda(), prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                                    special_token_mask=inp["special_token_mask"].cuda(), height=H, width=W, edit_latents=inp["edit_latents"].cuda(), is_train=False,
                                    blockwise_controlnet_conditioning=[ctrl[0].cuda()], blockwise_controlnet_inputs=off, progress_id=pid, num_inference_steps=steps)[0]
    assert torch.equal(y_off, y_plain)


@gpu
def test_masked_attention_matches_sdpa_with_a_boolean_mask():
    """DiTEngine.masked_attention (batched score GEMM -> masked softmax -> PV GEMM) vs fp32 softmax attention with the same mask, S not a multiple of 8."""
    from test_parity_depth_gpu import device_model
    pipe, _, _ = device_model(1, seed=1)
    eng = pipe.dit.engine()
    S, H = 333, 24
    g = torch.Generator(device="cuda").manual_seed(2)
    q, k, v = (torch.randn(S, H * 128, device="cuda", generator=g).bfloat16() for _ in range(3))
    mask = (torch.rand(S, S, device="cuda", generator=g) > 0.4)
    mask |= torch.eye(S, dtype=torch.bool, device="cuda")
    o = torch.empty_like(q)
    eng.masked_attention(q, k, v, o, mask.to(torch.uint8))
    hm = lambda t: t.view(S, H, 128).transpose(0, 1).float()
    sc = hm(q) @ hm(k).transpose(1, 2) / math.sqrt(128)
    ref = (torch.softmax(sc.masked_fill(~mask, float("-inf")), dim=-1) @ hm(v)).transpose(0, 1).reshape(S, H * 128)
    assert rel_l2(o, ref) < 5e-3
    eng.nat.check_async()


@gpu
def test_eligen_entity_control_forward():
    """model_fn with entity prompts / masks (:1360-1364): native (segment RoPE tables, masked attention) vs the fp32 oracle, whose EliGen branch is
    pinned to the reference by tests/golden/f5.pt; and the boolean mask equals the oracle's additive one."""
    from test_parity_depth_gpu import device_model
    from physicedit_b200.model_fn import entity_attention_mask, model_fn_qwen_image
    L, H, W, T = 2, 128, 128, 40
    pipe, sd, ad = device_model(L, seed=8)
    inp = O.synth_inputs(H, W, T, seed=53, dtype=torch.bfloat16, n_special=16)
    g = torch.Generator().manual_seed(10)
    ents = [(3 * torch.randn(1, n, 3584, generator=g)).bfloat16() for n in (9, 14)]
    masks = torch.zeros(1, 2, 1, H // 8, W // 8, dtype=torch.bfloat16)
    masks[0, 0, 0, :7, :9] = 1
    masks[0, 1, 0, 6:, 5:] = 1
    t = torch.tensor([640.0]).to(torch.bfloat16)
    with torch.no_grad():
        y = model_fn_qwen_image(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"].cuda(), timestep=t.cuda(),
                                prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                                special_token_mask=inp["special_token_mask"].cuda(), height=H, width=W, edit_latents=inp["edit_latents"].cuda(), is_train=False,
                                entity_prompt_emb=[e.cuda() for e in ents], entity_prompt_emb_mask=[torch.ones(1, e.shape[1], dtype=torch.long).cuda() for e in ents],
                                entity_masks=masks.cuda())[0]
        y_plain = model_fn_qwen_image(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"].cuda(), timestep=t.cuda(),
                                      prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                                      special_token_mask=inp["special_token_mask"].cuda(), height=H, width=W, edit_latents=inp["edit_latents"].cuda(), is_train=False)[0]

    def oracle(dtype):
        return _oracle(sd, ad, inp, t, H, W, dtype, L, entity=dict(prompt_emb=[e.cuda().to(dtype) for e in ents], masks=masks.cuda().to(dtype)))
    y32, y16 = oracle(torch.float32), oracle(torch.bfloat16)
    err, floor = rel_l2(y, y32), rel_l2(y16, y32)
    print(f"\nEliGen: native vs fp32 {err:.3e}, floor {floor:.3e}; vs no entity control {rel_l2(y, y_plain):.3e}")
    assert err <= floor + TOL_EXTRA
    assert not torch.equal(y, y_plain)
    lat = [inp["latents"].cuda(), inp["edit_latents"].cuda()]
    m_nat = entity_attention_mask(masks.cuda(), [9, 14, T], lat)
    seq = [9, 14, T]
    n_img = 2 * (H // 16) * (W // 16)
    patched = [torch.nn.functional.max_pool2d(masks[:, i].float(), 2).flatten(1)[0] > 0 for i in range(2)] + [torch.ones((H // 16) * (W // 16), dtype=torch.bool)]
    total = sum(seq) + n_img
    want = torch.ones(total, total, dtype=torch.bool)
    cum = [0, 9, 23, 23 + T]
    for i in range(3):
        im = patched[i].repeat(2)[None, :].expand(seq[i], -1)
        want[cum[i]:cum[i + 1], cum[3]:] = im
        want[cum[3]:, cum[i]:cum[i + 1]] = im.t()
        for j in range(3):
            if j != i:
                want[cum[i]:cum[i + 1], cum[j]:cum[j + 1]] = False
    assert torch.equal(m_nat.cpu().bool(), want)


@gpu
def test_fp8_attention_option():
    """enable_fp8_attention (qwen_image_dit.py:24-35): q / k / v scaled by their std and rounded to e4m3, scale q_std k_std / sqrt(d), output x v_std --
    vs the same arithmetic spelled out in fp32 (FlashAttention-3 is not installed here: the reference itself cannot run this branch on this box)."""
    from test_parity_depth_gpu import device_model
    pipe, _, _ = device_model(1, seed=1)
    eng = pipe.dit.engine()
    S, H = 1000, 24
    g = torch.Generator(device="cuda").manual_seed(3)
    q, k, v = (torch.randn(S, H * 128, device="cuda", generator=g).bfloat16() * s for s in (1.3, 0.8, 2.0))
    o, o_bf16 = torch.empty_like(q), torch.empty_like(q)
    eng.fp8_attention(q, k, v, o)
    eng.nat.attention(q, k, v, o_bf16, H, 1 / math.sqrt(128))
    qs, ks, vs = q.std(), k.std(), v.std()
    f8 = lambda t, s_: (t / s_).to(torch.float8_e4m3fn).float().view(S, H, 128).transpose(0, 1)
    p = torch.softmax(f8(q, qs) @ f8(k, ks).transpose(1, 2) * (qs * ks).float() / math.sqrt(128), dim=-1)
    ref = ((p @ f8(v, vs)).to(torch.bfloat16) * vs).transpose(0, 1).reshape(S, H * 128)
    e = rel_l2(o, ref)
    print(f"\nfp8 attention: native vs fp32 emulation {e:.3e}; vs bf16 attention {rel_l2(o, o_bf16):.3e}")
    assert e < 6e-3 and rel_l2(o, o_bf16) > 2e-2
