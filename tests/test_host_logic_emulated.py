"""Host logic of the training path (physicedit_b200/autograd.py) on the CPU: the autograd Functions are run against tests/abi_emulator.py -- a
contract-level emulation of the C-ABI entry points they call -- and their gradients compared with torch autograd on the same math in fp32.  What this
pins without a GPU: which operand is transposed / padded for dX and dW, the bias-gradient-as-extra-column trick, the seven batched products and the
row / column statistics of the attention backward, zero-padding of ragged sequence lengths, LoRA composition.  The GPU tests run the same host code on
the real library (tests/test_training_gpu.py, tests/test_f5_gpu.py).  The file also covers the EliGen masked attention and the blockwise controlnet the same way."""
import math
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from abi_emulator import EmulatedNative  # noqa: E402


@pytest.fixture()
def ag(monkeypatch):
    from physicedit_b200 import autograd
    emu = EmulatedNative()
    monkeypatch.setattr(autograd, "_nat", lambda t: emu)
    autograd.weight_transposes.clear()
    autograd.emu = emu
    return autograd


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,N,K,bias", [(13, 16, 24, True), (40, 8, 64, False), (1, 32, 16, True), (64, 24, 8, True)])
def test_linear_function_gradients(ag, M, N, K, bias):
    g = torch.Generator().manual_seed(M * 100 + N)
    x = torch.randn(3, M, K, generator=g).bfloat16().requires_grad_()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16().requires_grad_()
    b = (torch.randn(N, generator=g) * 0.1).bfloat16().requires_grad_() if bias else None
    dy = torch.randn(3, M, N, generator=g).bfloat16()
    y = ag.linear(x, w, b)
    assert y.shape == (3, M, N)
    y.backward(dy)
    x32, w32 = x.detach().float().requires_grad_(), w.detach().float().requires_grad_()
    b32 = b.detach().float().requires_grad_() if bias else None
    F.linear(x32, w32, b32).backward(dy.float())
    assert rel(x.grad, x32.grad) < 6e-3 and rel(w.grad, w32.grad) < 6e-3 and x.grad.shape == x.shape and w.grad.is_contiguous()
    if bias:
        assert rel(b.grad, b32.grad) < 6e-3
    names = [c[0] for c in ag.emu.calls]
    assert names.count("pe_gemm") == 3 and names.count("pe_transpose") == 3          # Y, dX (+ W^T), dW|db (+ dY^T, X^T)


def test_frozen_weight_transposes_are_cached_per_live_parameter(ag):
    w = torch.nn.Parameter((torch.randn(16, 24) / 5).bfloat16(), requires_grad=False)
    for _ in range(2):
        x = torch.randn(9, 24).bfloat16().requires_grad_()
        ag.linear(x, w, None).sum().backward()
    assert [c[0] for c in ag.emu.calls].count("pe_transpose") == 1 and ag.weight_transposes.used == 16 * 24 * 2
    with torch.no_grad():
        w.add_(1)                                              # an in-place update bumps the version: a new transpose
    ag.linear(torch.randn(9, 24).bfloat16().requires_grad_(), w, None).sum().backward()
    assert [c[0] for c in ag.emu.calls].count("pe_transpose") == 2
    # a write through .data (what GeneralLoRALoader.load does) is invisible to the version counter: the engine's invalidate() clears the cache
    from physicedit_b200.dit import DiTEngine
    w.data.add_(1)
    DiTEngine.invalidate(type("E", (), {})())
    assert ag.weight_transposes.used == 0
    ag.linear(torch.randn(9, 24).bfloat16().requires_grad_(), w, None).sum().backward()
    assert [c[0] for c in ag.emu.calls].count("pe_transpose") == 3
    n_before = len(ag.weight_transposes.d)
    del w
    import gc
    gc.collect()
    assert len(ag.weight_transposes.d) < n_before              # the entry dies with the parameter (its address may be recycled)


@pytest.mark.parametrize("S,H", [(37, 2), (64, 1), (21, 3)])
def test_attention_function_gradients(ag, monkeypatch, S, H):
    monkeypatch.setattr(ag, "HEAD_DIM", 16)
    g = torch.Generator().manual_seed(S)
    q, k, v, do = (torch.randn(S, H * 16, generator=g).bfloat16() for _ in range(4))

    def run(native):
        qq, kk, vv = (t.detach().clone().to(torch.bfloat16 if native else torch.float32).requires_grad_() for t in (q, k, v))
        if native:
            o = ag.attention(qq, kk, vv, H)
        else:
            hm = lambda t: t.view(S, H, 16).transpose(0, 1)
            o = (torch.softmax(hm(qq) @ hm(kk).transpose(1, 2) / 4.0, dim=-1) @ hm(vv)).transpose(0, 1).reshape(S, H * 16)
        o.backward(do.to(o.dtype))
        return o.detach(), qq.grad, kk.grad, vv.grad
    got, want = run(True), run(False)
    for name, a, b in zip(("o", "dq", "dk", "dv"), got, want):
        assert a.shape == b.shape and rel(a, b) < 1.2e-2, (name, rel(a, b))
    names = [c[0] for c in ag.emu.calls]
    assert names.count("pe_gemm_batched") == 7 and names.count("pe_attention_fwd_lse") == 1 and names.count("pe_attention_bwd_delta") == 1


def test_lora_and_hot_lora_linear_match_their_formulas(ag):
    from physicedit_b200.lora import HotLoRALinear, LoRALinear
    torch.manual_seed(0)
    base = torch.nn.Linear(24, 16).bfloat16()
    m = LoRALinear(base, r=8, lora_alpha=16)
    m.lora_B["default"].weight.data = (torch.randn(16, 8) * 0.3).bfloat16()
    x = torch.randn(5, 24).bfloat16()
    y = m(x)
    A, B = m.lora_A["default"].weight.float(), m.lora_B["default"].weight.float()
    want = F.linear(x.float(), base.weight.float(), base.bias.float()) + (x.float() @ A.t() @ B.t()) * 2.0
    assert m.scaling == 2.0 and rel(y, want) < 6e-3
    y.sum().backward()
    assert m.lora_A["default"].weight.grad is not None and m.lora_B["default"].weight.grad is not None and base.weight.grad is None
    h = HotLoRALinear(torch.nn.Linear(24, 16).bfloat16())
    h.lora_A_weights.append(A.bfloat16()); h.lora_B_weights.append(B.bfloat16())
    want_h = F.linear(x.float(), h.weight.float(), h.bias.float()) + x.float() @ A.t() @ B.t()
    assert rel(h(x), want_h) < 6e-3


def test_perceiver_attention_with_a_ragged_key_count(ag):
    """n + m keys not a multiple of 8: the zero padding of k / v rows must not change the softmax."""
    from physicedit_b200.adapters import PerceiverAttention
    torch.manual_seed(1)
    att = PerceiverAttention(dim=32, dim_head=8, heads=2).bfloat16()
    x, lat = torch.randn(13, 32).bfloat16(), torch.randn(6, 32).bfloat16().requires_grad_()
    out = ag.perceiver_attention(att, x, lat)
    xn = F.layer_norm(x.float(), (32,), att.norm_media.weight.float(), att.norm_media.bias.float())
    ln = F.layer_norm(lat.detach().float(), (32,), att.norm_latents.weight.float(), att.norm_latents.bias.float())
    q = (ln @ att.to_q.weight.float().t()).view(6, 2, 8).transpose(0, 1)
    k, v = (torch.cat((xn, ln)) @ att.to_kv.weight.float().t()).chunk(2, dim=-1)
    k, v = k.view(19, 2, 8).transpose(0, 1), v.view(19, 2, 8).transpose(0, 1)
    want = (torch.softmax(q @ k.transpose(1, 2) * att.scale, dim=-1) @ v).transpose(0, 1).reshape(6, 16) @ att.to_out.weight.float().t()
    assert out.shape == (6, 32) and rel(out, want) < 2e-2
    out.sum().backward()
    assert lat.grad is not None and torch.isfinite(lat.grad.float()).all()


def test_masked_attention_host_logic():
    """DiTEngine.masked_attention (EliGen): head-major zero-padded operands, one byte mask shared by all heads of a batched score matrix, head chunking --
    on the emulated ABI vs masked softmax attention in fp32; S = 21 is not a multiple of 8 and the scratch budget forces 5 chunks of heads."""
    from physicedit_b200.dit import DiTEngine, NUM_HEADS
    emu = EmulatedNative()
    eng = type("E", (), {"nat": emu, "MASKED_ATTN_SCRATCH_BYTES": 6 * 24 * 24 * 5})()
    S, H, D = 21, NUM_HEADS, 128
    g = torch.Generator().manual_seed(3)
    q, k, v = (torch.randn(S, H * D, generator=g).bfloat16() for _ in range(3))
    mask = torch.rand(S, S, generator=g) > 0.4
    mask |= torch.eye(S, dtype=torch.bool)
    o = torch.empty_like(q)
    DiTEngine.masked_attention(eng, q, k, v, o, mask.to(torch.uint8))
    hm = lambda t: t.float().view(S, H, D).transpose(0, 1)
    sc = (hm(q) @ hm(k).transpose(1, 2) / math.sqrt(D)).masked_fill(~mask, float("-inf"))
    want = (torch.softmax(sc, dim=-1) @ hm(v)).transpose(0, 1).reshape(S, H * D)
    assert rel(o, want) < 8e-3
    names = [c[0] for c in emu.calls]
    assert names.count("pe_gemm_batched") == 2 * 5 and names.count("pe_softmax_rows") == 5          # ceil(24 heads / 5 per chunk) = 5 chunks


def test_blockwise_controlnet_host_logic(monkeypatch):
    """QwenImageBlockwiseMultiControlNet.apply_ / preprocess on the emulated ABI vs the oracle's controlnet_sum (pinned to the reference by
    tests/golden/f5.pt): activity windows, one vs several controlnets (the summed correction is rounded before it is added), token order of preprocess."""
    from oracle import dit_oracle as O
    from physicedit_b200 import native as nv
    from physicedit_b200.compat import ControlNetInput
    from physicedit_b200.controlnet import QwenImageBlockWiseControlNet, QwenImageBlockwiseMultiControlNet
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    L, n = 1, 16
    sds = [{k: v.to(torch.bfloat16) for k, v in O.synth_weights(O.controlnet_param_shapes(L), seed=s).items()} for s in (61, 62)]
    nets = []
    for sd in sds:
        m = QwenImageBlockWiseControlNet(num_layers=L).bfloat16()
        m.load_state_dict(sd)
        nets.append(m)
    multi = QwenImageBlockwiseMultiControlNet(nets)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, 3072, generator=g).bfloat16()
    lat = [torch.randn(1, 16, 8, 8, generator=g).bfloat16() for _ in range(2)]
    spec = [dict(scale=1.0, start=1.0, end=0.0), dict(scale=0.5, start=0.8, end=0.3)]
    inputs = [ControlNetInput(controlnet_id=i, **s) for i, s in enumerate(spec)]
    conds = multi.preprocess(inputs, lat)
    ora = [dict(weights={k: v.float() for k, v in sd.items()}, latents=l.float(), **s) for sd, l, s in zip(sds, lat, spec)]
    oconds = [O.controlnet_img_in(c["weights"], O.patchify(c["latents"])) for c in ora]
    assert rel(conds[0], oconds[0]) < 6e-3 and conds[0].shape == (1, n, 3072)
    for pid, n_active in ((0, 1), (2, 2), (4, 1)):                          # of 5 steps: progress 1.0, 0.5, 0.0
        assert sum(multi.active(ci, pid, 5) for ci in inputs) == n_active
        got = x.clone()
        multi.apply_(got, conds, inputs, pid, 5, 0)
        want = x.float() + O.controlnet_sum(ora, oconds, x.float()[None], 0, pid, 5)[0]
        assert rel(got, want) < 8e-3, pid
    off = [ControlNetInput(controlnet_id=0, scale=1.0, start=0.4, end=0.0)]
    same = x.clone()
    multi.apply_(same, conds[:1], off, 2, 5, 0)
    assert torch.equal(same, x)


def test_dit_block_host_sequencing_matches_the_oracle_block():
    """DiTEngine.run_block (9 C-ABI calls per double-stream block) on the emulated ABI vs oracle.block_forward in fp32: which modulation slice feeds which
    LN / gate, the (img, txt) segment order of the grouped GEMMs, the fused [q; k; v] weight packing, the joint [text; image] layout, `1 + scale` mask."""
    from oracle import dit_oracle as O
    from physicedit_b200.dit import DIM, DiTEngine, QwenImageDiT, Workspace
    emu = EmulatedNative()
    W = O.synth_weights(O.dit_param_shapes(1), seed=33)
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=1)
    dit.load_state_dict({k: v.to(torch.bfloat16) for k, v in W.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    eng = object.__new__(DiTEngine)
    eng.dit, eng.device, eng.nat, eng.use_cta_pair, eng.attn_flags, eng._ws, eng._rope, eng.sp = dit, torch.device("cpu"), emu, True, 0, {}, {}, None
    eng._pack()
    assert dit.transformer_blocks[0].attn.to_k.weight.data_ptr() == eng.qkv_w[0][0][DIM:].data_ptr()          # parameters are views of the fused buffer
    T, shapes = 24, [(1, 4, 4), (1, 4, 4)]
    S_img = 32
    g = torch.Generator().manual_seed(4)
    text = torch.randn(1, T, DIM, generator=g).bfloat16()
    image = torch.randn(1, S_img, DIM, generator=g).bfloat16()
    temb = torch.randn(1, DIM, generator=g).bfloat16()
    x = torch.cat([text[0], image[0]], dim=0).contiguous()
    with torch.no_grad():
        mods = eng.block_mods(temb, [0])
        eng.run_block(0, x, T, mods[0, 0], eng.rope(shapes, T), Workspace(S_img, T, "cpu"))
    W32 = {k: v.to(torch.bfloat16).float() for k, v in W.items()}
    t_ref, i_ref = O.block_forward(W32, 0, image.float(), text.float(), temb.float(), O.rope_tables(shapes, T))
    assert rel(x[:T], t_ref[0]) < 1.5e-2 and rel(x[T:], i_ref[0]) < 1.5e-2
    assert [c[0] for c in emu.calls].count("pe_gemm") == 4 and [c[0] for c in emu.calls].count("pe_layernorm_modulate2") == 2


def test_whole_dit_forward_host_sequencing_matches_the_oracle(monkeypatch):
    """DiTEngine.forward -- patchify of noise + edit latents, img_in, txt_norm + txt_in, timestep embedding + conditioning (modulation GEMVs with the
    `1 + scale` slots, norm_out's (scale, shift) order), two blocks, norm_out + proj_out on the noise tokens only, unpatchify -- on the emulated ABI vs
    oracle.model_fn in fp32 with the bf16 timestep bookkeeping; a non-square image and an edit image of another size."""
    from oracle import dit_oracle as O
    from physicedit_b200 import native as nv
    from physicedit_b200.dit import DiTEngine, QwenImageDiT
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))          # TimestepEmbeddings.forward asks for the device's handle
    L, H, Wd, T = 2, 64, 96, 40
    W = O.synth_weights(O.dit_param_shapes(L), seed=35)
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=L)
    dit.load_state_dict({k: v.to(torch.bfloat16) for k, v in W.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    eng = object.__new__(DiTEngine)
    eng.dit, eng.device, eng.nat, eng.use_cta_pair, eng.attn_flags, eng._ws, eng._rope, eng.sp = dit, torch.device("cpu"), emu, True, 0, {}, {}, None
    eng._pack()
    inp = O.synth_inputs(H, Wd, T, seed=36, dtype=torch.bfloat16, edit_hw=(80, 48))
    t = torch.tensor([431.0]).to(torch.bfloat16)
    out = torch.empty_like(inp["latents"])
    with torch.no_grad():
        eng.forward([inp["latents"], inp["edit_latents"]], t, inp["prompt_emb"][0].contiguous(), out, t_key=float(t[0]))
        W32 = {k: v.to(torch.bfloat16).float() for k, v in W.items()}
        want = O.model_fn(W32, None, inp["latents"].float(), t, inp["prompt_emb"].float().clone(), inp["prompt_emb_mask"], None, H, Wd,
                          edit_latents=inp["edit_latents"].float(), cuda_scalar_div=True)
    assert out.shape == want.shape == (1, 16, H // 8, Wd // 8)
    print(f"whole forward on the emulated ABI vs fp32 oracle: {rel(out, want):.3e}")
    assert rel(out, want) < 1e-2, rel(out, want)
    names = [c[0] for c in emu.calls]
    assert names.count("pe_patchify") == 2 and names.count("pe_unpatchify") == 1 and names.count("pe_attention_fwd") == L


def test_training_forward_and_backward_host_logic_matches_oracle_autograd(ag, monkeypatch):
    """The CPU twin of tests/test_training_gpu.py: autograd.dit_forward (un-merged LoRA on 6 targets, per-block checkpointing, adapter rows carrying a graph)
    + backward on the emulated ABI vs autograd through oracle.model_fn in fp32 with W_eff = W + B A."""
    from oracle import dit_oracle as O
    from physicedit_b200 import native as nv
    from physicedit_b200.dit import DiTEngine, QwenImageDiT
    from physicedit_b200.lora import inject_lora
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: ag.emu))
    H, Wd, T = 64, 64, 24
    W = O.synth_weights(O.dit_param_shapes(1), seed=37)
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=1)
    dit.load_state_dict({k: v.to(torch.bfloat16) for k, v in W.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    targets = ["to_q", "add_k_proj", "to_out.0", "img_mlp.net.2", "txt_mod.1", "to_v"]
    inject_lora(dit, targets, r=8)
    g = torch.Generator().manual_seed(6)
    lora = {}
    for name, p in dit.named_parameters():
        if "lora_" in name:
            p.data = (torch.randn(p.shape, generator=g) * (0.5 / math.sqrt(p.shape[1]))).bfloat16()
            lora[name] = p
    eng = object.__new__(DiTEngine)
    eng.dit, eng.device, eng.nat, eng.use_cta_pair, eng.attn_flags, eng._ws, eng._rope, eng.sp = dit, torch.device("cpu"), ag.emu, True, 0, {}, {}, None
    eng._pack()
    object.__setattr__(dit, "_engine", eng)
    inp = O.synth_inputs(H, Wd, T, seed=38, dtype=torch.bfloat16)
    t = torch.tensor([520.0]).to(torch.bfloat16)
    target = torch.randn(1, 16, H // 8, Wd // 8, generator=g)
    pe = inp["prompt_emb"].clone().requires_grad_()                       # stands for the rows the adapter wrote
    pred = ag.dit_forward(dit, [inp["latents"], inp["edit_latents"]], t, pe, use_gradient_checkpointing=True)
    loss = F.mse_loss(pred.float(), target)
    loss.backward()
    # oracle: same function of (A, B, prompt_emb) in fp32
    W32 = {k: v.to(torch.bfloat16).float() for k, v in W.items()}
    leaves = {k: v.detach().float().requires_grad_() for k, v in lora.items()}
    for name in [n for n in leaves if ".lora_A." in n]:
        mod = name.split(".lora_A.")[0]
        W32[mod + ".weight"] = W32[mod + ".weight"] + leaves[mod + ".lora_B.default.weight"] @ leaves[name]
    pe32 = inp["prompt_emb"].float().clone().requires_grad_()
    want = O.model_fn(W32, None, inp["latents"].float(), t, pe32, inp["prompt_emb_mask"], None, H, Wd, edit_latents=inp["edit_latents"].float(),
                      cuda_scalar_div=True)
    loss32 = F.mse_loss(want, target)
    loss32.backward()
    assert abs(loss.item() - loss32.item()) < 5e-3 * loss32.item()
    dead = [k for k, p in lora.items() if p.grad is None]
    assert all("txt_mlp" in k or "to_add_out" in k for k in dead)          # the text tail of the last block has no gradient
    cat = lambda d, grad=True: torch.cat([(d[k].grad if d[k].grad is not None else torch.zeros_like(d[k])).float().flatten() for k in sorted(d)])
    e_lora, e_pe = rel(cat(lora), cat(leaves)), rel(pe.grad, pe32.grad)
    print(f"emulated training step: loss {loss.item():.5f} vs {loss32.item():.5f}; LoRA grads {e_lora:.3e}; d prompt_emb {e_pe:.3e}")
    assert e_lora < 3e-2 and e_pe < 3e-2


@pytest.mark.parametrize("with_image", [False, True])
def test_text_encoder_prefill_host_logic_matches_the_library(monkeypatch, with_image):
    """QwenImageTextEncoder.edit_forward (vision tower with window reordering and per-patch KV ranges, image-token scatter, mrope tables, the decoder
    layers with fused [q|k|v] / [gate|up] packing, grouped KV heads, causal ranges) on the emulated ABI vs the installed transformers model in fp32
    (oracle/vl_oracle.py: small seeded configuration with the real structure)."""
    from oracle import vl_oracle as VO
    from physicedit_b200 import native as nv
    from physicedit_b200.text_encoder import QwenImageTextEncoder
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    hf = VO.hf_model(torch.float32)
    with torch.device("meta"):
        te = QwenImageTextEncoder(VO.native_config(), rope_mode="mrope_hf55")
    te.load_state_dict({k: v.to(torch.bfloat16) for k, v in hf.state_dict().items()}, assign=True, strict=True)
    te.eval()
    inp = VO.inputs(with_image=with_image, seed=4)
    want, _ = VO.edit_forward(hf, inp)
    got = te.edit_forward(**{k: (v.to(torch.bfloat16) if k == "pixel_values" else v) for k, v in inp.items()})[-1]
    e = rel(got, want)
    print(f"text-encoder prefill on the emulated ABI (image: {with_image}) vs transformers fp32: {e:.3e}")
    assert got.shape == want.shape and e < 2.5e-2


@pytest.mark.parametrize("fused", [True, False])
def test_text_encoder_greedy_decode_host_logic(monkeypatch, fused):
    """QwenImageTextEncoder.generate on the emulated ABI (eager steps): KV-cache rows, device-side counters (cache row, rope row with the mrope delta,
    log slot), EOS handling, the fused and the unfused decode sequence -- token ids vs the installed transformers model's greedy search (a first
    difference is accepted only where the oracle's own top-2 logits are within 2 bf16 ulps: a tie, not a bug)."""
    from oracle import vl_oracle as VO
    from physicedit_b200 import native as nv
    from physicedit_b200.text_encoder import QwenImageTextEncoder
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    hf = VO.hf_model(torch.float32)
    with torch.device("meta"):
        te = QwenImageTextEncoder(VO.native_config(), rope_mode="mrope_hf55")
    te.load_state_dict({k: v.to(torch.bfloat16) for k, v in hf.state_dict().items()}, assign=True, strict=True)
    te.eval()
    te.use_cuda_graph, te.fused_decode = False, fused
    inp = VO.inputs(with_image=True, seed=4)
    n = 12
    seq = te.generate(**{k: (v.to(torch.bfloat16) if k == "pixel_values" else v) for k, v in inp.items()}, max_new_tokens=n)
    T = inp["input_ids"].shape[1]
    assert torch.equal(seq[0, :T], inp["input_ids"][0])
    mine = seq[0, T:].tolist()
    want, top2 = VO.generate(hf, inp, n)
    want = want.tolist()
    k = next((i for i, (a, b) in enumerate(zip(mine, want)) if a != b), None)
    if k is not None:
        gap, ulp = (top2[k, 0] - top2[k, 1]).item(), 2.0 ** (math.floor(math.log2(max(abs(top2[k, 0].item()), 1e-6))) - 7)
        assert gap <= 2 * ulp, f"token {k} differs ({mine[k]} vs {want[k]}) although the oracle's top-2 logits are {gap / ulp:.1f} bf16 ulps apart"
        mine, want = mine[:k], want[:k]
    assert mine == want[:len(mine)] and len(mine) >= 4
    assert te.last_generate_stats["cuda_graph"] is False


def test_cfg_denoise_loop_host_logic_matches_the_oracle_loop(monkeypatch):
    """pipe.denoise (set_timesteps with the dynamic shift, per step two model_fn forwards whose adapter rewrites the 64 special rows of EACH branch's
    prompt in place, the conditioning table shared by the two branches, CFG combine + Euler update) on the emulated ABI vs oracle.denoise_loop in fp32
    with the bf16 timestep bookkeeping: 3 steps, 1 block."""
    from oracle import dit_oracle as O
    from physicedit_b200 import adapters, native as nv
    from physicedit_b200.dit import DiTEngine, QwenImageDiT
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    monkeypatch.setattr(adapters, "_nat", lambda t: emu)
    H = Wd = 64
    W = O.synth_weights(O.dit_param_shapes(1), seed=39)
    A = O.synth_weights(O.adapter_param_shapes(), seed=40)
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=1)
    dit.load_state_dict({k: v.to(torch.bfloat16) for k, v in W.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    eng = object.__new__(DiTEngine)
    eng.dit, eng.device, eng.nat, eng.use_cta_pair, eng.attn_flags, eng._ws, eng._rope, eng.sp = dit, torch.device("cpu"), emu, True, 0, {}, {}, None
    eng._pack()
    object.__setattr__(dit, "_engine", eng)
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit
    pipe.visual_thinking_adapter.load_state_dict({k: v.to(torch.bfloat16) for k, v in A.items()})
    pipe.visual_thinking_adapter.to(torch.bfloat16)
    pipe.cfg_streams = 1
    posi = O.synth_inputs(H, Wd, 88, seed=41, dtype=torch.bfloat16)
    nega = O.synth_inputs(H, Wd, 72, seed=42, dtype=torch.bfloat16)
    keys = ("prompt_emb", "prompt_emb_mask", "special_token_mask")
    ip, in_ = {k: posi[k].clone() for k in keys}, {k: nega[k].clone() for k in keys}
    out = pipe.denoise(posi["latents"], ip, in_, posi["edit_latents"], height=H, width=Wd, num_inference_steps=3, cfg_scale=4.0)
    ad = pipe.visual_thinking_adapter
    with torch.no_grad():
        W32 = {k: v.to(torch.bfloat16).float() for k, v in W.items()}
        A32 = {k: v.to(torch.bfloat16).float() for k, v in A.items()}
        op = {k: (posi[k].float().clone() if posi[k].is_floating_point() else posi[k]) for k in keys}
        on = {k: (nega[k].float().clone() if nega[k].is_floating_point() else nega[k]) for k in keys}
        want = O.denoise_loop(W32, A32, posi["latents"].float(), op, on, posi["edit_latents"].float(), H, Wd, 3, cuda_scalar_div=True, timestep_dtype=torch.bfloat16)
    e = rel(out, want)
    sm = posi["special_token_mask"][0]
    e_special = rel(ip["prompt_emb"][0][sm], op["prompt_emb"][0][sm])
    print(f"3-step CFG loop on the emulated ABI vs the fp32 oracle loop: latents {e:.3e}; mutated special rows {e_special:.3e}")
    assert e < 1e-2 and e_special < 1e-2
    assert torch.equal(ip["prompt_emb"][0][~sm], posi["prompt_emb"][0][~sm])               # only the special rows were written (:1336)
    names = [c[0] for c in emu.calls]
    assert names.count("pe_cfg_euler_step") == 3 and names.count("pe_special_blend_scatter") == 6
    assert names.count("pe_timestep_embedding") == 3                                        # conditioning computed once per timestep, shared by the CFG pair


def test_feature_extractor_host_logic_matches_the_aux_oracle(monkeypatch):
    """The pseudo-target branch of a training sample (rows a14-a16) on the emulated ABI vs oracle/aux_oracle.py in fp32: DINOv2-with-registers (im2col patch
    embed with K padded to 592, bicubic position table, CLS / register placement, LayerScale through the gate-residual epilogue, non-affine final norm,
    5 leading tokens dropped), both perceiver resamplers, their adapters, frame-index embeddings, middle - source deltas."""
    from oracle import aux_oracle as AO
    from physicedit_b200 import adapters, native as nv
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    monkeypatch.setattr(adapters, "_nat", lambda t: emu)
    P = AO.aux_synth(seed=5, dtype=torch.bfloat16)
    ain = AO.aux_inputs(seed=6, n_mid=2, lat_hw=(16, 16), dtype=torch.bfloat16)
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, dinov2_config=dict(hidden=768, layers=12, heads=12))
    pipe.dinov2.encoder.load_state_dict({k: v for k, v in P["dinov2"].items() if not k.startswith("layernorm.")}, strict=True)
    for name in ("dino_resampler", "vae_resampler", "dino_resampler_adapter", "vae_resampler_adapter", "dino_time_embed", "vae_time_embed"):
        getattr(pipe, name).load_state_dict(P[name])
    pipe.to(torch.bfloat16)
    pipe.eval()
    Pf = {k: {n: t.float() for n, t in v.items()} for k, v in P.items()}
    with torch.no_grad():
        d_nat = pipe.dinov2(ain["dino_middle"])
        d_ref = AO.dinov2_with_norm(Pf["dinov2"], ain["dino_middle"].float())
        assert d_nat.shape == d_ref.shape == (2, 256, 768) and rel(d_nat, d_ref) < 2e-2
        out = pipe.physical_visual_embeddings(**ain)
        ed, ev = AO.physical_visual_embeddings(Pf, **{k: v.float() for k, v in ain.items()})
    e_d, e_v = rel(out["pseudo_special_emb_dino"], ed), rel(out["pseudo_special_emb_vae"], ev)
    print(f"feature extractors on the emulated ABI vs the fp32 aux oracle: dinov2 {rel(d_nat, d_ref):.3e}; targets dino {e_d:.3e} vae {e_v:.3e}")
    assert out["pseudo_special_emb_dino"].shape == (1, 64, 3584) and e_d < 3e-2 and e_v < 3e-2


def test_denoise_with_unmerged_lora_matches_the_merged_weights(monkeypatch):
    """An evaluation in the middle of a training run (train_physicedit.py:39-169 calls `pipe(..., is_train=False)` with the LoRA wrappers still in
    the DiT): the loop's forwards go through the un-merged path and must land in the buffers the loop combines -- same latents as after
    `merge_lora` up to bf16 rounding of `W + B A` vs `W x + B (A x)`."""
    from oracle import dit_oracle as O
    from physicedit_b200 import adapters, autograd, native as nv
    from physicedit_b200.dit import DiTEngine, QwenImageDiT
    from physicedit_b200.lora import inject_lora, merge_lora
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    monkeypatch.setattr(adapters, "_nat", lambda t: emu)
    monkeypatch.setattr(autograd, "_nat", lambda t: emu)
    autograd.weight_transposes.clear()
    H = Wd = 64
    W = O.synth_weights(O.dit_param_shapes(1), seed=61)
    A = O.synth_weights(O.adapter_param_shapes(), seed=62)
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=1)
    dit.load_state_dict({k: v.to(torch.bfloat16) for k, v in W.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)

    def bind():
        eng = object.__new__(DiTEngine)
        eng.dit, eng.device, eng.nat, eng.use_cta_pair, eng.attn_flags, eng._ws, eng._rope, eng.sp = dit, torch.device("cpu"), emu, True, 0, {}, {}, None
        eng._pack()
        object.__setattr__(dit, "_engine", eng)
    bind()
    inject_lora(dit, ["to_q", "to_out.0", "img_mlp.net.2", "img_mod.1", "add_v_proj"], r=8)
    g = torch.Generator().manual_seed(63)
    for name, p in dit.named_parameters():
        if "lora_" in name:
            p.data = (torch.randn(p.shape, generator=g) * (1.0 / math.sqrt(p.shape[1]))).bfloat16()
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit
    pipe.visual_thinking_adapter.load_state_dict({k: v.to(torch.bfloat16) for k, v in A.items()})
    pipe.visual_thinking_adapter.to(torch.bfloat16)
    pipe.cfg_streams = 1
    pipe.eval()
    posi = O.synth_inputs(H, Wd, 88, seed=64, dtype=torch.bfloat16)
    nega = O.synth_inputs(H, Wd, 72, seed=65, dtype=torch.bfloat16)
    keys = ("prompt_emb", "prompt_emb_mask", "special_token_mask")
    run = lambda: pipe.denoise(posi["latents"], {k: posi[k].clone() for k in keys}, {k: nega[k].clone() for k in keys}, posi["edit_latents"], height=H, width=Wd,
                               num_inference_steps=2, cfg_scale=4.0)
    unmerged = run()
    assert torch.isfinite(unmerged.float()).all()
    merge_lora(dit)
    bind()
    merged = run()
    e = rel(unmerged, merged)
    print(f"2-step CFG loop, un-merged LoRA vs merged weights: {e:.3e}")
    assert e < 2e-2 and not torch.equal(merged, posi["latents"])
