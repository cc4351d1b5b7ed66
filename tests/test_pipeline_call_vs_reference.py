"""`pipe(prompt, edit_image=..., is_train=False)` end to end: the REFERENCE pipeline object (its own __init__, units, loop, stock PyTorch DiT on the
CPU, fp32 and bf16) against this package's pipeline (units, native loop on the emulated C ABI), same weights, same stub text encoder / VAE, same
tokenizer / processor files.  Compared: the latents each pipeline hands to `vae.decode` -- this package has to be as close to the reference's
fp32 result as the reference's own bf16 run is (floor + 1e-3), the protocol of the GPU parity tests.  What this adds to the per-part tests: the
glue of `__call__` itself (how the unit outputs reach `model_fn`, the noise draw, height / width rounding, scheduler shift, CFG, the persistent
per-branch prompt mutation)."""
import os
import sys

import pytest
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abi_emulator import EmulatedNative  # noqa: E402
from oracle import dit_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402
from test_units_vs_reference import PROC, TOK, picture  # noqa: E402

pytestmark = pytest.mark.skipif(ref_import.reference_root() is None or not os.path.isdir(TOK), reason="no reference tree / tokenizer files on this machine")


class StubVL:
    def __init__(self, reply_ids):
        self.reply_ids = reply_ids

    def edit_forward(self, input_ids=None, **kw):
        g = torch.Generator().manual_seed(int(input_ids.sum()) % 9973)
        return (torch.randn(input_ids.shape[0], input_ids.shape[1], 3584, generator=g) * 2.0,)

    def generate(self, input_ids=None, **kw):
        return torch.cat([input_ids, self.reply_ids.unsqueeze(0)], dim=1)


class StubVAE:
    def __init__(self):
        self.decoded = []

    def encode(self, x, **kw):
        g = torch.Generator().manual_seed(int(x.shape[-1]) * 31 + int(x.shape[-2]))
        return torch.randn(x.shape[0], 16, x.shape[2] // 8, x.shape[3] // 8, generator=g).to(x.dtype)

    def decode(self, z, **kw):
        self.decoded.append(z.detach().float().cpu().clone())
        return torch.tanh(z[:, :3].float()).repeat_interleave(8, dim=2).repeat_interleave(8, dim=3).to(z.dtype)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_call_matches_the_reference_pipeline(tmp_path, monkeypatch):
    from transformers import Qwen2Tokenizer, Qwen2VLProcessor
    from physicedit_b200 import adapters, native as nv
    from physicedit_b200.dit import DiTEngine, QwenImageDiT
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    Wsd = O.synth_weights(O.dit_param_shapes(1), seed=71, dtype=torch.bfloat16)
    Asd = O.synth_weights(O.adapter_param_shapes(), seed=72, dtype=torch.bfloat16)
    tok = Qwen2Tokenizer.from_pretrained(TOK)
    base = Qwen2VLProcessor.from_pretrained(PROC)
    reply = tok('{"middle_transition_prompt": "The glass tips and the water runs out."}', return_tensors="pt").input_ids[0]
    image = picture(320, 256, 5)
    call = dict(prompt="tip the glass over", edit_image=image, seed=3, num_inference_steps=3, height=64, width=96, is_train=False,
                edit_image_auto_resize=False)            # the edit image keeps its 320 x 256 (320 tokens instead of 4096): a CPU-sized request

    # img2img + inpainting (:612-613, BasePipeline.step): start from a noised `input_image`, keep it outside the mask
    hole = Image.new("L", (96, 64), 0)
    hole.paste(255, (24, 16, 72, 48))
    inpaint_call = dict(call, input_image=picture(96, 64, 6), denoising_strength=0.8, inpaint_mask=hole.convert("RGB"), inpaint_blur_size=1, inpaint_blur_sigma=0.8, seed=4)

    def processor():
        return Qwen2VLProcessor(image_processor=base.image_processor, tokenizer=Qwen2Tokenizer.from_pretrained(TOK), video_processor=base.video_processor,
                                chat_template=base.chat_template)
    ref_import.tiny_dinov2_folder(str(tmp_path / "dino"))
    out = {}
    with ref_import.ReferenceModules() as ref:
        for dtype in (torch.float32, torch.bfloat16):
            pipe = ref.phys.QwenImagePhysicPipeline(device="cpu", torch_dtype=dtype, dinov2_path=str(tmp_path / "dino"))
            pipe.dit = ref_import.build_reference_dit(ref, Wsd, 1, dtype, "cpu")
            pipe.visual_thinking_adapter.load_state_dict({k: v.to(dtype) for k, v in Asd.items()})
            pipe.visual_thinking_adapter.to(dtype)
            pipe.text_encoder, pipe.vae, pipe.tokenizer, pipe.processor = StubVL(reply), StubVAE(), tok, processor()
            pipe.processor.tokenizer.add_special_tokens({"additional_special_tokens": ["<begin_of_img>", "<end_of_img>"] + [f"<img{i}>" for i in range(64)]})
            pipe.boi_token_id = pipe.processor.tokenizer.convert_tokens_to_ids("<begin_of_img>")            # from_pretrained :528-539
            pipe.eoi_token_id = pipe.processor.tokenizer.convert_tokens_to_ids("<end_of_img>")
            if dtype == torch.float32:
                # NoiseInitializer draws in the pipeline dtype (:689): an fp32 pipeline would start from different noise.  Give the fp32 truth the bf16 draw.
                draw = pipe.generate_noise
                pipe.generate_noise = lambda shape, **kw: draw(shape, **dict(kw, rand_torch_dtype=torch.bfloat16)).float()
            img = pipe(**call)
            out[dtype] = (pipe.vae.decoded[-1], img)
            if dtype == torch.bfloat16:
                pipe(**inpaint_call)
                out["inpaint"] = pipe.vae.decoded[-1]
    # this package: same weights, native loop on the emulated ABI
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    monkeypatch.setattr(adapters, "_nat", lambda t: emu)
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=1)
    dit.load_state_dict({k: v.clone() for k, v in Wsd.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit
    pipe.visual_thinking_adapter.load_state_dict(Asd)
    pipe.text_encoder, pipe.vae = StubVL(reply), StubVAE()
    pipe.batch_cfg_generation = False
    pipe.attach_tokenizer(tokenizer=tok, processor=processor())
    pipe.to(torch.bfloat16)
    pipe.cfg_streams = 1
    eng = object.__new__(DiTEngine)
    eng.dit, eng.device, eng.nat, eng.use_cta_pair, eng.attn_flags, eng._ws, eng._rope, eng.sp = dit, torch.device("cpu"), emu, True, 0, {}, {}, None
    eng._pack()
    object.__setattr__(dit, "_engine", eng)
    img = pipe(**call)
    got = pipe.vae.decoded[-1]
    l32, l16 = out[torch.float32][0], out[torch.bfloat16][0]
    floor, err = rel(l16, l32), rel(got, l32)
    print(f"pipe(...) latents: this package vs reference fp32 {err:.3e}; reference bf16 vs fp32 {floor:.3e}; this package vs reference bf16 {rel(got, l16):.3e}")
    assert got.shape == l32.shape == (1, 16, 8, 12) and floor < 0.1 and err <= floor + 1e-3 and rel(got, l16) < 1e-2, (err, floor, rel(got, l16))
    assert isinstance(img, Image.Image) and img.size == out[torch.float32][1].size == (96, 64)
    names = [c[0] for c in emu.calls]
    assert names.count("pe_cfg_euler_step") == 3 and names.count("pe_special_blend_scatter") == 6
    pipe(**inpaint_call)
    e_inpaint = rel(pipe.vae.decoded[-1], out["inpaint"])
    print(f"pipe(..., input_image, inpaint_mask) latents: this package vs reference bf16 {e_inpaint:.3e}")
    assert e_inpaint < 1e-2


def test_training_loss_matches_the_reference_pipeline(tmp_path, monkeypatch):
    """`pipe.training_loss(global_step=, **models, **inputs)` as scripts/train/train_physicedit.py:309-310 calls it: the reference pipeline object (stock
    PyTorch autograd on the CPU, bf16) against this package's (autograd Functions on the emulated C ABI), same weights, same `inputs`, same RNG stream for
    the two draws (timestep id, noise; :314-317).  Trainable here: the dual adapter and the two pseudo targets (the reference's un-merged LoRA needs
    peft, which is absent -- that part is checked against the oracle).  Compared: the loss, `special_token_loss`, every adapter gradient, d loss / d targets."""
    from physicedit_b200 import adapters, autograd, native as nv
    from physicedit_b200.dit import DiTEngine, QwenImageDiT
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    Wsd = O.synth_weights(O.dit_param_shapes(1), seed=81, dtype=torch.bfloat16)
    Asd = O.synth_weights(O.adapter_param_shapes(), seed=82, dtype=torch.bfloat16)
    H, Wd, T = 64, 96, 120
    inp = O.synth_inputs(H, Wd, T, seed=83, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(84)
    base = dict(input_latents=torch.randn(1, 16, H // 8, Wd // 8, generator=g).bfloat16(), height=H, width=Wd, edit_latents=inp["edit_latents"],
                prompt_emb_mask=inp["prompt_emb_mask"], special_token_mask=inp["special_token_mask"], use_gradient_checkpointing=True,
                use_gradient_checkpointing_offload=False, cfg_scale=1, is_train=True)
    gt = [(torch.randn(1, 64, 3584, generator=g) * 0.5).bfloat16() for _ in range(2)]

    def run(pipe):
        pipe.scheduler.set_timesteps(1000, training=True)
        pipe.freeze_except(["visual_thinking_adapter"])
        targets = [t.clone().requires_grad_() for t in gt]
        inputs = dict(base, prompt_emb=inp["prompt_emb"].clone(), pseudo_special_emb_dino=targets[0], pseudo_special_emb_vae=targets[1])
        models = {name: getattr(pipe, name) for name in pipe.in_iteration_models}
        torch.manual_seed(4321)
        loss = pipe.training_loss(global_step=10, **models, **inputs)
        loss.backward()
        grads = {n: p.grad.detach().float().clone() for n, p in pipe.visual_thinking_adapter.named_parameters()}
        return loss.item(), pipe.special_token_loss, grads, [t.grad.float().clone() for t in targets]

    ref_import.tiny_dinov2_folder(str(tmp_path / "dino"))
    with ref_import.ReferenceModules() as ref:
        rp = ref.phys.QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, dinov2_path=str(tmp_path / "dino"))
        rp.dit = ref_import.build_reference_dit(ref, Wsd, 1, torch.bfloat16, "cpu")
        rp.visual_thinking_adapter.load_state_dict(Asd)
        rp.visual_thinking_adapter.to(torch.bfloat16)
        want = run(rp)
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    monkeypatch.setattr(adapters, "_nat", lambda t: emu)
    monkeypatch.setattr(autograd, "_nat", lambda t: emu)
    autograd.weight_transposes.clear()
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=1)
    dit.load_state_dict({k: v.clone() for k, v in Wsd.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit
    pipe.visual_thinking_adapter.load_state_dict(Asd)
    pipe.to(torch.bfloat16)
    eng = object.__new__(DiTEngine)
    eng.dit, eng.device, eng.nat, eng.use_cta_pair, eng.attn_flags, eng._ws, eng._rope, eng.sp = dit, torch.device("cpu"), emu, True, 0, {}, {}, None
    eng._pack()
    object.__setattr__(dit, "_engine", eng)
    got = run(pipe)
    cat = lambda d: torch.cat([d[k].flatten() for k in sorted(d)])
    e_ad, e_t = rel(cat(got[2]), cat(want[2])), max(rel(a, b) for a, b in zip(got[3], want[3]))
    print(f"training_loss: {got[0]:.5f} vs reference {want[0]:.5f}; special_token_loss {got[1]:.5f} vs {want[1]:.5f}; adapter grads {e_ad:.3e}; d targets {e_t:.3e}")
    assert set(got[2]) == set(want[2]) and abs(got[0] - want[0]) < 1e-2 * abs(want[0]) and abs(got[1] - want[1]) < 1e-2 * abs(want[1])
    assert e_ad < 3e-2 and e_t < 3e-2


def test_direct_distill_loss_matches_the_reference_pipeline(tmp_path, monkeypatch):
    """`pipe.direct_distill_loss(**inputs)` (:332-340, task "direct_distill" of the train script): the whole 2-step sampler under autograd -- per step one
    model_fn forward (adapter rows rewritten in place) and `pipe.step` -- regressed on `input_latents`; reference pipeline (stock autograd) vs this package
    (emulated ABI), same weights and inputs.  Trainable: the DiT itself (full-weight gradients stand in for the un-merged LoRA, which needs peft on the
    reference side); the adapter stays frozen -- with a trainable adapter the reference's own multi-step loop fails in autograd (the in-place write of
    :1336 bumps the version of a tensor the previous step's graph saved).  Compared: the loss and the gradients of five DiT weights."""
    from physicedit_b200 import adapters, autograd, native as nv
    from physicedit_b200.dit import DiTEngine, QwenImageDiT
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    Wsd = O.synth_weights(O.dit_param_shapes(1), seed=91, dtype=torch.bfloat16)
    Asd = O.synth_weights(O.adapter_param_shapes(), seed=92, dtype=torch.bfloat16)
    H, Wd, T = 64, 64, 96
    inp = O.synth_inputs(H, Wd, T, seed=93, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(94)
    gt = [(torch.randn(1, 64, 3584, generator=g) * 0.5).bfloat16() for _ in range(2)]
    target = torch.randn(1, 16, H // 8, Wd // 8, generator=g).bfloat16()

    def run(pipe):
        pipe.freeze_except(["dit"])
        inputs = dict(latents=inp["latents"].clone(), input_latents=target, height=H, width=Wd, edit_latents=inp["edit_latents"], prompt_emb=inp["prompt_emb"].clone(),
                      prompt_emb_mask=inp["prompt_emb_mask"], special_token_mask=inp["special_token_mask"], num_inference_steps=2, use_gradient_checkpointing=False,
                      pseudo_special_emb_dino=gt[0], pseudo_special_emb_vae=gt[1], is_train=True)
        loss = pipe.direct_distill_loss(**inputs)
        loss.backward()
        names = ("img_in.weight", "transformer_blocks.0.attn.to_q.weight", "transformer_blocks.0.attn.add_k_proj.bias", "transformer_blocks.0.img_mlp.net.2.weight",
                 "transformer_blocks.0.img_mod.1.weight", "proj_out.weight")
        params = dict(pipe.dit.named_parameters())
        return loss.item(), {n: params[n].grad.detach().float().clone() for n in names}

    ref_import.tiny_dinov2_folder(str(tmp_path / "dino"))
    with ref_import.ReferenceModules() as ref:
        rp = ref.phys.QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, dinov2_path=str(tmp_path / "dino"))
        rp.dit = ref_import.build_reference_dit(ref, Wsd, 1, torch.bfloat16, "cpu")
        rp.visual_thinking_adapter.load_state_dict(Asd)
        rp.visual_thinking_adapter.to(torch.bfloat16)
        want = run(rp)
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    monkeypatch.setattr(adapters, "_nat", lambda t: emu)
    monkeypatch.setattr(autograd, "_nat", lambda t: emu)
    autograd.weight_transposes.clear()
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=1)
    dit.load_state_dict({k: v.clone() for k, v in Wsd.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit
    pipe.visual_thinking_adapter.load_state_dict(Asd)
    pipe.to(torch.bfloat16)
    eng = object.__new__(DiTEngine)
    eng.dit, eng.device, eng.nat, eng.use_cta_pair, eng.attn_flags, eng._ws, eng._rope, eng.sp = dit, torch.device("cpu"), emu, True, 0, {}, {}, None
    eng._pack()
    object.__setattr__(dit, "_engine", eng)
    got = run(pipe)
    cat = lambda d: torch.cat([d[k].flatten() for k in sorted(d)])
    e = rel(cat(got[1]), cat(want[1]))
    print(f"direct_distill_loss: {got[0]:.5f} vs reference {want[0]:.5f}; DiT weight grads {e:.3e}")
    assert abs(got[0] - want[0]) < 1e-2 * abs(want[0]) and e < 5e-2


def test_training_loss_with_a_trainable_blockwise_controlnet_matches_the_reference(tmp_path, monkeypatch):
    """SURVEY 8f5 under autograd: `training_loss` with `blockwise_controlnet_conditioning` / `_inputs` (what QwenImageUnit_BlockwiseControlNet produces) and the
    controlnet as the trainable model -- reference pipeline (its QwenImageBlockwiseMultiControlNet, stock autograd) vs this package (autograd Functions on the
    emulated ABI): loss and every controlnet gradient.  Two requests, one of them outside its [end, start] window at this progress."""
    from physicedit_b200 import adapters, autograd, native as nv
    from physicedit_b200.compat import ControlNetInput
    from physicedit_b200.controlnet import QwenImageBlockWiseControlNet, QwenImageBlockwiseMultiControlNet
    from physicedit_b200.dit import DiTEngine, QwenImageDiT
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    Wsd = O.synth_weights(O.dit_param_shapes(1), seed=101, dtype=torch.bfloat16)
    Asd = O.synth_weights(O.adapter_param_shapes(), seed=102, dtype=torch.bfloat16)
    H, Wd, T = 64, 64, 80
    inp = O.synth_inputs(H, Wd, T, seed=103, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(104)
    cn = QwenImageBlockWiseControlNet(num_layers=1)
    csd = {k: (torch.randn(v.shape, generator=g) * (0.02 if v.dim() > 1 else 0.5) + (1.0 if "rms" in k else 0.0)).bfloat16() for k, v in cn.state_dict().items()}
    conds = [torch.randn(1, 16, H // 8, Wd // 8, generator=g).bfloat16() for _ in range(2)]
    cinputs = [ControlNetInput(controlnet_id=0, scale=0.7), ControlNetInput(controlnet_id=0, scale=1.3, start=0.4, end=0.0)]       # progress 1.0 at step 0: inactive
    gt = [(torch.randn(1, 64, 3584, generator=g) * 0.5).bfloat16() for _ in range(2)]
    target = torch.randn(1, 16, H // 8, Wd // 8, generator=g).bfloat16()

    def run(pipe, multi):
        pipe.blockwise_controlnet = multi
        pipe.scheduler.set_timesteps(1000, training=True)
        pipe.freeze_except(["blockwise_controlnet"])
        inputs = dict(input_latents=target, height=H, width=Wd, edit_latents=inp["edit_latents"], prompt_emb=inp["prompt_emb"].clone(), prompt_emb_mask=inp["prompt_emb_mask"],
                      special_token_mask=inp["special_token_mask"], use_gradient_checkpointing=True, use_gradient_checkpointing_offload=False, cfg_scale=1, is_train=True,
                      pseudo_special_emb_dino=gt[0], pseudo_special_emb_vae=gt[1], blockwise_controlnet_conditioning=[c.clone() for c in conds],
                      blockwise_controlnet_inputs=cinputs, progress_id=0, num_inference_steps=5)
        models = {name: getattr(pipe, name) for name in pipe.in_iteration_models}
        torch.manual_seed(77)
        loss = pipe.training_loss(global_step=3, **models, **inputs)
        loss.backward()
        return loss.item(), {n: p.grad.detach().float().clone() for n, p in multi.named_parameters() if p.grad is not None}

    ref_import.tiny_dinov2_folder(str(tmp_path / "dino"))
    with ref_import.ReferenceModules() as ref:
        import importlib
        rcn = importlib.import_module("diffsynth.models.qwen_image_controlnet").QwenImageBlockWiseControlNet(num_layers=1)
        rcn.load_state_dict(csd)
        rp = ref.phys.QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, dinov2_path=str(tmp_path / "dino"))
        rp.dit = ref_import.build_reference_dit(ref, Wsd, 1, torch.bfloat16, "cpu")
        rp.visual_thinking_adapter.load_state_dict(Asd)
        rp.visual_thinking_adapter.to(torch.bfloat16)
        want = run(rp, ref.phys.QwenImageBlockwiseMultiControlNet([rcn.to(torch.bfloat16)]))
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    monkeypatch.setattr(adapters, "_nat", lambda t: emu)
    monkeypatch.setattr(autograd, "_nat", lambda t: emu)
    autograd.weight_transposes.clear()
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=1)
    dit.load_state_dict({k: v.clone() for k, v in Wsd.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit
    pipe.visual_thinking_adapter.load_state_dict(Asd)
    cn.load_state_dict(csd)
    pipe.to(torch.bfloat16)
    eng = object.__new__(DiTEngine)
    eng.dit, eng.device, eng.nat, eng.use_cta_pair, eng.attn_flags, eng._ws, eng._rope, eng.sp = dit, torch.device("cpu"), emu, True, 0, {}, {}, None
    eng._pack()
    object.__setattr__(dit, "_engine", eng)
    got = run(pipe, QwenImageBlockwiseMultiControlNet([cn.to(torch.bfloat16)]))
    assert set(got[1]) == set(want[1]) and len(got[1]) == 8              # img_in (2) + x_rms, y_rms, input_proj (2), output_proj (2) of block 0
    cat = lambda d: torch.cat([d[k].flatten() for k in sorted(d)])
    e = rel(cat(got[1]), cat(want[1]))
    print(f"training_loss with a blockwise controlnet: {got[0]:.5f} vs reference {want[0]:.5f}; controlnet grads {e:.3e}")
    assert abs(got[0] - want[0]) < 1e-2 * abs(want[0]) and e < 3e-2
