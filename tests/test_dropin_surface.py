"""The reference's scripts resolve against this framework (CPU; no kernels run here).

* `compat.install()` serves every `diffsynth.*` import of scripts/inference/validate.py and scripts/train/train_physicedit.py,
* validate.py's own `load_finetuned_into_pipe` (:33-65) -- LoRA keys -> `pipe.load_lora`, `pipe.*` keys -> `pipe.load_state_dict(strict=False)`
  -- works on this pipeline object with a synthetic checkpoint in the training script's key layout,
* `pipe(prompt, edit_image=PIL, is_train=False)` walks the unit list through `pipe.unit_runner` with the reference's contracts
  (processor / tokenizer from the reference tree, a stub VL model and a stub VAE) and reaches the denoise loop with the tensors the
  reference would hand to `model_fn` (shapes, masks, the 64 special positions, bf16 CPU-generator noise).
The reference tree (for the scripts and the vendored tokenizer files) is optional: tests that need it skip without it.
"""
import importlib.util
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "scripts", "inference", "validate.py")), reason="reference tree not present")


@pytest.fixture()
def compat_installed():
    from physicedit_b200 import compat
    saved = {k: v for k, v in sys.modules.items() if k == "diffsynth" or k.startswith("diffsynth.")}
    for k in saved:
        del sys.modules[k]
    compat.install()
    yield
    for k in [k for k in sys.modules if k == "diffsynth" or k.startswith("diffsynth.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def _cpu_pipe(layers=1):
    from physicedit_b200.dit import QwenImageDiT
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = QwenImageDiT(num_layers=layers).to(torch.bfloat16)
    return pipe


def test_every_diffsynth_import_of_the_scripts_resolves(compat_installed):
    src = """
from diffsynth import load_state_dict
from diffsynth.pipelines.qwen_image_physical import QwenImagePhysicPipeline, ModelConfig
from diffsynth.pipelines.flux_image_new import ControlNetInput
from diffsynth.trainers.utils import DiffusionTrainingModule, ModelLogger, qwen_image_parser, launch_training_task, launch_data_process_task, PhysicalEditingDataset
from diffsynth.trainers.unified_dataset import UnifiedDataset
from diffsynth.utils import PipelineUnit, PipelineUnitRunner
"""
    ns = {}
    exec(src, ns)                                                   # lines 2-6 of train_physicedit.py, :17-18 of validate.py
    args = ns["qwen_image_parser"]().parse_args(["--dataset_base_path", "x", "--dinov2_path", "y", "--lora_rank", "128", "--use_gradient_checkpointing"])
    assert args.remove_prefix_in_ckpt == "pipe.dit." and args.lora_rank == 128 and args.use_gradient_checkpointing and args.resume_type == "auto"
    from physicedit_b200.unified_dataset import UnifiedDataset
    assert ns["UnifiedDataset"] is UnifiedDataset                   # imported by the train script (:6): tests/test_datasets.py
    from physicedit_b200.datasets import PhysicalEditingDataset
    assert ns["PhysicalEditingDataset"] is PhysicalEditingDataset   # the dataset the script builds (:420): tests/test_datasets.py
    pipe = _cpu_pipe()
    names = [type(u).__name__ for u in pipe.units]
    assert names == ["QwenImageUnit_ShapeChecker", "QwenImageUnit_NoiseInitializer", "QwenImageUnit_InputImageEmbedder", "QwenImageUnit_Inpaint",
                     "QwenImageUnit_EditImageEmbedder", "QwenImageUnit_ContextImageEmbedder", "QwenImageUnit_PhysicalVisualEmbedder",
                     "QwenImageUnit_PhysicalVerbalEmbedder", "QwenImageUnit_PromptEmbedder", "QwenImageUnit_EntityControl",
                     "QwenImageUnit_BlockwiseControlNet"]                # qwen_image_physical.py:233-245, same order
    assert pipe.in_iteration_models == ("dit", "blockwise_controlnet", "visual_thinking_adapter") and callable(pipe.unit_runner)


def test_training_module_exports_the_checkpoint_layout_validate_py_splits():
    from physicedit_b200.trainers import DiffusionTrainingModule

    class M(DiffusionTrainingModule):
        def __init__(self):
            super().__init__()
            self.pipe = _cpu_pipe()
    m = M()
    m.pipe.freeze_except([])
    blk = m.pipe.dit.transformer_blocks[0]
    blk.attn.to_q.register_parameter("lora_A_default", torch.nn.Parameter(torch.zeros(4, 3072)))     # stands in for a PEFT-injected tensor
    for p in m.pipe.visual_thinking_adapter.parameters():
        p.requires_grad_(True)
    sd = m.export_trainable_state_dict(m.state_dict(), remove_prefix="pipe.dit.")
    assert "transformer_blocks.0.attn.to_q.lora_A_default" in sd                                       # `pipe.dit.` stripped
    assert "pipe.visual_thinking_adapter.head_dino.0.weight" in sd and len(sd) == 1 + 8                 # adapter keys keep `pipe.`
    assert m.mapping_lora_state_dict({"a.lora_A.weight": 1, "a.lora_B.default.weight": 2, "b.bias": 3}) == {"a.lora_A.default.weight": 1, "a.lora_B.default.weight": 2}


@needs_ref
def test_validate_py_loads_a_finetuned_checkpoint_into_this_pipeline(compat_installed, tmp_path):
    spec = importlib.util.spec_from_file_location("ref_validate", os.path.join(REF, "scripts", "inference", "validate.py"))
    sys.dont_write_bytecode = True
    validate = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(validate)                               # the reference script itself, imports served by compat.install()
    assert validate.QwenImagePhysicPipeline.__module__.startswith("physicedit_b200")
    pipe = _cpu_pipe()
    g = torch.Generator().manual_seed(0)
    w_q = pipe.dit.transformer_blocks[0].attn.to_q.weight.detach().clone()
    w_mod = pipe.dit.transformer_blocks[0].img_mod[1].weight.detach().clone()
    ck = {}
    for name, (o, i) in (("transformer_blocks.0.attn.to_q", (3072, 3072)), ("transformer_blocks.0.img_mod.1", (18432, 3072))):
        ck[f"{name}.lora_A.default.weight"] = (torch.randn(8, i, generator=g) * 0.05).bfloat16()         # key layout of train_multigpu.sh:27-31
        ck[f"{name}.lora_B.default.weight"] = (torch.randn(o, 8, generator=g) * 0.05).bfloat16()
    ad_new = {k: torch.randn(v.shape, generator=g).bfloat16() for k, v in pipe.visual_thinking_adapter.state_dict().items()}
    ck.update({f"pipe.visual_thinking_adapter.{k}": v for k, v in ad_new.items()})
    ck["stray.key"] = torch.zeros(1)                                                                     # not `pipe.`-prefixed: ignored (:57-59)
    from safetensors.torch import save_file
    path = str(tmp_path / "step-8000.safetensors")
    save_file(ck, path)
    with pytest.raises(FileNotFoundError):
        validate.load_finetuned_into_pipe(pipe, str(tmp_path / "missing.safetensors"))
    validate.load_finetuned_into_pipe(pipe, path)
    for name, w0 in (("transformer_blocks.0.attn.to_q", w_q), ("transformer_blocks.0.img_mod.1", w_mod)):
        want = w0 + torch.mm(ck[f"{name}.lora_B.default.weight"], ck[f"{name}.lora_A.default.weight"])    # bf16 mm + bf16 add, alpha = 1
        got = dict(pipe.dit.named_modules())[name].weight
        assert torch.equal(got, want)
    for k, v in ad_new.items():
        assert torch.equal(pipe.visual_thinking_adapter.state_dict()[k], v)
    assert validate.resize_image(__import__("PIL.Image", fromlist=["Image"]).new("RGB", (640, 480))).size == (1184, 896)


class _StubVL:
    """Stands in for the Qwen2.5-VL encoder with its two entry points (models/qwen_image_text_encoder_withdecode.py:188, HF generate)."""

    def __init__(self, reply_ids):
        self.reply_ids, self.calls = reply_ids, []

    def edit_forward(self, input_ids=None, attention_mask=None, pixel_values=None, image_grid_thw=None, output_hidden_states=True, **kw):
        self.calls.append(("edit_forward", input_ids.shape, None if pixel_values is None else tuple(pixel_values.shape)))
        g = torch.Generator().manual_seed(int(input_ids.sum()) % 1000)
        return (torch.randn(input_ids.shape[0], input_ids.shape[1], 3584, generator=g),)

    def generate(self, input_ids=None, max_new_tokens=None, **kw):
        self.calls.append(("generate", input_ids.shape, max_new_tokens))
        return torch.cat([input_ids, self.reply_ids.unsqueeze(0)], dim=1)


class _StubVAE:
    def encode(self, x, **kw):
        return torch.zeros(x.shape[0], 16, x.shape[2] // 8, x.shape[3] // 8, dtype=x.dtype)


@needs_ref
def test_call_walks_the_units_and_reaches_the_denoise_loop(monkeypatch):
    from PIL import Image
    from transformers import Qwen2Tokenizer, Qwen2VLProcessor
    from physicedit_b200.pipeline import ModelConfig
    tok_dir = os.path.join(REF, "DiffSynth-Studio", "models", "Qwen", "Qwen-Image", "tokenizer")
    proc_dir = os.path.join(REF, "DiffSynth-Studio", "models", "Qwen", "Qwen-Image-Edit", "processor")
    tok = Qwen2Tokenizer.from_pretrained(tok_dir)
    base = Qwen2VLProcessor.from_pretrained(proc_dir)
    # the vendored processor folder carries no vocabulary (its tokenizer knows only the added tokens): give it the real one
    proc = Qwen2VLProcessor(image_processor=base.image_processor, tokenizer=Qwen2Tokenizer.from_pretrained(tok_dir),
                            video_processor=base.video_processor, chat_template=base.chat_template)
    pipe = _cpu_pipe()
    reply = tok('{"middle_transition_prompt": "The cup tips over and the water spreads."}', return_tensors="pt").input_ids[0]
    pipe.text_encoder = _StubVL(reply)
    pipe.vae = _StubVAE()
    pipe.attach_tokenizer(tokenizer=tok, processor=proc)
    assert pipe.eoi_token_id - pipe.boi_token_id == 1 and pipe.boi_token_id >= len(tok)                 # the 66 added tokens (:528-539)
    seen = {}

    def fake_denoise(latents, inputs_posi, inputs_nega, edit_latents=None, context_latents=None, **kw):
        seen.update(latents=latents, posi=inputs_posi, nega=inputs_nega, edit_latents=edit_latents, kw=kw)
        return latents
    monkeypatch.setattr(pipe, "denoise", fake_denoise)
    img = Image.new("RGB", (800, 600), (90, 120, 30))
    out = pipe("knock the cup over", edit_image=img, seed=7, num_inference_steps=4, height=500, width=760, is_train=False, output_type="latent")
    # ShapeChecker rounds up to multiples of 16; NoiseInitializer draws bf16 on the CPU generator (NOT fp32 -> bf16)
    assert out.shape == (1, 16, 64, 96) and seen["kw"]["height"] == 512 and seen["kw"]["width"] == 768
    want = torch.randn((1, 16, 64, 96), generator=torch.Generator("cpu").manual_seed(7), dtype=torch.bfloat16)
    assert torch.equal(out, want)
    # EditImageEmbedder: auto-resized to ~1024^2 on a 32-pixel grid before the VAE
    assert seen["edit_latents"].shape == (1, 16, 896 // 8, 1184 // 8)
    # VerbalEmbedder generated once per CFG branch (<= 1000 new tokens), PromptEmbedder encoded once per branch with the image
    kinds = [c[0] for c in pipe.text_encoder.calls]
    assert kinds == ["generate", "generate", "edit_forward", "edit_forward"] and pipe.text_encoder.calls[0][2] == 1000
    posi, nega = seen["posi"], seen["nega"]
    T = posi["prompt_emb"].shape[1]
    assert posi["prompt_emb"].dtype == torch.bfloat16 and posi["prompt_emb_mask"].shape == (1, T) and posi["prompt_emb_mask"].all()
    sp = posi["special_token_mask"]
    assert sp.shape == (1, T) and int(sp.sum()) == 64 and sp[0].nonzero().flatten().diff().eq(1).all()   # 64 consecutive <imgN> positions
    assert nega["prompt_emb"].shape[1] < T and int(nega["special_token_mask"].sum()) == 64               # empty negative prompt, same tail
    # the generated JSON was parsed and appended to the positive prompt as "\nkey: value" (:859-872, :813-814)
    edit_calls = [c for c in pipe.text_encoder.calls if c[0] == "edit_forward"]
    assert edit_calls[0][1][1] > edit_calls[1][1][1] and edit_calls[0][2][1] == 1176


def test_unit_runner_contracts():
    from physicedit_b200.units import PipelineUnit, PipelineUnitRunner

    class Both(PipelineUnit):
        def __init__(self):
            super().__init__(seperate_cfg=True, input_params_posi={"p": "prompt"}, input_params_nega={"p": "negative_prompt"}, input_params=("k",))

        def process(self, pipe, p, k):
            return {"emb": f"{p}|{k}"}
    run = PipelineUnitRunner()
    sh, po, ne = run(Both(), None, {"cfg_scale": 4.0, "k": 1}, {"prompt": "a"}, {"negative_prompt": "b"})
    assert po["emb"] == "a|1" and ne["emb"] == "b|1"
    sh, po, ne = run(Both(), None, {"cfg_scale": 1, "k": 1}, {"prompt": "a"}, {"negative_prompt": "b"})
    assert ne["emb"] == "a|1"                                        # cfg_scale == 1: the negative side inherits the positive outputs (:268-269)

    class Take(PipelineUnit):
        def __init__(self):
            super().__init__(take_over=True)

        def process(self, pipe, inputs_shared, inputs_posi, inputs_nega):
            inputs_shared["seen"] = True
            return inputs_shared, inputs_posi, inputs_nega
    assert run(Take(), None, {}, {}, {})[0]["seen"]


def test_lora_injection_layout_matches_what_the_loaders_read_back():
    """add_lora_to_model (trainers/utils.py:799-808) without peft: 12 target linears per block wrapped, PEFT's parameter names, base frozen;
    the exported trainable-only state dict is exactly what GeneralLoRALoader / validate.py:44-65 fold back, bit-identically to merge()."""
    import copy
    from physicedit_b200.lora import GeneralLoRALoader, LoRALinear, merge_lora
    from physicedit_b200.trainers import DiffusionTrainingModule
    targets = "to_q,to_k,to_v,add_q_proj,add_k_proj,add_v_proj,to_out.0,to_add_out,img_mlp.net.2,img_mod.1,txt_mlp.net.2,txt_mod.1".split(",")

    class M(DiffusionTrainingModule):
        def __init__(self):
            super().__init__()
            self.pipe = _cpu_pipe(layers=1)
    m = M()
    plain = copy.deepcopy(m.pipe.dit)
    m.pipe.freeze_except([])
    m.pipe.dit = m.add_lora_to_model(m.pipe.dit, target_modules=targets, lora_rank=8, upcast_dtype=torch.bfloat16)
    wrapped = [n for n, mod in m.pipe.dit.named_modules() if isinstance(mod, LoRALinear)]
    assert len(wrapped) == 12 and "transformer_blocks.0.attn.to_out.0" in wrapped and "transformer_blocks.0.img_mlp.net.0.proj" not in wrapped
    names = m.trainable_param_names()
    assert len(names) == 24 and all(".lora_A.default.weight" in n or ".lora_B.default.weight" in n for n in names)
    assert all(not p.requires_grad for n, p in m.pipe.dit.named_parameters() if "lora_" not in n)
    a = m.pipe.dit.transformer_blocks[0].attn.to_q
    assert a.lora_A["default"].weight.shape == (8, 3072) and a.lora_B["default"].weight.shape == (3072, 8) and a.scaling == 1.0
    assert torch.count_nonzero(a.lora_B["default"].weight) == 0 and a.weight is a.base_layer.weight          # PEFT init; Linear-like surface
    torch.manual_seed(0)
    for n, p in m.pipe.dit.named_parameters():
        if ".lora_B." in n:
            p.data.copy_(torch.randn_like(p) * 0.05)
    sd = m.export_trainable_state_dict(m.state_dict(), remove_prefix="pipe.dit.")
    assert set(sd) == {n[len("pipe.dit."):] for n in names} and "transformer_blocks.0.attn.to_q.lora_A.default.weight" in sd
    GeneralLoRALoader(torch_dtype=torch.bfloat16).load(plain, sd, alpha=1.0)                                 # the inference-side fold of that checkpoint
    merge_lora(m.pipe.dit)
    assert not any(isinstance(mod, LoRALinear) for mod in m.pipe.dit.modules())
    merged, folded = m.pipe.dit.state_dict(), plain.state_dict()
    assert set(merged) == set(folded) and all(torch.equal(merged[k], folded[k]) for k in merged)
    with pytest.raises(ValueError, match="not found"):
        m.add_lora_to_model(m.pipe.dit, target_modules=["no_such_linear"], lora_rank=8)


def test_from_pretrained_recognises_files_by_content(tmp_path, capsys):
    """from_pretrained (:497-541) on local files: a blockwise-controlnet checkpoint is recognised by its keys and appended to pipe.blockwise_controlnet
    (two files -> two entries, the inpaint variant with its 4 extra input channels), an unknown file prints the reference's message and loads nothing
    (model_manager.py:375-376), dinov2_path is mandatory (:198) and may be a local HF folder."""
    from safetensors.torch import save_file
    from oracle import dit_oracle as O
    from oracle import ref_import
    import physicedit_b200 as pe
    dino = ref_import.tiny_dinov2_folder(str(tmp_path / "dino"))
    cn = {k: v.to(torch.bfloat16) for k, v in O.synth_weights(O.controlnet_param_shapes(1), seed=1).items()}
    save_file(cn, str(tmp_path / "cn.safetensors"))
    inpaint = dict(cn)
    inpaint["img_in.weight"] = torch.zeros(3072, 68, dtype=torch.bfloat16)
    save_file(inpaint, str(tmp_path / "cn_inpaint.safetensors"))
    save_file({"something.weight": torch.zeros(4, 4)}, str(tmp_path / "junk.safetensors"))
    with pytest.raises(AssertionError, match="dinov2_path"):
        pe.QwenImagePhysicPipeline.from_pretrained(device="cpu", model_configs=[])
    pipe = pe.QwenImagePhysicPipeline.from_pretrained(torch_dtype=torch.bfloat16, device="cpu", dinov2_path=dino,
                                                      model_configs=[pe.ModelConfig(path=str(tmp_path / n)) for n in ("cn.safetensors", "cn_inpaint.safetensors", "junk.safetensors")])
    assert pipe.dit is None and pipe.vae is None and pipe.text_encoder is None
    nets = pipe.blockwise_controlnet.models
    assert len(nets) == 2 and nets[0].img_in.weight.shape == (3072, 64) and nets[1].img_in.weight.shape == (3072, 68)
    assert torch.equal(nets[0].controlnet_blocks[0].input_proj.weight, cn["controlnet_blocks.0.input_proj.weight"])
    assert "cannot detect the model type" in capsys.readouterr().out
    # the scripts' way of naming files (inference_pica.py:229-241): ModelConfig(model_id, origin_file_pattern, local_model_path) -> a glob over local
    # files; several shards of one model are merged before the model is recognised
    shard_dir = tmp_path / "models" / "Org" / "Net" / "controlnet"
    shard_dir.mkdir(parents=True)
    keys = sorted(cn)
    save_file({k: cn[k] for k in keys[:3]}, str(shard_dir / "part-00001-of-00002.safetensors"))
    save_file({k: cn[k] for k in keys[3:]}, str(shard_dir / "part-00002-of-00002.safetensors"))
    cfg = pe.ModelConfig(model_id="Org/Net", origin_file_pattern="controlnet/part-*.safetensors", local_model_path=str(tmp_path / "models"))
    pipe2 = pe.QwenImagePhysicPipeline.from_pretrained(torch_dtype=torch.bfloat16, device="cpu", dinov2_path=dino, model_configs=[cfg])
    assert isinstance(cfg.path, list) and len(cfg.path) == 2 and len(pipe2.blockwise_controlnet.models) == 1
    assert torch.equal(pipe2.blockwise_controlnet.models[0].controlnet_blocks[0].output_proj.weight, cn["controlnet_blocks.0.output_proj.weight"])
    folder = pe.ModelConfig(model_id="Org/Net", origin_file_pattern="controlnet/", local_model_path=str(tmp_path / "models"))
    folder.download_if_necessary()
    assert folder.path == os.path.join(str(tmp_path / "models" / "Org" / "Net"), "controlnet/")         # folders keep the pattern's trailing slash (:216)
    with pytest.raises(ValueError, match="No valid model files"):
        pe.ModelConfig().download_if_necessary()


def test_hot_lora_surface():
    """enable_lora_magic / load_lora(hotload=True) / clear_lora (qwen_image_physical.py:265-305): wrappers around every linear of the DiT, factors
    attached by PEFT key name and scaled by alpha, the un-merged flag follows the lists; get_special_divisor (:308-311)."""
    from physicedit_b200.lora import HotLoRALinear
    pipe = _cpu_pipe(layers=1)
    n_linear = sum(isinstance(m, torch.nn.Linear) for m in pipe.dit.modules())
    pipe.load_lora(pipe.dit, state_dict={"transformer_blocks.0.attn.to_q.lora_A.default.weight": torch.ones(4, 3072),
                                         "transformer_blocks.0.attn.to_q.lora_B.default.weight": torch.ones(3072, 4)}, hotload=True)
    assert not getattr(pipe.dit, "_lora_injected", False)                # no wrappers yet: nothing attaches (the reference's isinstance filter)
    pipe.enable_lora_magic()
    pipe.enable_lora_magic()                                             # idempotent
    wrappers = {n: m for n, m in pipe.dit.named_modules() if isinstance(m, HotLoRALinear)}
    assert len(wrappers) == n_linear and "transformer_blocks.0.attn.to_out.0" in wrappers
    q = wrappers["transformer_blocks.0.attn.to_q"]
    assert q.weight is q.base_layer.weight and q.in_features == 3072 and not pipe.dit._lora_injected
    sd = {"transformer_blocks.0.attn.to_q.lora_A.default.weight": torch.ones(4, 3072), "transformer_blocks.0.attn.to_q.lora_B.default.weight": torch.ones(3072, 4),
          "transformer_blocks.0.attn.to_k.lora_A.default.weight": torch.ones(4, 3072)}          # B missing: skipped
    pipe.load_lora(pipe.dit, state_dict=sd, alpha=0.5, hotload=True)
    pipe.load_lora(pipe.dit, state_dict=sd, alpha=2.0, hotload=True)
    assert len(q.lora_A_weights) == 2 and float(q.lora_A_weights[0][0, 0]) == 0.5 and float(q.lora_A_weights[1][0, 0]) == 2.0 and float(q.lora_B_weights[0][0, 0]) == 1.0
    assert len(wrappers["transformer_blocks.0.attn.to_k"].lora_A_weights) == 0 and pipe.dit._lora_injected
    pipe.clear_lora()
    assert len(q.lora_A_weights) == 0 and not pipe.dit._lora_injected
    assert pipe.get_special_divisor(global_step=0) == 10 and pipe.get_special_divisor(global_step=5000) == 5.5 and pipe.get_special_divisor(global_step=20000) == 1.0


def test_product_entry_points_fail_loudly_without_a_gpu():
    """No CPU path anywhere in the product: the step function, the text encoder and the training Functions raise NativeUnavailable when the handle cannot be
    created (no sm_100 GPU / no libpe_b200.so) -- they never fall back to torch ops or to the oracle."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from physicedit_b200 import native as nv
    from physicedit_b200 import autograd as ag
    from physicedit_b200.model_fn import model_fn_qwen_image
    pipe = _cpu_pipe(layers=1)
    lat = torch.zeros(1, 16, 8, 8, dtype=torch.bfloat16)
    kw = dict(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=lat, timestep=torch.tensor([500.0]).bfloat16(),
              prompt_emb=torch.zeros(1, 8, 3584, dtype=torch.bfloat16), prompt_emb_mask=torch.ones(1, 8, dtype=torch.long), special_token_mask=None, height=64,
              width=64, is_train=False)
    with pytest.raises(nv.NativeUnavailable):
        model_fn_qwen_image(**kw)
    with pytest.raises(nv.NativeUnavailable):
        model_fn_qwen_image(**dict(kw, latents=lat.float()))
    with pytest.raises(nv.NativeUnavailable):
        ag.linear(torch.zeros(4, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    from physicedit_b200.text_encoder import QwenImageTextEncoder, VLConfig
    te = QwenImageTextEncoder(VLConfig(hidden=64, layers=1, heads=2, kv_heads=1, head_dim=32, intermediate=64, vocab=32, v_hidden=32, v_depth=1, v_heads=1,
                                       v_intermediate=32, v_out=64, fullatt=(0,))).bfloat16()
    with pytest.raises(nv.NativeUnavailable):
        te.edit_forward(input_ids=torch.zeros(1, 4, dtype=torch.long), attention_mask=torch.ones(1, 4, dtype=torch.long))


def test_base_pipeline_step_with_and_without_an_inpaint_mask():
    """BasePipeline.step / blend_with_mask (utils/__init__.py:146-156), the torch-arithmetic update `direct_distill_loss` and the reference's loop body use."""
    pipe = _cpu_pipe()
    pipe.scheduler.set_timesteps(4, dynamic_shift_len=16)
    g = torch.Generator().manual_seed(0)
    lat, pred, clean = (torch.randn(1, 16, 4, 4, generator=g) for _ in range(3))
    sig = pipe.scheduler.sigmas
    want = lat + pred * (sig[2] - sig[1])
    assert torch.allclose(pipe.step(pipe.scheduler, latents=lat, progress_id=1, noise_pred=pred, input_latents=clean, height=64), want)
    mask = (torch.rand(1, 1, 4, 4, generator=g) > 0.5).float()
    expected = (lat - clean) / sig[1]                                # the prediction that returns to `clean` (flow_match.py:85-91)
    want = lat + (expected * (1 - mask) + pred * mask) * (sig[2] - sig[1])
    assert torch.allclose(pipe.step(pipe.scheduler, latents=lat, progress_id=1, noise_pred=pred, input_latents=clean, inpaint_mask=mask), want)


def test_load_lora_from_a_model_config_given_by_id_and_pattern(tmp_path):
    """load_lora(module, ModelConfig(model_id=, origin_file_pattern=, local_model_path=)) resolves the file like the reference (:258-263) and folds it."""
    from safetensors.torch import save_file
    from physicedit_b200.pipeline import ModelConfig
    pipe = _cpu_pipe()
    w0 = pipe.dit.transformer_blocks[0].attn.to_k.weight.detach().clone()
    g = torch.Generator().manual_seed(1)
    sd = {"transformer_blocks.0.attn.to_k.lora_A.default.weight": (torch.randn(8, 3072, generator=g) * 0.05).bfloat16(),
          "transformer_blocks.0.attn.to_k.lora_B.default.weight": (torch.randn(3072, 8, generator=g) * 0.05).bfloat16()}
    folder = tmp_path / "models" / "me" / "my-lora"
    folder.mkdir(parents=True)
    save_file(sd, str(folder / "physicedit.safetensors"))
    pipe.load_lora(pipe.dit, ModelConfig(model_id="me/my-lora", origin_file_pattern="*.safetensors", local_model_path=str(tmp_path / "models")), alpha=0.5)
    want = w0 + 0.5 * torch.mm(sd["transformer_blocks.0.attn.to_k.lora_B.default.weight"], sd["transformer_blocks.0.attn.to_k.lora_A.default.weight"])
    assert torch.equal(pipe.dit.transformer_blocks[0].attn.to_k.weight, want.to(torch.bfloat16)) or \
        torch.allclose(pipe.dit.transformer_blocks[0].attn.to_k.weight.float(), want.float(), atol=2e-2)


def test_launcher_runs_a_script_against_the_alias_package(tmp_path):
    """`python -m physicedit_b200 <script> <args>`: the script's `diffsynth` imports resolve to this package, it runs as __main__ with its own argv."""
    import subprocess
    script = tmp_path / "probe.py"
    script.write_text("import sys\n"
                      "from diffsynth.pipelines.qwen_image_physical import QwenImagePhysicPipeline, ModelConfig\n"
                      "from diffsynth.trainers.utils import PhysicalEditingDataset, launch_training_task\n"
                      "from diffsynth.trainers.unified_dataset import UnifiedDataset, LoadImage\n"
                      "if __name__ == '__main__':\n"
                      "    print('ARGV', sys.argv[1:], QwenImagePhysicPipeline.__module__, PhysicalEditingDataset.__module__, UnifiedDataset.__module__)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "physicedit_b200", str(script), "--prompt", "x y"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "ARGV ['--prompt', 'x y'] physicedit_b200.pipeline physicedit_b200.datasets physicedit_b200.unified_dataset" in r.stdout
    r = subprocess.run([sys.executable, "-m", "physicedit_b200"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 2 and "Launcher" in r.stdout
