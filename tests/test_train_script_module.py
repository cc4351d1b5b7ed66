"""The reference train script's OWN training module on this framework, on the CPU.

`QwenImageTrainingModule` is compiled, unmodified, from scripts/train/train_physicedit.py:191-325 (only that class statement: the script's
other top-level code needs accelerate / a wandb login); its `diffsynth.*` imports are served by `physicedit_b200.compat.install()`.  The test
walks what `train_physicedit.py` does per sample -- constructor (`from_pretrained`, `switch_pipe_to_training_mode`: scheduler training table,
freezing, un-merged LoRA), a sample of `PhysicalEditingDataset`, `forward_preprocess` (every pipeline unit, training flavour: input / edit image
latents, DINOv2 + VAE pseudo targets from the middle key frames, the rule-conditioned prompt, prompt embedding), `training_loss` and
`backward()` -- with the C ABI emulated (tests/abi_emulator.py), a stub Qwen2.5-VL and a stub VAE.  What it pins: the host-side contract between
the script, the dataset's sample dictionary, the units and the training path (names, keyword flow, which parameters receive gradients)."""
import ast
import os
import sys
import types

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abi_emulator import EmulatedNative  # noqa: E402

REF = "/root/reference"
SCRIPT = os.path.join(REF, "scripts", "train", "train_physicedit.py")
needs_ref = pytest.mark.skipif(not os.path.isfile(SCRIPT), reason="the reference's scripts are not on this machine")
cv2 = pytest.importorskip("cv2")

TRAINABLE = "visual_thinking_adapter,vae_time_embed,vae_resampler,vae_resampler_adapter,dino_time_embed,dino_resampler,dino_resampler_adapter"   # train_multigpu.sh:38
TARGETS = "to_q,to_k,to_v,add_q_proj,add_k_proj,add_v_proj,to_out.0,to_add_out,img_mlp.net.2,img_mod.1,txt_mlp.net.2,txt_mod.1"                     # :30
EXTRA = "edit_image,supported_rules,contradicted_rules,middle_key_frames,stitched_image,state,transition,triplet"                                 # :20


class _StubVL:
    """Stands in for the Qwen2.5-VL encoder (models/qwen_image_text_encoder_withdecode.py): deterministic hidden states, a canned generation."""

    def __init__(self, reply_ids):
        self.reply_ids, self.calls = reply_ids, []

    def edit_forward(self, input_ids=None, attention_mask=None, pixel_values=None, image_grid_thw=None, output_hidden_states=True, **kw):
        self.calls.append("edit_forward")
        g = torch.Generator().manual_seed(int(input_ids.sum()) % 1000)
        return (torch.randn(input_ids.shape[0], input_ids.shape[1], 3584, generator=g),)

    def generate(self, input_ids=None, max_new_tokens=None, **kw):
        self.calls.append("generate")
        return torch.cat([input_ids, self.reply_ids.unsqueeze(0)], dim=1)

    def parameters(self):
        return iter(())


class _StubVAE(torch.nn.Module):
    def encode(self, x, **kw):
        g = torch.Generator().manual_seed(int(x.shape[-1]) * 7 + int(x.shape[-2]))
        return torch.randn(x.shape[0], 16, x.shape[2] // 8, x.shape[3] // 8, generator=g).to(x.dtype)

    def decode(self, z, **kw):
        return torch.tanh(z[:, :3].float()).repeat_interleave(8, dim=2).repeat_interleave(8, dim=3).to(z.dtype)


@pytest.fixture()
def script_module(monkeypatch):
    from physicedit_b200 import compat
    saved = {k: v for k, v in sys.modules.items() if k == "diffsynth" or k.startswith("diffsynth.")}
    compat.install()
    ns = {"wandb": types.SimpleNamespace(log=lambda *a, **k: None, init=lambda *a, **k: None)}
    exec("import torch, os, json, time, shutil\n"
         "from diffsynth import load_state_dict\n"
         "from diffsynth.pipelines.qwen_image_physical import QwenImagePhysicPipeline, ModelConfig\n"
         "from diffsynth.pipelines.flux_image_new import ControlNetInput\n"
         "from diffsynth.trainers.utils import DiffusionTrainingModule, ModelLogger, qwen_image_parser, launch_training_task, launch_data_process_task, PhysicalEditingDataset\n"
         "from diffsynth.trainers.unified_dataset import UnifiedDataset\n", ns)                 # lines 1-6 of the script
    tree = ast.parse(open(SCRIPT, encoding="utf-8").read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("QwenImageTrainingModule", "WandbModelLogger")]
    exec(compile(ast.Module(body=cls, type_ignores=[]), SCRIPT, "exec"), ns)
    yield ns
    for k in [k for k in sys.modules if k == "diffsynth" or k.startswith("diffsynth.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def small_edit_images(monkeypatch, area=256 * 256):
    """CPU sizing of the script-level tests: the edit image is auto-resized to ~256^2 instead of ~1024^2 (256 instead of 4096 tokens per forward on the
    emulated ABI).  The 1024^2 rule itself is pinned against the reference's unit in tests/test_units_vs_reference.py."""
    from physicedit_b200 import units as U
    monkeypatch.setattr(U.QwenImageUnit_EditImageEmbedder, "edit_image_auto_resize", lambda self, image: U.resize_to_area(image, area))


def _pipe_on_the_emulator(monkeypatch, emu):
    """A CPU pipeline with a 1-block DiT bound to the emulated ABI, a 1-layer DINOv2, stub VL / VAE and the real tokenizer / processor files."""
    from oracle import dit_oracle as O
    from transformers import Qwen2Tokenizer, Qwen2VLProcessor
    from physicedit_b200 import adapters, autograd, native as nv
    from physicedit_b200.dit import DiTEngine, QwenImageDiT
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    monkeypatch.setattr(adapters, "_nat", lambda t: emu)
    monkeypatch.setattr(autograd, "_nat", lambda t: emu)
    autograd.weight_transposes.clear()
    W = O.synth_weights(O.dit_param_shapes(1), seed=51)
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=1)
    dit.load_state_dict({k: v.to(torch.bfloat16) for k, v in W.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, dinov2_config=dict(hidden=768, layers=1, heads=12))
    pipe.dit = dit
    tok_dir = os.path.join(REF, "DiffSynth-Studio", "models", "Qwen", "Qwen-Image", "tokenizer")
    proc_dir = os.path.join(REF, "DiffSynth-Studio", "models", "Qwen", "Qwen-Image-Edit", "processor")
    tok = Qwen2Tokenizer.from_pretrained(tok_dir)
    base = Qwen2VLProcessor.from_pretrained(proc_dir)
    proc = Qwen2VLProcessor(image_processor=base.image_processor, tokenizer=Qwen2Tokenizer.from_pretrained(tok_dir), video_processor=base.video_processor,
                            chat_template=base.chat_template)
    reply = tok('{"middle_transition_prompt": "The ball rolls off the table."}', return_tensors="pt").input_ids[0]
    pipe.text_encoder = _StubVL(reply)
    pipe.vae = _StubVAE()
    pipe.attach_tokenizer(tokenizer=tok, processor=proc)
    pipe.to(torch.bfloat16)
    eng = object.__new__(DiTEngine)               # after the move: `.to()` drops a DiT's engine (its packed buffers follow the storage)
    eng.dit, eng.device, eng.nat, eng.use_cta_pair, eng.attn_flags, eng._ws, eng._rope, eng.sp = dit, torch.device("cpu"), emu, True, 0, {}, {}, None
    eng._pack()
    object.__setattr__(dit, "_engine", eng)
    return pipe


@needs_ref
def test_the_train_scripts_module_runs_a_sample_end_to_end(script_module, monkeypatch, tmp_path):
    from test_datasets import meta, write_clip
    import json
    from physicedit_b200 import datasets as D
    from physicedit_b200.lora import LoRALinear
    clip_dir = tmp_path / "clips" / "scene"
    clip_dir.mkdir(parents=True)
    write_clip(clip_dir / "4.mp4", 50, 96, 64, 3)
    (clip_dir / D.METADATA_FILE).write_text(json.dumps(meta(4)) + "\n", encoding="utf-8")
    args = script_module["qwen_image_parser"]().parse_args(["--dataset_base_path", str(tmp_path / "clips"), "--dinov2_path", "unused", "--height", "64",
                                                            "--width", "96", "--num_frames", "49"])
    dataset = script_module["PhysicalEditingDataset"](args=args)                                  # train_physicedit.py:420
    emu = EmulatedNative()
    pipe = _pipe_on_the_emulator(monkeypatch, emu)
    small_edit_images(monkeypatch)
    Pipeline = script_module["QwenImagePhysicPipeline"]
    seen = {}

    def from_pretrained(**kw):                    # the checkpoint files are not on this machine: hand the constructor the prepared pipeline
        seen.update(kw)
        return pipe
    monkeypatch.setattr(Pipeline, "from_pretrained", staticmethod(from_pretrained))
    Module = script_module["QwenImageTrainingModule"]
    module = Module(trainable_models=TRAINABLE, lora_base_model="dit", lora_target_modules=TARGETS, lora_rank=8, use_gradient_checkpointing=True,
                    extra_inputs=EXTRA, dinov2_path="unused")                                     # :424-439 with the flags of train_multigpu.sh
    assert seen["device"] == "cpu" and seen["torch_dtype"] == torch.bfloat16 and seen["dinov2_path"] == "unused" and seen["model_configs"] == []
    assert seen["tokenizer_config"].origin_file_pattern == "tokenizer/" and seen["processor_config"].model_id == "Qwen/Qwen-Image-Edit"
    # switch_pipe_to_training_mode (trainers/utils.py:856-888)
    assert len(pipe.scheduler.timesteps) == 1000 and hasattr(pipe.scheduler, "linear_timesteps_weights")
    wrapped = [n for n, m in pipe.dit.named_modules() if isinstance(m, LoRALinear)]
    assert len(wrapped) == 12 and all(any(n.endswith(t) for t in TARGETS.split(",")) for n in wrapped)
    names = module.trainable_param_names()
    assert sum(".lora_A." in n or ".lora_B." in n for n in names) == 24
    assert all(n.startswith("pipe.dit.") and ".lora_" in n or n.split(".")[1] in TRAINABLE.split(",") for n in names)
    assert not any(n.startswith("pipe.dinov2.") for n in names)
    sd = module.export_trainable_state_dict(module.state_dict(), remove_prefix="pipe.dit.")       # what ModelLogger writes / validate.py reads back
    assert "transformer_blocks.0.attn.to_q.lora_A.default.weight" in sd and "pipe.visual_thinking_adapter.head_dino.0.weight" in sd

    data = dataset[0]
    assert len(data["middle_key_frames"]) == 6 and data["stitched_image"] is not None
    inputs = module.forward_preprocess(data)                                                      # :255-295: every unit, training flavour
    assert inputs["input_latents"].shape == (1, 16, 8, 12) and inputs["latents"].shape == (1, 16, 8, 12)
    assert inputs["height"] == 64 and inputs["width"] == 96 and inputs["cfg_scale"] == 1
    edit = inputs["edit_latents"][0] if isinstance(inputs["edit_latents"], list) else inputs["edit_latents"]
    assert edit.shape[0] == 1 and edit.shape[1] == 16
    assert inputs["pseudo_special_emb_dino"].shape == (1, 64, 3584) and inputs["pseudo_special_emb_vae"].shape == (1, 64, 3584)
    assert inputs["pseudo_special_emb_dino"].requires_grad and inputs["pseudo_special_emb_vae"].requires_grad       # the resampler stack trains
    T = inputs["prompt_emb"].shape[1]
    assert inputs["prompt_emb"].shape == (1, T, 3584) and int(inputs["special_token_mask"].sum()) == 64 and inputs["prompt_emb_mask"].shape == (1, T)
    assert "generate" not in pipe.text_encoder.calls                   # training: the transition text comes from the rules, not from a generation
    torch.manual_seed(0)
    loss = module(data, inputs=inputs)                                                            # :298-325 -> pipe.training_loss(**models, **inputs)
    assert loss.ndim == 0 and torch.isfinite(loss) and loss.item() > 0 and pipe.special_token_loss > 0
    loss.backward()
    grads = {n: p.grad for n, p in module.named_parameters() if p.requires_grad}
    missing = [n for n, g in grads.items() if g is None]
    # one block: its text tail (txt_mlp, to_add_out and the text modulation that only feeds them) has no path to the loss -- find_unused_parameters
    assert all(any(t in n for t in ("txt_mlp", "to_add_out", "txt_mod")) for n in missing), missing
    for part in ("visual_thinking_adapter", "dino_resampler.", "vae_resampler.", "dino_resampler_adapter", "vae_resampler_adapter", "dino_time_embed",
                 "vae_time_embed", "attn.to_q.lora_B", "img_mlp.net.2.lora_B", "img_mod.1.lora_B"):       # lora_B starts at zero (PEFT init): dA = 0 on step one
        hit = [g for n, g in grads.items() if part in n and g is not None]
        assert hit and all(torch.isfinite(g.float()).all() for g in hit) and any(g.float().abs().sum() > 0 for g in hit), part

    # ---- task "data_process" (:312-313 + trainers/utils.py:980-1002): the units' outputs of every sample cached to <output_path>/<rank>/<i>.pth
    module.task = "data_process"
    cache_logger = script_module["ModelLogger"](str(tmp_path / "cache"))
    script_module["launch_data_process_task"](dataset, module, cache_logger, num_workers=0)
    cached = torch.load(str(tmp_path / "cache" / "0" / "0.pth"), weights_only=False)
    assert set(inputs) == set(cached) and cached["prompt_emb"].shape == inputs["prompt_emb"].shape and cached["input_latents"].shape == (1, 16, 8, 12)
    assert not cached["pseudo_special_emb_dino"].requires_grad                                    # cached under no_grad
    module.task = "sft"
    replay = script_module["UnifiedDataset"](base_path=str(tmp_path / "cache"))                   # no metadata file: the cache reader
    assert replay.load_from_cache and len(replay) == 1
    torch.manual_seed(0)
    cached_loss = module({}, inputs=replay[0])                                                    # what launch_training_task does with a cached dataset
    assert torch.isfinite(cached_loss) and cached_loss.item() > 0
    # ---- the script's logger on this framework's stand-in for accelerate.Accelerator (trainers._Ranks): checkpoint + mid-training evaluation
    from safetensors.torch import load_file
    from physicedit_b200.trainers import _Ranks
    logger = script_module["WandbModelLogger"](str(tmp_path / "out"), remove_prefix_in_ckpt="pipe.dit.", eval_every_n_steps=1, eval_data=dataset)
    path = logger.save_checkpoint(_Ranks(), module, 3)                                            # :171-186
    ck = load_file(path)
    assert path.endswith("step-3.safetensors") and set(ck) == set(sd) and ck["transformer_blocks.0.attn.to_q.lora_B.default.weight"].shape == (3072, 8)
    # resume_type "model" (:527-548): keys without `pipe.` get the stripped prefix back, then a non-strict load into the training module
    before = {k: v.detach().clone() for k, v in module.state_dict().items() if k in module.trainable_param_names()}
    with torch.no_grad():
        for p_ in module.trainable_modules():
            p_.zero_()
    missing, unexpected = module.load_state_dict({(k if k.startswith("pipe.") else f"pipe.dit.{k}"): v for k, v in ck.items()}, strict=False)
    assert unexpected == [] and not (set(missing) & set(before))
    assert all(torch.equal(module.state_dict()[k], v) for k, v in before.items())
    calls = {}

    def fake_denoise(latents, inputs_posi, inputs_nega, edit_latents=None, context_latents=None, **kw):
        calls.update(kw, T_posi=inputs_posi["prompt_emb"].shape[1], has_nega=inputs_nega is not None, grad=torch.is_grad_enabled(), training=module.training)
        return latents
    monkeypatch.setattr(pipe, "denoise", fake_denoise)
    module.train()
    metrics = logger.evaluate_model(module, _Ranks())                                             # :39-169 -> pipe(prompt, edit_image, is_train=False)
    assert "eval_error" not in metrics, metrics
    assert os.path.isfile(metrics["eval_edit_image_path"]) and metrics["eval_edit_image_path"].endswith("_idx_4.jpg")
    assert calls["height"] == 480 and calls["width"] == 832 and calls["num_inference_steps"] == 40 and calls["has_nega"] and not calls["grad"] and not calls["training"]
    assert module.training and len(pipe.scheduler.timesteps) == 1000                              # training mode and the training table restored
    assert "generate" in pipe.text_encoder.calls                                                  # inference: the transition text is generated


@needs_ref
def test_the_whole_train_script_runs_as_main_in_a_fresh_process(tmp_path):
    """scripts/train/train_physicedit.py executed as `__main__`, unmodified (tests/run_train_script.py: compat aliases, a stand-in for the absent
    `accelerate`, models on the emulated ABI): flags of train_multigpu.sh at a small size, one clip, one epoch -> trainable-parameter report, one
    optimizer step, the epoch checkpoint in the layout validate.py reads and its metadata file."""
    import json
    import subprocess
    from safetensors.torch import load_file
    from test_datasets import meta, write_clip
    from physicedit_b200 import datasets as D
    clip_dir = tmp_path / "clips" / "scene"
    clip_dir.mkdir(parents=True)
    write_clip(clip_dir / "0.mp4", 50, 96, 64, 10)
    (clip_dir / D.METADATA_FILE).write_text(json.dumps(meta(0)) + "\n", encoding="utf-8")
    out = tmp_path / "run"
    flags = ["--dataset_base_path", str(tmp_path / "clips"), "--height", "64", "--width", "96", "--num_frames", "49", "--data_file_keys", "image",
             "--extra_inputs", EXTRA, "--max_pixels", "1048576", "--dataset_repeat", "1", "--dinov2_path", "unused", "--learning_rate", "5e-5",
             "--num_epochs", "1", "--remove_prefix_in_ckpt", "pipe.dit.", "--output_path", str(out), "--lora_base_model", "dit",
             "--lora_target_modules", TARGETS, "--lora_rank", "8", "--use_gradient_checkpointing", "--dataset_num_workers", "0", "--find_unused_parameters",
             "--trainable_models", TRAINABLE]
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", WANDB_MODE="disabled")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_train_script.py"), SCRIPT] + flags,
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MODEL TRAINABLE PARAMETERS REPORT" in r.stdout and "collected 1 samples" in r.stdout and "[HARNESS] emulated launches" in r.stdout
    ck = load_file(str(out / "epoch-0.safetensors"))
    assert "transformer_blocks.0.attn.to_q.lora_A.default.weight" in ck and "pipe.dino_resampler.latents" in ck and "pipe.vae_time_embed.weight" in ck
    assert not any(k.startswith("pipe.dit.") or k.startswith("pipe.dinov2.") for k in ck)
    assert ck["transformer_blocks.0.attn.to_q.lora_B.default.weight"].float().abs().sum() > 0           # the AdamW step moved B off its zero init
    md = json.loads((out / "epoch-0.json").read_text())
    assert md["global_step"] == 1 and md["epoch"] == 0 and md["save_type"] == "epoch" and md["num_processes"] == 1 and md["batches_per_epoch_total"] == 1
