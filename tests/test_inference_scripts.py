"""The reference's benchmark-inference driver on this framework, on the CPU: scripts/inference/inference_pica.py is executed unmodified (its
`main()`: argument parsing, dataset loop, `from_pretrained`, `load_finetuned_into_pipe`, `enable_vram_management`, `pipe(prompt, edit_image=...,
is_train=False)`, image files written) with `diffsynth.*` served by `physicedit_b200.compat.install()`, the C ABI emulated
(tests/abi_emulator.py), a stub Qwen2.5-VL / VAE, and the PICABench download replaced by two in-memory records."""
import importlib.util
import os
import sys
import types

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abi_emulator import EmulatedNative  # noqa: E402

REF = "/root/reference"
SCRIPT = os.path.join(REF, "scripts", "inference", "inference_pica.py")
needs_ref = pytest.mark.skipif(not os.path.isfile(SCRIPT), reason="the reference's scripts are not on this machine")


@needs_ref
def test_inference_pica_main_runs_on_this_framework(monkeypatch, tmp_path, capsys):
    from PIL import Image
    from safetensors.torch import save_file
    from physicedit_b200 import compat
    from test_train_script_module import _pipe_on_the_emulator, small_edit_images
    saved = {k: v for k, v in sys.modules.items() if k == "diffsynth" or k.startswith("diffsynth.")}
    compat.install()
    try:
        # third-party imports of the script that are not installed here and not used by main()
        iio = types.ModuleType("imageio.v3")
        imageio = types.ModuleType("imageio")
        imageio.v3 = iio
        openai = types.ModuleType("openai")
        openai.OpenAI = lambda **kw: None
        for name, mod in (("imageio", imageio), ("imageio.v3", iio), ("openai", openai)):
            if name not in sys.modules:
                monkeypatch.setitem(sys.modules, name, mod)
        spec = importlib.util.spec_from_file_location("ref_inference_pica", SCRIPT)
        sys.dont_write_bytecode = True
        script = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(script)
        assert script.QwenImagePhysicPipeline.__module__.startswith("physicedit_b200")

        emu = EmulatedNative()
        pipe = _pipe_on_the_emulator(monkeypatch, emu)
        small_edit_images(monkeypatch)
        pipe.cfg_streams = 1
        w_q = pipe.dit.transformer_blocks[0].attn.to_q.weight.detach().clone()
        # a checkpoint as train_physicedit.py writes it: LoRA keys with `pipe.dit.` stripped + `pipe.*` keys of the trained modules
        g = torch.Generator().manual_seed(3)
        ck = {"transformer_blocks.0.attn.to_q.lora_A.default.weight": (torch.randn(8, 3072, generator=g) * 0.05).bfloat16(),
              "transformer_blocks.0.attn.to_q.lora_B.default.weight": (torch.randn(3072, 8, generator=g) * 0.05).bfloat16()}
        ad_new = {k: (torch.randn(v.shape, generator=g) * 0.02).bfloat16() for k, v in pipe.visual_thinking_adapter.state_dict().items()}
        ck.update({f"pipe.visual_thinking_adapter.{k}": v for k, v in ad_new.items()})
        ck_path = str(tmp_path / "epoch-4.safetensors")
        save_file(ck, ck_path)
        seen = {}

        def from_pretrained(**kw):
            seen.update(kw)
            return pipe
        monkeypatch.setattr(script.QwenImagePhysicPipeline, "from_pretrained", staticmethod(from_pretrained))
        records = [dict(superficial_prompt="s", intermediate_prompt=f"tip the glass {i}", explicit_prompt="e",
                        input_image=Image.new("RGB", (96, 64), (40 * i, 90, 200))) for i in range(3)]
        monkeypatch.setattr(script, "load_dataset", lambda name, cache_dir=None: {"picabench": records})
        out_dir = tmp_path / "out"
        monkeypatch.setattr(sys, "argv", ["inference_pica.py", "--base_model_path", str(tmp_path / "base"), "--dinov2_path", "unused", "--data_path", str(tmp_path),
                                          "--lora_path", ck_path, "--output_path", str(out_dir), "--num_inference_steps", "1", "--start_idx", "2", "--end_idx", "900",
                                          "--seed", "11"])
        script.main()
        assert seen["device"] == "cuda" and len(seen["model_configs"]) == 3 and seen["model_configs"][0].origin_file_pattern.startswith("transformer/")
        # load_finetuned_into_pipe (:142-175): the LoRA folded into the DiT weight, the adapter replaced
        want = w_q + torch.mm(ck["transformer_blocks.0.attn.to_q.lora_B.default.weight"], ck["transformer_blocks.0.attn.to_q.lora_A.default.weight"])
        assert torch.equal(pipe.dit.transformer_blocks[0].attn.to_q.weight, want)
        assert all(torch.equal(pipe.visual_thinking_adapter.state_dict()[k], v) for k, v in ad_new.items())
        # record 2 (start_idx .. min(end_idx, len)) was edited at its own size and saved as <idx>.jpg
        files = sorted(os.listdir(out_dir))
        assert files == ["00002.jpg"]
        assert Image.open(out_dir / "00002.jpg").size == (96, 64)
        names = [c[0] for c in emu.calls]
        assert names.count("pe_cfg_euler_step") == 1 and names.count("pe_special_blend_scatter") == 2          # 1 image x 1 step x 2 CFG branches
        assert pipe.text_encoder.calls.count("generate") == 2 and pipe.text_encoder.calls.count("edit_forward") == 2
        assert "[DONE] Generated" in capsys.readouterr().out
    finally:
        for k in [k for k in sys.modules if k == "diffsynth" or k.startswith("diffsynth.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@needs_ref
def test_validate_py_main_256x256_four_steps(monkeypatch, tmp_path, capsys):
    """BASELINE.json configs[0]: scripts/inference/validate.py, one 256 x 256 edit, 4 denoise steps -- the script's own `main()` (argument parsing,
    from_pretrained, `load_finetuned_into_pipe` without a checkpoint, image load, `pipe(...)`, save) on this framework with the C ABI emulated.
    The script's `resize_image` (which would blow the input up to ~1024^2) is pinned to the 256 x 256 the config names."""
    from PIL import Image
    from physicedit_b200 import compat
    from test_train_script_module import _pipe_on_the_emulator, small_edit_images
    script_path = os.path.join(REF, "scripts", "inference", "validate.py")
    saved = {k: v for k, v in sys.modules.items() if k == "diffsynth" or k.startswith("diffsynth.")}
    compat.install()
    try:
        spec = importlib.util.spec_from_file_location("ref_validate_main", script_path)
        sys.dont_write_bytecode = True
        script = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(script)
        emu = EmulatedNative()
        pipe = _pipe_on_the_emulator(monkeypatch, emu)
        small_edit_images(monkeypatch)
        pipe.cfg_streams = 1
        monkeypatch.setattr(script.QwenImagePhysicPipeline, "from_pretrained", staticmethod(lambda **kw: pipe))
        monkeypatch.setattr(script, "resize_image", lambda image, target_area=None: image.resize((256, 256)))
        src = tmp_path / "in.png"
        Image.new("RGB", (300, 280), (200, 120, 40)).save(src)
        dst = tmp_path / "results" / "edited.png"
        monkeypatch.setattr(sys, "argv", ["validate.py", "--prompt", "let the ice melt", "--image_path", str(src), "--save_path", str(dst), "--seed", "5",
                                          "--num_inference_steps", "4"])
        script.main()
        out = capsys.readouterr().out
        assert "No checkpoint path provided" in out and "[DONE] Saved result" in out
        assert Image.open(dst).size == (256, 256)
        names = [c[0] for c in emu.calls]
        assert names.count("pe_cfg_euler_step") == 4 and names.count("pe_special_blend_scatter") == 8 and names.count("pe_timestep_embedding") == 4
        assert names.count("pe_attention_fwd") == 8                      # 4 steps x 2 CFG branches x 1 block
    finally:
        for k in [k for k in sys.modules if k == "diffsynth" or k.startswith("diffsynth.")]:
            del sys.modules[k]
        sys.modules.update(saved)
