"""ORACLE-side test infrastructure (not product code): import the REFERENCE's own modules (unmodified) when a copy is around.

Search order: `baseline/_ref/diffsynth` (the pip --target install of /root/reference/DiffSynth-Studio that DESIGN.md records;
git-ignored, travels to the GPU box with the snapshot) and `/root/reference/DiffSynth-Studio/diffsynth` (authoring container
only).  `import diffsynth` itself fails here (its __init__ pulls imageio / modelscope), so the package __init__ is skipped by
registering an empty namespace module whose __path__ points at the tree plus a stub `modelscope`.

Only tests/ and bench.py's baseline legs (`stock_gpu`, `--impl reference`) use this; nothing under physicedit_b200/ does.
"""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref", "diffsynth"), "/root/reference/DiffSynth-Studio/diffsynth")


def reference_root():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "pipelines", "qwen_image_physical.py")):
            return c
    return None


class ReferenceModules:
    """Context manager: inside, `diffsynth.*` resolves to the reference tree; on exit sys.modules is restored (so that
    physicedit_b200.compat.install(), which registers its own `diffsynth` alias package, can be used by other tests)."""

    def __init__(self):
        self.root = reference_root()
        self._saved = {}

    def __enter__(self):
        if self.root is None:
            raise FileNotFoundError("no reference tree (baseline/_ref or /root/reference)")
        os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
        sys.dont_write_bytecode = True
        self._saved = {k: v for k, v in sys.modules.items() if k == "diffsynth" or k.startswith("diffsynth.") or k == "modelscope"}
        for k in self._saved:
            del sys.modules[k]
        pkg = types.ModuleType("diffsynth")
        pkg.__path__ = [self.root]
        sys.modules["diffsynth"] = pkg
        if "modelscope" not in sys.modules:
            ms = types.ModuleType("modelscope")
            ms.snapshot_download = lambda *a, **k: None
            sys.modules["modelscope"] = ms
        self.phys = importlib.import_module("diffsynth.pipelines.qwen_image_physical")
        self.dit = importlib.import_module("diffsynth.models.qwen_image_dit")
        self.helpers = importlib.import_module("diffsynth.pipelines.helpers")
        self.flow_match = importlib.import_module("diffsynth.schedulers.flow_match")
        return self

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k == "diffsynth" or k.startswith("diffsynth.")]:
            del sys.modules[k]
        sys.modules.update(self._saved)
        return False


def build_reference_dit(ref: ReferenceModules, state_dict, num_layers, dtype, device):
    """The reference's QwenImageDiT (models/qwen_image_dit.py:404) holding `state_dict`, on `device` in `dtype`."""
    import torch
    with torch.device("meta"):
        m = ref.dit.QwenImageDiT(num_layers=num_layers)
    m.load_state_dict({k: v.to(device=device, dtype=dtype).clone() for k, v in state_dict.items()}, assign=True)
    m.pos_embed = ref.dit.QwenEmbedRope(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)      # plain attributes built in the ctor
    return m.eval()


def tiny_dinov2_folder(path):
    """A loadable HF folder for `Dinov2withNorm(dinov2_path=...)` (pipelines/dinov2.py:17): random weights, tiny config --
    the reference pipeline's constructor insists on one (qwen_image_physical.py:198-199) even at inference."""
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    cfg = Dinov2WithRegistersConfig(hidden_size=64, num_hidden_layers=1, num_attention_heads=2, mlp_ratio=2, patch_size=14, image_size=28,
                                    num_register_tokens=4)
    Dinov2WithRegistersModel(cfg).save_pretrained(path)
    return path
