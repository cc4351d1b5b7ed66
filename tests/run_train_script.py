"""Test harness (not product code): runs the reference's scripts/train/train_physicedit.py as `__main__`, unmodified, on this framework in a
fresh process.  What it substitutes, all of it outside the script:
  * `diffsynth.*`            -> physicedit_b200 (compat.install())
  * `accelerate`             -> a single-process stand-in with the eight members the script's loop touches (absent in this image; a real run uses
                                the real package -- the training module is an ordinary nn.Module, `launch_training_task` shows the same loop on DDP)
  * `from_pretrained`        -> a 1-block DiT on the emulated C ABI (tests/abi_emulator.py), 1-layer DINOv2, stub Qwen2.5-VL / VAE: the checkpoint files
                                are not on this machine
Usage: python tests/run_train_script.py <train_physicedit.py> <its command-line flags ...>"""
import contextlib
import importlib.machinery
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import torch  # noqa: E402
from abi_emulator import EmulatedNative  # noqa: E402
from physicedit_b200 import compat, trainers  # noqa: E402
from physicedit_b200.pipeline import QwenImagePhysicPipeline  # noqa: E402
from test_train_script_module import _pipe_on_the_emulator, small_edit_images  # noqa: E402


class Accelerator(trainers._Ranks):
    def __init__(self, gradient_accumulation_steps=1, kwargs_handlers=None, **kw):
        super().__init__()
        self.num_processes = self.world
        self.gradient_accumulation_steps = gradient_accumulation_steps

    def prepare(self, *objects):
        return objects

    def accumulate(self, model):
        return contextlib.nullcontext()

    def backward(self, loss):
        loss.backward()


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def main():
    script, flags = sys.argv[1], sys.argv[2:]
    compat.install()
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:  # noqa: BLE001
            sys.modules["wandb"] = types.SimpleNamespace(log=lambda *a, **k: None, init=lambda *a, **k: None, finish=lambda: None, Image=lambda p: p)
    emu = EmulatedNative()
    pipe = _pipe_on_the_emulator(_Patch(), emu)
    small_edit_images(_Patch())
    QwenImagePhysicPipeline.from_pretrained = staticmethod(lambda **kw: pipe)
    # after transformers has been imported (it probes `accelerate` once, at import): the stand-in serves the script only
    acc, acc_utils = types.ModuleType("accelerate"), types.ModuleType("accelerate.utils")
    acc.__spec__ = importlib.machinery.ModuleSpec("accelerate", None)
    acc.Accelerator, acc.utils = Accelerator, acc_utils
    acc_utils.DistributedDataParallelKwargs = lambda **kw: kw
    sys.modules.setdefault("accelerate", acc)
    sys.modules.setdefault("accelerate.utils", acc_utils)
    torch.manual_seed(0)
    sys.argv = [script] + flags
    runpy.run_path(script, run_name="__main__")
    names = [c[0] for c in emu.calls]
    print(f"[HARNESS] emulated launches: {len(names)}; pe_gemm {names.count('pe_gemm')}; pe_attention_fwd_lse {names.count('pe_attention_fwd_lse')}")


if __name__ == "__main__":
    main()
