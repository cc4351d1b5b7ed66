"""Glue for running the reference's own scripts / pipeline objects on the native path.

* `adopt_dit(ref_dit)` / `adopt_adapter(ref_adapter)` / `adopt_vae(ref_vae)`: wrap modules that the REFERENCE loaded (its ModelManager, LoRA loader,
  `load_state_dict(strict=False)`) into the native classes without copying weights (`load_state_dict(assign=True)`).
* `install()`: registers a `diffsynth` alias package so that the imports used by scripts/inference/*.py and
  scripts/train/train_physicedit.py (`from diffsynth import load_state_dict`, `from diffsynth.pipelines.qwen_image_physical import
  QwenImagePhysicPipeline, ModelConfig`, `from diffsynth.pipelines.flux_image_new import ControlNetInput`,
  `from diffsynth.trainers.utils import DiffusionTrainingModule, ModelLogger, qwen_image_parser, ...`,
  `from diffsynth.trainers.unified_dataset import UnifiedDataset`) resolve here.
"""
from __future__ import annotations

import sys
import types
from dataclasses import dataclass
from typing import Optional

import torch


def adopt_dit(ref_dit: torch.nn.Module):
    from .dit import QwenImageDiT
    n = len(ref_dit.transformer_blocks)
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=n)
    dit.load_state_dict(ref_dit.state_dict(), assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    for i, b in enumerate(dit.transformer_blocks):
        object.__setattr__(b, "_owner", (dit, i))
    return dit.eval()


def adopt_adapter(ref_adapter: torch.nn.Module):
    from .adapters import VisualThinkingDualAdapter
    with torch.device("meta"):
        ad = VisualThinkingDualAdapter(3584, 3584, ref_adapter.t_min, ref_adapter.t_max)
    ad.load_state_dict(ref_adapter.state_dict(), assign=True)
    return ad.eval()


def adopt_vae(ref_vae: torch.nn.Module):
    """The reference's loaded QwenImageVAE (models/qwen_image_vae.py:640) -> the native module on the same parameter storage."""
    from .vae import QwenImageVAE
    with torch.device("meta"):
        vae = QwenImageVAE()
    vae.load_state_dict(ref_vae.state_dict(), assign=True)
    return vae.eval()


def adopt_text_encoder(ref_te: torch.nn.Module):
    """The reference's loaded QwenImageTextEncoderWithDecode (models/qwen_image_text_encoder_withdecode.py:6, a transformers
    Qwen2_5_VLForConditionalGeneration) -> the native module on the same parameter storage (7B config of the wrapper)."""
    from .text_encoder import QwenImageTextEncoder
    with torch.device("meta"):
        te = QwenImageTextEncoder()
    te.load_state_dict(ref_te.state_dict(), assign=True)
    return te.eval()


@dataclass
class ControlNetInput:
    """pipelines/flux_image_new.py:5-13: one blockwise-controlnet request (physicedit_b200/controlnet.py)."""
    controlnet_id: int = 0
    scale: float = 1.0
    start: float = 1.0
    end: float = 0.0
    image: Optional[object] = None
    inpaint_mask: Optional[object] = None
    processor_id: Optional[str] = None


def install() -> None:
    from . import pipeline, scheduler, lora, dit, adapters, model_fn, vae, units, trainers

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    root = mod("diffsynth", load_state_dict=pipeline.load_state_dict, ModelConfig=pipeline.ModelConfig)
    root.__path__ = []
    mod("diffsynth.pipelines").__path__ = []
    mod("diffsynth.pipelines.qwen_image_physical", QwenImagePhysicPipeline=pipeline.QwenImagePhysicPipeline, ModelConfig=pipeline.ModelConfig,
        model_fn_qwen_image=model_fn.model_fn_qwen_image, SPECIAL_TOKEN_NUM=pipeline.SPECIAL_TOKEN_NUM,
        **{k: getattr(units, k) for k in dir(units) if k.startswith("QwenImageUnit_") or k in ("PipelineUnit", "PipelineUnitRunner", "SYSTEM_PROMPT_SAMPLE")})
    mod("diffsynth.utils", BasePipeline=pipeline.QwenImagePhysicPipeline, ModelConfig=pipeline.ModelConfig, PipelineUnit=units.PipelineUnit,
        PipelineUnitRunner=units.PipelineUnitRunner)
    mod("diffsynth.trainers").__path__ = []
    mod("diffsynth.trainers.utils", **{k: getattr(trainers, k) for k in ("DiffusionTrainingModule", "ModelLogger", "qwen_image_parser",
                                                                          "launch_training_task", "launch_data_process_task", "PhysicalEditingDataset", "Pica100kDataset")})
    from . import unified_dataset
    mod("diffsynth.trainers.unified_dataset", **{n: getattr(unified_dataset, n) for n in dir(unified_dataset)
                                                 if isinstance(getattr(unified_dataset, n), type) and getattr(unified_dataset, n).__module__ == unified_dataset.__name__})
    mod("diffsynth.pipelines.flux_image_new", ControlNetInput=ControlNetInput)
    mod("diffsynth.pipelines.helpers", **{k: getattr(adapters, k) for k in ("FeedForward", "PerceiverAttention", "PerceiverResampler",
                                                                              "VisualThinkingAdapter", "VisualThinkingDualAdapter")})
    mod("diffsynth.pipelines.dinov2", Dinov2withNorm=adapters.Dinov2withNorm)
    mod("diffsynth.models").__path__ = []
    mod("diffsynth.models.qwen_image_dit", QwenImageDiT=dit.QwenImageDiT, QwenImageTransformerBlock=dit.QwenImageTransformerBlock,
        QwenEmbedRope=dit.QwenEmbedRope, RMSNorm=dit.RMSNorm)
    mod("diffsynth.models.qwen_image_vae", QwenImageVAE=vae.QwenImageVAE, QwenImageVAEStateDictConverter=vae.QwenImageVAEStateDictConverter)
    mod("diffsynth.models.utils", load_state_dict=pipeline.load_state_dict, hash_state_dict_keys=pipeline.hash_state_dict_keys,
        RMSNorm=dit.RMSNorm, AdaLayerNorm=dit.AdaLayerNorm, TimestepEmbeddings=dit.TimestepEmbeddings)
    mod("diffsynth.schedulers").__path__ = []
    mod("diffsynth.schedulers.flow_match", FlowMatchScheduler=scheduler.FlowMatchScheduler)
    mod("diffsynth.lora", GeneralLoRALoader=lora.GeneralLoRALoader)
