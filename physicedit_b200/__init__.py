"""physicedit_b200 -- B200-native (sm_100a) hot path of PhysicEdit / Qwen-Image-Edit behind DiffSynth-Studio's API surface.

Public entry points:
    model_fn_qwen_image      drop-in for pipelines/qwen_image_physical.py:1302-1403 (install as `pipe.model_fn`)
    QwenImagePhysicPipeline  the pipeline with the reference's constructor / loader / LoRA / denoise surface
    QwenImageDiT, adopt_dit  the DiT with the reference's parameter layout; adopt a loaded reference module
    QwenImageVAE, load_vae   the VAE either side of the loop (encode / decode of single images), same parameter layout
    QwenImageTextEncoder     the Qwen2.5-VL encoder / greedy generator in front of the loop
    inject_lora, launch_training_task, DiffusionTrainingModule   un-merged LoRA and the training loop (forward + backward on the native kernels)
    PhysicalEditingDataset, Pica100kDataset, UnifiedDataset, launch_data_process_task   the training-data formats either side of that loop
    QwenImageBlockWiseControlNet / QwenImageBlockwiseMultiControlNet   the optional blockwise controlnet
    FlowMatchScheduler, GeneralLoRALoader, ModelConfig, load_state_dict
Everything numeric runs in physicedit_b200/lib/libpe_b200.so (include/pe_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

_LAZY = {
    "model_fn_qwen_image": ("model_fn", "model_fn_qwen_image"),
    "QwenImagePhysicPipeline": ("pipeline", "QwenImagePhysicPipeline"),
    "ModelConfig": ("pipeline", "ModelConfig"),
    "load_state_dict": ("pipeline", "load_state_dict"),
    "load_dit": ("pipeline", "load_dit"),
    "load_vae": ("pipeline", "load_vae"),
    "QwenImageVAE": ("vae", "QwenImageVAE"),
    "QwenImageDiT": ("dit", "QwenImageDiT"),
    "FlowMatchScheduler": ("scheduler", "FlowMatchScheduler"),
    "GeneralLoRALoader": ("lora", "GeneralLoRALoader"),
    "VisualThinkingDualAdapter": ("adapters", "VisualThinkingDualAdapter"),
    "adopt_dit": ("compat", "adopt_dit"),
    "adopt_adapter": ("compat", "adopt_adapter"),
    "adopt_vae": ("compat", "adopt_vae"),
    "adopt_text_encoder": ("compat", "adopt_text_encoder"),
    "QwenImageTextEncoder": ("text_encoder", "QwenImageTextEncoder"),
    "load_text_encoder": ("text_encoder", "load_text_encoder"),
    "QwenImageBlockWiseControlNet": ("controlnet", "QwenImageBlockWiseControlNet"),
    "QwenImageBlockwiseMultiControlNet": ("controlnet", "QwenImageBlockwiseMultiControlNet"),
    "inject_lora": ("lora", "inject_lora"),
    "merge_lora": ("lora", "merge_lora"),
    "DiffusionTrainingModule": ("trainers", "DiffusionTrainingModule"),
    "launch_training_task": ("trainers", "launch_training_task"),
    "launch_data_process_task": ("trainers", "launch_data_process_task"),
    "PhysicalEditingDataset": ("datasets", "PhysicalEditingDataset"),
    "Pica100kDataset": ("datasets", "Pica100kDataset"),
    "UnifiedDataset": ("unified_dataset", "UnifiedDataset"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)
