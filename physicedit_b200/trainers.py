"""The `diffsynth.trainers.*` names scripts/train/train_physicedit.py imports (:1-6), so that the script resolves against this
framework under `compat.install()`.

In scope here is only what touches the hot path's checkpoint / module contract: `DiffusionTrainingModule` (which parameters are
trainable, LoRA injection, the `pipe.dit.`-stripped trainable-only state dict that scripts/inference/validate.py:44-65 later splits into
LoRA and `pipe.*` keys), `ModelLogger` (who writes that file), the argument parser's flag set and `launch_training_task` (the optimizer loop,
on plain torch.distributed DDP instead of accelerate; forward and backward on the native kernels) and `PhysicalEditingDataset`,
the training-data format the script instantiates (`physicedit_b200/datasets.py`), `launch_data_process_task` and -- the reader of what that task caches --
`UnifiedDataset` (`physicedit_b200/unified_dataset.py`).
Reference: DiffSynth-Studio/diffsynth/trainers/utils.py:777-1115.
"""
from __future__ import annotations

import argparse
import json
import os

import torch
import torch.distributed as dist

from .pipeline import ModelConfig


class DiffusionTrainingModule(torch.nn.Module):
    """trainers/utils.py:777-888."""

    def to(self, *args, **kwargs):
        for _, model in self.named_children():
            model.to(*args, **kwargs)
        return self

    def trainable_modules(self):
        return (p for p in self.parameters() if p.requires_grad)

    def trainable_param_names(self):
        return {n for n, p in self.named_parameters() if p.requires_grad}

    def add_lora_to_model(self, model, target_modules, lora_rank, lora_alpha=None, upcast_dtype=None):
        """Injection of un-merged LoRA (:799-808: `inject_adapter_in_model(LoraConfig(r, lora_alpha, target_modules), model)`).  peft is not a
        dependency here: physicedit_b200.lora.inject_lora restates its Linear layer (same parameter names `lora_A.default.weight` /
        `lora_B.default.weight`, same init, same target matching, base weights frozen) and the forward / backward of the wrapped linears run
        on the native GEMMs (physicedit_b200.autograd)."""
        from .lora import inject_lora
        model = inject_lora(model, target_modules, lora_rank, lora_alpha)
        if upcast_dtype is not None:
            for p in model.parameters():
                if p.requires_grad:
                    p.data = p.to(upcast_dtype)
        return model

    def mapping_lora_state_dict(self, state_dict):
        """PEFT keys without an adapter name get `.default` (:811-819): the layout GeneralLoRALoader.get_name_dict expects."""
        out = {}
        for k, v in state_dict.items():
            if "lora_A.default.weight" in k or "lora_B.default.weight" in k:
                out[k] = v
            elif "lora_A.weight" in k or "lora_B.weight" in k:
                out[k.replace("lora_A.weight", "lora_A.default.weight").replace("lora_B.weight", "lora_B.default.weight")] = v
        return out

    def export_trainable_state_dict(self, state_dict, remove_prefix=None):
        """Trainable parameters only, `remove_prefix` (normally "pipe.dit.") stripped (:822-832): LoRA keys come out as
        `transformer_blocks.N...lora_A.default.weight`, adapter keys keep `pipe.visual_thinking_adapter...`."""
        names = self.trainable_param_names()
        out = {}
        for k, v in state_dict.items():
            if k in names:
                out[k[len(remove_prefix):] if (remove_prefix and k.startswith(remove_prefix)) else k] = v
        return out

    def transfer_data_to_device(self, data, device, torch_float_dtype=None):
        for k, v in data.items():
            if isinstance(v, torch.Tensor):
                v = v.to(device)
                if torch_float_dtype is not None and v.is_floating_point():
                    v = v.to(torch_float_dtype)
                data[k] = v
        return data

    def parse_model_configs(self, model_paths, model_id_with_origin_paths, enable_fp8_training=False, local_model_path=None, skip_download=False):
        if enable_fp8_training:
            raise NotImplementedError("fp8 weight storage (enable_fp8_training) is outside the hot path (SURVEY 8f5)")
        cfgs = [ModelConfig(path=p) for p in json.loads(model_paths)] if model_paths is not None else []
        for spec in (model_id_with_origin_paths.split(",") if model_id_with_origin_paths else []):
            mid, pattern = spec.split(":")
            cfgs.append(ModelConfig(model_id=mid, origin_file_pattern=pattern, local_model_path=local_model_path, skip_download=skip_download))
        return cfgs


    def switch_pipe_to_training_mode(self, pipe, trainable_models, lora_base_model, lora_target_modules, lora_rank, lora_checkpoint=None,
                                     enable_fp8_training=False):
        """:856-888, called from the train script's module constructor (scripts/train/train_physicedit.py:216-220): the 1000-step training
        table of the scheduler, everything frozen except `trainable_models`, un-merged LoRA injected into `lora_base_model` (optionally
        initialised from a checkpoint in either PEFT key layout)."""
        if enable_fp8_training:
            raise NotImplementedError("fp8 weight storage (enable_fp8_training) is outside the hot path (SURVEY 8f5)")
        pipe.scheduler.set_timesteps(1000, training=True)
        pipe.freeze_except([] if trainable_models is None else trainable_models.split(","))
        if lora_base_model is not None:
            model = self.add_lora_to_model(getattr(pipe, lora_base_model), target_modules=lora_target_modules.split(","), lora_rank=lora_rank,
                                           upcast_dtype=pipe.torch_dtype)
            if lora_checkpoint is not None:
                from .pipeline import load_state_dict
                state_dict = self.mapping_lora_state_dict(load_state_dict(lora_checkpoint))
                result = model.load_state_dict(state_dict, strict=False)
                print(f"LoRA checkpoint loaded: {lora_checkpoint}, total {len(state_dict)} keys")
                if len(result.unexpected_keys) > 0:
                    print(f"Warning, LoRA key mismatch! Unexpected keys in LoRA checkpoint: {result.unexpected_keys}")
            setattr(pipe, lora_base_model, model)


class ModelLogger:
    """trainers/utils.py:891-929: rank 0 writes the trainable-only state dict as safetensors at epoch end / every save_steps."""

    def __init__(self, output_path, remove_prefix_in_ckpt=None, state_dict_converter=lambda x: x):
        self.output_path = output_path
        self.remove_prefix_in_ckpt = remove_prefix_in_ckpt
        self.state_dict_converter = state_dict_converter
        self.num_steps = 0

    def on_step_end(self, accelerator, model, save_steps=None):
        self.num_steps += 1
        if save_steps is not None and self.num_steps % save_steps == 0:
            self.save_model(accelerator, model, f"step-{self.num_steps}.safetensors")

    def on_epoch_end(self, accelerator, model, epoch_id):
        self.save_model(accelerator, model, f"epoch-{epoch_id}.safetensors")

    def on_training_end(self, accelerator, model, save_steps=None):
        if save_steps is not None and self.num_steps % save_steps != 0:
            self.save_model(accelerator, model, f"step-{self.num_steps}.safetensors")

    def save_model(self, accelerator, model, file_name):
        accelerator.wait_for_everyone()
        if accelerator.is_main_process:
            sd = accelerator.unwrap_model(model).export_trainable_state_dict(accelerator.get_state_dict(model), remove_prefix=self.remove_prefix_in_ckpt)
            os.makedirs(self.output_path, exist_ok=True)
            accelerator.save(self.state_dict_converter(sd), os.path.join(self.output_path, file_name), safe_serialization=True)


# flag, type (None = store_true), default, required  -- the flag set of qwen_image_parser (:1072-1115)
_FLAGS = [
    ("dataset_base_path", str, "", True), ("dataset_metadata_path", str, None, False), ("max_pixels", int, 1024 * 1024, False), ("height", int, None, False),
    ("width", int, None, False), ("data_file_keys", str, "image", False), ("dataset_repeat", int, 1, False), ("model_paths", str, None, False),
    ("model_id_with_origin_paths", str, None, False), ("tokenizer_path", str, None, False), ("learning_rate", float, 1e-4, False), ("num_epochs", int, 1, False),
    ("output_path", str, "./models", False), ("remove_prefix_in_ckpt", str, "pipe.dit.", False), ("trainable_models", str, None, False),
    ("lora_base_model", str, None, False), ("lora_target_modules", str, "q,k,v,o,ffn.0,ffn.2", False), ("lora_rank", int, 32, False),
    ("lora_checkpoint", str, None, False), ("extra_inputs", str, None, False), ("use_gradient_checkpointing", None, False, False),
    ("use_gradient_checkpointing_offload", None, False, False), ("gradient_accumulation_steps", int, 1, False), ("find_unused_parameters", None, False, False),
    ("save_steps", int, None, False), ("dataset_num_workers", int, 0, False), ("weight_decay", float, 0.01, False), ("processor_path", str, None, False),
    ("enable_fp8_training", None, False, False), ("task", str, "sft", False), ("num_frames", int, 81, False), ("wandb_project", str, None, False),
    ("wandb_run_name", str, None, False), ("save_every_n_steps", int, None, False), ("eval_every_n_steps", int, None, False), ("resume_from", str, None, False),
    ("resume_original_num_processes", int, 4, False), ("local_model_path", str, None, False), ("dinov2_path", str, None, True),
]


def qwen_image_parser():
    ap = argparse.ArgumentParser(description="Qwen-Image / PhysicEdit training arguments (flag-compatible with the reference's qwen_image_parser).")
    for name, typ, default, required in _FLAGS:
        if typ is None:
            ap.add_argument(f"--{name}", default=False, action="store_true")
        else:
            ap.add_argument(f"--{name}", type=typ, default=default, required=required)
    ap.add_argument("--resume_type", type=str, choices=["auto", "full", "model"], default="auto")
    return ap


class _Ranks:
    """The four things the reference asks of `accelerate.Accelerator` in its loop and its ModelLogger (:891-977), on plain torch.distributed."""

    def __init__(self):
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.rank = dist.get_rank() if self.distributed else 0
        self.world = dist.get_world_size() if self.distributed else 1
        self.is_main_process = self.rank == 0

    def wait_for_everyone(self):
        if self.distributed:
            dist.barrier()

    def unwrap_model(self, model):
        return model.module if isinstance(model, torch.nn.parallel.DistributedDataParallel) else model

    def get_state_dict(self, model):
        return self.unwrap_model(model).state_dict()

    def save(self, state_dict, path, safe_serialization=True):
        from safetensors.torch import save_file
        save_file({k: v.detach().contiguous().cpu() for k, v in state_dict.items()}, path)


def launch_training_task(dataset, model, model_logger, learning_rate: float = 1e-5, weight_decay: float = 1e-2, num_workers: int = 8, save_steps: int = None,
                         num_epochs: int = 1, gradient_accumulation_steps: int = 1, find_unused_parameters: bool = False, args=None):
    """trainers/utils.py:932-977 without accelerate: AdamW over `model.trainable_modules()`, constant LR, one sample per step per rank, DDP
    (NCCL gradient all-reduce, `find_unused_parameters` as the reference passes it: the last block's text tail has no gradient) when a process
    group with more than one rank is up, gradient accumulation through `no_sync`, ModelLogger hooks at the same points.  The backward runs on
    the native kernels (physicedit_b200/autograd.py)."""
    import contextlib
    if args is not None:
        learning_rate, weight_decay, num_workers = args.learning_rate, args.weight_decay, args.dataset_num_workers
        save_steps, num_epochs = args.save_steps, args.num_epochs
        gradient_accumulation_steps, find_unused_parameters = args.gradient_accumulation_steps, args.find_unused_parameters
    ranks = _Ranks()
    optimizer = torch.optim.AdamW(model.trainable_modules(), lr=learning_rate, weight_decay=weight_decay)
    scheduler = torch.optim.lr_scheduler.ConstantLR(optimizer)
    sampler = torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=ranks.world, rank=ranks.rank, shuffle=True) if ranks.distributed else None
    loader = torch.utils.data.DataLoader(dataset, shuffle=sampler is None, sampler=sampler, collate_fn=lambda x: x[0], num_workers=num_workers)
    wrapped = model
    if ranks.distributed:
        cuda = next((p.device for p in model.parameters() if p.is_cuda), None)
        wrapped = torch.nn.parallel.DistributedDataParallel(model, device_ids=None if cuda is None else [cuda.index],
                                                            find_unused_parameters=find_unused_parameters)
    micro = 0
    for epoch_id in range(num_epochs):
        if sampler is not None:
            sampler.set_epoch(epoch_id)
        for data in loader:
            if data is None:                                  # a clip that could not be decoded (the dataset has warned)
                if ranks.distributed:                         # skipping on one rank only would leave the others waiting in the gradient all-reduce
                    raise RuntimeError("the dataset returned no sample (undecodable clip); remove it from the dataset before a multi-rank run")
                continue
            if micro % gradient_accumulation_steps == 0:
                optimizer.zero_grad()                         # at the start of an accumulation window, like the reference's loop (:966)
            micro += 1
            boundary = micro % gradient_accumulation_steps == 0
            sync = contextlib.nullcontext() if (boundary or not ranks.distributed) else wrapped.no_sync()
            with sync:
                loss = wrapped({}, inputs=data) if getattr(dataset, "load_from_cache", False) else wrapped(data)
                (loss / gradient_accumulation_steps).backward()
            if boundary:
                optimizer.step()
                scheduler.step()
            model_logger.on_step_end(ranks, wrapped, save_steps)
        if save_steps is None:
            model_logger.on_epoch_end(ranks, wrapped, epoch_id)
    model_logger.on_training_end(ranks, wrapped, save_steps)
    return wrapped


def launch_data_process_task(dataset, model, model_logger, num_workers: int = 8, args=None):
    """trainers/utils.py:980-1002 without accelerate: run the module's pre-processing (`model(data, return_inputs=True)`: every pipeline unit -- VAE /
    text-encoder / DINOv2 forwards on the native kernels) over the dataset once and cache each sample's inputs as
    `<output_path>/<rank>/<index>.pth`; ranks take the samples `rank, rank + world, ...` and number their own files from 0, like a dataloader
    sharded by `accelerator.prepare`."""
    if args is not None:
        num_workers = args.dataset_num_workers
    ranks = _Ranks()
    sampler = torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=ranks.world, rank=ranks.rank, shuffle=False) if ranks.distributed else None
    loader = torch.utils.data.DataLoader(dataset, shuffle=False, sampler=sampler, collate_fn=lambda x: x[0], num_workers=num_workers)
    folder = os.path.join(model_logger.output_path, str(ranks.rank))
    os.makedirs(folder, exist_ok=True)
    for data_id, data in enumerate(loader):
        with torch.no_grad():
            torch.save(model(data, return_inputs=True), os.path.join(folder, f"{data_id}.pth"))


from .datasets import PhysicalEditingDataset, Pica100kDataset  # noqa: E402,F401  (trainers/utils.py:369-683, :685-775)


from .unified_dataset import UnifiedDataset  # noqa: E402,F401  (trainers/unified_dataset.py)
