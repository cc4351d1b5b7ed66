"""QwenImagePhysicPipeline with the reference's API surface, the denoise loop on the native path.

Mirrors DiffSynth-Studio/diffsynth/pipelines/qwen_image_physical.py (class QwenImagePhysicPipeline:
__init__ :185-247, load_lora :250-276, training_loss :313-329, enable_vram_management :375-494,
from_pretrained :497-541, __call__ :544-669) and diffsynth/utils/__init__.py (BasePipeline, ModelConfig).

In scope here: the in-iteration models (`dit`, `visual_thinking_adapter`, `blockwise_controlnet`), the scheduler,
the CFG denoise loop, LoRA fold / hot-load, checkpoint key layout, the training-path feature extractors (DINOv2, resamplers) and
`training_loss`, and either side of the loop the VAE (8f1, `vae.py`), the Qwen2.5-VL text encoder (8f2, `text_encoder.py`) and the
pre-loop units (8b, `units.py`) -- each loaded by registry hash like the DiT.  `denoise(...)` also takes pre-computed prompt
embeddings / latents directly, which is what bench.py and the parity tests drive.
"""
from __future__ import annotations

import glob
import os
from dataclasses import dataclass
from typing import Optional, Union

import torch
import torch.nn as nn

from . import native as nv
from .adapters import Dinov2withNorm, PerceiverResampler, VisualThinkingAdapter, VisualThinkingDualAdapter
from .dit import QwenImageDiT
from .lora import GeneralLoRALoader
from .model_fn import model_fn_qwen_image, prompt_lengths
from .scheduler import FlowMatchScheduler
from .units import PipelineUnit, PipelineUnitRunner, QwenImageUnit_PhysicalVerbalEmbedder, QwenImageUnit_PhysicalVisualEmbedder, default_units
from .vae import QwenImageVAE

SPECIAL_TOKEN_NUM = 64        # qwen_image_physical.py:28
DIT_KEY_HASH = "0319a1cb19835fb510907dd3367c95ff"      # configs/model_config.py:21
VAE_KEY_HASH = "ed4ea5824d55ec3107b09815e318123a"      # configs/model_config.py:24


@dataclass
class ModelConfig:
    """diffsynth/utils/__init__.py:160-220 minus the downloaders (no network): resolves local files only."""
    path: Union[str, list] = None
    model_id: str = None
    origin_file_pattern: Union[str, list] = None
    download_resource: str = "ModelScope"
    offload_device: Optional[Union[str, torch.device]] = None
    offload_dtype: Optional[torch.dtype] = None
    local_model_path: str = None
    skip_download: bool = True

    def download_if_necessary(self, use_usp=False):
        if self.path is not None:
            return
        if self.model_id is None:
            raise ValueError("No valid model files. Please use `ModelConfig(path=...)` or `ModelConfig(model_id=..., origin_file_pattern=...)`.")
        base = self.local_model_path or "./models"
        pattern = self.origin_file_pattern or ""
        target = os.path.join(base, self.model_id, pattern)
        if isinstance(pattern, str) and (pattern == "" or pattern.endswith("/")):
            self.path = target
            return
        found = sorted(glob.glob(target))
        self.path = found[0] if len(found) == 1 else found


def load_state_dict(file_path, torch_dtype=None, device="cpu"):
    """diffsynth.load_state_dict (models/utils.py:65-88)."""
    if isinstance(file_path, (list, tuple)):
        sd = {}
        for p in file_path:
            sd.update(load_state_dict(p, torch_dtype, device))
        return sd
    if file_path.endswith(".safetensors"):
        from safetensors import safe_open
        sd = {}
        with safe_open(file_path, framework="pt", device=str(device)) as f:
            for k in f.keys():
                t = f.get_tensor(k)
                sd[k] = t.to(torch_dtype) if torch_dtype is not None else t
        return sd
    sd = torch.load(file_path, map_location=device, weights_only=True)
    if torch_dtype is not None:
        sd = {k: (v.to(torch_dtype) if isinstance(v, torch.Tensor) else v) for k, v in sd.items()}
    return sd


def hash_state_dict_keys(state_dict, with_shape=True):
    """models/utils.py:148-182."""
    import hashlib
    keys = []
    for key, value in state_dict.items():
        if isinstance(key, str) and isinstance(value, torch.Tensor):
            if with_shape:
                keys.append(key + ":" + "_".join(map(str, list(value.shape))))
            keys.append(key)
    keys.sort()
    return hashlib.md5(",".join(keys).encode("UTF-8")).hexdigest()


def load_dit(path, torch_dtype=torch.bfloat16, device="cuda", state_dict=None) -> Optional[QwenImageDiT]:
    """ModelManager.load_model for the one registry row on this path (model_manager.py:350-376): detect by key hash,
    init on the meta device, `load_state_dict(assign=True)`, move.  Unknown files print and return None like the reference."""
    sd = state_dict if state_dict is not None else load_state_dict(path, torch_dtype=torch_dtype, device="cpu")
    if hash_state_dict_keys(sd) != DIT_KEY_HASH:
        print(f"    We cannot detect the model type. No models are loaded ({path}).")
        return None
    with torch.device("meta"):
        dit = QwenImageDiT()
    dit.load_state_dict(sd, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    for i, b in enumerate(dit.transformer_blocks):
        object.__setattr__(b, "_owner", (dit, i))
    return dit.to(dtype=torch_dtype, device=device).eval()


def load_vae(path, torch_dtype=torch.bfloat16, device="cuda", state_dict=None) -> Optional[QwenImageVAE]:
    """The `qwen_image_vae` registry row (configs/model_config.py:24; converter `from_diffusers` is the identity,
    models/qwen_image_vae.py:737-742): detect by key hash, meta-init, assign.  Unknown files print and return None."""
    sd = state_dict if state_dict is not None else load_state_dict(path, torch_dtype=torch_dtype, device="cpu")
    if hash_state_dict_keys(sd) != VAE_KEY_HASH:
        print(f"    We cannot detect the model type. No models are loaded ({path}).")
        return None
    with torch.device("meta"):
        vae = QwenImageVAE()
    vae.load_state_dict(sd, assign=True)
    return vae.to(dtype=torch_dtype, device=device).eval()


class QwenImagePhysicPipeline(nn.Module):
    # Set to a 2-rank torch.distributed group (parallel.make_cfg_pairs) to split the two CFG branches of ONE image over two GPUs.
    cfg_parallel_group = None
    # 2: the two CFG branches of a denoise step run concurrently on two CUDA streams (see run_cfg_branches); 1: back to back.
    # Bit-identical results either way; two streams measured +3 % steps/s at 1024^2 (r1, power-capped B200).
    cfg_streams = 2

    def __init__(self, device="cuda", torch_dtype=torch.bfloat16, dinov2_path=None, dinov2_config: dict = None, build_training_path: bool = True):
        super().__init__()
        self.device, self.torch_dtype = device, torch_dtype
        self.height_division_factor = self.width_division_factor = 16
        self.scheduler = FlowMatchScheduler(sigma_min=0, sigma_max=1, extra_one_step=True, exponential_shift=True,
                                            exponential_shift_mu=0.8, shift_terminal=0.02)
        self.text_encoder = None
        self.dit: QwenImageDiT = None
        self.vae = None
        if build_training_path:
            assert dinov2_path is not None or dinov2_config is not None, "dinov2_path must be provided (path to DINOv2-with-registers-base)"
            self.dinov2 = Dinov2withNorm(dinov2_path=dinov2_path, config=dinov2_config).to(device=device, dtype=torch_dtype)
            self.dinov2_mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
            self.dinov2_std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
            self.dino_resampler = PerceiverResampler(dim=768, num_latents=SPECIAL_TOKEN_NUM, depth=2).to(device=device, dtype=torch_dtype)
            self.dino_time_embed = nn.Embedding(6, 768).to(device=device, dtype=torch_dtype)
            self.dino_resampler_adapter = VisualThinkingAdapter(in_dim=768, out_dim=3584).to(device=device, dtype=torch_dtype)
            self.dino_input_size = 224
            self.vae_resampler = PerceiverResampler(dim=64, num_latents=SPECIAL_TOKEN_NUM, depth=2, max_num_media_tokens=10240).to(device=device, dtype=torch_dtype)
            self.vae_time_embed = nn.Embedding(6, 64).to(device=device, dtype=torch_dtype)
            self.vae_resampler_adapter = VisualThinkingAdapter(in_dim=64, out_dim=3584).to(device=device, dtype=torch_dtype)
        self.visual_thinking_adapter = VisualThinkingDualAdapter(in_dim=3584, out_dim=3584, t_min=self.scheduler.timesteps.min().item(),
                                                                 t_max=self.scheduler.timesteps.max().item()).to(device=device, dtype=torch_dtype)
        self.blockwise_controlnet = None
        self.tokenizer = None
        self.processor = None
        self.boi_token_id = self.eoi_token_id = None
        self.unit_runner = PipelineUnitRunner()
        self.in_iteration_models = ("dit", "blockwise_controlnet", "visual_thinking_adapter")
        self.units = default_units()
        self.model_fn = model_fn_qwen_image
        self.special_token_loss = 0.0
        self.vram_management_enabled = False

    # ---- reference API ------------------------------------------------------------------------------
    @staticmethod
    def from_pretrained(torch_dtype=torch.bfloat16, device="cuda", model_configs=(), tokenizer_config="default", processor_config=None,
                        dinov2_path=None):
        """:497-541.  Model files are recognised by the md5 of their state-dict keys + shapes like the reference's ModelManager
        (models/model_manager.py:350-376; registry rows configs/model_config.py:21-24): DiT, VAE and the Qwen2.5-VL text encoder.
        `tokenizer_config` / `processor_config` point at local folders (no network): Qwen2Tokenizer / Qwen2VLProcessor, plus the 66
        added special tokens `<begin_of_img> <end_of_img> <img0..63>` (:528-539)."""
        pipe = QwenImagePhysicPipeline(device=device, torch_dtype=torch_dtype, dinov2_path=dinov2_path)
        for cfg in model_configs:
            cfg.download_if_necessary()
            paths = cfg.path if isinstance(cfg.path, list) else [cfg.path]
            dtype = cfg.offload_dtype or torch_dtype
            try:
                sd = load_state_dict(paths if len(paths) > 1 else paths[0], torch_dtype=dtype, device="cpu")
                key_hash = hash_state_dict_keys(sd)
                if key_hash == VAE_KEY_HASH:
                    pipe.vae = load_vae(paths, torch_dtype=dtype, device=device, state_dict=sd)
                elif key_hash == DIT_KEY_HASH:
                    pipe.dit = load_dit(paths, torch_dtype=dtype, device=device, state_dict=sd)
                elif "controlnet_blocks.0.x_rms.weight" in sd and "img_in.weight" in sd:
                    # blockwise controlnets (models/qwen_image_controlnet.py; fetched with index="all" at :516-518): every file is one more entry
                    from .controlnet import QwenImageBlockWiseControlNet, QwenImageBlockwiseMultiControlNet
                    layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("controlnet_blocks."))
                    with torch.device("meta"):
                        cn = QwenImageBlockWiseControlNet(num_layers=layers, additional_in_dim=sd["img_in.weight"].shape[1] - 64)
                    cn.load_state_dict({k: v.to(dtype) for k, v in sd.items()}, assign=True)
                    cn = cn.to(device).eval()
                    nets = list(pipe.blockwise_controlnet.models) if pipe.blockwise_controlnet is not None else []
                    pipe.blockwise_controlnet = QwenImageBlockwiseMultiControlNet(nets + [cn])
                else:
                    from .text_encoder import load_text_encoder
                    te = load_text_encoder(sd, torch_dtype=dtype, device=device)
                    if te is not None:
                        pipe.text_encoder = te
                    else:
                        print(f"    We cannot detect the model type. No models are loaded ({paths}).")
            except Exception as e:  # noqa: BLE001  (the reference's loader prints and moves on, model_manager.py:375-376)
                print(f"    Loading {paths} failed: {e}")
        if isinstance(tokenizer_config, str) and tokenizer_config == "default":       # the reference's default argument (:502); only read with a text encoder
            tokenizer_config = ModelConfig(model_id="Qwen/Qwen-Image", origin_file_pattern="tokenizer/")
        pipe.attach_tokenizer(tokenizer_config, processor_config)
        return pipe

    def attach_tokenizer(self, tokenizer_config=None, processor_config=None, tokenizer=None, processor=None):
        """:522-539: tokenizer (only needed with a text encoder), processor, and the special tokens whose ids delimit the 64 rows the
        adapter rewrites.  Objects can be handed in directly (tests); otherwise they are loaded from the configs' local paths."""
        if tokenizer is None and tokenizer_config is not None and self.text_encoder is not None:
            tokenizer_config.download_if_necessary()
            from transformers import Qwen2Tokenizer
            tokenizer = Qwen2Tokenizer.from_pretrained(tokenizer_config.path)
        if processor is None and processor_config is not None:
            processor_config.download_if_necessary()
            from transformers import Qwen2VLProcessor
            processor = Qwen2VLProcessor.from_pretrained(processor_config.path)
        self.tokenizer = tokenizer if tokenizer is not None else self.tokenizer
        if processor is not None:
            self.processor = processor
            processor.tokenizer.add_special_tokens({"additional_special_tokens": ["<begin_of_img>", "<end_of_img>"] + [f"<img{i}>" for i in range(SPECIAL_TOKEN_NUM)]})
            self.boi_token_id = processor.tokenizer.convert_tokens_to_ids("<begin_of_img>")
            self.eoi_token_id = processor.tokenizer.convert_tokens_to_ids("<end_of_img>")
        return self

    def to(self, *args, **kwargs):
        """BasePipeline.to (utils/__init__.py:33-40): device / dtype of the intermediates follow the move."""
        device, dtype, _, _ = torch._C._nn._parse_to(*args, **kwargs)
        if device is not None:
            self.device = device
        if dtype is not None:
            self.torch_dtype = dtype
        super().to(*args, **kwargs)
        return self

    def load_lora(self, module: nn.Module, lora_config=None, alpha=1, hotload=False, state_dict=None):
        """:250-276.  hotload=False folds `alpha * B @ A` into the weights (GeneralLoRALoader); hotload=True attaches the factors un-merged to the
        wrappers `enable_lora_magic()` installed (none installed -> nothing attaches, like the reference's isinstance filter)."""
        if state_dict is None:
            if not isinstance(lora_config, str):
                lora_config.download_if_necessary()              # a ModelConfig given by model_id / origin_file_pattern resolves to local files here
                lora_config = lora_config.path
            state_dict = load_state_dict(lora_config, torch_dtype=self.torch_dtype, device=self.device)
        if hotload:
            from .lora import hotload_lora
            hotload_lora(module, state_dict, alpha=alpha)
            return
        GeneralLoRALoader(torch_dtype=self.torch_dtype, device=self.device).load(module, state_dict, alpha=alpha)

    def clear_lora(self):
        """:279-285: detach every hot-loaded LoRA."""
        from .lora import clear_hot_lora
        for _, child in self.named_children():
            clear_hot_lora(child)

    def enable_lora_magic(self):
        """:288-305: make the DiT's linears able to carry un-merged LoRAs (`load_lora(..., hotload=True)`)."""
        if self.dit is not None:
            from .lora import enable_hot_lora
            enable_hot_lora(self.dit)

    def get_special_divisor(self, global_step=None, warmup_steps=10000):
        """:308-311."""
        return 10 - 9.0 * min(1.0, global_step / float(warmup_steps))

    def direct_distill_loss(self, **inputs):
        """:332-340: run the whole sampler under autograd and regress its end point on `input_latents`."""
        self.scheduler.set_timesteps(inputs["num_inference_steps"])
        models = {name: getattr(self, name) for name in self.in_iteration_models if name not in inputs}
        for progress_id, timestep in enumerate(self.scheduler.timesteps):
            timestep = timestep.unsqueeze(0).to(dtype=self.torch_dtype, device=self.device)
            noise_pred, _ = self.model_fn(**models, **inputs, timestep=timestep, progress_id=progress_id)
            inputs["latents"] = self.step(self.scheduler, progress_id=progress_id, noise_pred=noise_pred, **inputs)
        return torch.nn.functional.mse_loss(inputs["latents"].float(), inputs["input_latents"].float())

    def enable_vram_management(self, *args, **kwargs):
        """Accepted for API compatibility (inference_pica.py:246): weights stay resident -- a B200 has 180 GB."""
        self.vram_management_enabled = False

    def load_models_to_device(self, model_names=()):
        return None

    def enable_sequence_parallel(self, group=None):
        """Latency mode for large images: ONE image on all ranks of `group` -- every rank denoises the same request, each DiT forward
        is split across the ranks (ulysses.py: rows sequence-parallel, attention head-parallel, the two all-to-alls fused into the QKV GEMM's and
        the attention kernel's epilogues as NVLink P2P stores).  Call on every rank after the weights are in place; `group=None` = WORLD."""
        from .ulysses import UlyssesContext
        self.dit.engine().sp = UlyssesContext(group, device=next(self.dit.parameters()).device)
        self._cfg_streams_before_sp = getattr(self, "cfg_streams", 1)
        self.cfg_streams = 1            # the per-block barriers are stream-ordered: both CFG branches run on one stream
        return self

    def disable_sequence_parallel(self):
        """Back to one whole forward per rank (the symmetric workspaces stay allocated with the context until it is collected)."""
        self.dit.engine().sp = None
        self.cfg_streams = getattr(self, "_cfg_streams_before_sp", self.cfg_streams)
        return self

    def freeze_except(self, model_names):
        for name, model in self.named_children():
            if name in model_names:
                model.train(); model.requires_grad_(True)
            else:
                model.eval(); model.requires_grad_(False)

    def check_resize_height_width(self, height, width):
        """diffsynth/utils/__init__.py:43-52: silently rounds UP to multiples of 16."""
        if height % 16 != 0:
            height = (height + 15) // 16 * 16
            print(f"height % 16 != 0. We round it up to {height}.")
        if width % 16 != 0:
            width = (width + 15) // 16 * 16
            print(f"width % 16 != 0. We round it up to {width}.")
        return height, width

    def preprocess_image(self, image, torch_dtype=None, device=None, pattern="B C H W", min_value=-1, max_value=1):
        """PIL.Image -> tensor in [min_value, max_value] (diffsynth/utils/__init__.py:60-66)."""
        import numpy as np
        t = torch.Tensor(np.array(image, dtype=np.float32)).to(dtype=torch_dtype or self.torch_dtype, device=device or self.device)
        t = t * ((max_value - min_value) / 255) + min_value
        t = t.permute(2, 0, 1)
        return t.unsqueeze(0) if "B" in pattern else t

    def vae_output_to_image(self, vae_output, pattern="B C H W", min_value=-1, max_value=1):
        """tensor -> PIL.Image (diffsynth/utils/__init__.py:76-83; the batch axis is averaged away as there)."""
        import numpy as np  # noqa: F401
        from PIL import Image
        if pattern == "B C H W":
            vae_output = vae_output.mean(dim=0).permute(1, 2, 0)
        image = ((vae_output - min_value) * (255 / (max_value - min_value))).clip(0, 255)
        return Image.fromarray(image.to(device="cpu", dtype=torch.uint8).numpy())

    @staticmethod
    def calculate_dimensions(target_area, ratio):
        """QwenImageUnit_EditImageEmbedder.calculate_dimensions (qwen_image_physical.py:1249-1255): (width, height) on a 32-pixel grid."""
        import math
        width = math.sqrt(target_area * ratio)
        height = width / ratio
        return round(width / 32) * 32, round(height / 32) * 32

    def auto_resize_edit_image(self, edit_image):
        """:1258-1260: resize a PIL image to ~1024 x 1024 pixels keeping its aspect ratio."""
        w, h = QwenImagePhysicPipeline.calculate_dimensions(1024 * 1024, edit_image.size[0] / edit_image.size[1])
        return edit_image.resize((w, h))

    def generate_noise(self, shape, seed=None, rand_device="cpu", rand_torch_dtype=torch.float32, device=None, torch_dtype=None):
        generator = None if seed is None else torch.Generator(rand_device).manual_seed(seed)
        noise = torch.randn(shape, generator=generator, device=rand_device, dtype=rand_torch_dtype)
        return noise.to(dtype=torch_dtype or self.torch_dtype, device=device or self.device)

    def blend_with_mask(self, base, addition, mask):
        """utils/__init__.py:146-147."""
        return base * (1 - mask) + addition * mask

    def step(self, scheduler, latents, progress_id, noise_pred, input_latents=None, inpaint_mask=None, **kwargs):
        """BasePipeline.step (utils/__init__.py:150-156): one Euler update in torch arithmetic -- the form `direct_distill_loss` differentiates
        through and the reference's own loop body calls.  With an inpaint mask the prediction outside the mask is replaced by the one that leads
        back to `input_latents`; `denoise` comes through here for inpainting requests (its mask-free steps use the fused CFG + Euler kernel instead)."""
        timestep = scheduler.timesteps[progress_id]
        if inpaint_mask is not None:
            noise_pred = self.blend_with_mask(scheduler.return_to_timestep(timestep, latents, input_latents), noise_pred, inpaint_mask)
        return scheduler.step(noise_pred, timestep, latents)

    def enable_cpu_offload(self):
        """utils/__init__.py:127-129 (deprecated there in favour of enable_vram_management): accepted, weights stay resident (180 GB HBM)."""
        import warnings
        warnings.warn("`enable_cpu_offload` will be deprecated. Please use `enable_vram_management`.")

    def get_vram(self):
        """utils/__init__.py:132-133: total device memory in GiB."""
        return torch.cuda.mem_get_info(self.device)[1] / (1024 ** 3)

    # ---- the hot loop -------------------------------------------------------------------------------
    @torch.no_grad()
    def denoise(self, latents, inputs_posi: dict, inputs_nega: Optional[dict], edit_latents=None, context_latents=None, *, height: int,
                width: int, num_inference_steps: int = 30, cfg_scale: float = 4.0, denoising_strength: float = 1.0,
                exponential_shift_mu=None, progress_bar_cmd=None, timesteps_device: Optional[torch.Tensor] = None,
                model_kwargs: Optional[dict] = None, inpaint_mask: Optional[torch.Tensor] = None, input_latents: Optional[torch.Tensor] = None):
        """Lines 600 and 646-661 of the reference __call__: set_timesteps, then per step two model_fn forwards
        (posi / nega, each with its own persistently-mutated prompt_emb), CFG combine and the Euler update
        (one fused kernel).  inputs_*: dicts with prompt_emb [1,T,3584], prompt_emb_mask, special_token_mask.
        `model_kwargs`: further shared model_fn arguments (edit_rope_interpolation, blockwise_controlnet_conditioning / _inputs, ...)."""
        self.scheduler.set_timesteps(num_inference_steps, denoising_strength=denoising_strength,
                                     dynamic_shift_len=(height // 16) * (width // 16), exponential_shift_mu=exponential_shift_mu)
        nat = nv.Native.get(latents.device.index or 0)
        # CFG needs both branches (:653-658 always runs the negative forward when cfg_scale != 1): refuse a half-specified request
        # instead of combining with an uninitialised negative prediction
        use_cfg = self._use_cfg(cfg_scale, inputs_nega)
        inputs_posi, inputs_nega = self._with_lengths(inputs_posi), self._with_lengths(inputs_nega) if use_cfg else None
        ts = self.scheduler.timesteps
        ts_dev = ts.to(dtype=self.torch_dtype).to(latents.device) if timesteps_device is None else timesteps_device   # one H2D for the whole table
        latents = latents.clone()
        vbuf = torch.empty((2,) + tuple(latents.shape), dtype=latents.dtype, device=latents.device)     # [posi | nega], one buffer: the
        vp, vn = vbuf[0], vbuf[1]                                                                         # CFG-parallel all-gather runs in place
        # timestep-only quantities for the whole schedule, batch-8 GEMVs (see DiTEngine.precompute_conditioning)
        unmerged = getattr(self.dit, "_lora_injected", False)     # evaluation in the middle of training: model_fn takes the un-merged path, which
        if not unmerged:                                          # computes its own (LoRA-carrying) modulation -- the engine's table would go unused
            self.dit.engine().precompute_conditioning(ts_dev, [float(t.to(self.torch_dtype)) for t in ts])
        it = enumerate(ts)
        if progress_bar_cmd is not None:
            it = progress_bar_cmd(list(it))
        for progress_id, t in it:
            t_host = float(t.to(self.torch_dtype))
            kw = dict(dit=self.dit, visual_thinking_adapter=self.visual_thinking_adapter, latents=latents, timestep=ts_dev[progress_id:progress_id + 1],
                      height=height, width=width, edit_latents=edit_latents, context_latents=context_latents, is_train=False,
                      progress_id=progress_id, timestep_host=t_host, num_inference_steps=num_inference_steps,
                      blockwise_controlnet=getattr(self, "blockwise_controlnet", None))
            if model_kwargs:
                kw.update(model_kwargs)
            self.run_cfg_branches(kw, inputs_posi, inputs_nega, vp, vn, ts_dev[progress_id:progress_id + 1], t_host)
            if inpaint_mask is not None:
                # inpainting (:612, BasePipeline.step utils/__init__.py:150-156): outside the mask the prediction is replaced by the one that leads
                # back to `input_latents`.  Rare option, 128 KB of latents: the reference's own torch expressions (same rounding points) on top of
                # the two native forwards instead of the fused CFG + Euler kernel.
                noise_pred = vn + cfg_scale * (vp - vn) if use_cfg else vp
                latents = self.step(self.scheduler, latents=latents, progress_id=progress_id, noise_pred=noise_pred, input_latents=input_latents,
                                    inpaint_mask=inpaint_mask)
                continue
            ds = float(self.scheduler.dsigma(t))
            nat.cfg_euler_step(latents, vp, vn if use_cfg else None, float(cfg_scale), ds)
        # a kernel whose bounded pipeline wait timed out leaves its output partly written and the handle's flag set: surface it here,
        # at the loop's natural sync point, instead of handing garbage latents to the VAE / the caller
        nat.check_async()
        return latents

    @staticmethod
    def _use_cfg(cfg_scale, inputs_nega) -> bool:
        if cfg_scale != 1.0 and inputs_nega is None:
            raise ValueError(f"cfg_scale={cfg_scale} needs the negative-prompt inputs (the reference always runs the negative branch when "
                             "cfg_scale != 1, qwen_image_physical.py:653-658); pass inputs_nega or cfg_scale=1.0")
        return cfg_scale != 1.0

    @staticmethod
    def _with_lengths(inputs: Optional[dict]) -> Optional[dict]:
        """Adds txt_len / n_special (read from the masks once per request) so the per-step forwards need no device->host sync."""
        if inputs is None or "txt_len" in inputs:
            return inputs
        out = dict(inputs)
        out.update(prompt_lengths(inputs.get("prompt_emb_mask"), inputs.get("special_token_mask"), inputs["prompt_emb"].shape[1]))
        return out

    @torch.no_grad()
    def run_cfg_branches(self, kw: dict, inputs_posi: dict, inputs_nega: Optional[dict], vp, vn, t_dev, t_host) -> None:
        """The positive (and, with CFG, the negative) DiT forward of one denoise step (:653-655) into vp / vn.
        With `self.cfg_streams == 2` the two branches -- independent until the combine -- run on two streams: every hot kernel is a
        persistent grid of one CTA (pair) per SM whose last wave is partial (816 attention items or 408 out-projection tiles on
        148 SMs = 5.51 waves), and with a second stream the other branch's CTAs take the SMs a finishing kernel releases.
        Results are unchanged: same kernels, same inputs, separate workspaces (cfg_branch)."""
        grp = getattr(self, "cfg_parallel_group", None)
        if inputs_nega is not None and grp is not None:
            # Latency mode: ONE image on a pair of GPUs.  Rank 0 of the pair runs the positive branch, rank 1 the
            # negative one (each keeps mutating only its own prompt_emb, as the reference does per branch), then the two
            # [1,16,h8,w8] predictions (512 KiB at 1024^2) are exchanged with one all-gather over NVLink and both ranks apply the
            # same CFG + Euler update, so their latents stay bit-identical without any further traffic.
            from . import parallel
            r = parallel.dist.get_rank(grp)
            self.model_fn(**kw, **(inputs_posi if r == 0 else inputs_nega), out=vp if r == 0 else vn)
            parallel.exchange_cfg_predictions(vp, vn, r, grp)
            return
        if inputs_nega is not None and getattr(self, "cfg_streams", 1) == 2:
            dev = vp.device
            main = torch.cuda.current_stream(dev)
            if getattr(self, "_side_stream", None) is None or self._side_stream.device != dev:
                self._side_stream = torch.cuda.Stream(dev)
            side = self._side_stream
            if not getattr(self.dit, "_lora_injected", False):
                self.dit.engine().conditioning(t_dev, t_host)   # timestep-only tensors: computed (or found cached) before the fork
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.model_fn(**kw, **inputs_nega, out=vn, cfg_branch=1)
            self.model_fn(**kw, **inputs_posi, out=vp, cfg_branch=0)
            main.wait_stream(side)
        else:
            self.model_fn(**kw, **inputs_posi, out=vp)
            if inputs_nega is not None:
                self.model_fn(**kw, **inputs_nega, out=vn)

    @torch.no_grad()
    def denoise_step(self, latents, inputs_posi: dict, inputs_nega: Optional[dict], edit_latents=None, context_latents=None, *, progress_id: int,
                     height: int, width: int, cfg_scale: float = 4.0):
        """One iteration of the loop above against the CURRENT scheduler table (call scheduler.set_timesteps first):
        updates `latents` in place and returns it.  This is the unit bench.py times end to end."""
        nat = nv.Native.get(latents.device.index or 0)
        use_cfg = self._use_cfg(cfg_scale, inputs_nega)
        inputs_posi, inputs_nega = self._with_lengths(inputs_posi), self._with_lengths(inputs_nega) if use_cfg else None
        t = self.scheduler.timesteps[progress_id]
        t_host = float(t.to(self.torch_dtype))
        t_dev = t.to(self.torch_dtype).reshape(1).to(latents.device, non_blocking=True)
        if not hasattr(self, "_vbuf") or self._vbuf.shape[1:] != latents.shape or self._vbuf.device != latents.device:
            self._vbuf = torch.empty((2,) + tuple(latents.shape), dtype=latents.dtype, device=latents.device)
        vp, vn = self._vbuf[0], self._vbuf[1]
        kw = dict(dit=self.dit, visual_thinking_adapter=self.visual_thinking_adapter, latents=latents, timestep=t_dev, height=height, width=width,
                  edit_latents=edit_latents, context_latents=context_latents, is_train=False, progress_id=progress_id, timestep_host=t_host)
        self.run_cfg_branches(kw, inputs_posi, inputs_nega, vp, vn, t_dev, t_host)
        nat.cfg_euler_step(latents, vp, vn if use_cfg else None, float(cfg_scale), float(self.scheduler.dsigma(t)))
        return latents

    @torch.no_grad()
    def __call__(self, prompt=None, negative_prompt="", cfg_scale=4.0, input_image=None, denoising_strength=1.0, inpaint_mask=None,
                 inpaint_blur_size=None, inpaint_blur_sigma=None, height=1328, width=1328, seed=None, rand_device="cpu", num_inference_steps=30,
                 exponential_shift_mu=None, blockwise_controlnet_inputs=None, eligen_entity_prompts=None, eligen_entity_masks=None,
                 eligen_enable_on_negative=False, edit_image=None, edit_image_auto_resize=True, edit_rope_interpolation=False, context_image=None,
                 enable_fp8_attention=False, tiled=False, tile_size=128, tile_stride=64, progress_bar_cmd="tqdm", supported_rules=None,
                 contradicted_rules=None, middle_key_frames=None, stitched_image=None, state=None, transition=None, triplet=None, is_train=True,
                 have_text_reasoning=True,
                 prompt_inputs_posi: dict = None, prompt_inputs_nega: dict = None, edit_latents=None, context_latents=None, output_type="pil"):
        """:544-669, same keyword surface.  The request is carried in three dictionaries (shared / positive / negative) through
        `self.units` by `self.unit_runner`; then the CFG denoise loop (native: `denoise`) and the VAE decode.  Extras beyond the
        reference: `prompt_inputs_posi/nega`, `edit_latents`, `context_latents` hand over pre-computed unit outputs (a unit whose
        model is absent returns nothing, so they survive), `output_type="latent"` skips the decode."""
        if isinstance(progress_bar_cmd, str):                # the reference's default is tqdm itself (:586); resolved here so that importing the package does not need it
            from tqdm import tqdm
            progress_bar_cmd = tqdm
        inputs_posi = dict(prompt_inputs_posi or {}, prompt=prompt)
        inputs_nega = dict(prompt_inputs_nega or {}, negative_prompt=negative_prompt)
        inputs_shared = {
            "cfg_scale": cfg_scale, "input_image": input_image, "denoising_strength": denoising_strength, "inpaint_mask": inpaint_mask,
            "inpaint_blur_size": inpaint_blur_size, "inpaint_blur_sigma": inpaint_blur_sigma, "height": height, "width": width, "seed": seed,
            "rand_device": rand_device, "enable_fp8_attention": enable_fp8_attention, "num_inference_steps": num_inference_steps,
            "blockwise_controlnet_inputs": blockwise_controlnet_inputs, "tiled": tiled, "tile_size": tile_size, "tile_stride": tile_stride,
            "eligen_entity_prompts": eligen_entity_prompts, "eligen_entity_masks": eligen_entity_masks,
            "eligen_enable_on_negative": eligen_enable_on_negative, "edit_image": edit_image, "edit_image_auto_resize": edit_image_auto_resize,
            "edit_rope_interpolation": edit_rope_interpolation, "context_image": context_image, "supported_rules": supported_rules,
            "contradicted_rules": contradicted_rules, "middle_key_frames": middle_key_frames, "stitched_image": stitched_image, "state": state,
            "transition": transition, "triplet": triplet, "is_train": is_train,
        }
        # the scheduler table must exist before the units run (InputImageEmbedder adds noise at timesteps[0], :600)
        self.scheduler.set_timesteps(num_inference_steps, denoising_strength=denoising_strength, dynamic_shift_len=(height // 16) * (width // 16),
                                     exponential_shift_mu=exponential_shift_mu)
        units = [u for u in self.units if u is not None]
        if not is_train:
            units = [u for u in units if not isinstance(u, QwenImageUnit_PhysicalVisualEmbedder)]
        if not have_text_reasoning:
            units = [u for u in units if not isinstance(u, QwenImageUnit_PhysicalVerbalEmbedder)]
        if self.vae is None and ((edit_image is not None and edit_latents is None) or (context_image is not None and context_latents is None)
                                 or input_image is not None):
            raise RuntimeError("no VAE loaded: pass edit_latents / context_latents, or load the qwen_image_vae checkpoint (load_vae / from_pretrained)")
        if edit_latents is not None and self.text_encoder is None:   # pre-computed latents AND prompt embeddings: the image itself is not needed
            inputs_shared["edit_image"] = None
        for unit in units:
            if self.vae is None and unit.onload_model_names == ("vae",) and not (unit.input_params and "noise" in unit.input_params):
                continue                                  # no VAE loaded: image-embedding units have nothing to run on (latents must be handed in)
            if (isinstance(unit, QwenImageUnit_PhysicalVerbalEmbedder) and cfg_scale != 1 and hasattr(self.text_encoder, "generate_batch")
                    and getattr(self, "batch_cfg_generation", True) and inputs_shared.get("edit_image") is not None
                    and None in (supported_rules, contradicted_rules, middle_key_frames, input_image)):
                # both CFG branches' generations decoded together (same outputs as the runner's two calls, about half the time)
                out_p, out_n = unit.process_both_branches(self, inputs_posi.get("prompt"), inputs_nega.get("negative_prompt"), inputs_shared["edit_image"])
                inputs_posi.update(out_p)
                inputs_nega.update(out_n)
                continue
            inputs_shared, inputs_posi, inputs_nega = self.unit_runner(unit, self, inputs_shared, inputs_posi, inputs_nega)
        if edit_latents is not None:
            inputs_shared["edit_latents"] = edit_latents
        if context_latents is not None:
            inputs_shared["context_latents"] = context_latents
        if "prompt_emb" not in inputs_posi:
            raise RuntimeError("no prompt embedding: load the Qwen2.5-VL text encoder (from_pretrained / load_text_encoder) together with the "
                               "tokenizer and processor, or pass prompt_inputs_posi / prompt_inputs_nega (prompt_emb, prompt_emb_mask, special_token_mask)")
        model_kwargs = {"edit_rope_interpolation": bool(edit_rope_interpolation), "enable_fp8_attention": bool(enable_fp8_attention)}
        if inputs_shared.get("blockwise_controlnet_conditioning"):
            if self.blockwise_controlnet is None:
                raise RuntimeError("blockwise_controlnet_inputs were given but pipe.blockwise_controlnet is not loaded")
            model_kwargs.update(blockwise_controlnet_conditioning=inputs_shared["blockwise_controlnet_conditioning"],
                                blockwise_controlnet_inputs=inputs_shared["blockwise_controlnet_inputs"])
        height, width = inputs_shared["height"], inputs_shared["width"]
        keys = ("prompt_emb", "prompt_emb_mask", "special_token_mask")
        ent = ("entity_prompt_emb", "entity_prompt_emb_mask", "entity_masks")                 # EliGen (QwenImageUnit_EntityControl), per branch
        pick = lambda d: dict({k: d.get(k) for k in keys}, **{k: d[k] for k in ent if d.get(k) is not None})
        posi = pick(inputs_posi)
        nega = pick(inputs_nega) if (cfg_scale != 1.0 and "prompt_emb" in inputs_nega) else None
        latents = self.denoise(inputs_shared["latents"], posi, nega, inputs_shared.get("edit_latents"), inputs_shared.get("context_latents"),
                               height=height, width=width, num_inference_steps=num_inference_steps, cfg_scale=cfg_scale,
                               denoising_strength=denoising_strength, exponential_shift_mu=exponential_shift_mu, progress_bar_cmd=progress_bar_cmd,
                               model_kwargs=model_kwargs, inpaint_mask=inputs_shared.get("inpaint_mask"), input_latents=inputs_shared.get("input_latents"))
        if output_type == "latent" or self.vae is None:
            return latents
        image = self.vae.decode(latents, device=self.device, tiled=tiled, tile_size=tile_size, tile_stride=tile_stride)
        return self.vae_output_to_image(image) if output_type == "pil" else image

    # ---- training path (forward only) -----------------------------------------
    def physical_visual_embeddings(self, dino_middle: torch.Tensor, dino_source: torch.Tensor, vae_middle_latents: torch.Tensor,
                                   vae_source_latents: torch.Tensor):
        """QwenImageUnit_PhysicalVisualEmbedder.process (:1057-1118) from pre-processed tensors:
        dino_* [F,3,224,224] normalised pixels, vae_* [F,16,h8,w8] latents.  Returns the regression targets
        pseudo_special_emb_dino / pseudo_special_emb_vae [1,64,3584].  With the resampler stack trainable and grad mode on (the train script's
        `forward_preprocess`, scripts/train/train_physicedit.py:290-295) the targets carry a graph, as in the reference: the adapter loss
        trains the resamplers through them."""
        from . import autograd as ag
        stack = (self.dino_time_embed, self.dino_resampler, self.dino_resampler_adapter, self.vae_time_embed, self.vae_resampler, self.vae_resampler_adapter)
        if ag.needs_grad(*stack) and any(m.training for m in stack):
            return self._physical_visual_embeddings_autograd(dino_middle, dino_source, vae_middle_latents, vae_source_latents)
        with torch.no_grad():
            return self._physical_visual_embeddings_forward(dino_middle, dino_source, vae_middle_latents, vae_source_latents)

    def _physical_visual_embeddings_forward(self, dino_middle, dino_source, vae_middle_latents, vae_source_latents):
        """Inference-mode body of the unit: everything on the native kernels, no graph."""
        nat = nv.Native.get(dino_middle.device.index or 0)

        def dino_branch(px, with_time):
            hs = self.dinov2(px)                                         # [F,256,768]
            Fn, L, Hd = hs.shape
            hs = hs.reshape(Fn * L, Hd).contiguous()
            if with_time:                                                  # + time_emb[f] on every token of frame f
                te = self.dino_time_embed.weight[:Fn].repeat_interleave(L, dim=0).contiguous()
                nat.add_rows(hs, te, Fn * L, 1.0)
            return self.dino_resampler_adapter(self.dino_resampler(hs.unsqueeze(0)))

        def vae_branch(lat, with_time):
            Fn = lat.shape[0]
            L = (lat.shape[2] // 2) * (lat.shape[3] // 2)
            tok = torch.empty(Fn * L, 64, dtype=torch.bfloat16, device=lat.device)
            for f in range(Fn):
                nat.patchify(lat[f].contiguous(), tok[f * L:(f + 1) * L])
            if with_time:
                te = self.vae_time_embed.weight[:Fn].repeat_interleave(L, dim=0).contiguous()
                nat.add_rows(tok, te, Fn * L, 1.0)
            return self.vae_resampler_adapter(self.vae_resampler(tok.unsqueeze(0)))

        def delta(a, b):
            out = a.reshape(-1, a.shape[-1]).contiguous().clone()
            nat.add_rows(out, b.reshape(-1, b.shape[-1]).contiguous(), out.shape[0], -1.0)
            return out.view(a.shape)

        return {"pseudo_special_emb_dino": delta(dino_branch(dino_middle, True), dino_branch(dino_source, False)),
                "pseudo_special_emb_vae": delta(vae_branch(vae_middle_latents, True), vae_branch(vae_source_latents, False))}

    def _physical_visual_embeddings_autograd(self, dino_middle, dino_source, vae_middle_latents, vae_source_latents):
        """The same unit (:1057-1118) when the resampler stack is trainable (scripts/train/train_multigpu.sh:38): DINOv2 and the VAE stay frozen
        (no gradient into pixels), everything after them runs under autograd on the native GEMMs (physicedit_b200/autograd.py)."""
        nat = nv.Native.get(dino_middle.device.index or 0)

        def dino_tokens(px, with_time):
            with torch.no_grad():
                hs = self.dinov2(px)                                      # [F, 256, 768]
            if with_time:
                hs = hs + self.dino_time_embed.weight[:hs.shape[0]].unsqueeze(1)
            return hs.reshape(1, -1, hs.shape[-1])

        def vae_tokens(lat, with_time):
            Fn, L = lat.shape[0], (lat.shape[2] // 2) * (lat.shape[3] // 2)
            tok = torch.empty(Fn, L, 64, dtype=torch.bfloat16, device=lat.device)
            with torch.no_grad():
                for f in range(Fn):
                    nat.patchify(lat[f].contiguous(), tok[f])
            if with_time:
                tok = tok + self.vae_time_embed.weight[:Fn].unsqueeze(1)
            return tok.reshape(1, Fn * L, 64)
        dino = lambda px, wt: self.dino_resampler_adapter(self.dino_resampler(dino_tokens(px, wt)))
        vae = lambda lat, wt: self.vae_resampler_adapter(self.vae_resampler(vae_tokens(lat, wt)))
        return {"pseudo_special_emb_dino": dino(dino_middle, True) - dino(dino_source, False),
                "pseudo_special_emb_vae": vae(vae_middle_latents, True) - vae(vae_source_latents, False)}

    def training_loss(self, global_step=None, timestep_id=None, noise=None, **inputs):
        """:313-329, forward value.  Called as the train script does (`pipe.training_loss(global_step=, **models, **inputs)`,
        scripts/train/train_physicedit.py:309-310: the in-iteration models arrive inside **inputs) or with the inputs only.
        `timestep_id` / `noise` pin the two random draws (:314, :317) for parity tests; by default they are drawn as in the reference."""
        if timestep_id is None:
            timestep_id = torch.randint(0, self.scheduler.num_train_timesteps, (1,))
        timestep_id = torch.as_tensor(timestep_id).reshape(-1).cpu()          # the scheduler's tables live on the host (flow_match.py:34-69)
        timestep = self.scheduler.timesteps[timestep_id].to(dtype=self.torch_dtype, device=self.device)
        if noise is None:
            noise = torch.randn_like(inputs["input_latents"])
        inputs["latents"] = self.scheduler.add_noise(inputs["input_latents"], noise, timestep)
        target = self.scheduler.training_target(inputs["input_latents"], noise, timestep)
        models = {name: getattr(self, name) for name in self.in_iteration_models if name not in inputs}
        noise_pred, special_token_loss = self.model_fn(**models, **inputs, timestep=timestep)
        loss = torch.nn.functional.mse_loss(noise_pred.float(), target.float())
        self.special_token_loss = float(special_token_loss.detach().mean().item()) if torch.is_tensor(special_token_loss) else 0.0
        loss = loss * self.scheduler.training_weight(timestep)
        return loss + special_token_loss
