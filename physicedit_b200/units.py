"""Pre-loop pipeline units of QwenImagePhysicPipeline with the reference's unit contract.

Mirrors `PipelineUnit` / `PipelineUnitRunner` (DiffSynth-Studio/diffsynth/utils/__init__.py:223-279) and the eleven units the
pipeline registers (pipelines/qwen_image_physical.py:225-246, bodies :673-1299).  A unit declares which request fields it reads
(`input_params`, or per CFG branch `input_params_posi` / `input_params_nega`), `process()` returns the fields it adds, and the
runner merges them into the shared / positive / negative dictionaries -- scripts/train/train_physicedit.py drives exactly this
interface (`pipe.units`, `pipe.unit_runner`, :285-300), which is why the names, the flags and the field names are kept verbatim.

Everything numeric inside a unit goes to the pipeline's NATIVE modules (`pipe.vae`, `pipe.text_encoder`, `pipe.dinov2`, the
resamplers): the units themselves are host-side bookkeeping (image resizing, chat templates, masks, padding).
"""
from __future__ import annotations

import json
import math
from typing import List

import torch

SPECIAL_TOKEN_NUM = 64        # qwen_image_physical.py:28

# The instruction the VL model answers before every edit (qwen_image_physical.py:158-176).  It is DATA, not logic: the generated
# reasoning text -- and through it prompt_emb -- depends on every character, so it is kept byte-identical.
SYSTEM_PROMPT_SAMPLE = """
You are a physics-aware visual editing assistant.
You will receive an "Edit Instruction" and an "Edit Image".
Your task is to generate a detailed description of the edit operations required to transform the image according to the instruction, ensuring all changes strictly follow physical laws.

INPUTS:
- Edit Instruction: The desired modification.
- Edit Image: The visual starting point.

REQUIREMENTS:
1. Physical Plausibility: All operations must respect physics (like gravity, inertia, material properties, light transport, collision, etc.).
2. Mechanism of Change: Describe *how* the change occurs visually (e.g., "The vase tilts and falls due to gravity," not just "The vase is on the floor").
3. Material Consistency: Ensure materials behave correctly (liquids flow, solids rigid/deform, cloth wrinkles).

OUTPUT FORMAT:
Return STRICT JSON ONLY:
{
  "middle_transition_prompt": "A multi-clause paragraph describing the step-by-step physical operations and visual transition."
}
""".strip()

# chat templates of the Qwen-Image text conditioning (:763, :774, :805): prompt-only (34 template tokens dropped), edit (64 dropped)
TEMPLATE_T2I = ("<|im_start|>system\nDescribe the image by detailing the color, shape, size, texture, quantity, text, spatial relationships of the "
                "objects and background:<|im_end|>\n<|im_start|>user\n{}<|im_end|>\n<|im_start|>assistant\n")
TEMPLATE_EDIT = ("<|im_start|>system\nDescribe the key features of the input image (color, shape, size, texture, objects, background), then explain how "
                 "the user's text instruction should alter or modify the image. Generate a new image that meets the user's requirements while "
                 "maintaining consistency with the original input where appropriate.<|im_end|>\n<|im_start|>user\n{}<|im_end|>\n<|im_start|>assistant\n")
VISION_SLOT = "<|vision_start|><|image_pad|><|vision_end|>"
DROP_T2I, DROP_EDIT = 34, 64


class PipelineUnit:
    """utils/__init__.py:223-243."""

    def __init__(self, seperate_cfg: bool = False, take_over: bool = False, input_params: tuple = None, input_params_posi: dict = None,
                 input_params_nega: dict = None, onload_model_names: tuple = None):
        self.seperate_cfg = seperate_cfg            # (sic) the reference's spelling is part of the interface
        self.take_over = take_over
        self.input_params = input_params
        self.input_params_posi = input_params_posi
        self.input_params_nega = input_params_nega
        self.onload_model_names = onload_model_names

    def process(self, pipe, **kwargs) -> dict:
        raise NotImplementedError("`process` is not implemented.")


class PipelineUnitRunner:
    """utils/__init__.py:247-279: three calling conventions -- take-over units see all three dictionaries, CFG-separated units run
    once per branch (the negative branch only when cfg_scale != 1, otherwise it inherits the positive outputs), plain units read
    and write the shared dictionary."""

    def __call__(self, unit: PipelineUnit, pipe, inputs_shared: dict, inputs_posi: dict, inputs_nega: dict):
        if unit.take_over:
            return unit.process(pipe, inputs_shared=inputs_shared, inputs_posi=inputs_posi, inputs_nega=inputs_nega)
        common = {name: inputs_shared.get(name) for name in (unit.input_params or ())}
        if not unit.seperate_cfg:
            inputs_shared.update(unit.process(pipe, **common))
            return inputs_shared, inputs_posi, inputs_nega
        out = unit.process(pipe, **{name: inputs_posi.get(src) for name, src in unit.input_params_posi.items()}, **common)
        inputs_posi.update(out)
        if inputs_shared["cfg_scale"] != 1:
            out = unit.process(pipe, **{name: inputs_nega.get(src) for name, src in unit.input_params_nega.items()}, **common)
        inputs_nega.update(out)
        return inputs_shared, inputs_posi, inputs_nega


# ---------------------------------------------------------------------------------------------------------------------------
def area_preserving_size(image, target_area: int):
    """(width, height) with width*height ~ target_area at the image's aspect ratio, on a 32-pixel grid (:751-757 and four more copies)."""
    ratio = image.size[0] / image.size[1]
    w = math.sqrt(target_area * ratio)
    return round(w / 32) * 32, round((w / ratio) / 32) * 32


def resize_to_area(image, target_area: int):
    return image.resize(area_preserving_size(image, target_area))


def valid_rows(hidden_states: torch.Tensor, mask: torch.Tensor, drop: int) -> List[torch.Tensor]:
    """extract_masked_hidden (:743-749) + the template-token drop: per batch row, the hidden states under the attention mask minus
    the first `drop` (system-prompt) positions."""
    m = mask.bool()[:, :hidden_states.shape[1]]
    return [h[keep][drop:] for h, keep in zip(hidden_states, m)]


def pad_stack(rows: List[torch.Tensor], pipe):
    """(:823-827) zero-pad to the longest row; the mask is int64 ones over the valid part."""
    n = max(r.shape[0] for r in rows)
    emb = torch.stack([torch.cat([r, r.new_zeros(n - r.shape[0], r.shape[1])]) for r in rows])
    mask = torch.stack([torch.cat([torch.ones(r.shape[0], dtype=torch.long, device=r.device),
                                   torch.zeros(n - r.shape[0], dtype=torch.long, device=r.device)]) for r in rows])
    return emb.to(dtype=pipe.torch_dtype, device=pipe.device), mask


# ---------------------------------------------------------------------------------------------------------------------------
class QwenImageUnit_ShapeChecker(PipelineUnit):
    def __init__(self):
        super().__init__(input_params=("height", "width"))

    def process(self, pipe, height, width):
        height, width = pipe.check_resize_height_width(height, width)
        return {"height": height, "width": width}


class QwenImageUnit_NoiseInitializer(PipelineUnit):
    """(:683-689) NOTE the draw happens in the PIPE dtype on the CPU generator (bf16 randn), not fp32-then-cast: a different stream."""

    def __init__(self):
        super().__init__(input_params=("height", "width", "seed", "rand_device"))

    def process(self, pipe, height, width, seed, rand_device):
        return {"noise": pipe.generate_noise((1, 16, height // 8, width // 8), seed=seed, rand_device=rand_device, rand_torch_dtype=pipe.torch_dtype)}


class QwenImageUnit_InputImageEmbedder(PipelineUnit):
    """(:693-710) image-to-image start: latents = add_noise(vae.encode(input_image)); in training mode the clean latents are the target."""

    def __init__(self):
        super().__init__(input_params=("input_image", "noise", "tiled", "tile_size", "tile_stride"), onload_model_names=("vae",))

    def process(self, pipe, input_image, noise, tiled, tile_size, tile_stride):
        if input_image is None:
            return {"latents": noise, "input_latents": None}
        pipe.load_models_to_device(["vae"])
        x0 = pipe.vae.encode(pipe.preprocess_image(input_image).to(device=pipe.device, dtype=pipe.torch_dtype), tiled=tiled, tile_size=tile_size,
                             tile_stride=tile_stride)
        if pipe.scheduler.training:
            return {"latents": noise, "input_latents": x0}
        return {"latents": pipe.scheduler.add_noise(x0, noise, timestep=pipe.scheduler.timesteps[0]), "input_latents": x0}


class QwenImageUnit_Inpaint(PipelineUnit):
    """(:714-730) the inpaint mask at latent resolution, optionally Gaussian-blurred."""

    def __init__(self):
        super().__init__(input_params=("inpaint_mask", "height", "width", "inpaint_blur_size", "inpaint_blur_sigma"))

    def process(self, pipe, inpaint_mask, height, width, inpaint_blur_size, inpaint_blur_sigma):
        if inpaint_mask is None:
            return {}
        m = pipe.preprocess_image(inpaint_mask.convert("RGB").resize((width // 8, height // 8)), min_value=0, max_value=1).mean(dim=1, keepdim=True)
        if inpaint_blur_size is not None and inpaint_blur_sigma is not None:
            from torchvision.transforms import GaussianBlur
            m = GaussianBlur(kernel_size=inpaint_blur_size * 2 + 1, sigma=inpaint_blur_sigma)(m)
        return {"inpaint_mask": m}


class QwenImageUnit_EditImageEmbedder(PipelineUnit):
    """(:1244-1285) edit image(s) -> VAE latents; each image first resized to ~1024^2 pixels unless edit_image_auto_resize=False."""

    def __init__(self):
        super().__init__(input_params=("edit_image", "tiled", "tile_size", "tile_stride", "edit_image_auto_resize"), onload_model_names=("vae",))

    def calculate_dimensions(self, target_area, ratio):
        w = math.sqrt(target_area * ratio)
        return round(w / 32) * 32, round((w / ratio) / 32) * 32

    def edit_image_auto_resize(self, edit_image):
        return resize_to_area(edit_image, 1024 * 1024)

    def process(self, pipe, edit_image, tiled, tile_size, tile_stride, edit_image_auto_resize=False):
        if edit_image is None:
            return {}
        pipe.load_models_to_device(["vae"])

        def one(image):
            image = self.edit_image_auto_resize(image) if edit_image_auto_resize else image
            lat = pipe.vae.encode(pipe.preprocess_image(image).to(device=pipe.device, dtype=pipe.torch_dtype), tiled=tiled, tile_size=tile_size,
                                  tile_stride=tile_stride)
            return image, lat
        if not isinstance(edit_image, (list, tuple)):
            image, lat = one(edit_image)
            return {"edit_latents": lat, "edit_image": image}
        pairs = [one(im) for im in edit_image]
        return {"edit_latents": [p[1] for p in pairs], "edit_image": [p[0] for p in pairs]}


class QwenImageUnit_ContextImageEmbedder(PipelineUnit):
    """(:1288-1299)"""

    def __init__(self):
        super().__init__(input_params=("context_image", "height", "width", "tiled", "tile_size", "tile_stride"), onload_model_names=("vae",))

    def process(self, pipe, context_image, height, width, tiled, tile_size, tile_stride):
        if context_image is None:
            return {}
        pipe.load_models_to_device(["vae"])
        x = pipe.preprocess_image(context_image.resize((width, height))).to(device=pipe.device, dtype=pipe.torch_dtype)
        return {"context_latents": pipe.vae.encode(x, tiled=tiled, tile_size=tile_size, tile_stride=tile_stride)}


class QwenImageUnit_PhysicalVerbalEmbedder(PipelineUnit):
    """(:837-988) the "physical thinking" text: with training annotations (`triplet`) it is assembled from them; at inference the VL
    model GENERATES it (<= 1000 new tokens, once per CFG branch) from the edit instruction and a ~384^2 copy of the edit image."""

    ACCEPTED = (("Reasoning",), ("physical_reasoning", "middle_transition_prompt", "final_state_prompt"), ("middle_transition_prompt",))

    def __init__(self):
        super().__init__(seperate_cfg=True, input_params_posi={"prompt": "prompt"}, input_params_nega={"prompt": "negative_prompt"},
                         input_params=("edit_image", "supported_rules", "contradicted_rules", "middle_key_frames", "input_image", "triplet"),
                         onload_model_names=("text_encoder", "tokenizer"))

    def resize_image(self, image, target_area=384 * 384):
        return resize_to_area(image, target_area)

    def _parse_generation_response(self, response: str) -> dict:
        """(:874-905) the outermost {...} of the answer must be JSON whose string fields form exactly one accepted field set."""
        lo, hi = response.find("{"), response.rfind("}")
        if lo == -1 or hi <= lo:
            raise ValueError(f"Cannot find JSON in response: {response}")
        try:
            data = json.loads(response[lo:hi + 1])
        except json.JSONDecodeError as exc:
            raise ValueError(f"Cannot parse JSON: {response[lo:hi + 1]}") from exc
        fields = {}
        for key in {k for group in self.ACCEPTED for k in group}:
            val = data.get(key)
            if val is None:
                continue
            if not isinstance(val, str):
                raise ValueError(f"Field {key} must be string, got {type(val)}: {data}")
            fields[key] = val.strip()
        if not any(set(fields) == set(group) for group in self.ACCEPTED):
            raise ValueError(f"Unsupported response format. Expected one of {self.ACCEPTED}, got keys {sorted(fields)}: {data}")
        return fields

    def _answer(self, pipe, model_inputs, out) -> str:
        """(:861-872) the prompt tokens trimmed off, decoded; a parsable JSON answer is flattened to "\\nkey: value" lines, anything
        else is passed through verbatim."""
        new_tokens = [o[len(i):] for i, o in zip(model_inputs.input_ids, out)]
        text = pipe.tokenizer.batch_decode(new_tokens, skip_special_tokens=True, clean_up_tokenization_spaces=False)[0]
        try:
            fields = self._parse_generation_response(text)
        except ValueError:
            return text
        return "".join(f"\n{k}: {v}" for k, v in fields.items())

    def generate_text(self, pipe, model_inputs) -> str:
        """(:859-860) greedy generation, at most 1000 new tokens."""
        return self._answer(pipe, model_inputs, pipe.text_encoder.generate(**model_inputs, max_new_tokens=1000))

    def sample_inputs(self, pipe, edit_image, prompt):
        """(:943-963) the chat request the VL model answers: system prompt, the edit instruction, a ~384^2 copy of the edit image."""
        user = [{"type": "input_text", "text": "Edit Instruction:"}, {"type": "input_text", "text": prompt},
                {"type": "input_text", "text": "Edit Image:"}, {"type": "image"}]
        chat = pipe.processor.apply_chat_template([{"role": "system", "content": SYSTEM_PROMPT_SAMPLE}, {"role": "user", "content": user}],
                                                  tokenize=False, add_generation_prompt=True, add_vision_id=True)
        return pipe.processor(text=[chat], images=self.resize_image(edit_image), padding=True, return_tensors="pt").to(pipe.device)

    def encode_physical_prompt_sample(self, pipe, edit_image, prompt) -> str:
        """(:943-967)"""
        return self.generate_text(pipe, self.sample_inputs(pipe, edit_image, prompt))

    def process_both_branches(self, pipe, prompt, negative_prompt, edit_image):
        """The positive and the negative branch's `process` (the runner calls it once per branch, utils/__init__.py:256-270) in ONE batched
        generation when the text encoder offers `generate_batch`: B=1 decode streams every weight per token, so two requests decoded
        together cost about one.  Same requests, same greedy tokens, same parsing -- returns (outputs_posi, outputs_nega)."""
        reqs = [self.sample_inputs(pipe, edit_image, p) for p in (prompt, negative_prompt)]
        keys = ("input_ids", "attention_mask", "pixel_values", "image_grid_thw")
        outs = pipe.text_encoder.generate_batch([{k: r[k] for k in keys if k in r} for r in reqs], max_new_tokens=1000)
        return tuple({"physical_txt": self._answer(pipe, r, o)} for r, o in zip(reqs, outs))

    def process(self, pipe, prompt, edit_image=None, supported_rules=None, contradicted_rules=None, middle_key_frames=None, input_image=None,
                triplet=None) -> dict:
        if pipe.text_encoder is None:
            return {}
        if supported_rules is not None and contradicted_rules is not None and middle_key_frames is not None and input_image is not None:
            # training samples carry the annotation (:976-983; the VL rewrite of it is commented out in the reference)
            return {"physical_txt": f"Middle Transition Prompt: {triplet.get('middle_transition_prompt', '')}\n"
                                    f"Final State Prompt: {triplet.get('final_state_prompt', '')}"}
        return {"physical_txt": self.encode_physical_prompt_sample(pipe, edit_image, prompt)}


class QwenImageUnit_PromptEmbedder(PipelineUnit):
    """(:732-835) prompt (+ physical thinking text) -> prompt_emb / prompt_emb_mask / special_token_mask.  With one edit image the
    prompt is followed by `<begin_of_img><img0>...<img63><end_of_img>`, whose 64 positions the adapter rewrites every step."""

    def __init__(self):
        super().__init__(seperate_cfg=True, input_params_posi={"prompt": "prompt", "physical_txt": "physical_txt"},
                         input_params_nega={"prompt": "negative_prompt"}, input_params=("edit_image",), onload_model_names=("text_encoder",))

    def resize_image(self, image, target_area=384 * 384):
        return resize_to_area(image, target_area)

    def _hidden(self, pipe, model_inputs, drop):
        kw = {k: model_inputs[k] for k in ("pixel_values", "image_grid_thw") if k in model_inputs}
        hs = pipe.text_encoder.edit_forward(input_ids=model_inputs.input_ids, attention_mask=model_inputs.attention_mask, output_hidden_states=True,
                                            **kw)[-1]
        return valid_rows(hs, model_inputs.attention_mask, drop)

    def encode_prompt(self, pipe, prompt: List[str]):
        txt = [TEMPLATE_T2I.format(p) for p in prompt]
        mi = pipe.tokenizer(txt, max_length=4096 + DROP_T2I, padding=True, truncation=True, return_tensors="pt").to(pipe.device)
        if mi.input_ids.shape[1] >= 1024:
            print(f"Warning!!! QwenImage model was trained on prompts up to 512 tokens. Current prompt requires {mi['input_ids'].shape[1] - DROP_T2I} "
                  "tokens, which may lead to unpredictable behavior.")
        return self._hidden(pipe, mi, DROP_T2I)

    def encode_prompt_edit(self, pipe, prompt: List[str], edit_image):
        tail = "\n<begin_of_img>" + "".join(f"<img{i}>" for i in range(SPECIAL_TOKEN_NUM)) + "<end_of_img><|im_end|>"
        txt = [TEMPLATE_EDIT.format(VISION_SLOT + p + tail) for p in prompt]
        mi = pipe.processor(text=txt, images=self.resize_image(edit_image), padding=True, return_tensors="pt").to(pipe.device)
        boi = torch.where(mi.input_ids == pipe.boi_token_id)[1]
        eoi = torch.where(mi.input_ids == pipe.eoi_token_id)[1]
        special = torch.zeros_like(mi.attention_mask, dtype=torch.bool)
        special[:, boi + 1:eoi] = True
        return self._hidden(pipe, mi, DROP_EDIT), special[:, DROP_EDIT:]

    def encode_prompt_edit_multi(self, pipe, prompt: List[str], edit_image):
        slots = "".join(f"Picture {i + 1}: {VISION_SLOT}" for i in range(len(edit_image)))
        txt = [TEMPLATE_EDIT.format(slots + p) for p in prompt]
        mi = pipe.processor(text=txt, images=[self.resize_image(im) for im in edit_image], padding=True, return_tensors="pt").to(pipe.device)
        return self._hidden(pipe, mi, DROP_EDIT)

    def process(self, pipe, prompt, edit_image=None, physical_txt=None, pseudo_special_emb=None) -> dict:
        if physical_txt is not None:
            prompt = prompt + physical_txt
        if pipe.text_encoder is None:
            return {}
        special = None
        if edit_image is None:
            rows = self.encode_prompt(pipe, [prompt])
        elif isinstance(edit_image, (list, tuple)):
            rows = self.encode_prompt_edit_multi(pipe, [prompt], edit_image)
        else:
            rows, special = self.encode_prompt_edit(pipe, [prompt], edit_image)
        emb, mask = pad_stack(rows, pipe)
        return {"prompt_emb": emb, "prompt_emb_mask": mask, "special_token_mask": special}


class QwenImageUnit_PhysicalVisualEmbedder(PipelineUnit):
    """(:991-1118) TRAINING only (dropped by `__call__` when is_train=False): the regression targets of the adapter's two heads,
    pseudo_special_emb_{dino,vae} = features(middle key frames, + frame-index embedding) - features(source image), each through its
    perceiver resampler (64 latents) and resampler adapter.  Runs on the native DINOv2 / resamplers / VAE."""

    def __init__(self):
        super().__init__(input_params=("middle_key_frames", "edit_image", "tiled", "tile_size", "tile_stride"),
                         onload_model_names=("text_encoder", "dino_resampler", "dino_resampler_adapter", "vae_resampler", "vae_resampler_adapter"))

    def dino_input_preprocess(self, pipe, frames, size):
        """(:1042-1054) Resize(1.5 x size, bicubic) -> RandomCrop(size) (global torch RNG, as in the reference) -> ImageNet normalisation."""
        from torchvision import transforms
        tf = transforms.Compose([transforms.Resize(int(size * 1.5), interpolation=transforms.InterpolationMode.BICUBIC),
                                 transforms.RandomCrop(size), transforms.ToTensor()])
        x = torch.stack([tf(im) for im in frames]).to(pipe.device)
        return ((x - pipe.dinov2_mean.to(pipe.device)) / pipe.dinov2_std.to(pipe.device)).to(pipe.torch_dtype)

    def process(self, pipe, middle_key_frames=None, edit_image=None, tiled=False, tile_size=None, tile_stride=None) -> dict:
        if pipe.text_encoder is None:
            return {}
        enc = lambda im: pipe.vae.encode(pipe.preprocess_image(im).to(pipe.device, pipe.torch_dtype), tiled=tiled, tile_size=tile_size, tile_stride=tile_stride)
        return pipe.physical_visual_embeddings(
            dino_middle=self.dino_input_preprocess(pipe, middle_key_frames, pipe.dino_input_size),
            dino_source=self.dino_input_preprocess(pipe, [edit_image], pipe.dino_input_size),
            vae_middle_latents=torch.cat([enc(f) for f in middle_key_frames]), vae_source_latents=enc(edit_image))


class QwenImageUnit_EntityControl(PipelineUnit):
    """(:1121-1197) EliGen entity prompts / masks -> per-entity prompt embeddings and latent-resolution masks (take-over unit)."""

    def __init__(self):
        super().__init__(take_over=True, onload_model_names=("text_encoder",))

    def get_prompt_emb(self, pipe, prompt) -> dict:
        if pipe.text_encoder is None:
            return {}
        mi = pipe.tokenizer([TEMPLATE_T2I.format(prompt)], max_length=1024 + DROP_T2I, padding=True, truncation=True, return_tensors="pt").to(pipe.device)
        hs = pipe.text_encoder.edit_forward(input_ids=mi.input_ids, attention_mask=mi.attention_mask, output_hidden_states=True)[-1]
        emb, mask = pad_stack(valid_rows(hs, mi.attention_mask, DROP_T2I), pipe)
        return {"prompt_emb": emb, "prompt_emb_mask": mask}

    def preprocess_masks(self, pipe, masks, height, width, dim):
        from PIL import Image
        return [(pipe.preprocess_image(m.resize((width, height), resample=Image.NEAREST)).mean(dim=1, keepdim=True) > 0)
                .repeat(1, dim, 1, 1).to(device=pipe.device, dtype=pipe.torch_dtype) for m in masks]

    def process(self, pipe, inputs_shared, inputs_posi, inputs_nega):
        prompts, masks = inputs_shared.get("eligen_entity_prompts"), inputs_shared.get("eligen_entity_masks")
        if not prompts or not masks:
            return inputs_shared, inputs_posi, inputs_nega
        pipe.load_models_to_device(self.onload_model_names)
        h, w = inputs_shared["height"], inputs_shared["width"]
        ent_masks = torch.cat(self.preprocess_masks(pipe, masks, h // 8, w // 8, 1), dim=0).unsqueeze(0)          # [1, n_entity, 1, h/8, w/8]
        embs = [self.get_prompt_emb(pipe, p) for p in prompts]
        inputs_posi.update({"entity_prompt_emb": [e["prompt_emb"] for e in embs], "entity_masks": ent_masks,
                            "entity_prompt_emb_mask": [e["prompt_emb_mask"] for e in embs]})
        if inputs_shared.get("cfg_scale", 1.0) != 1.0:
            if inputs_shared.get("eligen_enable_on_negative", False):
                inputs_nega.update({"entity_prompt_emb": [inputs_nega["prompt_emb"]] * len(embs), "entity_masks": ent_masks,
                                    "entity_prompt_emb_mask": [inputs_nega["prompt_emb_mask"]] * len(embs)})
            else:
                inputs_nega.update({"entity_prompt_emb": None, "entity_masks": None, "entity_prompt_emb_mask": None})
        return inputs_shared, inputs_posi, inputs_nega


class QwenImageUnit_BlockwiseControlNet(PipelineUnit):
    """(:1201-1241) control images -> VAE latents (+ an inverted inpaint-mask channel) for the blockwise controlnets."""

    def __init__(self):
        super().__init__(input_params=("blockwise_controlnet_inputs", "tiled", "tile_size", "tile_stride"), onload_model_names=("vae",))

    def process(self, pipe, blockwise_controlnet_inputs, tiled, tile_size, tile_stride):
        if blockwise_controlnet_inputs is None:
            return {}
        import numpy as np
        from PIL import Image
        pipe.load_models_to_device(self.onload_model_names)
        out = []
        for ci in blockwise_controlnet_inputs:
            image = ci.image
            if ci.inpaint_mask is not None:                       # black out the masked pixels before encoding
                hole = pipe.preprocess_image(ci.inpaint_mask.resize(image.size)).mean(dim=[0, 1]).cpu()
                arr = np.array(image)
                arr[hole > 0] = 0
                image = Image.fromarray(arr)
            lat = pipe.vae.encode(pipe.preprocess_image(image).to(device=pipe.device, dtype=pipe.torch_dtype), tiled=tiled, tile_size=tile_size,
                                  tile_stride=tile_stride)
            if ci.inpaint_mask is not None:
                m = ((pipe.preprocess_image(ci.inpaint_mask) + 1) / 2).mean(dim=1, keepdim=True)
                lat = torch.concat([lat, 1 - torch.nn.functional.interpolate(m, size=lat.shape[-2:])], dim=1)
            out.append(lat)
        return {"blockwise_controlnet_conditioning": out}


def default_units() -> List[PipelineUnit]:
    """The unit list of QwenImagePhysicPipeline.__init__ (:233-245), same order."""
    return [QwenImageUnit_ShapeChecker(), QwenImageUnit_NoiseInitializer(), QwenImageUnit_InputImageEmbedder(), QwenImageUnit_Inpaint(),
            QwenImageUnit_EditImageEmbedder(), QwenImageUnit_ContextImageEmbedder(), QwenImageUnit_PhysicalVisualEmbedder(),
            QwenImageUnit_PhysicalVerbalEmbedder(), QwenImageUnit_PromptEmbedder(), QwenImageUnit_EntityControl(),
            QwenImageUnit_BlockwiseControlNet()]
