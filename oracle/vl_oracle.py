"""ORACLE (test infrastructure, not product code) for the Qwen2.5-VL text-encoder path.

The arithmetic of this path is not in the reference tree: the wrapper `QwenImageTextEncoderWithDecode`
(DiffSynth-Studio/diffsynth/models/qwen_image_text_encoder_withdecode.py:6-275) subclasses transformers'
`Qwen2_5_VLForConditionalGeneration` and `edit_forward` (:188-275) is `self.model(..., output_hidden_states=True).hidden_states`;
`generate` is GenerationMixin's greedy search with `GenerationConfig.from_model_config(config.text_config)` (:144).  transformers is
an UN-PINNED dependency (requirements.txt:3; the wrapper's config string says 4.54.0; this image has 5.5.0) -> "parity unpinned" for
the dependency version, see DESIGN.md.  The oracle is therefore the installed library itself, driven exactly as the wrapper drives it,
on a SMALL seeded configuration with the same structure (GQA 2:1, head dim 128, mrope sections 16/24/24, vision head dim 80,
windowed + full-attention vision blocks, an MLP width that is not a multiple of 8 like the real 3420).
Only tests/ use this module; tests/golden/vl.pt (oracle/make_golden_vl.py) pins its outputs in this image.
"""
from __future__ import annotations

import math

import torch

IMAGE_TOKEN, VISION_START, VISION_END, EOS = 555, 553, 554, 2
TEXT = dict(hidden_size=512, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2, vocab_size=640,
            rms_norm_eps=1e-6, max_position_embeddings=4096, hidden_act="silu", tie_word_embeddings=False,
            rope_parameters={"rope_type": "default", "mrope_section": [16, 24, 24], "rope_theta": 1000000.0},
            bos_token_id=1, eos_token_id=EOS, pad_token_id=EOS)
VISION = dict(depth=3, hidden_size=160, num_heads=2, intermediate_size=220, out_hidden_size=512, patch_size=14, spatial_merge_size=2,
              temporal_patch_size=2, window_size=112, fullatt_block_indexes=[1], in_channels=3, tokens_per_second=2, hidden_act="silu")


def native_config():
    """The same sizes as a physicedit_b200.text_encoder.VLConfig."""
    from physicedit_b200.text_encoder import VLConfig
    return VLConfig(hidden=512, layers=3, heads=4, kv_heads=2, head_dim=128, intermediate=1024, vocab=640, v_hidden=160, v_depth=3, v_heads=2,
                    v_intermediate=220, v_out=512, fullatt=(1,), image_token_id=IMAGE_TOKEN, eos_token_id=EOS)


def hf_model(dtype=torch.float32, device="cpu"):
    from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration
    cfg = Qwen2_5_VLConfig(text_config=dict(TEXT), vision_config=dict(VISION), image_token_id=IMAGE_TOKEN, video_token_id=556,
                           vision_start_token_id=VISION_START, vision_end_token_id=VISION_END)
    cfg._attn_implementation = "sdpa"
    with torch.device("meta"):
        m = Qwen2_5_VLForConditionalGeneration(cfg)
    m = m.to_empty(device="cpu")
    sd = synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=7)
    m.load_state_dict(sd, strict=True)
    for mod in m.modules():                                    # non-persistent rotary buffers were created on meta: rebuild them
        if hasattr(mod, "inv_freq") and hasattr(mod, "original_inv_freq"):
            inv, _ = mod.compute_default_rope_parameters(mod.config, "cpu")
            mod.inv_freq, mod.original_inv_freq = inv, inv.clone()
        elif hasattr(mod, "inv_freq") and hasattr(mod, "theta"):
            mod.inv_freq = 1.0 / (mod.theta ** (torch.arange(0, mod.dim, 2, dtype=torch.float) / mod.dim))
    m.generation_config.eos_token_id = EOS
    m.generation_config.pad_token_id = EOS
    return m.to(device=device, dtype=dtype).eval()


def synth_state_dict(shapes, seed):
    """Seeded weights (bf16-representable): matrices ~ U(-b, b) with b = 1.6 / sqrt(fan_in) (keeps the residual stream O(1) at this
    depth), biases small, norm weights 1 + 0.1 N(0, 1), embeddings N(0, 1)."""
    out = {}
    for n, key in enumerate(sorted(shapes)):
        shp = shapes[key]
        g = torch.Generator().manual_seed(seed * 7919 + n)
        if "embed_tokens" in key:
            t = torch.randn(shp, generator=g)
        elif len(shp) >= 2:
            fan_in = math.prod(shp[1:])
            t = (torch.rand(shp, generator=g) * 2 - 1) * (1.6 / math.sqrt(fan_in))
        elif key.endswith(".bias"):
            t = (torch.rand(shp, generator=g) * 2 - 1) * 0.05
        else:
            t = 1 + 0.1 * torch.randn(shp, generator=g)
        out[key] = t.to(torch.bfloat16).to(torch.float32)
    return out


def inputs(with_image=True, seed=3, n_text=23, grid=(1, 12, 16)):
    """A processor-shaped request: [text..., <vision_start>, <image_pad> x (h*w/4), <vision_end>, text...] + flattened patches."""
    g = torch.Generator().manual_seed(seed)
    words = lambda n: torch.randint(3, 500, (n,), generator=g)
    if not with_image:
        ids = words(n_text + 9)
        return dict(input_ids=ids.view(1, -1), attention_mask=torch.ones(1, ids.numel(), dtype=torch.long))
    t, h, w = grid
    n_img = t * h * w // 4
    ids = torch.cat([words(7), torch.tensor([VISION_START]), torch.full((n_img,), IMAGE_TOKEN), torch.tensor([VISION_END]), words(n_text)])
    px = torch.randn(t * h * w, 3 * 2 * 14 * 14, generator=g).to(torch.bfloat16).to(torch.float32)
    return dict(input_ids=ids.view(1, -1), attention_mask=torch.ones(1, ids.numel(), dtype=torch.long), pixel_values=px,
                image_grid_thw=torch.tensor([list(grid)]))


def _mm(inp):
    return (inp["input_ids"] == IMAGE_TOKEN).to(torch.int32)


@torch.no_grad()
def edit_forward(model, inp, mrope=True):
    """The wrapper's edit_forward (:188-275): final entry of `self.model(...).hidden_states`.  `mrope=True` hands over the 5.x
    `mm_token_type_ids` (3-D positions, transformers 5.5 flavour); False reproduces what the reference's call does under 5.5
    (no token types -> sequential positions)."""
    dev = next(model.parameters()).device
    kw = {k: v.to(dev) for k, v in inp.items()}
    if "pixel_values" in kw:
        kw["pixel_values"] = kw["pixel_values"].to(next(model.parameters()).dtype)
        if mrope:
            kw["mm_token_type_ids"] = _mm(inp).to(dev)
    model.model.rope_deltas = None
    out = model.model(**kw, output_hidden_states=True, return_dict=True, use_cache=False)
    return out.hidden_states[-1], out.last_hidden_state


@torch.no_grad()
def generate(model, inp, max_new_tokens):
    """pipe.text_encoder.generate(**model_inputs, max_new_tokens=...) (qwen_image_physical.py:860): greedy; returns the new ids and
    the per-step fp32 logits of the top two candidates (to tell a tie from a bug when two implementations part ways)."""
    dev = next(model.parameters()).device
    kw = {k: v.to(dev) for k, v in inp.items()}
    if "pixel_values" in kw:
        kw["pixel_values"] = kw["pixel_values"].to(next(model.parameters()).dtype)
        kw["mm_token_type_ids"] = _mm(inp).to(dev)
    model.model.rope_deltas = None
    out = model.generate(**kw, max_new_tokens=max_new_tokens, do_sample=False, output_scores=True, return_dict_in_generate=True)
    new = out.sequences[0, inp["input_ids"].shape[1]:]
    top2 = torch.stack([s[0].float().topk(2).values for s in out.scores])
    return new.cpu(), top2.cpu()
