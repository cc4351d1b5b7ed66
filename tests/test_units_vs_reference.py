"""Pre-loop units: this package's `physicedit_b200/units.py` against the REFERENCE's unit classes, unit by unit.

Both sets of eleven units are run by their own `PipelineUnitRunner` over copies of the same request, on the same pipeline object (this
package's, on the CPU) whose text encoder / VAE are recording stubs: what is compared is everything a unit contributes -- the fields it
returns (tensors bit for bit, images pixel for pixel, strings) and what it asked of the text encoder (token ids, attention masks, pixel
values, grids, `max_new_tokens`).  That pins the chat templates, the 34 / 64 dropped template tokens, the special-token mask, the 384^2 /
1024^2 resize rules, the parsing of the generated JSON, the training-time transition text, the noise draw, the EliGen mask preparation.
Needs the reference tree (skipped on the GPU box, where only the built package travels)."""
import copy
import os
import sys

import pytest
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_import import ReferenceModules, reference_root  # noqa: E402

REF = "/root/reference"
TOK = os.path.join(REF, "DiffSynth-Studio", "models", "Qwen", "Qwen-Image", "tokenizer")
PROC = os.path.join(REF, "DiffSynth-Studio", "models", "Qwen", "Qwen-Image-Edit", "processor")
pytestmark = pytest.mark.skipif(reference_root() is None or not os.path.isdir(TOK), reason="no reference tree / tokenizer files on this machine")


class RecordingVL:
    """Qwen2.5-VL stand-in: hidden states are a deterministic function of the token ids, `generate` appends a canned reply; every call is logged."""

    def __init__(self, reply_ids):
        self.reply_ids, self.log = reply_ids, []

    def _rec(self, kind, kw):
        self.log.append((kind, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in kw.items() if v is not None}))

    def edit_forward(self, **kw):
        self._rec("edit_forward", kw)
        ids = kw["input_ids"]
        g = torch.Generator().manual_seed(int(ids.sum()) % 9973)
        return (torch.randn(ids.shape[0], ids.shape[1], 3584, generator=g),)

    def generate(self, **kw):
        self._rec("generate", kw)
        return torch.cat([kw["input_ids"], self.reply_ids.unsqueeze(0).expand(kw["input_ids"].shape[0], -1)], dim=1)


class RecordingVAE:
    def __init__(self):
        self.log = []

    def encode(self, x, **kw):
        self.log.append((tuple(x.shape), x.dtype, float(x.float().sum()), dict(kw)))
        g = torch.Generator().manual_seed(int(x.shape[-1]) * 31 + int(x.shape[-2]) + int(abs(float(x.float().mean())) * 1000))
        return torch.randn(x.shape[0], 16, x.shape[2] // 8, x.shape[3] // 8, generator=g).to(x.dtype)


@pytest.fixture(scope="module")
def ref():
    with ReferenceModules() as r:
        yield r


@pytest.fixture(scope="module")
def tok_proc():
    from transformers import Qwen2Tokenizer, Qwen2VLProcessor
    tok = Qwen2Tokenizer.from_pretrained(TOK)
    base = Qwen2VLProcessor.from_pretrained(PROC)
    proc = Qwen2VLProcessor(image_processor=base.image_processor, tokenizer=Qwen2Tokenizer.from_pretrained(TOK), video_processor=base.video_processor,
                            chat_template=base.chat_template)
    return tok, proc


def make_pipe(tok_proc, reply: str, training=False):
    from physicedit_b200.dit import QwenImageDiT  # noqa: F401
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    tok, proc = tok_proc
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.text_encoder = RecordingVL(tok(reply, return_tensors="pt").input_ids[0])
    pipe.vae = RecordingVAE()
    pipe.attach_tokenizer(tokenizer=tok, processor=proc)
    if training:
        pipe.scheduler.set_timesteps(1000, training=True)
    else:
        pipe.scheduler.set_timesteps(4, dynamic_shift_len=24)
    return pipe


def same(a, b, path="out"):
    if torch.is_tensor(a) or torch.is_tensor(b):
        assert torch.is_tensor(a) and torch.is_tensor(b), path
        assert a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b), f"{path}: tensors differ ({a.shape} {a.dtype} vs {b.shape} {b.dtype})"
    elif isinstance(a, Image.Image) or isinstance(b, Image.Image):
        assert isinstance(a, Image.Image) and isinstance(b, Image.Image) and a.size == b.size and a.tobytes() == b.tobytes(), path
    elif isinstance(a, dict):
        assert isinstance(b, dict) and set(a) == set(b), f"{path}: keys {sorted(a)} vs {sorted(b)}"
        for k in a:
            same(a[k], b[k], f"{path}.{k}")
    elif isinstance(a, (list, tuple)):
        assert isinstance(b, (list, tuple)) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            same(x, y, f"{path}[{i}]")
    elif hasattr(a, "__dict__") and not callable(a) and type(a).__name__ == "ControlNetInput":
        same(vars(a), vars(b), path)
    else:
        assert a == b, f"{path}: {a!r} vs {b!r}"


def run_units(units, runner, pipe, shared, posi, nega, skip=()):
    pipe.text_encoder.log, pipe.vae.log = [], []
    shared, posi, nega = copy.copy(shared), copy.copy(posi), copy.copy(nega)
    for u in units:
        if type(u).__name__ in skip:
            continue
        shared, posi, nega = runner(u, pipe, shared, posi, nega)
    return dict(shared=shared, posi=posi, nega=nega, vl=list(pipe.text_encoder.log), vae=list(pipe.vae.log))


def both(ref, pipe, shared, posi, nega, skip=()):
    from physicedit_b200 import units as U
    names = [type(u).__name__ for u in U.default_units()]
    ref_units = [getattr(ref.phys, n)() for n in names]                 # the order pipe.units registers them in (:233-245)
    torch.manual_seed(123)
    r = run_units(ref_units, ref.phys.PipelineUnitRunner(), pipe, shared, posi, nega, skip)
    torch.manual_seed(123)
    o = run_units(U.default_units(), U.PipelineUnitRunner(), pipe, shared, posi, nega, skip)
    return r, o


def request(**over):
    shared = {"cfg_scale": 4.0, "input_image": None, "denoising_strength": 1.0, "inpaint_mask": None, "inpaint_blur_size": None, "inpaint_blur_sigma": None,
              "height": 100, "width": 150, "seed": 7, "rand_device": "cpu", "enable_fp8_attention": False, "num_inference_steps": 4,
              "blockwise_controlnet_inputs": None, "tiled": False, "tile_size": 128, "tile_stride": 64, "eligen_entity_prompts": None,
              "eligen_entity_masks": None, "eligen_enable_on_negative": False, "edit_image": None, "edit_image_auto_resize": True,
              "edit_rope_interpolation": False, "context_image": None, "supported_rules": None, "contradicted_rules": None, "middle_key_frames": None,
              "stitched_image": None, "state": None, "transition": None, "triplet": None, "is_train": False}
    shared.update(over)
    return shared


def picture(w, h, seed):
    g = torch.Generator().manual_seed(seed)
    return Image.fromarray((torch.rand(h, w, 3, generator=g) * 255).to(torch.uint8).numpy())


VISUAL = ("QwenImageUnit_PhysicalVisualEmbedder",)          # dropped by __call__ at inference (is_train=False); covered on the training side below
NO_REASONING = VISUAL + ("QwenImageUnit_PhysicalVerbalEmbedder",)      # `have_text_reasoning=False`: the reference's unit needs ONE edit image (:962)


@pytest.mark.parametrize("reply", ['{"middle_transition_prompt": " The cup tips over; water spreads. "}',
                                   '{"physical_reasoning": "gravity", "middle_transition_prompt": "falls", "final_state_prompt": "on the floor"}',
                                   'Sure! {"Reasoning": "ice melts above 0 C"} hope that helps',
                                   'no json here at all', '{"middle_transition_prompt": 5}', '{"middle_transition_prompt": "a", "Reasoning": "b"}'])
def test_inference_request_with_one_edit_image(ref, tok_proc, reply):
    pipe = make_pipe(tok_proc, reply)
    r, o = both(ref, pipe, request(edit_image=picture(800, 600, 1)), {"prompt": "knock the cup over"}, {"negative_prompt": ""}, skip=VISUAL)
    same(r["shared"], o["shared"], "shared"); same(r["posi"], o["posi"], "posi"); same(r["nega"], o["nega"], "nega")
    same(r["vl"], o["vl"], "text-encoder calls"); same(r["vae"], o["vae"], "vae calls")
    assert [c[0] for c in o["vl"]] == ["generate", "generate", "edit_forward", "edit_forward"] and o["vl"][0][1]["max_new_tokens"] == 1000
    assert o["shared"]["height"] == 112 and o["shared"]["width"] == 160 and int(o["posi"]["special_token_mask"].sum()) == 64
    assert isinstance(o["posi"]["physical_txt"], str) and o["posi"]["prompt_emb"].dtype == torch.bfloat16


def test_prompt_only_multi_image_context_and_img2img_requests(ref, tok_proc):
    pipe = make_pipe(tok_proc, '{"middle_transition_prompt": "x"}')
    # no edit image: the text-to-image template (34 dropped tokens); cfg_scale 1: the negative branch inherits the positive fields
    r, o = both(ref, pipe, request(cfg_scale=1, height=64, width=64), {"prompt": "a red cube on a glass table"}, {"negative_prompt": "blurry"}, skip=NO_REASONING)
    for k in ("shared", "posi", "nega", "vl", "vae"):
        same(r[k], o[k], k)
    assert o["posi"]["special_token_mask"] is None and [c[0] for c in o["vl"]].count("edit_forward") == 1
    # a list of edit images ("Picture 1: ... Picture 2: ..."), a context image, an input image to start from (noised at timesteps[0])
    pipe.scheduler.set_timesteps(4, denoising_strength=0.7, dynamic_shift_len=70)       # what __call__ does before the units run (:600)
    r, o = both(ref, pipe, request(edit_image=[picture(640, 480, 2), picture(300, 500, 3)], context_image=picture(200, 120, 4), input_image=picture(160, 112, 5),
                                   edit_image_auto_resize=False, denoising_strength=0.7),
                {"prompt": "swap the two objects"}, {"negative_prompt": ""}, skip=NO_REASONING)
    for k in ("shared", "posi", "nega", "vl", "vae"):
        same(r[k], o[k], k)
    assert isinstance(o["shared"]["edit_latents"], list) and len(o["shared"]["edit_latents"]) == 2 and o["shared"]["context_latents"].shape == (1, 16, 14, 20)
    assert o["shared"]["input_latents"] is not None and not torch.equal(o["shared"]["latents"], o["shared"]["noise"])


def test_inpaint_mask_and_eligen_units(ref, tok_proc):
    pipe = make_pipe(tok_proc, '{"middle_transition_prompt": "x"}')
    masks = [picture(160, 112, 8).convert("L").point(lambda v: 255 if v > 128 else 0).convert("RGB"), picture(160, 112, 9).convert("RGB")]
    req = request(edit_image=picture(512, 512, 6), height=112, width=160, inpaint_mask=picture(160, 112, 7), inpaint_blur_size=2, inpaint_blur_sigma=1.5,
                  input_image=picture(160, 112, 10), eligen_entity_prompts=["a green bottle", "a wooden spoon"], eligen_entity_masks=masks,
                  eligen_enable_on_negative=True)
    r, o = both(ref, pipe, req, {"prompt": "put the spoon in the bottle"}, {"negative_prompt": "low quality"}, skip=VISUAL)
    for k in ("shared", "posi", "nega", "vl", "vae"):
        same(r[k], o[k], k)
    assert o["shared"]["inpaint_mask"].shape == (1, 1, 14, 20) and len(o["posi"]["entity_prompt_emb"]) == 2 and o["posi"]["entity_masks"].shape[1] == 2
    assert "entity_prompt_emb" in o["nega"]


def test_blockwise_controlnet_unit(ref, tok_proc):
    from physicedit_b200.compat import ControlNetInput
    pipe = make_pipe(tok_proc, '{"middle_transition_prompt": "x"}')
    hole = picture(160, 112, 41).convert("L").point(lambda v: 255 if v > 100 else 0).convert("RGB")
    inputs = [ControlNetInput(image=picture(160, 112, 40)), ControlNetInput(image=picture(160, 112, 42), inpaint_mask=hole, scale=0.5)]
    r, o = both(ref, pipe, request(edit_image=picture(512, 384, 43), height=112, width=160, blockwise_controlnet_inputs=inputs),
                {"prompt": "follow the sketch"}, {"negative_prompt": ""}, skip=VISUAL)
    for k in ("shared", "posi", "nega", "vl", "vae"):
        same(r[k], o[k], k)
    cond = o["shared"]["blockwise_controlnet_conditioning"]
    assert [c.shape for c in cond] == [(1, 16, 14, 20), (1, 17, 14, 20)]             # the masked request carries its mask as a 17th channel


def test_training_request(ref, tok_proc):
    """The sample dictionary of PhysicalEditingDataset as train_physicedit.py::forward_preprocess lays it out (:257-295): cfg_scale 1, the target
    image as `input_image`, rules + key frames + triplet -> the transition text is assembled from the triplet, no generation."""
    pipe = make_pipe(tok_proc, '{"middle_transition_prompt": "never used"}', training=True)
    frames = [picture(96, 64, 20 + i) for i in range(6)]
    req = {"input_image": picture(96, 64, 30), "height": 64, "width": 96, "cfg_scale": 1, "rand_device": "cpu", "use_gradient_checkpointing": True,
           "use_gradient_checkpointing_offload": False, "edit_image_auto_resize": True, "edit_image": picture(96, 64, 31),
           "supported_rules": [{"id": "r1", "instruction": "things fall", "matched_cues": ["down"]}], "contradicted_rules": [], "middle_key_frames": frames,
           "stitched_image": picture(192, 192, 32), "state": "solid", "transition": "melting",
           "triplet": {"middle_transition_prompt": "the ice softens", "final_state_prompt": "a puddle"}}
    r, o = both(ref, pipe, req, {"prompt": "melt the ice"}, {"negative_prompt": ""}, skip=VISUAL)
    for k in ("shared", "posi", "nega", "vl", "vae"):
        same(r[k], o[k], k)
    assert o["posi"]["physical_txt"] == "Middle Transition Prompt: the ice softens\nFinal State Prompt: a puddle"
    assert [c[0] for c in o["vl"]] == ["edit_forward"] and torch.equal(o["shared"]["latents"], o["shared"]["noise"]) and o["shared"]["input_latents"] is not None


def test_physical_visual_embedder_unit_on_the_emulated_abi(ref, tok_proc, monkeypatch):
    """The training-only unit (:991-1118): same RandomCrop draws (global torch RNG), DINOv2 / resamplers / adapters are this package's native modules
    on the emulated C ABI for BOTH units -- what differs is the glue (the reference's torch ops and einops vs `physical_visual_embeddings`), so the two
    pseudo targets agree to bf16 rounding.  The reference unit also encodes a description of the middle frames with the text encoder and discards it
    (:1059-1065); this package skips that dead forward, so the text-encoder logs are not compared."""
    from abi_emulator import EmulatedNative
    from physicedit_b200 import adapters, native as nv
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    from physicedit_b200 import units as U
    emu = EmulatedNative()
    monkeypatch.setattr(nv.Native, "get", classmethod(lambda cls, idx=0: emu))
    monkeypatch.setattr(adapters, "_nat", lambda t: emu)
    tok, proc = tok_proc
    torch.manual_seed(9)
    pipe = QwenImagePhysicPipeline(device="cpu", torch_dtype=torch.bfloat16, dinov2_config=dict(hidden=768, layers=1, heads=12))
    pipe.text_encoder = RecordingVL(tok("x", return_tensors="pt").input_ids[0])
    pipe.vae = RecordingVAE()
    pipe.attach_tokenizer(tokenizer=tok, processor=proc)
    pipe.to(torch.bfloat16)
    pipe.eval()
    frames = [picture(96, 64, 50 + i) for i in range(3)]
    kw = dict(middle_key_frames=frames, edit_image=picture(96, 64, 60), tiled=False, tile_size=128, tile_stride=64)
    with torch.no_grad():
        torch.manual_seed(77)
        want = ref.phys.QwenImageUnit_PhysicalVisualEmbedder().process(pipe, **kw)
        torch.manual_seed(77)
        got = U.QwenImageUnit_PhysicalVisualEmbedder().process(pipe, **kw)
    assert set(got) == set(want) == {"pseudo_special_emb_dino", "pseudo_special_emb_vae"}
    for k in got:
        e = ((got[k].float() - want[k].float()).norm() / want[k].float().norm()).item()
        print(f"{k}: this package's unit vs the reference's unit on the same native modules: rel-L2 {e:.3e}")
        assert got[k].shape == want[k].shape == (1, 64, 3584) and got[k].dtype == want[k].dtype and e < 1e-2, (k, e)
