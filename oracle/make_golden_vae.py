#!/usr/bin/env python
"""Generates tests/golden/vae.pt by importing and running the REFERENCE QwenImageVAE itself (authoring container only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_vae.py

Same import recipe as oracle/make_golden.py (namespace stub for `diffsynth`, stub `modelscope`).  The reference ships no VAE
weights or test vectors, so the class is run on the seeded synthetic weights of oracle/vae_oracle.py::vae_synth_weights
(load_state_dict(strict=True) proves the key / shape inventory) in fp32 and in bf16 on small ragged images.
Nothing in tests/, bench.py or smoke() reads /root/reference at run time.
"""
import hashlib
import importlib
import os
import sys
import types

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vae_oracle as VO  # noqa: E402

REF = "/root/reference/DiffSynth-Studio/diffsynth"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference_vae():
    pkg = types.ModuleType("diffsynth")
    pkg.__path__ = [REF]
    sys.modules["diffsynth"] = pkg
    ms = types.ModuleType("modelscope")
    ms.snapshot_download = lambda *a, **k: None
    sys.modules["modelscope"] = ms
    return importlib.import_module("diffsynth.models.qwen_image_vae")


def key_hash(sd) -> str:
    return hashlib.md5(",".join(f"{k}:{'_'.join(map(str, v.shape))}" for k, v in sorted(sd.items())).encode()).hexdigest()


@torch.no_grad()
def main():
    vae_mod = import_reference_vae()
    W = {k: v.to(torch.bfloat16).float() for k, v in VO.vae_synth_weights(seed=21).items()}   # bf16-representable in both runs
    out = {"meta": dict(w_seed=21, cases={}), "cases": {}}
    with torch.device("meta"):
        probe = vae_mod.QwenImageVAE()
    out["meta"]["key_hash"] = key_hash(probe.state_dict())
    out["meta"]["n_tensors"] = len(probe.state_dict())
    out["meta"]["n_params"] = sum(v.numel() for v in probe.state_dict().values())
    for dtype, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        m = vae_mod.QwenImageVAE()
        m.load_state_dict({k: v.to(dtype) for k, v in W.items()}, strict=True)
        m = m.to(dtype).eval()
        for (h8, w8, seed) in ((4, 4, 31), (6, 10, 32)):
            inp = VO.vae_inputs(h8, w8, seed, dtype=torch.bfloat16)
            img, lat = inp["image"].to(dtype), inp["latents"].to(dtype)
            enc = m.encode(img, tiled=False, tile_size=(30, 52), tile_stride=(15, 26))
            dec = m.decode(lat, device="cpu", tiled=False)
            c = out["cases"].setdefault(f"{h8}x{w8}", dict(h8=h8, w8=w8, seed=seed))
            c[tag] = dict(encode=enc.clone(), decode=dec.clone())
            print(tag, h8, w8, "encode", tuple(enc.shape), float(enc.float().abs().mean()), "decode", tuple(dec.shape), float(dec.float().abs().mean()))
    torch.save(out, os.path.join(GOLD, "vae.pt"))
    print("vae.pt", os.path.getsize(os.path.join(GOLD, "vae.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
