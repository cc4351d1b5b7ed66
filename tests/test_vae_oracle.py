"""Pins oracle/vae_oracle.py against tests/golden/vae.pt, which oracle/make_golden_vae.py produced by running the reference's
own QwenImageVAE (DiffSynth-Studio/diffsynth/models/qwen_image_vae.py) on the same seeded synthetic weights.  CPU only."""
import hashlib

import pytest
import torch

from oracle import vae_oracle as VO


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item()


@pytest.fixture(scope="module")
def weights():
    return {k: v.to(torch.bfloat16).float() for k, v in VO.vae_synth_weights(seed=21).items()}


def test_vae_param_inventory_matches_reference(golden):
    g = golden("vae")["meta"]
    shapes = VO.vae_param_shapes()
    h = hashlib.md5(",".join(f"{k}:{'_'.join(map(str, v))}" for k, v in sorted(shapes.items())).encode()).hexdigest()
    assert h == g["key_hash"]
    assert len(shapes) == g["n_tensors"] == 194
    n = 0
    for v in shapes.values():
        m = 1
        for d in v:
            m *= d
        n += m
    assert n == g["n_params"]


def test_vae_oracle_fp32_matches_reference(golden, weights):
    g = golden("vae")
    for key, c in g["cases"].items():
        inp = VO.vae_inputs(c["h8"], c["w8"], c["seed"], dtype=torch.bfloat16)
        enc = VO.encode(weights, inp["image"].float())
        dec = VO.decode(weights, inp["latents"].float())
        assert enc.shape == c["fp32"]["encode"].shape and dec.shape == c["fp32"]["decode"].shape
        # a 2-D conv with the live temporal slice vs the reference's zero-padded 3-D conv: same products, different summation order
        assert rel_l2(enc, c["fp32"]["encode"]) < 2e-5, key
        assert rel_l2(dec, c["fp32"]["decode"]) < 2e-5, key


def test_vae_oracle_bf16_within_reference_noise(golden, weights):
    """bf16: the oracle's distance to the reference's bf16 output is of the order of the reference's own bf16-vs-fp32 error."""
    g = golden("vae")
    Wb = {k: v.to(torch.bfloat16) for k, v in weights.items()}
    for key, c in g["cases"].items():
        inp = VO.vae_inputs(c["h8"], c["w8"], c["seed"], dtype=torch.bfloat16)
        enc = VO.encode(Wb, inp["image"])
        dec = VO.decode(Wb, inp["latents"])
        floor_e = rel_l2(c["bf16"]["encode"], c["fp32"]["encode"])
        floor_d = rel_l2(c["bf16"]["decode"], c["fp32"]["decode"])
        assert rel_l2(enc, c["fp32"]["encode"]) <= floor_e * 1.5 + 1e-3, key
        assert rel_l2(dec, c["fp32"]["decode"]) <= floor_d * 1.5 + 1e-3, key
