"""Host logic of the native VAE (physicedit_b200/vae.py) on CPU: parameter inventory, weight repacking and the orchestration of
the C-ABI calls, run against tests/abi_emulator.py (a contract-level emulation of the entry points) and compared with the oracle /
the reference goldens.  The GPU tests (tests/test_vae_gpu.py) run the same host code on the real library."""
import hashlib

import pytest
import torch

from oracle import vae_oracle as VO
from physicedit_b200 import vae as V
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from abi_emulator import EmulatedNative  # noqa: E402


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item()


@pytest.fixture(scope="module")
def model():
    W = {k: v.to(torch.bfloat16) for k, v in VO.vae_synth_weights(seed=21).items()}
    with torch.device("meta"):
        m = V.QwenImageVAE()
    m.load_state_dict(W, assign=True, strict=True)
    m.prepare()
    return m


def test_state_dict_inventory_is_the_reference_s(golden):
    with torch.device("meta"):
        m = V.QwenImageVAE()
    sd = m.state_dict()
    h = hashlib.md5(",".join(f"{k}:{'_'.join(map(str, v.shape))}" for k, v in sorted(sd.items())).encode()).hexdigest()
    g = golden("vae")["meta"]
    assert h == g["key_hash"] and len(sd) == g["n_tensors"]
    assert list(sd) == list(VO.vae_param_shapes()) or set(sd) == set(VO.vae_param_shapes())


def test_refuses_cpu_tensors(model):
    from physicedit_b200 import native as nv
    with pytest.raises(nv.NativeUnavailable):
        model.decode(torch.zeros(1, 16, 4, 4, dtype=torch.bfloat16))
    with pytest.raises(nv.NativeUnavailable):
        model.encode(torch.zeros(1, 3, 32, 32))


def test_host_logic_decode_and_encode_match_reference(golden, model):
    g = golden("vae")
    for key, c in g["cases"].items():
        inp = VO.vae_inputs(c["h8"], c["w8"], c["seed"], dtype=torch.bfloat16)
        emu = EmulatedNative()
        dec = torch.empty_like(c["bf16"]["decode"][0])
        model._decode_one(emu, model._packed, inp["latents"][0].contiguous(), dec)
        enc = torch.empty_like(c["bf16"]["encode"][0])
        model._encode_one(emu, model._packed, inp["image"][0].contiguous(), enc)
        floor_d = rel_l2(c["bf16"]["decode"], c["fp32"]["decode"])
        floor_e = rel_l2(c["bf16"]["encode"], c["fp32"]["encode"])
        err_d = rel_l2(dec, c["fp32"]["decode"][0])
        err_e = rel_l2(enc, c["fp32"]["encode"][0])
        assert err_d <= floor_d * 1.5 + 1e-3, (key, err_d, floor_d)
        assert err_e <= floor_e * 1.5 + 1e-3, (key, err_e, floor_e)
        kinds = {k for k, _ in emu.calls}
        assert {"pe_conv2d", "pe_gemm", "pe_channel_rmsnorm", "pe_upsample2x", "pe_space_to_depth", "pe_softmax_rows", "pe_transpose"} <= kinds


def test_load_vae_by_registry_hash(tmp_path, capsys):
    """configs/model_config.py:24: the VAE checkpoint is recognised by its key/shape hash; anything else prints and returns None
    (model_manager.py:375-376)."""
    from physicedit_b200.pipeline import VAE_KEY_HASH, hash_state_dict_keys, load_vae
    W = {k: v.to(torch.bfloat16) for k, v in VO.vae_synth_weights(seed=21).items()}
    assert hash_state_dict_keys(W) == VAE_KEY_HASH == "ed4ea5824d55ec3107b09815e318123a"
    good = tmp_path / "vae.pt"
    torch.save(W, good)
    m = load_vae(str(good), torch_dtype=torch.bfloat16, device="cpu")
    assert isinstance(m, V.QwenImageVAE) and m.decoder.conv_out.weight.shape == (3, 96, 3, 3, 3)
    assert torch.equal(m.decoder.conv_out.weight, W["decoder.conv_out.weight"])
    bad = tmp_path / "other.pt"
    torch.save({"x.weight": torch.zeros(2, 2)}, bad)
    assert load_vae(str(bad), device="cpu") is None
    assert "cannot detect the model type" in capsys.readouterr().out
