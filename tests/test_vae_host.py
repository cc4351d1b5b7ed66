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


def test_argument_errors_match_the_scope(model):
    """Single frames only (the pipeline never passes video to the VAE, qwen_image_physical.py:665,1273): T > 1 is refused loudly, and
    encode insists on the 8-pixel grid the three stride-2 stages need."""
    with pytest.raises(NotImplementedError):
        model.decode(torch.zeros(1, 16, 2, 4, 4, dtype=torch.bfloat16))
    with pytest.raises(NotImplementedError):
        model.encode(torch.zeros(1, 3, 5, 32, 32, dtype=torch.bfloat16))
    with pytest.raises(ValueError):
        model.encode(torch.zeros(1, 3, 36, 32, dtype=torch.bfloat16))
    with pytest.raises(ValueError):
        V.QwenImageVAE(z_dim=4)


def test_packed_weights_follow_weight_updates(model):
    """prepare() caches repacked weights; load_state_dict / .to() must drop the cache (otherwise a LoRA-style in-place update or a
    checkpoint reload would silently keep running on the old weights)."""
    assert model._packed is not None
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model.load_state_dict(sd, assign=True)
    assert model._packed is None
    model.prepare()
    w2d, b, n8, kh, kw, cin = model._packed["decoder.conv_out"]
    assert (n8, kh, kw, cin) == (8, 3, 3, 96) and w2d.shape == (8, 9 * 128) and float(w2d[3:].abs().max()) == 0.0
    # tap-major, channel-minor packing of the LAST temporal slice
    w = model.decoder.conv_out.weight
    assert torch.equal(w2d[:3].reshape(3, 3, 3, 128)[..., :96], w[:, :, -1].permute(0, 2, 3, 1))
    # stride-2 conv as a 2x2 conv over the space-to-depth map: tap (a, b), phase (py, px) <- kernel element (2a+py, 2b+px)
    w2d, b, n8, kh, kw, cin = model._packed["encoder.down_blocks.2"]
    wd = model.encoder.down_blocks[2].resample[1].weight
    v = w2d.reshape(96, 2, 2, 4, 96)
    assert (kh, kw, cin) == (2, 2, 384)
    assert torch.equal(v[:, 1, 0, 0], wd[:, :, 2, 0]) and torch.equal(v[:, 0, 1, 2], wd[:, :, 1, 2])
    assert torch.equal(v[:, 0, 0, 3], wd[:, :, 1, 1]) and torch.equal(v[:, 1, 1, 0], wd[:, :, 2, 2])
    assert float(v[:, 1, :, 2:].abs().max()) == 0.0 and float(v[:, :, 1, 1::2].abs().max()) == 0.0     # kernel rows / columns 3 do not exist


def test_edit_image_auto_resize_dimensions():
    """QwenImageUnit_EditImageEmbedder.calculate_dimensions / edit_image_auto_resize (qwen_image_physical.py:1249-1260): ~1024^2 pixels on
    a 32-pixel grid, aspect ratio kept.  Known answers computed from the reference formula."""
    from PIL import Image
    from physicedit_b200.pipeline import QwenImagePhysicPipeline as P
    assert P.calculate_dimensions(1024 * 1024, 1.0) == (1024, 1024)
    assert P.calculate_dimensions(1024 * 1024, 16 / 9) == (1376, 768)
    assert P.calculate_dimensions(1024 * 1024, 3 / 4) == (896, 1184)
    assert P.calculate_dimensions(1024 * 1024, 832 / 480) == (1344, 768)
    img = Image.new("RGB", (640, 360))
    assert P.auto_resize_edit_image(None, img).size == (1376, 768)
