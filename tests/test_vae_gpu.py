"""Parity of the native QwenImageVAE path through the C ABI on a B200: every new entry point against the contract-level
emulation (tests/abi_emulator.py) on seeded ragged inputs, then whole encode / decode against the goldens the reference class produced
(tests/golden/vae.pt) and against the fp32 oracle at a mid size.  Run with `-m gpu`.

Tolerances: layout kernels bit-exact; the convolution / GEMM differ from an fp32 CPU conv only by summation order, i.e. by at most one
bf16 rounding of a value (checked as: <= 1 bf16 ulp everywhere, and rel-L2 <= 2e-3); whole-model error vs the fp32 oracle must stay
within the reference's own bf16-vs-fp32 noise floor (measured in the goldens, ~1.1e-2) + 1e-3 -- the DiT's rule."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from abi_emulator import EmulatedNative  # noqa: E402
from oracle import vae_oracle as VO  # noqa: E402

gpu = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item()


def assert_within_one_ulp(got, want, what, frac_exact=0.90, floor=1e-30):
    """bf16 tensors equal up to one unit in the last place (summation-order noise), most entries identical.  `floor`: magnitude of the
    operands of a final add (residual epilogue), whose one-ulp flips survive cancellation."""
    g, w = got.float().cpu(), want.float().cpu()
    # entries that cancel to ~0 carry the fp32 summation-order noise of their O(rms) partial sums: measure them against 2 % of the rms
    floor = max(floor, 0.02 * w.pow(2).mean().sqrt().item())
    ulp = torch.maximum(w.abs(), g.abs()).clamp_min(floor) * 2.0 ** -7      # >= 1 bf16 ulp of the larger magnitude
    bad = (g - w).abs() > ulp
    assert not bad.any(), f"{what}: {int(bad.sum())} of {g.numel()} entries differ by more than 1 bf16 ulp; max abs diff {(g - w).abs().max().item()}"
    assert (g == w).float().mean().item() >= frac_exact, f"{what}: only {(g == w).float().mean().item():.3f} of the entries are identical"


@pytest.fixture(scope="module")
def nat():
    from physicedit_b200 import native as nv
    return nv.Native.get(0)


@pytest.fixture(scope="module")
def emu():
    return EmulatedNative()


def _rand(shape, seed, scale=1.0):
    g = torch.Generator("cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16)


@gpu
@pytest.mark.parametrize("pair", [0, 1, 2])       # pe_conv2d flags: 0 default (two patches per tile when N <= 128), 1 CTA pair, 2 single patch
@pytest.mark.parametrize("H,W,C,N,kh,kw,pad,ld_extra,residual", [
    (13, 21, 96, 96, 3, 3, 1, 0, False),      # ragged patch grid, C not a multiple of 64, trimmed n-tile
    (16, 16, 64, 384, 3, 3, 1, 0, False),     # two n-tiles (256 + 128)
    (8, 8, 384, 384, 3, 3, 1, 0, True),       # narrow map (16 x 8 pixel patches), residual epilogue
    (9, 30, 192, 8, 3, 3, 1, 0, False),       # padded 3-channel output (decoder.conv_out)
    (12, 20, 384, 96, 2, 2, 0, 0, False),     # the downsample form: 2x2 taps, zeros only on the bottom / right
    (7, 5, 64, 16, 3, 3, 1, 64 - 16, False),  # narrow output inside a wider map (ldo > N), tiny map
    (40, 56, 96, 192, 3, 3, 1, 0, True),      # several m-tiles per row and column
    (24, 33, 96, 96, 3, 3, 1, 0, True),       # dual-patch tiles with a ragged last tile row (24 = 16 + 8) and residual
    (50, 16, 128, 128, 3, 3, 1, 0, False),    # N = 128 exactly; lower patch of the last tile entirely outside the map
])
def test_conv2d_matches_contract(nat, emu, H, W, C, N, kh, kw, pad, ld_extra, residual, pair):
    cpad = (C + 63) // 64 * 64
    x = _rand((H * W, C), 1 + H)
    w = _rand((N, kh * kw * cpad), 2 + W, scale=(kh * kw * C) ** -0.5)
    b = _rand((N,), 3, 0.1)
    out0 = _rand((H * W, N + ld_extra), 4)
    gate = torch.ones(N, dtype=torch.bfloat16)
    epi = 4 if residual else 0
    want = out0.clone()
    emu.conv2d(x, H, W, C, w, b, want, N, kh, kw, pad, epi, gate=gate if residual else None)
    got = out0.clone().cuda()
    nat.conv2d(x.cuda(), H, W, C, w.cuda(), b.cuda(), got, N, kh, kw, pad, epi, gate=gate.cuda() if residual else None, flags=pair)
    nat.check_async()
    assert_within_one_ulp(got[:, :N], want[:, :N], "conv2d", floor=4.0 if residual else 1e-30)
    assert rel_l2(got[:, :N], want[:, :N]) < 2e-3
    if ld_extra:
        assert torch.equal(got[:, N:].cpu(), out0[:, N:]), "conv2d wrote outside its N columns"


@gpu
@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("M,N,K", [(300, 384, 384), (77, 64, 96), (1000, 1048, 384), (130, 16, 64), (600, 96, 192)])
def test_gemm_trim_n_and_f32_epilogue(nat, emu, M, N, K, pair):
    from physicedit_b200 import native as nv
    a, w, b = _rand((M, K), 5), _rand((N, K), 6, K ** -0.5), _rand((N,), 7, 0.1)
    want = torch.empty((M, N), dtype=torch.bfloat16)
    emu.gemm([dict(a=a, w=w, bias=b, out=want)], N, K, 0)
    got = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    nat.gemm([dict(a=a.cuda(), w=w.cuda(), bias=b.cuda(), out=got)], N, K, nv.EPI_BIAS, nv.GEMM_FLAG_TRIM_N | pair)
    nat.check_async()
    assert_within_one_ulp(got, want, "gemm trim_n")
    wantf = torch.empty((M, N), dtype=torch.float32)
    emu.gemm([dict(a=a, w=w, bias=None, out=wantf)], N, K, 7)
    gotf = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    nat.gemm([dict(a=a.cuda(), w=w.cuda(), bias=None, out=gotf)], N, K, nv.EPI_F32, nv.GEMM_FLAG_TRIM_N | pair)
    nat.check_async()
    assert torch.allclose(gotf.cpu(), wantf, rtol=1e-4, atol=1e-4), (gotf.cpu() - wantf).abs().max()


@gpu
@pytest.mark.parametrize("rows,C,act", [(1000, 96, True), (333, 192, True), (4097, 384, False), (5, 384, True)])
def test_channel_rmsnorm(nat, emu, rows, C, act):
    x = _rand((rows, C), 8, 2.0)
    gamma = (1 + 0.1 * torch.randn(C, generator=torch.Generator().manual_seed(9))).to(torch.bfloat16)
    want = torch.empty_like(x)
    emu.channel_rmsnorm(x, want, C, gamma, act)
    got = torch.empty_like(x, device="cuda")
    nat.channel_rmsnorm(x.cuda(), got, C, gamma.cuda(), act)
    nat.check_async()
    # the norm is a 96..384-term fp32 sum rounded to bf16: a different summation order can move it by one ulp for a whole row
    assert_within_one_ulp(got, want, "channel_rmsnorm", frac_exact=0.97)


@gpu
def test_layout_kernels_bit_exact(nat, emu):
    H, W, C = 10, 14, 96
    x = _rand((H * W, C), 10)
    for name, shape_out, args in (("upsample2x", (4 * H * W, C), (H, W, C)), ("space_to_depth", (H * W // 4, 4 * C), (H, W, C))):
        want = torch.empty(shape_out, dtype=torch.bfloat16)
        getattr(emu, name)(x, want, *args)
        got = torch.empty(shape_out, dtype=torch.bfloat16, device="cuda")
        getattr(nat, name)(x.cuda(), got, *args)
        nat.check_async()
        assert torch.equal(got.cpu(), want), name
    mean = torch.tensor(VO.LATENT_MEAN).to(torch.bfloat16)
    stdinv = (1 / torch.tensor(VO.LATENT_STD)).to(torch.bfloat16)
    lat = _rand((16, 6, 10), 11)
    for op in (0, 1, 2):
        want = torch.zeros((60, 64), dtype=torch.bfloat16)
        emu.nchw_to_nhwc(lat, want, 16, op, mean, stdinv)
        got = torch.zeros((60, 64), dtype=torch.bfloat16, device="cuda")
        nat.nchw_to_nhwc(lat.cuda(), got, 16, op, mean.cuda(), stdinv.cuda())
        nat.check_async()
        assert torch.equal(got.cpu(), want), f"nchw_to_nhwc op {op}"
        back_w = torch.empty((16, 6, 10), dtype=torch.bfloat16)
        emu.nhwc_to_nchw(want, back_w, 16, op, mean, stdinv)
        back_g = torch.empty((16, 6, 10), dtype=torch.bfloat16, device="cuda")
        nat.nhwc_to_nchw(got, back_g, 16, op, mean.cuda(), stdinv.cuda())
        nat.check_async()
        assert torch.equal(back_g.cpu(), back_w), f"nhwc_to_nchw op {op}"
    # the decode de-normalisation equals the reference expression on bf16 tensors (qwen_image_vae.py:724-725)
    ref = (lat.cuda().unsqueeze(0) / stdinv.cuda().view(1, 16, 1, 1) + mean.cuda().view(1, 16, 1, 1))[0]
    got = torch.zeros((60, 64), dtype=torch.bfloat16, device="cuda")
    nat.nchw_to_nhwc(lat.cuda(), got, 16, 1, mean.cuda(), stdinv.cuda())
    assert torch.equal(got[:, :16].t().reshape(16, 6, 10), ref)
    v = _rand((61, 384), 12)
    vt = torch.zeros((384, 64), dtype=torch.bfloat16, device="cuda")
    nat.transpose(v.cuda(), vt[:, :61])
    nat.check_async()
    assert torch.equal(vt[:, :61].cpu(), v.t()) and float(vt[:, 61:].abs().max()) == 0.0


@gpu
@pytest.mark.parametrize("rows,n,n_pad", [(60, 60, 64), (300, 4096, 4096), (3, 17, 24)])
def test_softmax_rows(nat, rows, n, n_pad):
    s = torch.randn((rows, n_pad), generator=torch.Generator().manual_seed(13)) * 20
    want = torch.softmax(s[:, :n] * 384 ** -0.5, dim=-1)
    got = torch.full((rows, n_pad), 7.0, dtype=torch.bfloat16, device="cuda")
    nat.softmax_rows(s.cuda(), got, n, 384 ** -0.5)
    nat.check_async()
    assert torch.allclose(got[:, :n].float().cpu(), want, rtol=1e-2, atol=1e-6)
    assert float(got[:, n:].abs().max()) == 0.0 if n_pad > n else True


def _build_vae(seed=21):
    from physicedit_b200.vae import QwenImageVAE
    W = {k: v.to(torch.bfloat16) for k, v in VO.vae_synth_weights(seed=seed).items()}
    with torch.device("meta"):
        m = QwenImageVAE()
    m.load_state_dict(W, assign=True, strict=True)
    return m.to("cuda").eval(), W


@pytest.fixture(scope="module")
def vae():
    return _build_vae()


@gpu
def test_vae_encode_decode_match_reference_goldens(nat, vae, golden):
    m, _ = vae
    g = golden("vae")
    for key, c in g["cases"].items():
        inp = VO.vae_inputs(c["h8"], c["w8"], c["seed"], dtype=torch.bfloat16)
        dec = m.decode(inp["latents"].cuda(), device="cuda", tiled=False)
        enc = m.encode(inp["image"].cuda(), tiled=False, tile_size=(30, 52), tile_stride=(15, 26))
        nat.check_async()
        assert dec.shape == c["bf16"]["decode"].shape and enc.shape == c["bf16"]["encode"].shape
        floor_d = rel_l2(c["bf16"]["decode"], c["fp32"]["decode"])
        floor_e = rel_l2(c["bf16"]["encode"], c["fp32"]["encode"])
        err_d, err_e = rel_l2(dec, c["fp32"]["decode"]), rel_l2(enc, c["fp32"]["encode"])
        print(f"vae {key}: decode err {err_d:.3e} (reference bf16 floor {floor_d:.3e}), encode err {err_e:.3e} (floor {floor_e:.3e}), "
              f"vs reference bf16: decode {rel_l2(dec, c['bf16']['decode']):.3e} encode {rel_l2(enc, c['bf16']['encode']):.3e}")
        assert err_d <= floor_d + 1e-3, (key, err_d, floor_d)
        assert err_e <= floor_e + 1e-3, (key, err_e, floor_e)


@gpu
def test_vae_mid_size_vs_fp32_oracle_and_batch(nat, vae):
    """192 x 320 image (latent 24 x 40, 960 attention positions, several conv m-tiles per level) against the fp32 oracle; a batch of two."""
    m, W = vae
    Wf = {k: v.float() for k, v in W.items()}
    inp = VO.vae_inputs(24, 40, 41, dtype=torch.bfloat16)
    inp2 = VO.vae_inputs(24, 40, 42, dtype=torch.bfloat16)
    lat = torch.cat([inp["latents"], inp2["latents"]]).cuda()
    dec = m.decode(lat)
    enc = m.encode(inp["image"].cuda())
    nat.check_async()
    assert dec.shape == (2, 3, 192, 320) and enc.shape == (1, 16, 24, 40)
    assert torch.isfinite(dec.float()).all() and torch.isfinite(enc.float()).all()
    want_d = VO.decode(Wf, inp["latents"].float())
    want_e = VO.encode(Wf, inp["image"].float())
    err_d, err_e = rel_l2(dec[:1], want_d), rel_l2(enc, want_e)
    # floor at THIS size: the oracle's bf16 mode (the reference's arithmetic op by op, pinned by tests/test_vae_oracle.py) vs its fp32 mode
    floor_d, floor_e = rel_l2(VO.decode(W, inp["latents"]), want_d), rel_l2(VO.encode(W, inp["image"]), want_e)
    print(f"vae 24x40: decode err {err_d:.3e} (floor {floor_d:.3e}), encode err {err_e:.3e} (floor {floor_e:.3e}) vs fp32 oracle")
    assert err_d <= floor_d + 1e-3 and err_e <= floor_e + 1e-3, (err_d, floor_d, err_e, floor_e)
    assert torch.equal(m.decode(lat[1:])[0], dec[1]), "batch element 1 differs from a single-image call"
    # 5-D (B, C, 1, H, W) inputs keep their frame axis, as in the reference (:706-735)
    assert m.decode(lat[:1].unsqueeze(2)).shape == (1, 3, 1, 192, 320)


@gpu
def test_vae_full_size_properties(nat, vae):
    """1024 x 1024 (BASELINE config #2's image size; latent 128 x 128, 16384 attention positions in two query chunks): shapes, finite
    values, run-to-run determinism."""
    m, _ = vae
    inp = VO.vae_inputs(128, 128, 51, dtype=torch.bfloat16)
    img = inp["image"].cuda()
    enc = m.encode(img)
    dec = m.decode(inp["latents"].cuda())
    nat.check_async()
    assert enc.shape == (1, 16, 128, 128) and dec.shape == (1, 3, 1024, 1024)
    assert torch.isfinite(enc.float()).all() and torch.isfinite(dec.float()).all()
    assert torch.equal(m.encode(img), enc) and torch.equal(m.decode(inp["latents"].cuda()), dec), "not deterministic"


@gpu
def test_pipeline_call_pil_in_pil_out(nat, vae):
    """QwenImagePhysicPipeline.__call__ (qwen_image_physical.py:544-669) end to end at toy size with the native VAE either side of the
    native denoise loop: PIL edit image -> preprocess -> vae.encode -> 2 CFG steps of a 1-block DiT -> vae.decode -> PIL image."""
    import numpy as np
    from PIL import Image
    from oracle import dit_oracle as O
    from physicedit_b200.dit import QwenImageDiT
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    m, _ = vae
    W = {k: v.to(torch.bfloat16) for k, v in O.synth_weights(O.dit_param_shapes(1), seed=4).items()}
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=1)
    dit.load_state_dict(W, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    pipe = QwenImagePhysicPipeline(device="cuda", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit.to("cuda").eval()
    pipe.vae = m
    A = {k: v.to(torch.bfloat16) for k, v in O.synth_weights(O.adapter_param_shapes(), seed=2).items()}
    pipe.visual_thinking_adapter.load_state_dict(A)
    pipe.visual_thinking_adapter.to(device="cuda", dtype=torch.bfloat16)
    posi = O.synth_inputs(64, 96, 72, seed=6, dtype=torch.bfloat16)
    nega = O.synth_inputs(64, 96, 69, seed=7, dtype=torch.bfloat16)
    keys = ("prompt_emb", "prompt_emb_mask", "special_token_mask")
    rng = np.random.default_rng(0)
    edit = Image.fromarray(rng.integers(0, 256, size=(64, 96, 3), dtype=np.uint8))
    kw = dict(prompt_inputs_posi={k: posi[k].cuda() for k in keys}, prompt_inputs_nega={k: nega[k].cuda() for k in keys}, edit_image=edit,
              edit_image_auto_resize=False, context_image=edit.resize((48, 32)), height=64, width=96, seed=3, num_inference_steps=2, is_train=False)
    img = pipe(**kw)
    nat.check_async()
    assert isinstance(img, Image.Image) and img.size == (96, 64)
    lat = pipe(output_type="latent", **{k: (v if not isinstance(v, dict) else {a: b.clone() for a, b in v.items()}) for k, v in kw.items()})
    assert lat.shape == (1, 16, 8, 12) and torch.isfinite(lat.float()).all()
    # the edit latents the call used are the VAE encoding of the pre-processed image (qwen_image_physical.py:1271-1273)
    enc = m.encode(pipe.preprocess_image(edit))
    want = VO.encode({k: v.float() for k, v in vae[1].items()}, pipe.preprocess_image(edit).float().cpu())
    assert rel_l2(enc, want) <= 2.0e-2


@gpu
@pytest.mark.parametrize("pair", [0, 1])
def test_gemm_epilogue_store_width_paths_agree(nat, pair):
    """The bf16 epilogues store 32 bytes per lane when the output is 32-byte aligned (ldo, N multiples of 16) and fall back to 16-byte
    stores otherwise: the same GEMM into an aligned buffer and into a view that starts 16 bytes into a wider buffer must be identical,
    for the plain and the in-place residual epilogue, and must not touch the columns around the view."""
    from physicedit_b200 import native as nv
    M, N, K = 300, 256, 192
    a, w, b = _rand((M, K), 21).cuda(), _rand((N, K), 22, K ** -0.5).cuda(), _rand((N,), 23, 0.1).cuda()
    gate = _rand((N,), 24).cuda()
    res = _rand((M, N), 25).cuda()
    for epi in (nv.EPI_BIAS, nv.EPI_GATE_RESIDUAL):
        aligned = res.clone()
        nat.gemm([dict(a=a, w=w, bias=b, out=aligned, gate=gate)], N, K, epi, pair)
        wide = torch.full((M, N + 16), 3.0, dtype=torch.bfloat16, device="cuda")
        view = wide[:, 8:8 + N]                      # 16-byte aligned rows, not 32-byte aligned
        view.copy_(res)
        nat.gemm([dict(a=a, w=w, bias=b, out=view, gate=gate)], N, K, epi, pair)
        nat.check_async()
        assert torch.equal(view, aligned), f"epilogue {epi}: 16-byte and 32-byte store paths differ"
        assert float((wide[:, :8] - 3.0).abs().max()) == 0.0 and float((wide[:, 8 + N:] - 3.0).abs().max()) == 0.0
