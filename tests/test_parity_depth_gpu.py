"""Parity at depth and at the benchmark's own shapes (VERDICT r1 "next" item 1), `-m gpu`.

The fp32 oracle (oracle/dit_oracle.py) runs ON THE GPU here -- same functions, fp32 tensors, TF32 off -- so that the
configurations bench.py times can be checked, not only toy sizes:

  * the full 60-block forward at 1024^2 (S = 8192 + 512): native vs fp32 oracle, bar = floor + 1e-3 where floor is the
    reference arithmetic in bf16 (the oracle's bf16 mode = stock PyTorch ops on the same GPU) vs the same fp32 run;
    the native-vs-stock-bf16 distance (north_star's "1e-3" figure) is reported beside it,
  * a 12-step CFG loop at 8 blocks (compounding in-place prompt mutation),
  * config #4 (2048^2: S = 16384 + 4096 + 512) attention and one block; config #5 (512^2, 1536^2) one block,
  * the reference pipeline object itself driven through the native step function (the INTEGRATION.md seam).

Numbers are appended to gpurun_out/parity_depth.json when that directory exists (copied to profiles/ by hand).
"""
import json
import math
import os

import pytest
import torch

from oracle import dit_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gpu = pytest.mark.gpu
TOL_EXTRA = 1e-3


def rel_l2(a, b):
    a, b = a.float(), b.float().to(a.device)
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item()


def record(name, **vals):
    out = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out):
        return
    path = os.path.join(out, "parity_depth.json")
    try:
        data = json.load(open(path))
    except (OSError, ValueError):
        data = {}
    data[name] = vals
    json.dump(data, open(path, "w"), indent=1)


@pytest.fixture(autouse=True)
def exact_fp32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    torch.cuda.empty_cache()


class Upcast(dict):
    """The native model's bf16 state dict seen as fp32 weights, one tensor at a time (60 blocks in fp32 would be 81.7 GB)."""

    def __getitem__(self, k):
        return dict.__getitem__(self, k).float()

    def get(self, k, default=None):
        return self[k] if k in self else default


def device_model(layers, seed, gain=1.0):
    """bf16 weights of the real architecture generated on the device (like bench.build_model: U(-1/sqrt(K), 1/sqrt(K)) matrices,
    small biases) + the adapter; returns (pipe, dit state dict, adapter state dict)."""
    from physicedit_b200.dit import QwenImageDiT
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(seed)
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=layers)
    sd = {}
    for k, v in dit.state_dict().items():
        if v.dim() == 2:
            t = (torch.rand(v.shape, generator=g, device=dev) * 2 - 1) * (gain / math.sqrt(v.shape[1]))
        elif k.endswith(".bias"):
            t = (torch.rand(v.shape, generator=g, device=dev) * 2 - 1) * 0.02
        else:
            t = 1 + 0.1 * torch.randn(v.shape, generator=g, device=dev)
        sd[k] = t.to(torch.bfloat16)
    dit.load_state_dict({k: v.clone() for k, v in sd.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    for i, b in enumerate(dit.transformer_blocks):
        object.__setattr__(b, "_owner", (dit, i))
    pipe = QwenImagePhysicPipeline(device="cuda", torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit.eval()
    ad = {}
    for k, p in pipe.visual_thinking_adapter.state_dict().items():
        b = 1.0 / math.sqrt(p.shape[-1]) if p.dim() == 2 else 0.02
        ad[k] = ((torch.rand(p.shape, generator=g, device=dev) * 2 - 1) * b).to(torch.bfloat16)
    pipe.visual_thinking_adapter.load_state_dict(ad)
    return pipe, sd, ad


def oracle_forward(sd, ad, inp, t, H, W, dtype, layers, bf16_timestep=False):
    """O.model_fn on the GPU in `dtype` (fp32: the oracle proper; bf16: the reference's arithmetic = stock PyTorch ops).
    `bf16_timestep=True` with fp32: the timestep BOOKKEEPING of the bf16 path (t -> bf16 -> /1000 -> bf16, bf16 frequencies, bf16
    alpha -- bit-exact by contract) feeding fp32 arithmetic.  It removes the common-mode difference between the
    reference's fp32 and bf16 modes (a phase shift of up to ~2 rad in the top sinusoid components), so the floor it gives is
    rounding noise only: the stricter of the two protocols."""
    Wd = Upcast(sd) if dtype == torch.float32 else sd
    Ad = Upcast(ad) if (dtype == torch.float32 and ad is not None) else ad
    c = lambda x: x.cuda().to(dtype) if x.is_floating_point() else x.cuda()
    tt = t.cuda().to(torch.bfloat16 if bf16_timestep else dtype)
    sp = inp.get("special_token_mask")
    with torch.no_grad():
        return O.model_fn(Wd, Ad if sp is not None else None, c(inp["latents"]), tt, c(inp["prompt_emb"]).clone(), c(inp["prompt_emb_mask"]),
                          c(sp) if sp is not None else None, H, W, edit_latents=c(inp["edit_latents"]), num_layers=layers, cuda_scalar_div=True)


# ---------------------------------------------------------------------------------------------------------------------------
@gpu
def test_full_depth_forward_1024_vs_fp32_oracle_on_gpu():
    """BASELINE config #2, the forward bench.py times: 60 blocks, 4096 noise + 4096 edit tokens, T = 512, adapter on 64 tokens."""
    from physicedit_b200.model_fn import model_fn_qwen_image
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~70 GB of device memory")
    L, H, W, T = 60, 1024, 1024, 512
    pipe, sd, ad = device_model(L, seed=0)
    inp = O.synth_inputs(H, W, T, seed=100, dtype=torch.bfloat16, edit_hw=(1024, 1024))
    t = torch.tensor([744.611382484436]).to(torch.bfloat16)
    pe = inp["prompt_emb"].cuda().clone()
    y, _ = model_fn_qwen_image(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"].cuda(), timestep=t.cuda(),
                               prompt_emb=pe, prompt_emb_mask=inp["prompt_emb_mask"].cuda(), special_token_mask=inp["special_token_mask"].cuda(),
                               height=H, width=W, edit_latents=inp["edit_latents"].cuda(), is_train=False)
    pipe.dit.engine().nat.check_async()
    assert torch.isfinite(y.float()).all()
    y32 = oracle_forward(sd, ad, inp, t, H, W, torch.float32, L)
    y32b = oracle_forward(sd, ad, inp, t, H, W, torch.float32, L, bf16_timestep=True)
    y16 = oracle_forward(sd, ad, inp, t, H, W, torch.bfloat16, L)
    floor, err, vs_stock = rel_l2(y16, y32), rel_l2(y, y32), rel_l2(y, y16)
    floor_b, err_b = rel_l2(y16, y32b), rel_l2(y, y32b)
    record("forward_60_blocks_1024", err_native_vs_fp32=err, floor_stock_bf16_vs_fp32=floor, native_vs_stock_bf16=vs_stock,
           err_native_vs_fp32_bf16_timestep=err_b, floor_stock_bf16_vs_fp32_bf16_timestep=floor_b,
           out_abs_mean=y32.abs().mean().item(), blocks=L, S=8192 + T)
    print(f"\n60-block 1024^2 forward: native vs fp32 {err:.4e} (floor: stock bf16 vs fp32 {floor:.4e}); with the bf16 timestep bookkeeping in the "
          f"fp32 oracle {err_b:.4e} (floor {floor_b:.4e}); native vs stock bf16 {vs_stock:.4e}")
    assert err <= floor + TOL_EXTRA, (err, floor)
    assert err_b <= floor_b + TOL_EXTRA, (err_b, floor_b)


@gpu
def test_cfg_loop_12_steps_8_blocks_vs_fp32_loop():
    """12 denoise steps x 2 CFG forwards at 8 blocks: the adapter output of step k feeds step k+1 through the in-place prompt
    mutation (separately per branch), the Euler updates accumulate in bf16 -- error growth over the loop against the fp32 loop."""
    L, H, T1, T2, N = 8, 256, 96, 80, 12
    pipe, sd, ad = device_model(L, seed=3)
    posi = O.synth_inputs(H, H, T1, seed=11, dtype=torch.bfloat16)
    nega = O.synth_inputs(H, H, T2, seed=12, dtype=torch.bfloat16)
    keys = ("prompt_emb", "prompt_emb_mask", "special_token_mask")
    ip, in_ = {k: posi[k].cuda().clone() for k in keys}, {k: nega[k].cuda().clone() for k in keys}
    lat = pipe.denoise(posi["latents"].cuda(), ip, in_, posi["edit_latents"].cuda(), height=H, width=H, num_inference_steps=N, cfg_scale=4.0)

    def loop(dtype):
        Wd, Ad = (Upcast(sd), Upcast(ad)) if dtype == torch.float32 else (sd, ad)
        c = lambda d: {k: (v.cuda().to(dtype) if v.is_floating_point() else v.cuda()) for k, v in d.items()}
        p, n = c(posi), c(nega)
        with torch.no_grad():
            return O.denoise_loop(Wd, Ad, p["latents"].clone(), p, n, p["edit_latents"], H, H, N, num_layers=L, cuda_scalar_div=True,
                                  timestep_dtype=torch.bfloat16), p, n
    l32, p32, _ = loop(torch.float32)            # fp32 arithmetic on the bit-exact bf16 timestep bookkeeping (see oracle_forward)
    l16, p16, _ = loop(torch.bfloat16)
    floor, err = rel_l2(l16, l32), rel_l2(lat, l32)
    sm = posi["special_token_mask"][0].cuda()
    sp_err, sp_floor = rel_l2(ip["prompt_emb"][0][sm], p32["prompt_emb"][0][sm]), rel_l2(p16["prompt_emb"][0][sm], p32["prompt_emb"][0][sm])
    record("cfg_loop_12_steps_8_blocks", err_native_vs_fp32=err, floor_stock_bf16_vs_fp32=floor, native_vs_stock_bf16=rel_l2(lat, l16),
           special_tokens_err=sp_err, special_tokens_floor=sp_floor)
    print(f"\n12-step CFG loop, 8 blocks: native vs fp32 {err:.4e} (floor {floor:.4e}); compounded special tokens {sp_err:.4e} (floor {sp_floor:.4e})")
    assert err <= floor + TOL_EXTRA, (err, floor)
    assert torch.equal(ip["prompt_emb"][0][~sm].cpu(), posi["prompt_emb"][0][~sm.cpu()])       # non-special rows never touched
    assert sp_err <= sp_floor + 5e-3, (sp_err, sp_floor)


def _block_vs_oracle(S_hw, T, seed):
    """One block at S_img = sum(h*w) image tokens + T text tokens against the fp32 oracle block on the GPU (every row)."""
    from physicedit_b200.dit import QwenImageDiT
    pipe, sd, _ = device_model(1, seed=seed)
    eng = pipe.dit.engine()
    S_img = sum(h * w for _, h, w in S_hw)
    g = torch.Generator(device="cuda").manual_seed(seed)
    x0 = torch.randn(T + S_img, 3072, device="cuda", generator=g).bfloat16()
    temb = torch.randn(1, 3072, device="cuda", generator=g).bfloat16()
    rope = eng.rope(S_hw, T)
    ws = eng.workspace(S_img, T)
    mods = eng.block_mods(temb, [0])[0]
    x = x0.clone()
    eng.run_block(0, x, T, mods[0], rope, ws)
    eng.nat.check_async()
    tables = O.rope_tables(S_hw, T)
    with torch.no_grad():
        t32, i32 = O.block_forward(Upcast(sd), 0, x0[T:].float().unsqueeze(0), x0[:T].float().unsqueeze(0), temb.float(), tables)
        t16, i16 = O.block_forward(sd, 0, x0[T:].unsqueeze(0), x0[:T].unsqueeze(0), temb, tables)
    ref32, ref16 = torch.cat([t32[0], i32[0]]), torch.cat([t16[0], i16[0]])
    # the block is a residual update: compare the UPDATE (x_out - x_in), where a wrong kernel cannot hide behind the skip path
    d32 = ref32 - x0.float()
    return rel_l2(x.float() - x0.float(), d32), rel_l2(ref16.float() - x0.float(), d32), rel_l2(x, ref32)


@gpu
@pytest.mark.parametrize("name,S_hw,T", [("512", [(1, 32, 32), (1, 64, 64)], 512), ("1536", [(1, 96, 96), (1, 64, 64)], 512),
                                         ("2048", [(1, 128, 128), (1, 64, 64)], 512), ("2048_nega", [(1, 128, 128), (1, 64, 64)], 288)])
def test_one_block_at_config_4_and_5_shapes(name, S_hw, T):
    """BASELINE configs #4 (2048^2: 16384 + 4096 image tokens, 164 KV tiles) and #5 (512^2 / 1536^2 outputs; the edit image is
    always ~1024^2 = 4096 tokens): one whole block, every row, against the fp32 oracle on the GPU."""
    err, floor, full = _block_vs_oracle(S_hw, T, seed=len(name) + T)
    record(f"block_{name}_T{T}", update_err_native_vs_fp32=err, update_floor_stock_bf16_vs_fp32=floor, output_err=full)
    print(f"\nblock @{name} (T={T}): update err {err:.4e} (floor {floor:.4e}), output err {full:.4e}")
    assert err <= floor + TOL_EXTRA, (err, floor)


@gpu
@pytest.mark.parametrize("flags", [0, 16, 32])
def test_attention_config4_sequence_vs_fp32(flags):
    """S = 16384 + 4096 + 512 = 20992, 24 heads: every output element against exact fp32 softmax attention (one head at a time)."""
    from physicedit_b200 import native as nv
    nat = nv.Native.get(0)
    S, H = 20992, 24
    g = torch.Generator(device="cuda").manual_seed(5)
    q, k, v = (torch.randn(S, H * 128, device="cuda", generator=g).bfloat16() for _ in range(3))
    k[:, :128] *= 3.0                                   # one head with sharply peaked rows
    o = torch.empty_like(q)
    nat.attention(q, k, v, o, H, 1 / math.sqrt(128), flags)
    nat.check_async()
    worst = 0.0
    for h in range(H):
        sl = slice(h * 128, (h + 1) * 128)
        ref = torch.softmax(q[:, sl].float() @ k[:, sl].float().t() / math.sqrt(128), dim=-1) @ v[:, sl].float()
        worst = max(worst, rel_l2(o[:, sl], ref))
    record(f"attention_S20992_flags{flags}", worst_head_rel_l2=worst)
    assert worst < 5e-3, worst


# ---------------------------------------------------------------------------------------------------------------------------
@gpu
def test_reference_pipeline_object_runs_on_the_native_step_function(tmp_path):
    """The INTEGRATION.md seam on the REAL reference pipeline: its own `QwenImagePhysicPipeline` object (constructed by its own
    __init__), its own loop body (qwen_image_physical.py:646-661: CUDA bf16 timestep, progress_id, `self.step`), first with the
    stock `model_fn_qwen_image` + `QwenImageDiT` (stock PyTorch on this GPU, bf16 and fp32), then with `adopt_dit` / `adopt_adapter`
    and `pipe.model_fn = physicedit_b200.model_fn_qwen_image`."""
    from oracle import ref_import
    if ref_import.reference_root() is None:
        pytest.skip("no reference tree on this box (baseline/_ref is git-ignored; see DESIGN.md section 5)")
    import physicedit_b200 as pe
    L, H, T1, T2, N = 2, 128, 88, 72, 4
    Wsd = O.synth_weights(O.dit_param_shapes(L), seed=17, dtype=torch.bfloat16)
    Asd = O.synth_weights(O.adapter_param_shapes(), seed=18, dtype=torch.bfloat16)
    posi = O.synth_inputs(H, H, T1, seed=21, dtype=torch.bfloat16)
    nega = O.synth_inputs(H, H, T2, seed=22, dtype=torch.bfloat16)
    keys = ("prompt_emb", "prompt_emb_mask", "special_token_mask")
    ref_import.tiny_dinov2_folder(str(tmp_path / "dino"))

    with ref_import.ReferenceModules() as ref:
        def run(dtype, native):
            pipe = ref.phys.QwenImagePhysicPipeline(device="cuda", torch_dtype=dtype, dinov2_path=str(tmp_path / "dino"))
            pipe.dit = ref_import.build_reference_dit(ref, Wsd, L, dtype, "cuda")
            pipe.visual_thinking_adapter.load_state_dict({k: v.to(dtype) for k, v in Asd.items()})
            if native:                                                # INTEGRATION.md section 1, verbatim
                pipe.dit = pe.adopt_dit(pipe.dit)
                pipe.visual_thinking_adapter = pe.adopt_adapter(pipe.visual_thinking_adapter)
                pipe.model_fn = pe.model_fn_qwen_image
            c = lambda x: x.cuda().to(dtype) if x.is_floating_point() else x.cuda()
            shared = dict(latents=c(posi["latents"]), height=H, width=H, edit_latents=c(posi["edit_latents"]), is_train=False, cfg_scale=4.0)
            ip, in_ = {k: c(posi[k]).clone() for k in keys}, {k: c(nega[k]).clone() for k in keys}
            pipe.scheduler.set_timesteps(N, dynamic_shift_len=(H // 16) * (H // 16))
            launches0 = pe.native.Native.get(0).launches
            from torch.nn.attention import SDPBackend, sdpa_kernel
            backends = [SDPBackend.MATH] if dtype == torch.float32 else [SDPBackend.FLASH_ATTENTION, SDPBackend.CUDNN_ATTENTION, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]
            with torch.no_grad(), sdpa_kernel(backends):                                     # the reference's loop body, line for line (:646-661)
                models = {name: getattr(pipe, name) for name in pipe.in_iteration_models}
                for progress_id, timestep in enumerate(pipe.scheduler.timesteps):
                    timestep = timestep.unsqueeze(0).to(dtype=pipe.torch_dtype, device=pipe.device)
                    vp, _ = pipe.model_fn(**models, **shared, **ip, timestep=timestep, progress_id=progress_id)
                    vn, _ = pipe.model_fn(**models, **shared, **in_, timestep=timestep, progress_id=progress_id)
                    noise_pred = vn + 4.0 * (vp - vn)
                    shared["latents"] = pipe.step(pipe.scheduler, progress_id=progress_id, noise_pred=noise_pred, **shared)
            return shared["latents"], ip["prompt_emb"], pe.native.Native.get(0).launches - launches0
        l32, p32, _ = run(torch.float32, False)
        l16, p16, n_stock = run(torch.bfloat16, False)
        lnat, pnat, n_nat = run(torch.bfloat16, True)
    pe.native.Native.get(0).check_async()
    assert n_stock == 0 and n_nat > N * 2 * L * 9              # the adopted pipeline really ran libpe_b200 kernels
    floor, err = rel_l2(l16, l32), rel_l2(lnat, l32)
    record("reference_pipeline_seam", err_native_vs_ref_fp32=err, floor_ref_bf16_vs_ref_fp32=floor, native_vs_ref_bf16=rel_l2(lnat, l16),
           native_launches=n_nat)
    print(f"\nseam: native vs reference fp32 {err:.4e}; reference bf16 vs fp32 {floor:.4e}; native vs reference bf16 {rel_l2(lnat, l16):.4e}")
    assert err <= floor + TOL_EXTRA, (err, floor)
    sm = posi["special_token_mask"][0].cuda()
    assert torch.equal(pnat[0][~sm], posi["prompt_emb"].cuda()[0][~sm])
    assert rel_l2(pnat[0][sm], p32[0][sm]) <= rel_l2(p16[0][sm], p32[0][sm]) + 5e-3


# ---------------------------------------------------------------------------------------------------------------------------
@gpu
def test_special_token_count_is_not_fixed_at_64():
    """The reference gathers however many rows the mask selects (:1334).  70 special tokens: parity with the fp32 oracle;
    a caller-declared count smaller than the mask's: PE_ERR_INVALID_ARGUMENT through pe_check_async_error, never a silent drop."""
    from physicedit_b200 import native as nv
    from physicedit_b200.model_fn import model_fn_qwen_image
    pipe, sd, ad = device_model(1, seed=9)
    H, T = 64, 96
    inp = O.synth_inputs(H, H, T, seed=31, dtype=torch.bfloat16, n_special=70)
    t = torch.tensor([426.6734719276428]).to(torch.bfloat16)
    kw = dict(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"].cuda(), timestep=t.cuda(),
              prompt_emb_mask=inp["prompt_emb_mask"].cuda(), special_token_mask=inp["special_token_mask"].cuda(), height=H, width=H,
              edit_latents=inp["edit_latents"].cuda(), is_train=False)
    pe = inp["prompt_emb"].cuda().clone()
    y, _ = model_fn_qwen_image(prompt_emb=pe, **kw)
    nat = nv.Native.get(0)
    nat.check_async()
    y32 = oracle_forward(sd, ad, inp, t, H, H, torch.float32, 1)
    y16 = oracle_forward(sd, ad, inp, t, H, H, torch.bfloat16, 1)
    assert rel_l2(y, y32) <= rel_l2(y16, y32) + TOL_EXTRA
    sm = inp["special_token_mask"][0]
    assert not torch.equal(pe[0].cpu()[sm], inp["prompt_emb"][0][sm]) and torch.equal(pe[0].cpu()[~sm], inp["prompt_emb"][0][~sm])
    assert (pe[0].cpu()[sm] != inp["prompt_emb"][0][sm]).any(dim=1).all()          # all 70 rows were rewritten
    model_fn_qwen_image(prompt_emb=inp["prompt_emb"].cuda().clone(), n_special=64, **kw)
    with pytest.raises(nv.NativeError, match="selects 70 rows"):
        nat.check_async()
    nat.check_async()                                                               # the flag is cleared once reported


@gpu
def test_training_loss_matches_oracle_with_pinned_draws():
    """training_loss (:313-329) with the two random draws pinned: flow-matching MSE x training weight + adapter loss against the
    oracle's fp32 value, called the way scripts/train/train_physicedit.py:309-310 calls it (models inside **inputs)."""
    pipe, sd, ad = device_model(2, seed=13)
    pipe.scheduler.set_timesteps(1000, training=True)
    H, T = 128, 80
    inp = O.synth_inputs(H, H, T, seed=41, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(2)
    gt_d, gt_v = torch.randn(1, 64, 3584, generator=g).bfloat16(), torch.randn(1, 64, 3584, generator=g).bfloat16()
    noise = torch.randn(inp["latents"].shape, generator=g).bfloat16()
    tid = torch.tensor([371])
    models = {n: getattr(pipe, n) for n in pipe.in_iteration_models}
    loss = pipe.training_loss(global_step=0, timestep_id=tid, noise=noise.cuda(), **models, input_latents=inp["latents"].cuda(),
                              prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                              special_token_mask=inp["special_token_mask"].cuda(), height=H, width=H, edit_latents=inp["edit_latents"].cuda(),
                              pseudo_special_emb_dino=gt_d.cuda(), pseudo_special_emb_vae=gt_v.cuda(), is_train=True)
    pipe.dit.engine().nat.check_async()

    def oracle(dtype):
        s = O.FlowMatchOracle()
        s.set_timesteps(1000, training=True)
        ts = s.timesteps[tid].to(dtype)
        c = lambda x: x.to(dtype)
        x0 = c(inp["latents"])
        lat = s.add_noise(x0, c(noise), ts)
        col = {}
        Wd = {k: v.cpu().to(dtype) for k, v in sd.items()}
        Ad = {k: v.cpu().to(dtype) for k, v in ad.items()}
        with torch.no_grad():
            v = O.model_fn(Wd, Ad, lat.to(dtype), ts, c(inp["prompt_emb"]).clone(), inp["prompt_emb_mask"], inp["special_token_mask"], H, H,
                           edit_latents=c(inp["edit_latents"]), num_layers=2, collect=col)
        fm = torch.nn.functional.mse_loss(v.float(), (c(noise) - x0).float()) * s.training_weight(ts)
        sp = O.adapter_loss(col["dino_pred"], col["vae_pred"], c(gt_d), c(gt_v), ts, 19.999980926513672, 1000.0)
        return fm.item(), sp.float().item()
    fm32, sp32 = oracle(torch.float32)
    fm16, sp16 = oracle(torch.bfloat16)
    tot32, tot16 = fm32 + sp32, fm16 + sp16
    record("training_loss", native=loss.item(), oracle_fp32=tot32, oracle_bf16=tot16, flow_matching_fp32=fm32, adapter_fp32=sp32)
    assert abs(loss.item() - tot32) <= abs(tot16 - tot32) + 1e-2 * abs(tot32), (loss.item(), tot32, tot16)
    assert abs(pipe.special_token_loss - sp32) <= abs(sp16 - sp32) + 1e-2 * abs(sp32)


@gpu
def test_positive_only_request_at_default_cfg_scale_is_refused():
    """cfg_scale != 1 without negative inputs must not combine with an uninitialised negative prediction."""
    pipe, _, _ = device_model(1, seed=1)
    inp = O.synth_inputs(64, 64, 72, seed=6, dtype=torch.bfloat16)
    ip = {k: inp[k].cuda() for k in ("prompt_emb", "prompt_emb_mask", "special_token_mask")}
    with pytest.raises(ValueError, match="negative"):
        pipe.denoise(inp["latents"].cuda(), ip, None, inp["edit_latents"].cuda(), height=64, width=64, num_inference_steps=2)
    lat = pipe.denoise(inp["latents"].cuda(), dict(ip, prompt_emb=ip["prompt_emb"].clone()), None, inp["edit_latents"].cuda(), height=64, width=64,
                       num_inference_steps=2, cfg_scale=1.0)
    assert torch.isfinite(lat.float()).all()


@gpu
def test_in_place_weight_load_drops_the_conditioning_cache():
    """`dit.load_state_dict(sd)` without assign keeps every data_ptr; the timestep-keyed caches must still be dropped."""
    from physicedit_b200.model_fn import model_fn_qwen_image
    pipe, sd, ad = device_model(1, seed=2)
    inp = O.synth_inputs(64, 64, 72, seed=6, dtype=torch.bfloat16)
    t = torch.tensor([500.0]).to(torch.bfloat16)
    kw = dict(dit=pipe.dit, visual_thinking_adapter=None, latents=inp["latents"].cuda(), timestep=t.cuda(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
              special_token_mask=None, height=64, width=64, edit_latents=inp["edit_latents"].cuda(), is_train=False)
    y0, _ = model_fn_qwen_image(prompt_emb=inp["prompt_emb"].cuda(), **kw)
    y0 = y0.clone()
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2["transformer_blocks.0.img_mod.1.weight"] = sd2["transformer_blocks.0.img_mod.1.weight"] * 0.5
    sd2["time_text_embed.timestep_embedder.linear_2.bias"] = sd2["time_text_embed.timestep_embedder.linear_2.bias"] + 0.25
    pipe.dit.load_state_dict(sd2)                                                   # in place, same pointers, same timestep key
    y1, _ = model_fn_qwen_image(prompt_emb=inp["prompt_emb"].cuda(), **kw)
    assert not torch.equal(y0, y1)
    no_sp = dict(inp, special_token_mask=None)
    y32 = oracle_forward(sd2, None, no_sp, t, 64, 64, torch.float32, 1)
    y16 = oracle_forward(sd2, None, no_sp, t, 64, 64, torch.bfloat16, 1)
    assert rel_l2(y1, y32) <= rel_l2(y16, y32) + TOL_EXTRA


@gpu
def test_lora_fold_is_bit_identical_to_the_reference_loader_on_this_gpu():
    """`pipe.load_lora` (qwen_image_physical.py:250-276 -> lora/__init__.py:28-44): the REFERENCE's GeneralLoRALoader folding r=128 LoRA
    into its own QwenImageDiT on this GPU vs the native loader folding the same LoRA into the native DiT AFTER its engine fused the QKV
    weights (the fold must land in the fused buffer through the parameter views): every weight bit-identical, then one forward agrees."""
    from oracle import ref_import
    if ref_import.reference_root() is None:
        pytest.skip("no reference tree on this box")
    import physicedit_b200 as pe
    from physicedit_b200.lora import GeneralLoRALoader
    L = 2
    Wsd = O.synth_weights(O.dit_param_shapes(L), seed=23, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(5)
    targets = ["attn.to_q", "attn.to_k", "attn.to_v", "attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj", "attn.to_out.0", "attn.to_add_out",
               "img_mlp.net.2", "txt_mlp.net.2", "img_mod.1", "txt_mod.1"]                       # scripts/train/train_multigpu.sh:30
    lsd = {}
    for i in range(L):
        for t in targets:
            name = f"transformer_blocks.{i}.{t}"
            out_f, in_f = Wsd[name + ".weight"].shape
            lsd[f"{name}.lora_A.default.weight"] = (torch.randn(128, in_f, generator=g) * 0.02).bfloat16()
            lsd[f"{name}.lora_B.default.weight"] = (torch.randn(out_f, 128, generator=g) * 0.02).bfloat16()
    with ref_import.ReferenceModules() as ref:
        import importlib
        ref_lora = importlib.import_module("diffsynth.lora")
        rdit = ref_import.build_reference_dit(ref, Wsd, L, torch.bfloat16, "cuda")
        ref_lora.GeneralLoRALoader(device="cuda", torch_dtype=torch.bfloat16).load(rdit, lsd, alpha=1.0)
        want = {k: v.detach().clone() for k, v in rdit.state_dict().items()}
    from physicedit_b200.dit import QwenImageDiT
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=L)
    dit.load_state_dict({k: v.clone() for k, v in Wsd.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    dit = dit.to("cuda").eval()
    eng = dit.engine()                                                 # pack first: to_q / to_k / to_v become views of one buffer
    n = GeneralLoRALoader(device="cuda", torch_dtype=torch.bfloat16).load(dit, lsd, alpha=1.0)
    assert n == len(targets) * L
    got = dit.state_dict()
    for k, v in want.items():
        assert torch.equal(got[k], v), k
    assert torch.equal(eng.qkv_w[1][0][:3072], want["transformer_blocks.1.attn.to_q.weight"])      # the fused buffer holds the folded weights
    assert not torch.equal(want["transformer_blocks.0.attn.to_q.weight"].cpu(), Wsd["transformer_blocks.0.attn.to_q.weight"])
