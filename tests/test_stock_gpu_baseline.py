"""Times the reference's stock PyTorch GPU path (the oracle restatement = the same torch ops in the same order as
qwen_image_dit.py / qwen_image_physical.py: F.linear -> cuBLASLt, F.scaled_dot_product_attention, ATen elementwise)
next to the native path on the same B200, same shapes.  Opt-in: PE_STOCK_BASELINE=1 python -m pytest -m gpu -k stock -s
Writes gpurun_out/stock_gpu_baseline.json."""
import json
import os
import time

import pytest
import torch

from oracle import dit_oracle as O


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("PE_STOCK_BASELINE") != "1", reason="opt-in measurement (PE_STOCK_BASELINE=1)")
def test_stock_pytorch_gpu_path_vs_native():
    from physicedit_b200.dit import QwenImageDiT
    NL, H, T = 4, 1024, 512
    dev = "cuda"
    shapes = O.dit_param_shapes(NL)
    g = torch.Generator(device=dev).manual_seed(0)
    W = {}
    for k, shp in shapes.items():
        if len(shp) == 2:
            W[k] = ((torch.rand(shp, generator=g, device=dev) * 2 - 1) / shp[1] ** 0.5).bfloat16()
        elif k.endswith(".bias"):
            W[k] = ((torch.rand(shp, generator=g, device=dev) * 2 - 1) * 0.02).bfloat16()
        else:
            W[k] = torch.ones(shp, device=dev, dtype=torch.bfloat16)
    inp = {k: v.to(dev) for k, v in O.synth_inputs(H, H, T, seed=1, dtype=torch.bfloat16).items()}
    t = torch.tensor([500.0], device=dev).bfloat16()

    def stock():
        # rope tables are cached by the reference (rope_cache) -> built once outside the timed region
        return O.model_fn(W, None, inp["latents"], t, inp["prompt_emb"], inp["prompt_emb_mask"], None, H, H, edit_latents=inp["edit_latents"])

    rope = O.rope_tables([(1, 64, 64), (1, 64, 64)], T)
    rope = (rope[0].to(dev), rope[1].to(dev))
    _orig = O.rope_tables
    O.rope_tables = lambda *a, **k: rope
    try:
        with torch.no_grad():
            for _ in range(2):
                y_stock = stock()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                stock()
            e1.record()
            torch.cuda.synchronize()
            ms_stock = e0.elapsed_time(e1) / 5
    finally:
        O.rope_tables = _orig

    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=NL)
    dit.load_state_dict({k: v.clone() for k, v in W.items()}, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    dit = dit.to(dev).eval()
    eng = dit.engine()
    out = torch.empty_like(inp["latents"])
    lat = [inp["latents"].contiguous(), inp["edit_latents"].contiguous()]
    for _ in range(2):
        eng.forward(lat, t, inp["prompt_emb"][0], out, t_key=None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.forward(lat, t, inp["prompt_emb"][0], out, t_key=None)
    e1.record()
    torch.cuda.synchronize()
    ms_native = e0.elapsed_time(e1) / 5
    eng.nat.check_async()
    err = ((out.float() - y_stock.float()).norm() / y_stock.float().norm()).item()
    res = {"layers": NL, "S": 8192 + T, "stock_pytorch_ms_per_forward": ms_stock, "native_ms_per_forward": ms_native,
           "speedup": ms_stock / ms_native, "stock_ms_per_block": ms_stock / NL, "native_ms_per_block": ms_native / NL,
           "rel_l2_native_vs_stock_bf16": err, "torch": torch.__version__, "note": "4 of 60 blocks at the full 1024^2 sequence; per-block cost is depth-independent"}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/stock_gpu_baseline.json", "w"), indent=1)
    print(json.dumps(res))
    assert err < 3e-2
