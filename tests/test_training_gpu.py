"""Training path: gradients of the native differentiable forward against autograd through the ORACLE.

Protocol (the same as for the forward): the fp32 oracle differentiated by torch autograd is the truth; the same function differentiated in
bf16 with stock torch ops gives the noise floor of a bf16 implementation; the native path (bf16 tcgen05 GEMMs, flash attention forward,
composed attention backward) has to land within `FLOOR_GAIN * floor + GRAD_EXTRA` of the truth, per parameter group.  Un-merged LoRA enters
the oracle as W_eff = W + scaling * B @ A built under autograd -- the same function of (A, B) as PEFT's `base(x) + B(A(x)) * scaling`
(peft is absent here: parity of that layer is unpinned, see physicedit_b200/lora.py).
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import dit_oracle as O

gpu = pytest.mark.gpu
FLOOR_GAIN, GRAD_EXTRA = 1.25, 2e-3
TARGETS = "to_q,to_k,to_v,add_q_proj,add_k_proj,add_v_proj,to_out.0,to_add_out,img_mlp.net.2,img_mod.1,txt_mlp.net.2,txt_mod.1".split(",")   # train_multigpu.sh:30


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def nat():
    from physicedit_b200 import native as nv
    return nv.Native.get(0)


# ---------------------------------------------------------------------------------------------------------------------------
@gpu
@pytest.mark.parametrize("M,N,K,bias", [(77, 128, 3072, False), (300, 3072, 128, False), (64, 10752, 3584, True), (1, 18432, 3072, True),
                                        (1000, 64, 3072, True), (8704, 3072, 3072, True)])
def test_linear_forward_and_backward_on_the_gemm(nat, M, N, K, bias):
    """Y = X W^T + b, dX = dY W, dW = dY^T X, db = column sums of dY: all four from pe_gemm; truth = fp32 matmuls of the same bf16 data."""
    from physicedit_b200 import autograd as ag
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16().requires_grad_()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16().requires_grad_()
    b = (torch.randn(N, device="cuda", generator=g) * 0.1).bfloat16().requires_grad_() if bias else None
    dy = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    launches = nat.launches
    y = ag.linear(x, w, b)
    y.backward(dy)
    assert nat.launches - launches >= 6                                    # forward, dX (+ transpose), dW (+ 2 transposes)
    x32, w32 = x.detach().float().requires_grad_(), w.detach().float().requires_grad_()
    b32 = b.detach().float().requires_grad_() if bias else None
    y32 = F.linear(x32, w32, b32)
    y32.backward(dy.float())
    assert rel_l2(y, y32) < 4e-3
    assert rel_l2(x.grad, x32.grad) < 4e-3 and rel_l2(w.grad, w32.grad) < 4e-3
    if bias:
        assert rel_l2(b.grad, b32.grad) < 4e-3
    nat.check_async()


@gpu
@pytest.mark.parametrize("S,H", [(333, 2), (1024, 3), (2048 + 96, 24)])
def test_attention_backward_composed_from_the_gemm(nat, S, H):
    """dQ, dK, dV of the joint attention: native (flash forward + per-head GEMM / softmax / transpose backward) vs fp32 math attention
    under autograd; floor = the same in bf16 with torch's SDPA."""
    from physicedit_b200 import autograd as ag
    g = torch.Generator(device="cuda").manual_seed(S)
    q, k, v, do = (torch.randn(S, H * 128, device="cuda", generator=g).bfloat16() for _ in range(4))
    k = k * 1.5

    def run(dtype, native):
        qq, kk, vv = (t.detach().to(dtype).requires_grad_() for t in (q, k, v))
        if native:
            o = ag.attention(qq, kk, vv, H)
        else:
            hm = lambda t: t.view(S, H, 128).transpose(0, 1)
            if dtype == torch.float32:
                p = torch.softmax(hm(qq) @ hm(kk).transpose(1, 2) / math.sqrt(128), dim=-1)
                o = (p @ hm(vv)).transpose(0, 1).reshape(S, H * 128)
            else:
                o = F.scaled_dot_product_attention(hm(qq)[None], hm(kk)[None], hm(vv)[None])[0].transpose(0, 1).reshape(S, H * 128)
        o.backward(do.to(dtype))
        return o.detach(), qq.grad, kk.grad, vv.grad
    truth, floor, got = run(torch.float32, False), run(torch.bfloat16, False), run(torch.bfloat16, True)
    for name, t, f, n in zip(("o", "dq", "dk", "dv"), truth, floor, got):
        e_f, e_n = rel_l2(f, t), rel_l2(n, t)
        print(f"attention S={S} H={H} {name}: native {e_n:.3e} floor {e_f:.3e}")
        assert e_n <= FLOOR_GAIN * e_f + 2e-3, (name, e_n, e_f)
    nat.check_async()


# ---------------------------------------------------------------------------------------------------------------------------
def _native_training_pipe(layers, seed, rank):
    from test_parity_depth_gpu import device_model
    from physicedit_b200.trainers import DiffusionTrainingModule
    pipe, sd, ad = device_model(layers, seed)
    tm = DiffusionTrainingModule()
    pipe.scheduler.set_timesteps(1000, training=True)
    pipe.freeze_except(["visual_thinking_adapter"])
    pipe.dit = tm.add_lora_to_model(pipe.dit, target_modules=TARGETS, lora_rank=rank, upcast_dtype=torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    lora = {}
    for name, p in pipe.dit.named_parameters():
        if "lora_" in name:                                                # PEFT starts B at zero; a trained state has both factors alive
            with torch.no_grad():
                p.copy_((torch.randn(p.shape, device="cuda", generator=g) * (0.5 / math.sqrt(p.shape[1]))).bfloat16())
            lora[name] = p
    return pipe, sd, ad, lora


def _oracle_grads(sd, ad, lora_vals, inp, t, noisy, target, gt, weight, dtype, t_min, t_max):
    """Autograd through the oracle's model_fn with W_eff = W + B @ A, adapter weights as leaves; returns (loss, grads by native parameter name)."""
    dev = "cuda"
    W = {k: v.to(dev, dtype) for k, v in sd.items()}
    A = {k: v.to(dev, dtype).requires_grad_() for k, v in ad.items()}
    leaves = {k: v.detach().to(dev, dtype).requires_grad_() for k, v in lora_vals.items()}
    for name in [n for n in leaves if ".lora_A." in n]:
        mod = name.split(".lora_A.")[0]
        W[mod + ".weight"] = W[mod + ".weight"] + leaves[mod + ".lora_B.default.weight"] @ leaves[name]
    c = lambda x: x.to(dev, dtype) if x.is_floating_point() else x.to(dev)
    noisy, target = c(noisy), c(target)               # FlowMatchScheduler.add_noise / training_target (flow_match.py:84-96), evaluated once in bf16
    pe = c(inp["prompt_emb"]).clone()
    col = {}
    # the timestep BOOKKEEPING of the bf16 path (t -> bf16 -> /1000 -> bf16, bf16 frequencies, bf16 alpha: bit-exact by contract) also
    # feeds the fp32 run -- otherwise a phase shift of the top sinusoid components makes the two precisions different FUNCTIONS of the weights
    tt = t.to(dev, torch.bfloat16)
    pred = O.model_fn(W, A, noisy, tt, pe, c(inp["prompt_emb_mask"]), c(inp["special_token_mask"]), inp["H"], inp["H"],
                      edit_latents=c(inp["edit_latents"]), t_min=t_min, t_max=t_max, cuda_scalar_div=True, collect=col)
    loss = F.mse_loss(pred.float(), target.float()) * weight
    loss = loss + O.adapter_loss(col["dino_pred"], col["vae_pred"], c(gt[0]), c(gt[1]), tt, t_min, t_max)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    grads.update({"adapter." + k: v.grad for k, v in A.items()})
    return loss.item(), grads


@gpu
def test_training_loss_gradients_match_autograd_through_the_oracle():
    """pipe.training_loss(...).backward() (qwen_image_physical.py:313-329; train_physicedit.py:648-652) with LoRA r = 128 un-merged on the 12 target
    linears of every block and the dual adapter trainable: loss value and every gradient group vs the fp32 oracle."""
    layers, H, T, rank = 2, 128, 88, 128
    pipe, sd, ad, lora = _native_training_pipe(layers, seed=5, rank=rank)
    inp = O.synth_inputs(H, H, T, seed=41, dtype=torch.bfloat16)
    inp["H"] = H
    torch.manual_seed(3)
    gt = (torch.randn(1, 64, 3584).bfloat16(), torch.randn(1, 64, 3584).bfloat16())
    noise = torch.randn(1, 16, H // 8, H // 8).bfloat16()
    tid = torch.tensor([400])
    t = pipe.scheduler.timesteps[tid].to(torch.bfloat16)
    weight = float(pipe.scheduler.training_weight(t))
    noisy = pipe.scheduler.add_noise(inp["latents"].cuda(), noise.cuda(), t.cuda())
    target = pipe.scheduler.training_target(inp["latents"].cuda(), noise.cuda(), t.cuda())
    adp = pipe.visual_thinking_adapter
    from physicedit_b200 import native as nv
    l0 = nv.Native.get(0).launches
    loss = pipe.training_loss(input_latents=inp["latents"].cuda(), prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                              special_token_mask=inp["special_token_mask"].cuda(), height=H, width=H, edit_latents=inp["edit_latents"].cuda(),
                              pseudo_special_emb_dino=gt[0].cuda(), pseudo_special_emb_vae=gt[1].cuda(), is_train=True, use_gradient_checkpointing=True,
                              timestep_id=tid, noise=noise.cuda())
    loss.backward()
    n_launch = nv.Native.get(0).launches - l0
    nv.Native.get(0).check_async()
    assert n_launch > layers * 200, n_launch                               # forward + recompute + backward really ran on libpe_b200
    native = {k: p.grad for k, p in lora.items()}
    native.update({"adapter." + k: p.grad for k, p in adp.named_parameters()})
    # the text stream of the LAST block is dead after its attention (only image tokens reach proj_out, :1398-1400): those factors get no
    # gradient in the reference either (hence --find_unused_parameters, train_multigpu.sh:36)
    dead = [k for k, g in native.items() if g is None]
    assert all(k.startswith(f"transformer_blocks.{layers - 1}.") and ("to_add_out" in k or "txt_mlp" in k) for k in dead), dead
    native = {k: (g if g is not None else torch.zeros_like(lora[k])) for k, g in native.items()}
    assert all(torch.isfinite(g.float()).all() for g in native.values())
    lora_vals = {k: p.detach() for k, p in lora.items()}
    l32, g32 = _oracle_grads(sd, ad, lora_vals, inp, t, noisy, target, gt, weight, torch.float32, adp.t_min, adp.t_max)
    l16, g16 = _oracle_grads(sd, ad, lora_vals, inp, t, noisy, target, gt, weight, torch.bfloat16, adp.t_min, adp.t_max)
    print(f"\ntraining loss: native {loss.item():.6f} oracle fp32 {l32:.6f} oracle bf16 {l16:.6f}; launches {n_launch}")
    assert abs(loss.item() - l32) <= 1.5 * abs(l16 - l32) + 2e-3 * abs(l32)
    groups = {"lora_A": [k for k in native if ".lora_A." in k], "lora_B": [k for k in native if ".lora_B." in k],
              "mod_lora": [k for k in native if "_mod.1.lora" in k], "adapter_w": [k for k in native if k.startswith("adapter.") and k.endswith("weight")],
              "adapter_b": [k for k in native if k.startswith("adapter.") and k.endswith("bias")]}
    report = {}
    for gname, keys in groups.items():
        cat = lambda d: torch.cat([d[k].float().flatten().cpu() for k in keys])
        e_n, e_f = rel_l2(cat(native), cat(g32)), rel_l2(cat(g16), cat(g32))
        report[gname] = (e_n, e_f)
        print(f"  grad {gname:10s} ({len(keys):3d} tensors): native vs fp32 {e_n:.3e}; bf16 floor {e_f:.3e}")
    for gname, (e_n, e_f) in report.items():
        assert e_n <= FLOOR_GAIN * e_f + GRAD_EXTRA, (gname, e_n, e_f)
    # an optimizer step on these gradients changes the loss in the right direction
    opt = torch.optim.SGD([p for p in list(lora.values()) + list(adp.parameters())], lr=2e-2)
    opt.step()
    loss2 = pipe.training_loss(input_latents=inp["latents"].cuda(), prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                               special_token_mask=inp["special_token_mask"].cuda(), height=H, width=H, edit_latents=inp["edit_latents"].cuda(),
                               pseudo_special_emb_dino=gt[0].cuda(), pseudo_special_emb_vae=gt[1].cuda(), is_train=True, use_gradient_checkpointing=False,
                               timestep_id=tid, noise=noise.cuda())
    assert loss2.item() < loss.item(), (loss2.item(), loss.item())


@gpu
def test_unmerged_lora_forward_equals_the_folded_engine_within_rounding():
    """With LoRA injected, model_fn runs the un-merged path (also under no_grad: the reference evaluates with PEFT layers in place,
    train_physicedit.py:665); merging the factors and running the inference engine must give the same velocity up to bf16 rounding."""
    from physicedit_b200.lora import merge_lora
    from physicedit_b200.model_fn import model_fn_qwen_image
    pipe, sd, ad, lora = _native_training_pipe(2, seed=6, rank=128)
    H, T = 128, 72
    inp = O.synth_inputs(H, H, T, seed=42, dtype=torch.bfloat16)
    t = torch.tensor([603.0]).bfloat16().cuda()
    kw = lambda: dict(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"].cuda(), timestep=t,
                      prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                      special_token_mask=inp["special_token_mask"].cuda(), height=H, width=H, edit_latents=inp["edit_latents"].cuda(), is_train=False)
    with torch.no_grad():
        v_unmerged, _ = model_fn_qwen_image(**kw())
        merge_lora(pipe.dit)
        assert not any("lora_" in n for n, _ in pipe.dit.named_parameters())
        v_merged, _ = model_fn_qwen_image(**kw())
    e = rel_l2(v_unmerged, v_merged)
    print(f"\nun-merged LoRA forward vs folded engine: {e:.3e}")
    assert e < 1.5e-2


@gpu
def test_resampler_stack_gradients():
    """The pseudo-target branch (PerceiverResampler + VisualThinkingAdapter + time embeddings, helpers.py:21-121) under autograd on the native
    GEMMs vs the same modules evaluated with stock torch ops in fp32."""
    from physicedit_b200 import autograd as ag
    from physicedit_b200.adapters import PerceiverResampler, VisualThinkingAdapter
    torch.manual_seed(11)
    rs = PerceiverResampler(dim=64, depth=2, dim_head=64, heads=8, num_latents=64, max_num_media_tokens=2048).cuda().bfloat16()
    adp = VisualThinkingAdapter(64, 3584).cuda().bfloat16()
    x = torch.randn(1, 3 * 391, 64, device="cuda").bfloat16()              # 1173 media tokens: not a multiple of 8
    gt = torch.randn(1, 64, 3584, device="cuda").bfloat16()
    params = list(rs.parameters()) + list(adp.parameters())
    loss = F.mse_loss(adp(rs(x)).float(), gt.float())
    loss.backward()
    got = [p.grad.clone() for p in params]

    def stock(dtype):
        import copy
        r2, a2 = copy.deepcopy(rs).to(dtype), copy.deepcopy(adp).to(dtype)
        for p in list(r2.parameters()) + list(a2.parameters()):
            p.grad = None
        xm = x[0].to(dtype) + r2.pos_emb.weight[:x.shape[1]]
        lat = r2.latents
        for attn, ff in r2.layers:
            xn = F.layer_norm(xm, (64,), attn.norm_media.weight, attn.norm_media.bias)
            ln = F.layer_norm(lat, (64,), attn.norm_latents.weight, attn.norm_latents.bias)
            q = F.linear(ln, attn.to_q.weight).view(-1, 8, 64).transpose(0, 1)
            k, v = F.linear(torch.cat((xn, ln)), attn.to_kv.weight).chunk(2, dim=-1)
            k, v = k.view(-1, 8, 64).transpose(0, 1), v.view(-1, 8, 64).transpose(0, 1)
            dots = q @ k.transpose(1, 2) * attn.scale
            p = (dots - dots.amax(dim=-1, keepdim=True).detach()).softmax(dim=-1)
            lat = lat + F.linear((p @ v).transpose(0, 1).reshape(-1, 512), attn.to_out.weight)
            lat = lat + ff.net(lat)
        out = a2.net(F.layer_norm(lat, (64,), r2.norm.weight, r2.norm.bias)[None])
        F.mse_loss(out.float(), gt.float()).backward()
        return [p.grad for p in list(r2.parameters()) + list(a2.parameters())]
    t32, t16 = stock(torch.float32), stock(torch.bfloat16)
    cat = lambda gs: torch.cat([g.float().flatten().cpu() for g in gs])
    e_n, e_f = rel_l2(cat(got), cat(t32)), rel_l2(cat(t16), cat(t32))
    print(f"\nresampler stack grads: native {e_n:.3e} floor {e_f:.3e}")
    assert e_n <= FLOOR_GAIN * e_f + GRAD_EXTRA


class _Samples(torch.utils.data.Dataset):
    """Pre-computed unit outputs of n requests (what pipe.unit_runner hands training_loss after the frozen encoders ran)."""
    load_from_cache = False

    def __init__(self, n, H, T):
        self.items = []
        for i in range(n):
            inp = O.synth_inputs(H, H, T, seed=300 + i, dtype=torch.bfloat16)
            g = torch.Generator().manual_seed(900 + i)
            self.items.append(dict(input_latents=inp["latents"], prompt_emb=inp["prompt_emb"], prompt_emb_mask=inp["prompt_emb_mask"],
                                   special_token_mask=inp["special_token_mask"], edit_latents=inp["edit_latents"], height=H, width=H,
                                   pseudo_special_emb_dino=torch.randn(1, 64, 3584, generator=g).bfloat16(),
                                   pseudo_special_emb_vae=torch.randn(1, 64, 3584, generator=g).bfloat16(),
                                   noise=torch.randn(1, 16, H // 8, H // 8, generator=g).bfloat16(), timestep_id=torch.tensor([200 + 150 * i])))

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def _training_module(pipe):
    from physicedit_b200.trainers import DiffusionTrainingModule

    class Module(DiffusionTrainingModule):
        """The shape of scripts/train/train_physicedit.py's QwenImageTrainingModule.forward (:297-311): units' outputs -> pipe.training_loss."""

        def __init__(self):
            super().__init__()
            self.pipe = pipe
            self.losses = []

        def forward(self, data, inputs=None):
            d = self.transfer_data_to_device(dict(data), "cuda")
            d["prompt_emb"] = d["prompt_emb"].clone()
            loss = self.pipe.training_loss(**d, is_train=True, use_gradient_checkpointing=True)
            self.losses.append(float(loss.detach()))
            return loss
    return Module()


@gpu
def test_launch_training_task_trains_and_writes_the_reference_checkpoint_layout(tmp_path):
    """The optimizer loop (trainers/utils.py:932-977) on the native backward: AdamW over LoRA + adapter, 3 epochs over 2 fixed samples; the loss on
    those samples goes down and the checkpoint holds exactly the trainable keys in the layout validate.py:44-65 splits."""
    from physicedit_b200.trainers import ModelLogger, launch_training_task
    from safetensors.torch import load_file
    pipe, sd, ad, lora = _native_training_pipe(2, seed=7, rank=16)
    model = _training_module(pipe)
    logger = ModelLogger(str(tmp_path), remove_prefix_in_ckpt="pipe.dit.")
    launch_training_task(_Samples(2, 64, 80), model, logger, learning_rate=2e-4, weight_decay=0.0, num_workers=0, num_epochs=3)
    first, last = sum(model.losses[:2]), sum(model.losses[-2:])
    print(f"\ntraining: loss over the 2 samples {first:.4f} -> {last:.4f} after 3 epochs of AdamW steps")
    assert last < first
    ck = load_file(str(tmp_path / "epoch-2.safetensors"))
    assert any(k.startswith("transformer_blocks.0.attn.to_q.lora_A.default") for k in ck) and any(k.startswith("pipe.visual_thinking_adapter.head_dino.0") for k in ck)
    assert len(ck) == 2 * 12 * 2 + 8 and not any("base_layer" in k for k in ck)


@gpu
@pytest.mark.parametrize("batch,M,N,K", [(3, 200, 136, 128), (24, 333, 336, 128), (5, 300, 128, 304)])
def test_batched_gemm_matches_per_problem_products(nat, batch, M, N, K):
    """pe_gemm_batched: every problem of the launch equals its own product (tiles that overhang a problem read the neighbour's rows: they must
    never be stored), for the plain and the fp32 epilogue and the two attention-backward epilogues in row and column mode."""
    from physicedit_b200 import native as nv
    g = torch.Generator(device="cuda").manual_seed(batch * 1000 + M)
    Mp = (M + 7) // 8 * 8
    a = torch.randn(batch, Mp, K, device="cuda", generator=g).bfloat16()
    w = torch.randn(batch, N, K, device="cuda", generator=g).bfloat16() / math.sqrt(K)
    ref = torch.einsum("bmk,bnk->bmn", a.float(), w.float())[:, :M]
    flat = lambda t: t.reshape(-1, t.shape[-1])
    out = torch.full((batch, Mp, N), 7.0, device="cuda").bfloat16()
    nat.gemm_batched(flat(a), flat(w), flat(out), batch=batch, M=M, N=N, K=K, a_batch_rows=Mp, w_batch_rows=N, out_batch_rows=Mp)
    assert rel_l2(out[:, :M], ref) < 4e-3 and (out[:, M:] == 7.0).all()                       # rows >= M of every problem untouched
    o32 = torch.zeros(batch, Mp, N, device="cuda")
    nat.gemm_batched(flat(a), flat(w), flat(o32), batch=batch, M=M, N=N, K=K, a_batch_rows=Mp, w_batch_rows=N, out_batch_rows=Mp, epilogue=nv.EPI_F32)
    assert rel_l2(o32[:, :M], ref) < 1e-5
    for per_col in (False, True):
        vec = torch.randn(batch, max(Mp, N), device="cuda", generator=g)
        st = vec[:, None, :N] if per_col else vec[:, :M, None]
        p = torch.empty(batch, Mp, N, device="cuda").bfloat16()
        kw = dict(batch=batch, M=M, N=N, K=K, a_batch_rows=Mp, w_batch_rows=N, out_batch_rows=Mp, vec=vec, vec_batch_stride=vec.shape[1], vec_per_column=per_col)
        nat.gemm_batched(flat(a), flat(w), flat(p), epilogue=nv.EPI_ATTN_P, alpha=0.7, **kw)
        p_ref = torch.exp2(ref * 0.7 - st)
        assert rel_l2(p[:, :M], p_ref) < 4e-3
        ds = p.clone()
        nat.gemm_batched(flat(a), flat(w), flat(ds), epilogue=nv.EPI_ATTN_DS, alpha=0.25, **kw)
        assert rel_l2(ds[:, :M], p[:, :M].float() * (ref - st) * 0.25) < 4e-3
    nat.check_async()


@gpu
def test_attention_forward_row_statistics(nat):
    """pe_attention_fwd_lse: same output as pe_attention_fwd, and lse = log2 sum_j exp2(scale log2(e) s_j) per (head, row)."""
    S, H = 1000, 3
    g = torch.Generator(device="cuda").manual_seed(1)
    q, k, v = (torch.randn(S, H * 128, device="cuda", generator=g).bfloat16() for _ in range(3))
    o1, o2, lse = torch.empty_like(q), torch.empty_like(q), torch.empty(H, S, device="cuda")
    nat.attention(q, k, v, o1, H, 1 / math.sqrt(128))
    nat.attention_lse(q, k, v, o2, lse, H, 1 / math.sqrt(128))
    assert torch.equal(o1, o2)
    hm = lambda t: t.view(S, H, 128).transpose(0, 1).float()
    ref = torch.logsumexp(hm(q) @ hm(k).transpose(1, 2) / math.sqrt(128), dim=-1) * 1.4426950408889634
    assert (lse - ref).abs().max().item() < 2e-3


@gpu
def test_hot_loaded_lora_equals_the_folded_lora_within_rounding():
    """pipe.enable_lora_magic() + load_lora(hotload=True) (un-merged: out + x A^T B^T, vram_management/layers.py:177-179) vs load_lora() (folded) on
    the same factors; clear_lora() brings back the plain model bit for bit (and the inference engine)."""
    from test_parity_depth_gpu import device_model
    from physicedit_b200.model_fn import model_fn_qwen_image
    H, T = 128, 72
    inp = O.synth_inputs(H, H, T, seed=44, dtype=torch.bfloat16)
    t = torch.tensor([603.0]).bfloat16().cuda()

    def run(pipe):
        with torch.no_grad():
            return model_fn_qwen_image(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"].cuda(), timestep=t,
                                       prompt_emb=inp["prompt_emb"].cuda().clone(), prompt_emb_mask=inp["prompt_emb_mask"].cuda(),
                                       special_token_mask=inp["special_token_mask"].cuda(), height=H, width=H, edit_latents=inp["edit_latents"].cuda(), is_train=False)[0]
    g = torch.Generator().manual_seed(2)
    lora = {}
    for blk in range(2):
        for name, (o, i) in (("attn.to_q", (3072, 3072)), ("attn.add_v_proj", (3072, 3072)), ("img_mlp.net.2", (3072, 12288)), ("txt_mod.1", (18432, 3072))):
            lora[f"transformer_blocks.{blk}.{name}.lora_A.default.weight"] = (torch.randn(16, i, generator=g) / math.sqrt(i)).bfloat16()
            lora[f"transformer_blocks.{blk}.{name}.lora_B.default.weight"] = (torch.randn(o, 16, generator=g) * 0.2).bfloat16()
    pipe, _, _ = device_model(2, seed=12)
    plain = run(pipe)
    pipe.enable_lora_magic()
    assert torch.equal(run(pipe), plain)                                  # wrappers without LoRAs: still the engine, same bits
    pipe.load_lora(pipe.dit, state_dict=lora, alpha=0.7, hotload=True)
    hot = run(pipe)
    pipe.clear_lora()
    assert torch.equal(run(pipe), plain)
    pipe2, _, _ = device_model(2, seed=12)
    pipe2.load_lora(pipe2.dit, state_dict=lora, alpha=0.7)
    folded = run(pipe2)
    e, d = rel_l2(hot, folded), rel_l2(hot, plain)
    print(f"\nhot-loaded vs folded LoRA: {e:.3e}; LoRA effect {d:.3e}")
    assert e < 1.5e-2 and d > 5 * e


@gpu
def test_pseudo_targets_carry_gradients_to_the_resampler_stack(nat):
    """`pipe.physical_visual_embeddings` as the train script reaches it (units -> QwenImageUnit_PhysicalVisualEmbedder, grad mode on, the resampler
    stack in `trainable_models`, train_multigpu.sh:38): the regression targets carry a graph -- the adapter loss trains the resamplers through them
    (qwen_image_physical.py:1057-1118 has no no_grad) -- and equal the inference-mode values."""
    from oracle import aux_oracle as AO
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    torch.manual_seed(5)
    pipe = QwenImagePhysicPipeline(device="cuda", torch_dtype=torch.bfloat16, dinov2_config=dict(hidden=768, layers=2, heads=12))
    pipe.to("cuda")
    stack = ["dino_time_embed", "dino_resampler", "dino_resampler_adapter", "vae_time_embed", "vae_resampler", "vae_resampler_adapter"]
    ain = {k: v.cuda() for k, v in AO.aux_inputs(seed=8, n_mid=2, lat_hw=(16, 16), dtype=torch.bfloat16).items()}
    pipe.freeze_except([])
    with torch.no_grad():
        want = pipe.physical_visual_embeddings(**ain)
    assert not want["pseudo_special_emb_dino"].requires_grad
    frozen = pipe.physical_visual_embeddings(**ain)                      # grad mode on but nothing trainable: still the inference path
    assert not frozen["pseudo_special_emb_vae"].requires_grad and torch.equal(frozen["pseudo_special_emb_vae"], want["pseudo_special_emb_vae"])
    pipe.freeze_except(stack)
    got = pipe.physical_visual_embeddings(**ain)
    for k in ("pseudo_special_emb_dino", "pseudo_special_emb_vae"):
        assert got[k].requires_grad and got[k].shape == (1, 64, 3584)
        assert rel_l2(got[k], want[k]) < 2e-2, (k, rel_l2(got[k], want[k]))
    gt = torch.randn(1, 64, 3584, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    loss = F.mse_loss(got["pseudo_special_emb_dino"].float(), gt) + F.mse_loss(got["pseudo_special_emb_vae"].float(), gt)
    loss.backward()
    nat.check_async()
    for name in stack:
        grads = [p.grad for p in getattr(pipe, name).parameters()]
        assert all(g is not None and torch.isfinite(g.float()).all() for g in grads), name
        assert any(g.float().abs().sum() > 0 for g in grads), name
    assert all(p.grad is None for p in pipe.dinov2.parameters())
