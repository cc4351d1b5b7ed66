#!/usr/bin/env python
"""Headline benchmark: denoise steps/s of a 1024x1024 Qwen-Image-Edit / PhysicEdit edit (BASELINE.json).

One "step" = one denoise step of the reference loop (qwen_image_physical.py:648-661): TWO DiT forwards
(CFG 4.0: posi T=512, nega T=288; 4096 noise + 4096 edit-image tokens, 60 blocks, adapter on the 64 special
tokens) + CFG combine + Euler update.  Weights are random-init bf16 of the real architecture, inputs synthetic.

    python bench.py --gpus 1 --steps K --warmup W            # this framework (CUDA path through libpe_b200.so)
    python bench.py --impl reference ...                      # the reference algorithm on the host CPU (oracle port)

Under torchrun (--gpus N) every rank denoises its own image (one image per GPU, no per-step collective); value is
the whole-job steps/s, timed on the device with CUDA events between barriers, max over ranks.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM, LAYERS = 3072, 60
T_POSI, T_NEGA = 512, 288


def workload_config(resolution, layers):
    """The `config` object both arms print (identical dicts: the driver compares them).  Run-dependent figures go to `derived`."""
    return {"workload": f"{resolution}x{resolution} single-image edit, 50-step schedule, bf16, {layers} blocks, 4096 edit tokens, T=512/288, one image per GPU",
            "l2": "inputs larger than L2: 40.8 GB of weights stream per forward"}


def flops_forward(S_img, T, layers=LAYERS):
    S = S_img + T
    return layers * (24 * DIM * DIM * S + 4 * S * S * DIM + 24 * DIM * DIM) + 2 * S_img * 64 * DIM * 2 + 2 * T * 3584 * DIM


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def ncu_traffic_bytes(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu captures (profiles/r01_ncu_full_metrics.json for
    the attention and the MLP up-projection with a flushed L2, profiles/r02_gemm_raster_ab.json / r01_gemm_traffic.json for the other GEMM shapes)."""
    for name in ("r02_gemm_raster_ab.json", "r01_gemm_traffic.json"):      # r2: the down-projection's rasterisation changed, re-captured
        try:
            g = json.load(open(os.path.join(ROOT, "profiles", name))).get(kernel_substr)
            if g:
                return int((g["dram_read_MB"] + g["dram_write_MB"]) * 1e6)
        except (OSError, KeyError, ValueError, TypeError):
            pass
    try:        # r2 full capture of the attention kernel (profiles/r02_ncu_attention_full.json)
        if "attention_kernel" in kernel_substr:
            row = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_attention_full.json")))[0]
            return int((float(row["dram__bytes_read.sum"].split()[0]) + float(row["dram__bytes_write.sum"].split()[0])) * 1e6)
    except (OSError, KeyError, ValueError, IndexError):
        pass
    try:
        for row in json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_metrics.json"))):
            if kernel_substr in row["kernel"]:
                rd = float(row["dram__bytes_read.sum"].split()[0]) * 1e6
                wr = float(row["dram__bytes_write.sum"].split()[0]) * 1e6
                return int(rd + wr)
    except (OSError, KeyError, ValueError):
        pass
    return None


def build_model(device, layers, seed=0):
    """Random-init bf16 weights of the real architecture, generated on the device (uniform, PyTorch-default bound)."""
    from physicedit_b200.dit import QwenImageDiT
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.device("meta"):
        dit = QwenImageDiT(num_layers=layers)
    sd = {}
    for k, v in dit.state_dict().items():
        if v.dim() == 2:
            b = 1.0 / math.sqrt(v.shape[1])
            t = (torch.rand(v.shape, generator=g, device=device, dtype=torch.float32) * 2 - 1) * b
        elif k.endswith(".bias"):
            t = (torch.rand(v.shape, generator=g, device=device, dtype=torch.float32) * 2 - 1) * 0.02
        else:
            t = torch.ones(v.shape, device=device)
        sd[k] = t.to(torch.bfloat16)
    dit.load_state_dict(sd, assign=True)
    dit.pos_embed = type(dit.pos_embed)(theta=10000, axes_dim=[16, 56, 56], scale_rope=True)
    for i, b in enumerate(dit.transformer_blocks):
        object.__setattr__(b, "_owner", (dit, i))
    dit.eval()
    pipe = QwenImagePhysicPipeline(device=device, torch_dtype=torch.bfloat16, build_training_path=False)
    pipe.dit = dit
    for p in pipe.visual_thinking_adapter.parameters():
        b = 1.0 / math.sqrt(p.shape[-1]) if p.dim() == 2 else 0.02
        p.data = ((torch.rand(p.shape, generator=g, device=device, dtype=torch.float32) * 2 - 1) * b).to(torch.bfloat16)
    return pipe


def synth_inputs(height, width, T, seed, edit_hw=(1024, 1024), n_special=64):
    """Synthetic request in the shapes of SURVEY 8d (self-contained: the native arm must not touch oracle/): latents via the
    reference's generate_noise recipe (CPU generator, fp32 randn, cast to bf16), prompt_emb ~ 3 N(0,1), all-ones mask, special-token
    mask = 64 consecutive rows ending 5 before the end."""
    g = torch.Generator("cpu").manual_seed(seed)
    latents = torch.randn((1, 16, height // 8, width // 8), generator=g, dtype=torch.float32).to(torch.bfloat16)
    edit = torch.randn((1, 16, edit_hw[0] // 8, edit_hw[1] // 8), generator=g, dtype=torch.float32).to(torch.bfloat16)
    prompt = (3 * torch.randn((1, T, 3584), generator=g, dtype=torch.float32)).to(torch.bfloat16)
    mask = torch.ones((1, T), dtype=torch.int64)
    special = torch.zeros((1, T), dtype=torch.bool)
    special[0, T - 5 - n_special: T - 5] = True
    return dict(latents=latents, edit_latents=edit, prompt_emb=prompt, prompt_emb_mask=mask, special_token_mask=special)


def host_inputs(height, width, seed):
    """Pinned host buffers of one edit request (what a serving front-end would hand over)."""
    posi = synth_inputs(height, width, T_POSI, seed=seed)
    nega = synth_inputs(height, width, T_NEGA, seed=seed + 1)
    host = dict(latents=posi["latents"], edit_latents=posi["edit_latents"], pe_posi=posi["prompt_emb"], pe_nega=nega["prompt_emb"],
                mask_posi=posi["prompt_emb_mask"], mask_nega=nega["prompt_emb_mask"], sp_posi=posi["special_token_mask"], sp_nega=nega["special_token_mask"])
    return {k: v.pin_memory() for k, v in host.items()}


def run_native(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from physicedit_b200 import native as nv
    from physicedit_b200 import parallel
    nat = nv.Native.get(local)            # raises NativeUnavailable when libpe_b200.so / an sm_100 GPU is missing
    H = W = args.resolution
    pipe = build_model(device, args.layers, seed=0)
    coll = None
    if world > 1:
        # the one start-up collective: rank 0's weights to every GPU over NVLink (timed twice: the first call also builds NCCL's
        # channels, the second is the steady-state broadcast rate)
        coll = {}
        for tag in ("first_call", "steady"):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            nbytes = parallel.broadcast_weights(pipe, src=0)
            b1.record()
            torch.cuda.synchronize()
            bt = torch.tensor([b0.elapsed_time(b1)], device=device, dtype=torch.float64)
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
            coll[f"weight_broadcast_{tag}"] = {"seconds": round(bt.item() * 1e-3, 4), "bytes": nbytes, "GB_per_s": round(nbytes / (bt.item() * 1e-3) / 1e9, 1)}
    eng = pipe.dit.engine()
    eng.use_cta_pair = not args.no_cta_pair
    eng.attn_flags = args.attn_flags
    pipe.cfg_streams = args.cfg_streams
    n_images = world
    if args.cfg_parallel:
        # latency mode: one image per PAIR of GPUs, the two CFG branches of a step on the two ranks of the pair
        grp, pair, n_images = parallel.make_cfg_pairs()
        pipe.cfg_parallel_group = grp
    host = host_inputs(H, W, seed=100 + (rank // 2 if args.cfg_parallel else rank))             # a different image per rank (or per pair)
    dev = {k: v.to(device, non_blocking=True) for k, v in host.items()}
    # txt_len / n_special: request metadata the host already knows (prompt_lengths() would read them back from the masks once)
    ip = dict(prompt_emb=dev["pe_posi"], prompt_emb_mask=dev["mask_posi"], special_token_mask=dev["sp_posi"], txt_len=T_POSI, n_special=64)
    in_ = dict(prompt_emb=dev["pe_nega"], prompt_emb_mask=dev["mask_nega"], special_token_mask=dev["sp_nega"], txt_len=T_NEGA, n_special=64)
    sched = pipe.scheduler
    sched.set_timesteps(50, dynamic_shift_len=(H // 16) * (W // 16))
    ts_dev = sched.timesteps.to(torch.bfloat16).to(device)
    vbuf = torch.empty((2,) + tuple(dev["latents"].shape), dtype=dev["latents"].dtype, device=device)
    vp, vn = vbuf[0], vbuf[1]
    lat = dev["latents"].clone()

    def one_step(i, latents):
        pid = i % 50
        t_host = float(sched.timesteps[pid].to(torch.bfloat16))
        kw = dict(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=latents, timestep=ts_dev[pid:pid + 1], height=H, width=W,
                  edit_latents=dev["edit_latents"], is_train=False, timestep_host=t_host)
        pipe.run_cfg_branches(kw, ip, in_, vp, vn, ts_dev[pid:pid + 1], t_host)
        nat.cfg_euler_step(latents, vp, vn, 4.0, float(sched.dsigma(sched.timesteps[pid])))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i, lat)
    nat.check_async()
    # ---- device-resident timing (value) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = nat.launches
    # Per-launch CUDA events are taken in the timed region itself when the branches run on one stream.  With two streams the launches
    # of the two branches overlap on the GPU and a per-launch duration is ambiguous, so the attribution (roofline, kernel shares) then
    # comes from `attr_steps` extra single-stream steps run right after the timed region (same process, same clocks, same inputs).
    events_in_region = not args.no_kernel_events and (args.cfg_streams == 1 or args.cfg_parallel)
    nat.prof = {} if events_in_region else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    # timestep-only conditioning for the steps of this window, inside the timed region (pipe.denoise does the same for a whole
    # schedule): batch-K GEMVs stream the modulation weights once per <= 8 steps
    pids = [(args.warmup + i) % 50 for i in range(args.steps)]
    eng.precompute_conditioning(ts_dev[pids].contiguous(), [float(sched.timesteps[pid].to(torch.bfloat16)) for pid in pids])
    for i in range(args.steps):
        one_step(args.warmup + i, lat)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = nat.launches - l0
    prof = nat.profile_summary()
    nat.prof = None
    ms_attr, attr_steps = ms, args.steps
    if not args.no_kernel_events and not events_in_region:
        attr_steps = 2
        pipe.cfg_streams = 1
        nat.prof = {}
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        apids = [(args.warmup + args.steps + i) % 50 for i in range(attr_steps)]
        eng.precompute_conditioning(ts_dev[apids].contiguous(), [float(sched.timesteps[pid].to(torch.bfloat16)) for pid in apids])
        for i in range(attr_steps):
            one_step(args.warmup + args.steps + i, lat)
        a1.record()
        barrier()
        ms_attr = a0.elapsed_time(a1)
        prof = nat.profile_summary()
        nat.prof = None
        pipe.cfg_streams = args.cfg_streams
    nat.check_async()
    finite = bool(torch.isfinite(lat.float()).all().item())

    # ---- end to end through the public API with host buffers (e2e) ----
    h_lat = host["latents"].clone().pin_memory()
    d_in = {k: torch.empty_like(v, device=device) for k, v in host.items()}
    e2e_steps = max(2, min(args.steps, 4))
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(e2e_steps):
        for k in ("latents", "edit_latents", "pe_posi", "pe_nega", "mask_posi", "mask_nega", "sp_posi", "sp_nega"):
            d_in[k].copy_(h_lat if k == "latents" else host[k], non_blocking=True)
        ipe = dict(prompt_emb=d_in["pe_posi"], prompt_emb_mask=d_in["mask_posi"], special_token_mask=d_in["sp_posi"], txt_len=T_POSI, n_special=64)
        ine = dict(prompt_emb=d_in["pe_nega"], prompt_emb_mask=d_in["mask_nega"], special_token_mask=d_in["sp_nega"], txt_len=T_NEGA, n_special=64)
        out = pipe.denoise_step(d_in["latents"], ipe, ine, d_in["edit_latents"], progress_id=i % 50, height=H, width=W, cfg_scale=4.0)
        h_lat.copy_(out, non_blocking=True)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    d2h = h_lat.numel() * h_lat.element_size()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        # the final collective: every rank's latents to rank 0 (512 KiB per rank at 1024^2)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gathered = parallel.gather_latents(lat, dst=0)
        g1.record()
        torch.cuda.synchronize()
        gt = torch.tensor([g0.elapsed_time(g1)], device=device, dtype=torch.float64)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        coll["latent_gather"] = {"ms": round(gt.item(), 3), "bytes_per_rank": lat.numel() * lat.element_size(),
                                 "ok": bool(rank != 0 or (gathered is not None and len(gathered) == world))}
        coll["per_step_collectives"] = 0
        # latency-mode sub-leg: ONE image per pair of GPUs, the two CFG branches of a step on the two ranks of the
        # pair, one in-place NCCL all-gather of the two predictions per step
        if world % 2 == 0 and not args.cfg_parallel and not args.no_cfg_parallel_leg:
            grp, _, n_pairs = parallel.make_cfg_pairs()
            pipe.cfg_parallel_group = grp
            host_p = host_inputs(H, W, seed=100 + rank // 2)
            dev_p = {k: v.to(device, non_blocking=True) for k, v in host_p.items()}
            ipp = dict(prompt_emb=dev_p["pe_posi"], prompt_emb_mask=dev_p["mask_posi"], special_token_mask=dev_p["sp_posi"], txt_len=T_POSI, n_special=64)
            inp_ = dict(prompt_emb=dev_p["pe_nega"], prompt_emb_mask=dev_p["mask_nega"], special_token_mask=dev_p["sp_nega"], txt_len=T_NEGA, n_special=64)
            latp = dev_p["latents"].clone()
            ksteps = max(2, min(args.steps, 4))

            def pstep(i):
                pid = i % 50
                t_host = float(sched.timesteps[pid].to(torch.bfloat16))
                kw = dict(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=latp, timestep=ts_dev[pid:pid + 1], height=H,
                          width=W, edit_latents=dev_p["edit_latents"], is_train=False, timestep_host=t_host)
                pipe.run_cfg_branches(kw, ipp, inp_, vp, vn, ts_dev[pid:pid + 1], t_host)
                nat.cfg_euler_step(latp, vp, vn, 4.0, float(sched.dsigma(sched.timesteps[pid])))
            for i in range(2):
                pstep(i)
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for i in range(ksteps):
                pstep(2 + i)
            c1.record()
            barrier()
            ct = torch.tensor([c0.elapsed_time(c1)], device=device, dtype=torch.float64)
            dist.all_reduce(ct, op=dist.ReduceOp.MAX)
            # both ranks of a pair must hold bit-identical latents (each applied the same CFG + Euler update to the same inputs)
            peer = [torch.empty_like(latp) for _ in range(2)]
            dist.all_gather(peer, latp, group=grp)
            same = torch.tensor([float(torch.equal(peer[0], peer[1]))], device=device)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            coll["cfg_parallel"] = {"images_in_flight": n_pairs, "steps": ksteps, "ms_per_step": round(ct.item() / ksteps, 3),
                                    "steps_per_s_per_image": round(ksteps / (ct.item() * 1e-3), 4), "steps_per_s_job": round(n_pairs * ksteps / (ct.item() * 1e-3), 4),
                                    "pair_latents_bit_identical": bool(same.item() == 1.0),
                                    "per_step_collective": "one in-place all-gather of the two [1,16,h8,w8] predictions within each pair"}
            pipe.cfg_parallel_group = None
            nat.check_async()
        # second latency-mode sub-leg: ONE image on all N GPUs, every DiT forward split across the ranks (physicedit_b200/ulysses.py:
        # rows sequence-parallel, attention head-parallel, the two all-to-alls fused into the QKV GEMM's and the attention kernel's epilogues
        # as NVLink peer stores).  Every rank must end with bit-identical latents.
        if 24 % world == 0 and not args.cfg_parallel and not args.no_sequence_parallel_leg:
            host_s = host_inputs(H, W, seed=100)
            dev_s = {k: v.to(device, non_blocking=True) for k, v in host_s.items()}
            ips = dict(prompt_emb=dev_s["pe_posi"], prompt_emb_mask=dev_s["mask_posi"], special_token_mask=dev_s["sp_posi"], txt_len=T_POSI, n_special=64)
            ins = dict(prompt_emb=dev_s["pe_nega"], prompt_emb_mask=dev_s["mask_nega"], special_token_mask=dev_s["sp_nega"], txt_len=T_NEGA, n_special=64)
            lats = dev_s["latents"].clone()
            ksteps = max(2, min(args.steps, 4))
            streams_before = pipe.cfg_streams
            pipe.enable_sequence_parallel()
            for i in range(2):
                pipe.denoise_step(lats, ips, ins, dev_s["edit_latents"], progress_id=i, height=H, width=W)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(ksteps):
                pipe.denoise_step(lats, ips, ins, dev_s["edit_latents"], progress_id=2 + i, height=H, width=W)
            s1.record()
            barrier()
            st = torch.tensor([s0.elapsed_time(s1)], device=device, dtype=torch.float64)
            dist.all_reduce(st, op=dist.ReduceOp.MAX)
            every = [torch.empty_like(lats) for _ in range(world)]
            dist.all_gather(every, lats)
            same = torch.tensor([float(all(torch.equal(every[0], e) for e in every[1:]))], device=device)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            coll["sequence_parallel"] = {"images_in_flight": 1, "ranks": world, "steps": ksteps, "ms_per_step": round(st.item() / ksteps, 3),
                                         "steps_per_s_per_image": round(ksteps / (st.item() * 1e-3), 4), "speedup_vs_one_gpu_step": round((ms / args.steps) / (st.item() / ksteps), 3),
                                         "latents_bit_identical_on_all_ranks": bool(same.item() == 1.0),
                                         "per_block_exchange": "q/k/v head slices and attention rows stored into peer HBM by the QKV GEMM and attention epilogues "
                                                               "(NVLink P2P, torch symmetric memory); 2 device barriers per block"}
            pipe.disable_sequence_parallel()
            pipe.cfg_streams = streams_before
            nat.check_async()
            # both latency modes at once (world >= 4): one CFG branch per half of the node, sequence-parallel inside a half
            try:
                if world >= 4 and world % 2 == 0 and 24 % (world // 2) == 0:
                    sp_group, pair_group = parallel.make_cfg_sequence_groups()
                    lats2 = dev_s["latents"].clone()
                    # the adapter rewrites the special rows of prompt_emb in place every step: start again from the request's own embeddings
                    ips = dict(ips, prompt_emb=host_s["pe_posi"].to(device))
                    ins = dict(ins, prompt_emb=host_s["pe_nega"].to(device))
                    pipe.enable_sequence_parallel(sp_group)
                    pipe.cfg_parallel_group = pair_group
                    for i in range(2):
                        pipe.denoise_step(lats2, ips, ins, dev_s["edit_latents"], progress_id=i, height=H, width=W)
                    barrier()
                    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s0.record()
                    for i in range(ksteps):
                        pipe.denoise_step(lats2, ips, ins, dev_s["edit_latents"], progress_id=2 + i, height=H, width=W)
                    s1.record()
                    barrier()
                    st2 = torch.tensor([s0.elapsed_time(s1)], device=device, dtype=torch.float64)
                    dist.all_reduce(st2, op=dist.ReduceOp.MAX)
                    every = [torch.empty_like(lats2) for _ in range(world)]
                    dist.all_gather(every, lats2)
                    same2 = torch.tensor([float(all(torch.equal(every[0], e) for e in every[1:]) and torch.equal(lats2, lats))], device=device)
                    dist.all_reduce(same2, op=dist.ReduceOp.MIN)
                    coll["sequence_parallel_cfg_split"] = {"images_in_flight": 1, "ranks": world, "ranks_per_cfg_branch": world // 2, "steps": ksteps,
                                                           "ms_per_step": round(st2.item() / ksteps, 3), "steps_per_s_per_image": round(ksteps / (st2.item() * 1e-3), 4),
                                                           "speedup_vs_one_gpu_step": round((ms / args.steps) / (st2.item() / ksteps), 3),
                                                           "latents_bit_identical_on_all_ranks_and_to_the_unsplit_mode": bool(same2.item() == 1.0)}
                    pipe.cfg_parallel_group = None
                    pipe.disable_sequence_parallel()
                    pipe.cfg_streams = streams_before
                    nat.check_async()

            except Exception as e:  # noqa: BLE001  (a reported sub-leg: never takes the headline line down)
                coll["sequence_parallel_cfg_split"] = {"error": f"{type(e).__name__}: {e}"[:300]}
                pipe.cfg_parallel_group = None
                pipe.disable_sequence_parallel()

    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    S_img = (H // 16) * (W // 16) + 4096
    fl_step = flops_forward(S_img, T_POSI, args.layers) + flops_forward(S_img, T_NEGA, args.layers)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    # Roofline per hot kernel: ALGORITHMIC flops (or bytes) per launch / mean launch duration (CUDA events on the launch stream).
    # `roofline` is the kernel with the largest share of the step; `roofline_by_kernel` lists all of them.
    peak_gbs = peaks.get("hbm_gbs", 6500.0)
    S_avg = S_img + (T_POSI + T_NEGA) / 2
    S2_avg = ((S_img + T_POSI) ** 2 + (S_img + T_NEGA) ** 2) / 2
    hot = {  # tag: (description, bound, algorithmic flops per launch, algorithmic bytes per launch, ncu kernel name for the traffic figure)
        "attention": ("attention_kernel<2 query tiles, P in TMEM> (joint attention, S x S x 128 x 24 heads)", "tensor", 4 * S2_avg * 128 * 24,
                      4 * S_avg * DIM * 2, "attention_kernel<2, 1>"),
        "gemm_up": ("gemm_kernel<cta_pair, bias+gelu> (MLP up-projection, M=S N=12288 K=3072)", "tensor", 2 * S_avg * DIM * 4 * DIM,
                    2 * (S_avg * DIM + 4 * DIM * DIM + S_avg * 4 * DIM), "gemm_kernel<2, 2>"),
        "gemm_down": ("gemm_kernel<cta_pair, gate-residual> (MLP down-projection, M=S N=3072 K=12288)", "tensor", 2 * S_avg * DIM * 4 * DIM,
                      2 * (S_avg * 4 * DIM + 4 * DIM * DIM + 2 * S_avg * DIM), "gemm_down"),
        "gemm_qkv": ("gemm_kernel<cta_pair, qkv norm+rope> (fused QKV projection, M=S N=9216 K=3072)", "tensor", 2 * S_avg * DIM * 3 * DIM,
                     2 * (S_avg * DIM + 3 * DIM * DIM + 3 * S_avg * DIM), None),
        "gemm_out": ("gemm_kernel<cta_pair, gate-residual> (attention out-projection, M=S N=3072 K=3072)", "tensor", 2 * S_avg * DIM * DIM,
                     2 * (S_avg * DIM + DIM * DIM + 2 * S_avg * DIM), "gemm_out"),
        "ln_mod": ("layernorm_modulate2 (LN + AdaLN scale/shift, both streams)", "hbm", None, 2 * S_avg * DIM * 2, None),
    }
    roofs = {}
    for tag, (desc, bound, fl, by, ncu_name) in hot.items():
        if not prof.get(tag):
            continue
        n, tot = prof[tag]
        avg_s = tot / n * 1e-3
        if bound == "tensor":
            ach, pk, unit, src = fl / avg_s / 1e12, peak_tf, "TFLOP/s", "MEASURED_PEAKS.json bf16_tflops_sustained"
        else:
            ach, pk, unit, src = by / avg_s / 1e9, peak_gbs, "GB/s", "MEASURED_PEAKS.json hbm_gbs"
        roofs[tag] = {"bound": bound, "kernel": desc, "achieved": round(ach, 1), "peak": pk, "peak_source": src if peaks else "fallback", "unit": unit,
                      "frac": round(ach / pk, 4), "traffic": ncu_traffic_bytes(ncu_name) if ncu_name else None,
                      "traffic_unit": "bytes per launch (ncu: profiles/r02_ncu_attention_full.json, r01_ncu_full_metrics.json, r02_gemm_raster_ab.json, r01_gemm_traffic.json)", "algorithmic_bytes": int(by),
                      "launches": n, "avg_ms": round(tot / n, 4), "share_of_step": round(tot / ms_attr, 4)}
    roof = None
    if roofs:
        roof = dict(roofs[max(roofs, key=lambda k: roofs[k]["share_of_step"])])
        roof["whole_step_frac"] = round(fl_step * args.steps / (ms * 1e-3) / 1e12 / peak_tf, 4)
    shares = {k: round(v[1] / ms_attr, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    res = {
        "metric": "denoise steps/sec (1024x1024 edit, CFG: 2 DiT forwards/step)", "value": round(n_images * args.steps / (ms * 1e-3), 4), "unit": "steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random-init weights of the real architecture, seeded inputs)",
        "config": workload_config(args.resolution, args.layers),
        "derived": {"images_per_sec_50_steps": round(n_images * args.steps / (ms * 1e-3) / 50, 5),
                    "cfg_parallel": bool(args.cfg_parallel),
                    "cfg_streams": args.cfg_streams,
                    "attribution": ("per-launch CUDA events over the timed region" if events_in_region else
                                    f"per-launch CUDA events over {attr_steps} extra single-stream steps right after the timed region "
                                    f"({round(ms_attr / attr_steps, 2)} ms/step; with two streams the branches' launches overlap)"),
                    "tflops_per_step": round(fl_step / 1e12, 2), "achieved_tflops_per_gpu": round(fl_step * args.steps / (ms * 1e-3) / 1e12, 1)},
        "e2e": {"value": round(n_images * e2e_steps / (ms_e2e * 1e-3), 4), "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "gpu_launches": launches, "finite": finite, "clocks": clocks, "roofline": roof, "roofline_by_kernel": roofs, "kernel_time_share": shares,
    }
    if coll is not None:
        res["collectives"] = coll
    if not args.no_vae:
        res["vae"] = vae_leg(device, H, W, ms / args.steps)
    if world == 1 and not args.no_text_encoder:
        vae_ms = (res["vae"]["encode"]["ms"] + res["vae"]["decode"]["ms"]) if "vae" in res else 0.0
        try:
            res["text_encoder"] = text_encoder_leg(device, ms / args.steps, vae_ms)
        except Exception as e:  # noqa: BLE001  (a reported leg beside the metric: never takes the headline line down)
            res["text_encoder"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    if world == 1 and not args.no_stock_gpu:
        del pipe, eng, lat, dev, d_in
        torch.cuda.empty_cache()
        res["stock_gpu"] = stock_gpu_leg(device, H, W)
        res["vs_stock_gpu"] = res["stock_gpu"].get("speedup")
        res["parity"] = res["stock_gpu"].get("parity")
    if world == 1 and not args.no_training_leg:
        torch.cuda.empty_cache()
        try:
            res["training"] = training_leg(device, layers=args.train_layers)
        except Exception as e:  # noqa: BLE001  (a reported leg beside the metric: never takes the headline line down)
            res["training"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        res["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


def vae_leg(device, H, W, ms_per_step):
    """The step either side of the loop, outside the headline metric: native QwenImageVAE.encode of one HxW edit image from
    pinned host memory and .decode of the final latents back to host, random-init weights of the real architecture.  With it the
    per-image time of a 50-step edit (text encoder excluded) is encode + 50 steps + decode."""
    from physicedit_b200 import native as nv
    from physicedit_b200.vae import QwenImageVAE
    nat = nv.Native.get(device.index or 0)
    with torch.device("meta"):
        vae = QwenImageVAE()
    g = torch.Generator(device=device).manual_seed(0)
    sd = {}
    for k, v in vae.state_dict().items():
        if k.endswith("gamma"):
            t = 1 + 0.1 * torch.randn(v.shape, generator=g, device=device)
        elif k.endswith("weight"):
            t = (torch.rand(v.shape, generator=g, device=device) * 2 - 1) * 1.7 / (v.shape[1] * v.shape[-1] * v.shape[-2]) ** 0.5
        else:
            t = (torch.rand(v.shape, generator=g, device=device) * 2 - 1) * 0.05
        sd[k] = t.to(torch.bfloat16)
    vae.load_state_dict(sd, assign=True)
    vae.eval()
    img_h = (torch.rand(1, 3, H, W) * 2 - 1).to(torch.bfloat16).pin_memory()
    lat = torch.randn(1, 16, H // 8, W // 8, device=device).to(torch.bfloat16)
    out_h = torch.empty(1, 3, H, W, dtype=torch.bfloat16).pin_memory()
    times = {}
    for name in ("encode", "decode"):
        def run():
            if name == "encode":
                return vae.encode(img_h.to(device, non_blocking=True))
            out_h.copy_(vae.decode(lat), non_blocking=True)
            return out_h
        for _ in range(2):
            run()
        nat.check_async()
        n0 = nat.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r = run()
        e1.record()
        torch.cuda.synchronize(device)
        times[name] = {"ms": round(e0.elapsed_time(e1) / 5, 3), "launches": (nat.launches - n0) // 5, "finite": bool(torch.isfinite(r.float()).all())}
    per_image_ms = times["encode"]["ms"] + 50 * ms_per_step + times["decode"]["ms"]
    return {"what": f"QwenImageVAE encode (host image -> latents) / decode (latents -> host image) at {H}x{W}, native path, outside the headline metric",
            "encode": times["encode"], "decode": times["decode"],
            "image_50_steps": {"ms": round(per_image_ms, 1), "images_per_sec_per_gpu": round(1e3 / per_image_ms, 5),
                               "vae_share": round((times["encode"]["ms"] + times["decode"]["ms"]) / per_image_ms, 5),
                               "note": "encode + 50 x ms_per_step + decode; text encoder not included (SURVEY 8f2)"}}


def stock_gpu_leg(device, H, W, layers=4, iters=10, warmup=3):
    """GPU-vs-GPU baseline in the same run: the REFERENCE's own `model_fn_qwen_image` +
    `QwenImageDiT` (stock PyTorch ops: cuBLASLt + SDPA) when a reference tree is on the box (baseline/_ref), else the oracle's
    bf16 mode (the same ATen ops, restated) -- `kind` says which.  `layers` blocks at the benchmark's sequence (S = 8192 + 512),
    identical weights and inputs for both sides, CUDA events, `warmup` + `iters` forwards each.  Also the parity figure north_star
    quotes: rel-L2 between the native and the stock bf16 output of that forward."""
    from oracle import dit_oracle as O
    from oracle import ref_import
    from physicedit_b200.model_fn import model_fn_qwen_image
    pipe = build_model(device, layers, seed=0)
    sd = {k: v.detach() for k, v in pipe.dit.state_dict().items()}
    ad = {k: v.detach() for k, v in pipe.visual_thinking_adapter.state_dict().items()}
    inp = {k: v.to(device) for k, v in synth_inputs(H, W, T_POSI, seed=100).items()}
    t = torch.tensor([744.611382484436]).to(torch.bfloat16).to(device)
    nat_kw = dict(dit=pipe.dit, visual_thinking_adapter=pipe.visual_thinking_adapter, latents=inp["latents"], timestep=t, prompt_emb_mask=inp["prompt_emb_mask"],
                  special_token_mask=inp["special_token_mask"], height=H, width=W, edit_latents=inp["edit_latents"], is_train=False,
                  timestep_host=float(t[0]), txt_len=T_POSI, n_special=64)

    def timed(fn):
        for _ in range(warmup):
            out = fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / iters, out

    # every timed forward starts from the same prompt (the adapter rewrites the special rows in place)
    ms_nat, y_nat = timed(lambda: model_fn_qwen_image(prompt_emb=inp["prompt_emb"].clone(), **nat_kw)[0])
    y_nat = y_nat.clone()
    kind, how = "port", "oracle/dit_oracle.py in bf16 on the GPU (the reference's ATen ops restated; no reference tree on this box)"
    stock = None
    if ref_import.reference_root() is not None:
        try:
            ctx = ref_import.ReferenceModules()
            ref = ctx.__enter__()
            rdit = ref_import.build_reference_dit(ref, sd, layers, torch.bfloat16, device)
            rad = ref.helpers.VisualThinkingDualAdapter(3584, 3584, pipe.visual_thinking_adapter.t_min, pipe.visual_thinking_adapter.t_max)
            rad.load_state_dict(ad)
            rad = rad.to(device=device, dtype=torch.bfloat16).eval()
            ref_fn = ref.phys.model_fn_qwen_image

            def stock():
                with torch.no_grad():
                    return ref_fn(dit=rdit, blockwise_controlnet=None, visual_thinking_adapter=rad, latents=inp["latents"], timestep=t,
                                  prompt_emb=inp["prompt_emb"].clone(), prompt_emb_mask=inp["prompt_emb_mask"], special_token_mask=inp["special_token_mask"],
                                  height=H, width=W, edit_latents=inp["edit_latents"], is_train=False)[0]
            stock()
            kind, how = "reference", f"the reference's own model_fn_qwen_image + QwenImageDiT imported from {ref_import.reference_root()}"
        except Exception as e:  # noqa: BLE001
            stock, how = None, how + f" [importing the reference failed: {type(e).__name__}: {e}]"[:300]
    if stock is None:
        def stock():
            with torch.no_grad():
                return O.model_fn(sd, ad, inp["latents"], t, inp["prompt_emb"].clone(), inp["prompt_emb_mask"], inp["special_token_mask"], H, W,
                                  edit_latents=inp["edit_latents"], num_layers=layers, cuda_scalar_div=True)
    ms_stock, y_stock = timed(stock)
    # which attention kernel the stock path's SDPA dispatcher picked on this box
    sdpa_kernel_name = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            stock()
            torch.cuda.synchronize(device)
        names = sorted(((e.device_time_total, e.key) for e in prof.key_averages()), reverse=True)
        att = [k for _, k in names if any(s_ in k.lower() for s_ in ("flash", "fmha", "attention", "sdpa", "cudnn_generated"))]
        sdpa_kernel_name = att[0][:160] if att else None
    except Exception as e:  # noqa: BLE001
        sdpa_kernel_name = f"profiler unavailable: {type(e).__name__}"
    rel = ((y_nat.float() - y_stock.float()).norm() / y_stock.float().norm()).item()
    S_img = (H // 16) * (W // 16) + 4096
    fl = flops_forward(S_img, T_POSI, layers)
    return {"kind": kind, "what": how, "blocks": layers, "S": S_img + T_POSI, "iters": iters, "warmup": warmup,
            "stock_ms_per_forward": round(ms_stock, 3), "native_ms_per_forward": round(ms_nat, 3), "speedup": round(ms_stock / ms_nat, 3),
            "stock_tflops": round(fl / ms_stock / 1e9, 1), "native_tflops": round(fl / ms_nat / 1e9, 1), "stock_attention_kernel": sdpa_kernel_name,
            "note": "burst clocks (a 4-block forward is ~15 ms); the 60-block loop above runs at the power-capped clock",
            "parity": {"native_vs_stock_bf16_rel_l2": rel, "blocks": layers, "what": "rel-L2 of the model_fn output (latents) between the native and the stock bf16 forward, same weights and inputs; "
                       "the full-depth figure against the fp32 oracle is in profiles/r02_parity_depth.json (tests/test_parity_depth_gpu.py)"}}


def feature_extractor_leg(device, H, W, frames=6, iters=5, warmup=2):
    """The pseudo-target branch of a training sample (QwenImageUnit_PhysicalVisualEmbedder, qwen_image_physical.py:1057-1118) on the native kernels,
    forward only (inference mode of the frozen / evaluated stack): DINOv2-with-registers base on `frames` middle key frames + the source image at 224^2,
    the two perceiver resamplers (64 latents, keys over frames x 256 resp. frames x H/16 x W/16 tokens) and their adapters.  `pe_small_attention` is
    the CUDA-core attention of both; the GEMMs are `pe_gemm`."""
    from physicedit_b200 import native as nv
    from physicedit_b200.pipeline import QwenImagePhysicPipeline
    nat = nv.Native.get(device.index or 0)
    pipe = QwenImagePhysicPipeline(device=device, torch_dtype=torch.bfloat16, dinov2_config=dict(hidden=768, layers=12, heads=12))
    g = torch.Generator(device=device).manual_seed(3)
    for p in pipe.parameters():
        if p.dim() >= 2:
            p.data = ((torch.rand(p.shape, generator=g, device=device) * 2 - 1) / math.sqrt(p.shape[-1])).to(torch.bfloat16)
    pipe.to(device)
    pipe.eval()
    x = dict(dino_middle=torch.randn(frames, 3, 224, 224, device=device, generator=g).to(torch.bfloat16),
             dino_source=torch.randn(1, 3, 224, 224, device=device, generator=g).to(torch.bfloat16),
             vae_middle_latents=torch.randn(frames, 16, H // 8, W // 8, device=device, generator=g).to(torch.bfloat16),
             vae_source_latents=torch.randn(1, 16, H // 8, W // 8, device=device, generator=g).to(torch.bfloat16))

    def timed(fn):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = nat.launches
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / iters, (nat.launches - l0) // iters
    ms_all, n_all = timed(lambda: pipe.physical_visual_embeddings(**x))
    ms_dino, n_dino = timed(lambda: pipe.dinov2(x["dino_middle"]))
    # ViT-B/14 at 224^2: 261 tokens (256 patches + CLS + 4 registers), 12 layers: per token per layer 24 d^2 (4 projections + 8 d^2 of MLP pairs) + attention 4 n d
    tok, d = 261, 768
    fl_dino = frames * 12 * (tok * 24 * d * d + 4 * tok * tok * d)
    nat.check_async()
    return {"frames": frames, "ms_pseudo_targets": round(ms_all, 3), "own_kernel_launches": n_all, "ms_dinov2_middle_frames": round(ms_dino, 3),
            "dinov2_launches": n_dino, "dinov2_tflops": round(fl_dino / (ms_dino * 1e-3) / 1e12, 1),
            "note": "small-model regime: 1566 tokens x 768 wide, launch-bound (each launch is a few microseconds of work); < 1 % of a training step"}


TRAIN_TARGETS = "to_q,to_k,to_v,add_q_proj,add_k_proj,add_v_proj,to_out.0,to_add_out,img_mlp.net.2,img_mod.1,txt_mlp.net.2,txt_mod.1".split(",")


def training_leg(device, layers=8, H=480, W=832, T=512, rank=128, iters=3, warmup=2):
    """One training step (scripts/train/train_multigpu.sh: 480 x 832, LoRA r = 128 un-merged on 12 linears per block, the dual adapter
    trainable, gradient checkpointing per block): `pipe.training_loss(...).backward()` on the native path against the REFERENCE's own
    model_fn + QwenImageDiT under torch autograd with the same LoRA layer (stock PyTorch: cuBLASLt + SDPA forward / backward) on the same GPU,
    same weights, same inputs, `layers` blocks, CUDA events.  Reported beside the headline metric; the native attention backward is a
    composition of GEMM / softmax / transpose passes (HBM-bound), not yet a fused kernel -- this leg is where that shows."""
    import torch.nn.functional as F
    from oracle import ref_import
    from physicedit_b200 import native as nv
    from physicedit_b200.lora import inject_lora
    nat = nv.Native.get(device.index or 0)
    pipe = build_model(device, layers, seed=0)
    sd = {k: v.detach().clone() for k, v in pipe.dit.state_dict().items()}
    ad = {k: v.detach().clone() for k, v in pipe.visual_thinking_adapter.state_dict().items()}
    pipe.scheduler.set_timesteps(1000, training=True)
    pipe.freeze_except(["visual_thinking_adapter"])
    inject_lora(pipe.dit, TRAIN_TARGETS, rank)
    g = torch.Generator(device=device).manual_seed(1)
    lora = {}
    for name, p in pipe.dit.named_parameters():
        if "lora_" in name:
            p.data.copy_((torch.randn(p.shape, device=device, generator=g) * (0.5 / math.sqrt(p.shape[1]))).to(torch.bfloat16))
            lora[name] = p.detach().clone()
    inp = {k: v.to(device) for k, v in synth_inputs(H, W, T, seed=100, edit_hw=(H, W)).items()}
    gen = torch.Generator("cpu").manual_seed(5)
    gt = [torch.randn(1, 64, 3584, generator=gen).to(torch.bfloat16).to(device) for _ in range(2)]
    noise = torch.randn(1, 16, H // 8, W // 8, generator=gen).to(torch.bfloat16).to(device)
    tid = torch.tensor([400])
    params = [p for p in list(pipe.dit.parameters()) + list(pipe.visual_thinking_adapter.parameters()) if p.requires_grad]

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / iters, out

    def native_step():
        for p in params:
            p.grad = None
        loss = pipe.training_loss(input_latents=inp["latents"], prompt_emb=inp["prompt_emb"].clone(), prompt_emb_mask=inp["prompt_emb_mask"],
                                  special_token_mask=inp["special_token_mask"], height=H, width=W, edit_latents=inp["edit_latents"],
                                  pseudo_special_emb_dino=gt[0], pseudo_special_emb_vae=gt[1], is_train=True, use_gradient_checkpointing=True,
                                  timestep_id=tid, noise=noise)
        loss.backward()
        return loss.detach()
    l0 = nat.launches
    ms_nat, loss_nat = timed(native_step)
    launches = (nat.launches - l0) // (warmup + iters)
    nat.check_async()
    # forward-only and backward-only split of the native step
    ms_fwd, _ = timed(lambda: pipe.training_loss(input_latents=inp["latents"], prompt_emb=inp["prompt_emb"].clone(), prompt_emb_mask=inp["prompt_emb_mask"],
                                                  special_token_mask=inp["special_token_mask"], height=H, width=W, edit_latents=inp["edit_latents"],
                                                  pseudo_special_emb_dino=gt[0], pseudo_special_emb_vae=gt[1], is_train=True, use_gradient_checkpointing=True,
                                                  timestep_id=tid, noise=noise).detach())
    g_nat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).float().flatten() for n, p in pipe.dit.named_parameters() if "lora_" in n])
    out = {"config": {"blocks": layers, "height": H, "width": W, "S": 2 * (H // 16) * (W // 16) + T, "lora_rank": rank, "lora_targets_per_block": len(TRAIN_TARGETS),
                      "gradient_checkpointing": True, "trainable": "LoRA factors + dual adapter"},
           "native": {"ms_per_step": round(ms_nat, 2), "ms_forward": round(ms_fwd, 2), "ms_recompute_plus_backward": round(ms_nat - ms_fwd, 2),
                      "ms_per_block": round(ms_nat / layers, 3), "own_kernel_launches_per_step": launches, "loss": round(float(loss_nat), 5)},
           "iters": iters, "warmup": warmup}
    del pipe, params
    torch.cuda.empty_cache()
    try:
        out["feature_extractors"] = feature_extractor_leg(device, H, W)
    except Exception as e:  # noqa: BLE001
        out["feature_extractors"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if ref_import.reference_root() is None:
        out["stock_gpu"] = {"unavailable": "no reference tree on this box (baseline/_ref)"}
        return out
    try:
        with ref_import.ReferenceModules() as ref:
            rdit = ref_import.build_reference_dit(ref, sd, layers, torch.bfloat16, device)
            rad = ref.helpers.VisualThinkingDualAdapter(3584, 3584, 19.999980926513672, 1000.0)
            rad.load_state_dict(ad)
            rad = rad.to(device=device, dtype=torch.bfloat16).train()

            class StockLoRA(torch.nn.Module):                     # peft.tuners.lora.Linear at dropout 0 (peft itself is not installed here)
                def __init__(self, base, a, b):
                    super().__init__()
                    self.base_layer, self.a, self.b = base, torch.nn.Parameter(a.clone()), torch.nn.Parameter(b.clone())

                def forward(self, x):
                    return self.base_layer(x) + F.linear(F.linear(x, self.a), self.b)
            for p in rdit.parameters():
                p.requires_grad_(False)
            for name in sorted({n.split(".lora_A.")[0] for n in lora if ".lora_A." in n}):
                parent_name, _, leaf = name.rpartition(".")
                parent = rdit.get_submodule(parent_name)
                setattr(parent, leaf, StockLoRA(getattr(parent, leaf), lora[name + ".lora_A.default.weight"], lora[name + ".lora_B.default.weight"]))
            sparams = [p for p in list(rdit.parameters()) + list(rad.parameters()) if p.requires_grad]
            from physicedit_b200.scheduler import FlowMatchScheduler
            sch = FlowMatchScheduler()
            sch.set_timesteps(1000, training=True)
            t = sch.timesteps[tid].to(torch.bfloat16).to(device)
            noisy, target, weight = sch.add_noise(inp["latents"], noise, t), sch.training_target(inp["latents"], noise, t), float(sch.training_weight(t))
            ref_fn = ref.phys.model_fn_qwen_image

            def stock_step():
                for p in sparams:
                    p.grad = None
                pred, sl = ref_fn(dit=rdit, blockwise_controlnet=None, visual_thinking_adapter=rad, latents=noisy, timestep=t, prompt_emb=inp["prompt_emb"].clone(),
                                  prompt_emb_mask=inp["prompt_emb_mask"], special_token_mask=inp["special_token_mask"], height=H, width=W,
                                  edit_latents=inp["edit_latents"], is_train=True, use_gradient_checkpointing=True, pseudo_special_emb_dino=gt[0],
                                  pseudo_special_emb_vae=gt[1])
                loss = F.mse_loss(pred.float(), target.float()) * weight + sl
                loss.backward()
                return loss.detach()
            ms_stock, loss_stock = timed(stock_step)
            g_stock = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).float().flatten()
                                 for n, m in sorted(rdit.named_modules()) if isinstance(m, StockLoRA) for p in (m.a, m.b)])
            g_nat_sorted = g_nat                                   # named_parameters order: lora_A then lora_B per module, modules in definition order
            out["stock_gpu"] = {"kind": "reference", "what": f"the reference's model_fn_qwen_image + QwenImageDiT from {ref_import.reference_root()} under torch "
                                "autograd, LoRA layer restated (peft absent)", "ms_per_step": round(ms_stock, 2), "loss": round(float(loss_stock), 5)}
            out["speedup_vs_stock_gpu"] = round(ms_stock / ms_nat, 3)
            out["lora_grad_norm_native_over_stock"] = round((g_nat_sorted.norm() / g_stock.norm()).item(), 4)
    except Exception as e:  # noqa: BLE001
        out["stock_gpu"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def text_encoder_leg(device, ms_per_step, vae_ms, new_tokens=96):
    """The step in front of the loop, outside the headline metric: the Qwen2.5-VL text encoder at its real size (7B config of
    the reference's wrapper, random-init bf16 weights) on the native path -- `generate` (prefill + greedy KV-cache decode replayed from a
    CUDA graph) and `edit_forward` on a processor-shaped request (a 392x392 image = 196 image tokens + text).  B=1 decode is HBM-bound:
    the roofline is the bytes of weights one token must stream divided by the measured HBM bandwidth."""
    from physicedit_b200.text_encoder import QwenImageTextEncoder, VLConfig
    cfg = VLConfig()
    with torch.device("meta"):
        te = QwenImageTextEncoder(cfg, rope_mode="mrope")
    g = torch.Generator(device=device).manual_seed(1)
    sd = {}
    for k, v in te.state_dict().items():
        if "embed_tokens" in k:
            t = torch.randn(v.shape, generator=g, device=device, dtype=torch.float32)
        elif v.dim() >= 2:
            fan_in = math.prod(v.shape[1:])
            t = (torch.rand(v.shape, generator=g, device=device, dtype=torch.float32) * 2 - 1) * (1.0 / math.sqrt(fan_in))
        elif k.endswith(".bias"):
            t = (torch.rand(v.shape, generator=g, device=device, dtype=torch.float32) * 2 - 1) * 0.02
        else:
            t = torch.ones(v.shape, device=device)
        sd[k] = t.to(torch.bfloat16)
    te.load_state_dict(sd, assign=True)
    te.eval()
    te.cfg.eos_token_id = -1                      # random weights: never stop early, time exactly `new_tokens`
    n_img, n_txt = 196, 120
    ids = torch.cat([torch.randint(1000, 100000, (40,), generator=torch.Generator().manual_seed(0)), torch.full((n_img,), cfg.image_token_id),
                     torch.randint(1000, 100000, (n_txt,), generator=torch.Generator().manual_seed(1))]).view(1, -1)
    req = dict(input_ids=ids.to(device), attention_mask=torch.ones_like(ids).to(device),
               pixel_values=torch.randn(784, 1176, device=device).to(torch.bfloat16), image_grid_thw=torch.tensor([[1, 28, 28]], device=device))
    te.generate(**req, max_new_tokens=8)          # warm-up (packs the weights, compiles nothing: kernels are prebuilt)
    torch.cuda.synchronize(device)
    te.generate(**req, max_new_tokens=new_tokens)
    torch.cuda.synchronize(device)
    st = dict(te.last_generate_stats)
    # what the pipeline does: the positive and the negative branch decoded together (batch 2: one pass over the weights, two tokens)
    req2 = dict(req, input_ids=req["input_ids"][:, :-37].contiguous(), attention_mask=req["attention_mask"][:, :-37].contiguous())
    te.generate_batch([req, req2], max_new_tokens=new_tokens)
    torch.cuda.synchronize(device)
    st2 = dict(te.last_generate_stats)
    # per-kernel breakdown of one decode step: a few eager (graph-less) tokens with per-launch CUDA events
    from physicedit_b200 import native as nv
    nat = nv.Native.get(device.index or 0)
    te.use_cuda_graph = False
    nat.prof = {}
    te.generate(**req, max_new_tokens=6)
    torch.cuda.synchronize(device)
    breakdown = {k: round(v[1] * 1e3 / 5, 1) for k, v in sorted(nat.profile_summary().items(), key=lambda kv: -kv[1][1]) if k.startswith("te_")}
    nat.prof = None
    te.use_cuda_graph = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        h = te.edit_forward(**req)[-1]
    e1.record()
    torch.cuda.synchronize(device)
    ms_edit = e0.elapsed_time(e1) / 3
    c = cfg
    per_layer = (c.heads + 2 * c.kv_heads) * c.head_dim * c.hidden + c.hidden * c.heads * c.head_dim + 3 * c.hidden * c.intermediate
    bytes_per_token = 2 * (c.layers * per_layer + c.vocab * c.hidden)           # decoder weights + lm_head, bf16 (KV cache and norms are noise)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    gbs = bytes_per_token / (st["ms_per_token"] * 1e-3) / 1e9
    pre_loop_ms = st2["prefill_s"] * 1e3 + 1000 * st2["ms_per_token"] + 2 * ms_edit       # both CFG branches in one batch, 1000 new tokens each (the cap)
    image_ms = pre_loop_ms + vae_ms + 50 * ms_per_step
    return {"what": "Qwen2.5-VL 7B-config text encoder on the native path (random-init weights), outside the headline metric",
            "prompt_tokens": st["prompt_tokens"], "new_tokens_timed": st["tokens_computed"], "decode_ms_per_token": round(st["ms_per_token"], 3),
            "decode_ms_per_step_both_cfg_branches": round(st2["ms_per_token"], 3), "prefill_ms_both_cfg_branches": round(st2["prefill_s"] * 1e3, 2),
            "cuda_graph": st["cuda_graph"], "graph_capture_ms": round(st.get("graph_capture_s", 0) * 1e3, 1), "prefill_ms": round(st["prefill_s"] * 1e3, 2),
            "edit_forward_ms": round(ms_edit, 2), "decode_step_kernel_us": breakdown,
            "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peaks.get("hbm_gbs", 6500.0), "unit": "GB/s",
                         "frac": round(gbs / peaks.get("hbm_gbs", 6500.0), 4), "algorithmic_bytes": bytes_per_token},
            "finite": bool(torch.isfinite(h.float()).all()),
            "image_50_steps_with_text_encoder": {"pre_loop_ms": round(pre_loop_ms, 1), "ms": round(image_ms, 1), "images_per_sec_per_gpu": round(1e3 / image_ms, 5),
                                                 "note": "generate for both CFG branches in one batch (2 prefills + 1000 batch-2 decode steps, the reference's max_new_tokens) + 2 x "
                                                         "edit_forward + VAE encode/decode + 50 denoise steps; real prompts stop at EOS earlier"}}


def _cpu_threads():
    try:        # torchrun exports OMP_NUM_THREADS=1; the CPU arm is allowed every host thread it can use
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    return torch.get_num_threads()


class CpuBlockSample:
    """The reference algorithm (oracle port: torch CPU ops in bf16, like the reference's own CPU path) on the host cores.
    One sample = ONE block of the posi forward (S = 8192 + 512) + ONE block of the nega forward (S = 8192 + 288) at the
    benchmark's sequence lengths = exactly 1/60 of the block work of a CFG denoise step."""

    def __init__(self, args):
        from oracle import dit_oracle as O
        self.O = O
        torch.manual_seed(0)
        self.cores = _cpu_threads()
        H = W = args.resolution
        self.S_img = (H // 16) * (W // 16) + 4096
        shapes = {k: v for k, v in O.dit_param_shapes(1).items() if k.startswith("transformer_blocks.0.")}
        self.Wt = O.synth_weights(shapes, seed=0, dtype=torch.bfloat16)
        self.image = torch.randn(1, self.S_img, DIM).bfloat16()
        self.temb = torch.randn(1, DIM).bfloat16()
        self.branches = []
        for T in (T_POSI, T_NEGA):
            self.branches.append((torch.randn(1, T, DIM).bfloat16(), O.rope_tables([(1, H // 16, W // 16), (1, 64, 64)], T)))

    def run(self, branches=(0, 1)):
        t0 = time.time()
        with torch.no_grad():
            for b in branches:
                text, rope = self.branches[b]
                self.O.block_forward(self.Wt, 0, self.image, text, self.temb, rope)
        return time.time() - t0


def cpu_baseline(args, samples=3):
    """`cpu_baseline` of the native line: a bounded sample (a few x 1/60 of a step) of the same workload on the box's host cores."""
    smp = CpuBlockSample(args)
    smp.run()
    dts = [smp.run() for _ in range(samples)]
    dt = sum(dts) / len(dts)
    return {"value": round(1.0 / (dt * LAYERS), 6), "unit": "steps/s", "cores": smp.cores, "kind": "port",
            "sample": f"{samples} samples of (1 posi block at S={smp.S_img + T_POSI} + 1 nega block at S={smp.S_img + T_NEGA}) in bf16 = 1/60 of a CFG step each: "
                      f"{dt:.3f} s/sample -> x60 = {dt * LAYERS:.1f} s/step"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (the reference is pure PyTorch, so
    its CPU path IS these torch ops; oracle port, kind "port"), rank 0 only.  Each timed "step" is a bounded sample of one
    denoise step -- one posi block + one nega block at full sequence length = 1/60 of the step's block work -- so `ms_per_step`
    is the measured time of what was actually run and `value` = 1 / (60 x that) is the metric (steps/s) it extrapolates to.
    After the timed samples ONE real full-depth (60-block) posi forward is run and timed to check the extrapolation."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    smp = CpuBlockSample(args)
    for _ in range(args.warmup):
        smp.run()
    t0 = time.time()
    dts = [smp.run() for _ in range(args.steps)]
    wall = time.time() - t0
    dt = sum(dts) / len(dts)
    v = 1.0 / (dt * LAYERS)
    cb = {"value": round(v, 6), "unit": "steps/s", "cores": smp.cores, "kind": "port",
          "sample": f"each of the {args.steps} timed steps = 1 posi block (S={smp.S_img + T_POSI}) + 1 nega block (S={smp.S_img + T_NEGA}) in bf16 on the host CPU "
                    f"= 1/60 of a CFG denoise step: {dt:.3f} s/sample; value = 1 / (60 x sample time)",
          "sample_fraction_of_step": round(1.0 / LAYERS, 6), "timed_wall_s": round(wall, 2)}
    if not args.no_full_forward:
        # one REAL full-depth forward (60 blocks, posi branch; the same block weights re-used: 680 MB per block is far beyond the
        # caches either way): validates the x60 extrapolation
        t1 = time.time()
        with torch.no_grad():
            text, rope = smp.branches[0]
            image = smp.image
            for _ in range(LAYERS):
                text, image = smp.O.block_forward(smp.Wt, 0, image, text, smp.temb, rope)
        full = time.time() - t1
        posi_share = flops_forward(smp.S_img, T_POSI) / (flops_forward(smp.S_img, T_POSI) + flops_forward(smp.S_img, T_NEGA))
        cb["full_depth_posi_forward_s"] = round(full, 2)
        cb["extrapolated_posi_forward_s"] = round(dt * LAYERS * posi_share, 2)
    res = {"impl": "reference", "metric": "denoise steps/sec (1024x1024 edit, CFG: 2 DiT forwards/step)", "value": round(v, 6), "unit": "steps/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic (random-init weights of the real architecture, seeded inputs)",
           "config": workload_config(args.resolution, LAYERS),
           "cpu_baseline": cb, "e2e": {"value": round(v, 6), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--resolution", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=LAYERS)
    ap.add_argument("--no-cta-pair", action="store_true")
    ap.add_argument("--attn-flags", dest="attn_flags", type=int, default=0, help="PE_ATTN_FLAG_* bits (8 = split-row softmax kernel)")
    ap.add_argument("--cfg-streams", dest="cfg_streams", type=int, default=2, choices=[1, 2],
                    help="2 = the two CFG branches of a step run concurrently on two CUDA streams (fills partial last waves)")
    ap.add_argument("--cfg-parallel", dest="cfg_parallel", action="store_true",
                    help="latency mode: one image per pair of GPUs (positive branch on the even rank, negative on the odd one); needs an even --gpus")
    ap.add_argument("--no-kernel-events", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-training-leg", dest="no_training_leg", action="store_true", help="skip the training-step leg (native vs stock autograd)")
    ap.add_argument("--train-layers", dest="train_layers", type=int, default=8, help="blocks in the training-step leg")
    ap.add_argument("--no-text-encoder", dest="no_text_encoder", action="store_true", help="skip the Qwen2.5-VL text-encoder leg (7B config, random weights)")
    ap.add_argument("--no-stock-gpu", dest="no_stock_gpu", action="store_true", help="skip the stock-PyTorch GPU baseline leg (4 blocks, same inputs)")
    ap.add_argument("--no-cfg-parallel-leg", dest="no_cfg_parallel_leg", action="store_true", help="N>=2: skip the CFG-parallel latency sub-leg")
    ap.add_argument("--no-sequence-parallel-leg", dest="no_sequence_parallel_leg", action="store_true",
                    help="N>=2: skip the sequence-parallel (one image on all N GPUs) latency sub-leg")
    ap.add_argument("--no-full-forward", dest="no_full_forward", action="store_true", help="--impl reference: skip the one real 60-block CPU forward")
    ap.add_argument("--no-vae", dest="no_vae", action="store_true", help="skip the VAE encode/decode leg (reported beside, not inside, the metric)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_native(args)


if __name__ == "__main__":
    main()
