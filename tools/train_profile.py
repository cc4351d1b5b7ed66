#!/usr/bin/env python
"""Where one native training step spends GPU time (torch profiler, kernel table) and how much of the wall time the GPU is busy."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, torch
import bench
from physicedit_b200 import native as nv
from physicedit_b200.lora import inject_lora
layers, H, W, T = int(os.environ.get("LAYERS", 4)), 480, 832, 512
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
pipe = bench.build_model(dev, layers, seed=0)
pipe.scheduler.set_timesteps(1000, training=True)
pipe.freeze_except(["visual_thinking_adapter"])
inject_lora(pipe.dit, bench.TRAIN_TARGETS, 128)
g = torch.Generator(device=dev).manual_seed(1)
for n, p in pipe.dit.named_parameters():
    if "lora_" in n:
        p.data.copy_((torch.randn(p.shape, device=dev, generator=g) * (0.5 / math.sqrt(p.shape[1]))).to(torch.bfloat16))
inp = {k: v.to(dev) for k, v in bench.synth_inputs(H, W, T, seed=100, edit_hw=(H, W)).items()}
gt = [torch.randn(1, 64, 3584, device=dev).bfloat16() for _ in range(2)]
noise = torch.randn(1, 16, H // 8, W // 8, device=dev).bfloat16()
def step():
    loss = pipe.training_loss(input_latents=inp["latents"], prompt_emb=inp["prompt_emb"].clone(), prompt_emb_mask=inp["prompt_emb_mask"],
                              special_token_mask=inp["special_token_mask"], height=H, width=W, edit_latents=inp["edit_latents"], pseudo_special_emb_dino=gt[0],
                              pseudo_special_emb_vae=gt[1], is_train=True, use_gradient_checkpointing=True, timestep_id=torch.tensor([400]), noise=noise)
    loss.backward()
for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
rows = sorted(((e.device_time_total, e.count, e.key) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"), reverse=True)
tot = sum(r[0] for r in rows)
print(json.dumps({"layers": layers, "wall_ms": round(wall, 2), "gpu_busy_ms": round(tot / 1e3, 2)}))
for us, cnt, key in rows[:28]:
    print(f"{us / 1e3:9.3f} ms  {cnt:5d}x  {key[:110]}")
