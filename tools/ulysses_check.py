#!/usr/bin/env python
"""N-GPU check + timing of the sequence-parallel (Ulysses) DiT forward (run under torchrun --nproc-per-node N):
the velocity of a forward split across the ranks must be bit-identical to the single-GPU forward on every rank; then times K CFG steps of one
image at --resolution (2048 = BASELINE config #4) on the group against the same steps on one GPU.
    torchrun --nproc-per-node 2 tools/ulysses_check.py [--resolution 1024] [--layers 4] [--steps 3]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
ap = argparse.ArgumentParser()
ap.add_argument("--resolution", type=int, default=1024)
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--cfg-split", action="store_true", help="two halves of the world, one CFG branch each (sequence-parallel inside a half)")
args = ap.parse_args()
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = dist.get_world_size()
from physicedit_b200 import parallel
from physicedit_b200 import native as nv
from bench import build_model, host_inputs, T_POSI, T_NEGA
pipe = build_model(dev, args.layers, seed=0)
parallel.broadcast_weights(pipe, src=0)
nat = nv.Native.get(local)
H = W = args.resolution
host = host_inputs(H, W, seed=7)
d = {k: v.to(dev) for k, v in host.items()}
pipe.scheduler.set_timesteps(50, dynamic_shift_len=(H // 16) * (W // 16))
pipe.cfg_streams = 1

def inputs():
    return (dict(prompt_emb=d["pe_posi"].clone(), prompt_emb_mask=d["mask_posi"], special_token_mask=d["sp_posi"], txt_len=T_POSI, n_special=64),
            dict(prompt_emb=d["pe_nega"].clone(), prompt_emb_mask=d["mask_nega"], special_token_mask=d["sp_nega"], txt_len=T_NEGA, n_special=64))

def run(steps, timed=False):
    ip, in_ = inputs()
    lat = d["latents"].clone()
    for i in range(2 if timed else 0):            # warm-up
        pipe.denoise_step(lat, ip, in_, d["edit_latents"], progress_id=i, height=H, width=W)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        pipe.denoise_step(lat, ip, in_, d["edit_latents"], progress_id=2 + i, height=H, width=W)
    e1.record()
    torch.cuda.synchronize()
    nat.check_async()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return lat, t.item()

ref, _ = run(2)
ref_t = run(args.steps, timed=True)[1]
if args.cfg_split:
    sp_group, pair_group = parallel.make_cfg_sequence_groups()
    pipe.enable_sequence_parallel(sp_group)
    pipe.cfg_parallel_group = pair_group
else:
    pipe.enable_sequence_parallel()
out, _ = run(2)
same = torch.tensor([int(torch.equal(ref, out))], device=dev)
dist.all_reduce(same, op=dist.ReduceOp.MIN)
rel = ((out.float() - ref.float()).norm() / ref.float().norm()).item()
sp_t = run(args.steps, timed=True)[1]
if rank == 0:
    print(json.dumps({"ranks": N, "mode": "cfg branch per half x sequence-parallel inside" if args.cfg_split else "sequence-parallel", "resolution": args.resolution, "layers": args.layers, "bit_identical_on_all_ranks": bool(same.item()), "rel_l2_vs_single_gpu": rel,
                      "single_gpu_ms_per_step": round(ref_t, 2), "sequence_parallel_ms_per_step": round(sp_t, 2), "speedup": round(ref_t / sp_t, 3),
                      "finite": bool(torch.isfinite(out.float()).all())}))
dist.destroy_process_group()
sys.exit(0 if same.item() else 1)
