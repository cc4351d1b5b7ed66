#!/usr/bin/env python
"""2-GPU check of the CFG-parallel latency mode (run under torchrun --nproc-per-node 2): the latents of a 4-step denoise loop on a GPU
pair (positive branch on rank 0, negative on rank 1, one all-gather per step) must be bit-identical to the single-GPU loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from physicedit_b200 import parallel
from bench import build_model, host_inputs
pipe = build_model(torch.device("cuda", local), 2, seed=0)
parallel.broadcast_weights(pipe, src=0)
host = host_inputs(256, 256, seed=7)
dev = {k: v.cuda() for k, v in host.items()}
def run():
    ip = dict(prompt_emb=dev["pe_posi"].clone(), prompt_emb_mask=dev["mask_posi"], special_token_mask=dev["sp_posi"])
    in_ = dict(prompt_emb=dev["pe_nega"].clone(), prompt_emb_mask=dev["mask_nega"], special_token_mask=dev["sp_nega"])
    return pipe.denoise(dev["latents"], ip, in_, dev["edit_latents"], height=256, width=256, num_inference_steps=4, cfg_scale=4.0)
ref = run()
grp, pair, npairs = parallel.make_cfg_pairs()
pipe.cfg_parallel_group = grp
out = run()
pipe.dit.engine().nat.check_async()
ok = torch.equal(ref, out)
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("cfg-parallel latents bit-identical to the single-GPU loop on both ranks:", bool(flag.item()))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
