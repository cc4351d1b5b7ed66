#!/usr/bin/env python
"""2-rank check of the training loop on DDP (run under torchrun --nproc-per-node 2): one optimizer step of launch_training_task over 2 samples, one per
rank with the NCCL gradient all-reduce, must move the parameters like ONE process that accumulates the two samples' gradients
(gradient_accumulation_steps = 2): same mean gradient up to bf16 rounding of the reduction order, and both ranks end with identical
parameters.  Exit code 0 = ok."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import test_training_gpu as T
from physicedit_b200.trainers import ModelLogger, launch_training_task


class Quiet(ModelLogger):
    def save_model(self, *a, **k):
        pass


def run(distributed):
    pipe, sd, ad, lora = T._native_training_pipe(2, seed=7, rank=16)
    model = T._training_module(pipe)
    before = {k: p.detach().clone() for k, p in model.named_parameters() if p.requires_grad}
    data = T._Samples(2, 64, 80)
    if distributed:
        launch_training_task(data, model, Quiet("/tmp/x"), learning_rate=1e-3, weight_decay=0.0, num_workers=0, num_epochs=1, find_unused_parameters=True)
    else:
        saved = dist.get_world_size
        dist.get_world_size = lambda *a, **k: 1                      # the same loop as a single process
        try:
            launch_training_task(data, model, Quiet("/tmp/x"), learning_rate=1e-3, weight_decay=0.0, num_workers=0, num_epochs=1, gradient_accumulation_steps=2)
        finally:
            dist.get_world_size = saved
    grads = {k: (p.grad.float() if p.grad is not None else torch.zeros_like(p, dtype=torch.float32)) for k, p in model.named_parameters() if p.requires_grad}
    return grads, {k: (p.detach() - before[k]).float() for k, p in model.named_parameters() if p.requires_grad}


g_ddp, u_ddp = run(True)
g_one, u_one = run(False)
cat = lambda d: torch.cat([d[k].flatten() for k in sorted(d)])
a, b = cat(g_ddp), cat(g_one)
rel = ((a - b).norm() / b.norm()).item()
# every rank must hold the same parameters after the step
mine = cat(u_ddp)
other = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
dist.all_gather(other, mine)
same = all(torch.equal(other[0], o) for o in other[1:])
if rank == 0:
    print(json.dumps({"ranks": dist.get_world_size(), "grad_rel_l2_ddp_vs_accumulation": rel, "grad_norm": b.norm().item(), "update_norm": mine.norm().item(), "ranks_identical": same}))
dist.destroy_process_group()
sys.exit(0 if (same and rel < 2e-2 and b.norm().item() > 0 and mine.norm().item() > 0) else 1)
