#!/usr/bin/env python
"""2-GPU probe: torch symmetric memory on this box (peer-mapped buffers + device-side barrier) -- the transport of the Ulysses mode."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = dist.get_world_size()
t = symm.empty(1024, 3072, dtype=torch.bfloat16, device=f"cuda:{local}")
h = symm.rendezvous(t, dist.group.WORLD)
t.fill_(rank + 1)
h.barrier()
peer = h.get_buffer((rank + 1) % N, (1024, 3072), torch.bfloat16)
print(rank, "peer value", peer[0, 0].item(), "multicast", h.has_multicast_support, "ptrs", [hex(p) for p in h.buffer_ptrs][:2], flush=True)
peer[rank * 10:(rank + 1) * 10].fill_(100 + rank)          # P2P store into the peer's memory
h.barrier()
print(rank, "rows written by peer:", t[((rank + 1) % N) * 10, 0].item(), flush=True)
# P2P write bandwidth
big = symm.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
hb = symm.rendezvous(big, dist.group.WORLD)
pb = hb.get_buffer((rank + 1) % N, (256 << 20,), torch.uint8)
src = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
hb.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): pb.copy_(src)
e1.record(); torch.cuda.synchronize()
print(rank, "p2p copy GB/s", round(10 * (256 << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1), flush=True)
b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
b0.record()
for _ in range(100): hb.barrier()
b1.record(); torch.cuda.synchronize()
print(rank, "barrier us", round(b0.elapsed_time(b1) * 10, 1), flush=True)
dist.destroy_process_group()
