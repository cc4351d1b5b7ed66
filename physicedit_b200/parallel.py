"""Data parallelism for batched edits: one image per GPU, one process per GPU.

The reference pipeline is strictly batch 1 and shards benchmark inference by index range across independent
processes (scripts/inference/inference_pica.py:217-220, 252-263).  The B200-native equivalent keeps that
decomposition -- every rank runs the whole denoise loop on its own images, so there is NO per-step
collective -- and adds the two exchanges that make it one job: rank 0's (LoRA-folded) weights are broadcast
once over NVLink/NVSwitch (NCCL), and the final latents are gathered to rank 0.
Works with any torch.distributed backend (`gloo` in the CPU tests, `nccl` on the GPUs).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_indices(n_items: int, rank: Optional[int] = None, world: Optional[int] = None) -> List[int]:
    """Round-robin image -> rank assignment: rank r takes r, r+world, ... (every rank gets ceil or floor of n/world)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return list(range(rank, n_items, world))


def broadcast_weights(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 1 << 30) -> int:
    """Broadcast every parameter / buffer of `module` from `src`.  Tensors are sent in place (no staging copy);
    views into a shared storage (the engine's fused QKV buffers) are de-duplicated by storage pointer.  Returns the
    number of bytes this rank sent or received."""
    if not is_dist():
        return 0
    seen, total = set(), 0
    tensors = [p.data for p in module.parameters()] + [b for b in module.buffers()]
    for t in tensors:
        key = (t.untyped_storage().data_ptr(), t.storage_offset(), t.numel())
        if key in seen or t.numel() == 0:
            continue
        seen.add(key)
        if t.is_contiguous():
            dist.broadcast(t, src=src)
        else:
            c = t.contiguous()
            dist.broadcast(c, src=src)
            t.copy_(c)
        total += t.numel() * t.element_size()
    eng = getattr(getattr(module, "dit", module), "_engine", None)
    if eng is not None:
        eng.invalidate()
    return total


def gather_latents(latents: torch.Tensor, dst: int = 0) -> Optional[List[torch.Tensor]]:
    """Final gather of the per-rank results ([n_local,16,h8,w8], equal shapes) to `dst`; returns the list on dst, None elsewhere."""
    if not is_dist():
        return [latents]
    world = dist.get_world_size()
    if dist.get_backend() == "nccl":
        out = [torch.empty_like(latents) for _ in range(world)]
        dist.all_gather(out, latents.contiguous())
        return out if dist.get_rank() == dst else None
    bufs = [torch.empty_like(latents) for _ in range(world)] if dist.get_rank() == dst else None
    dist.gather(latents.contiguous(), bufs, dst=dst)
    return bufs


def run_batched_edits(pipe, requests: Sequence[dict], *, height: int, width: int, num_inference_steps: int, cfg_scale: float = 4.0):
    """Each request: dict(latents, inputs_posi, inputs_nega, edit_latents).  Rank r processes requests r, r+world, ...;
    returns (indices, latents) for the local shard.  Callers gather with `gather_latents`."""
    mine = shard_indices(len(requests)) if is_dist() else list(range(len(requests)))
    outs = []
    for i in mine:
        r = requests[i]
        outs.append(pipe.denoise(r["latents"], r["inputs_posi"], r["inputs_nega"], r.get("edit_latents"), height=height, width=width,
                                 num_inference_steps=num_inference_steps, cfg_scale=cfg_scale))
    return mine, outs


# ---- CFG-parallel latency mode: one image on a pair of GPUs ---------------------------------------------------
def make_cfg_pairs():
    """Splits the world into consecutive pairs (0,1), (2,3), ... and returns (this rank's pair group, pair index, number of pairs).
    Every rank must call it (new_group is collective).  Assign the group to `pipe.cfg_parallel_group`."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world % 2:
        raise ValueError(f"CFG-parallel needs an even number of ranks (world size {world})")
    mine = None
    for i in range(world // 2):
        g = dist.new_group([2 * i, 2 * i + 1])
        if rank // 2 == i:
            mine = g
    return mine, rank // 2, world // 2


def make_cfg_sequence_groups():
    """One image on ALL ranks, both latency modes at once: the world is cut into two halves, each half runs ONE CFG branch of every step
    sequence-parallel (ulysses.py), and rank i of the first half is paired with rank i of the second for the per-step exchange of the two
    predictions.  Returns (this rank's sequence-parallel group, its CFG pair group).  Every rank must call it (new_group is collective).
        sp, pair = parallel.make_cfg_sequence_groups(); pipe.enable_sequence_parallel(sp); pipe.cfg_parallel_group = pair"""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world % 2:
        raise ValueError(f"needs an even number of ranks (world size {world})")
    half = world // 2
    halves = [dist.new_group(list(range(0, half))), dist.new_group(list(range(half, world)))]
    mine_pair = None
    for i in range(half):
        g = dist.new_group([i, i + half])
        if rank % half == i:
            mine_pair = g
    return halves[rank // half], mine_pair


def exchange_cfg_predictions(vp: torch.Tensor, vn: torch.Tensor, rank_in_pair: int, group) -> None:
    """After rank 0 of the pair filled `vp` (positive branch) and rank 1 filled `vn`, makes both tensors valid on both ranks.
    vp / vn are the two halves of one contiguous [2, ...] buffer (pipeline.denoise allocates them so): with NCCL this is a single
    in-place all-gather (the per-step exchange of this mode: 2 x 512 KiB at 1024^2 against ~150 ms of compute per branch)."""
    mine = vp if rank_in_pair == 0 else vn
    adjacent = (vp.is_contiguous() and vn.is_contiguous() and vp.untyped_storage().data_ptr() == vn.untyped_storage().data_ptr()
                and vn.storage_offset() == vp.storage_offset() + vp.numel())
    if adjacent and dist.get_backend(group) == "nccl":
        both = torch.as_strided(vp, (2 * vp.numel(),), (1,), vp.storage_offset())
        dist.all_gather_into_tensor(both, mine.reshape(-1), group=group)
        return
    outs = [torch.empty_like(vp), torch.empty_like(vn)]
    dist.all_gather(outs, mine.contiguous(), group=group)
    (vn if rank_in_pair == 0 else vp).copy_(outs[1 - rank_in_pair])
