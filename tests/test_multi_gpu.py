"""Tests that need TWO GPUs (skipped on a 1-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).

CFG-parallel latency mode: one image on a pair of GPUs -- the positive branch of every denoise step on rank 0, the negative
one on rank 1, one in-place NCCL all-gather of the two predictions per step -- must give latents bit-identical to the single-GPU loop on
both ranks; the sequence-parallel (Ulysses) forward -- one image on N GPUs, rows split across the ranks, attention head-parallel, the two
all-to-alls fused into the QKV GEMM's and the attention kernel's epilogues as NVLink peer stores -- must give a velocity bit-identical to the
single-GPU forward on every rank; and the data-parallel plumbing (NCCL weight broadcast, final gather) must hand every rank rank 0's weights."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_cfg_parallel_pair_is_bit_identical_to_the_single_gpu_loop():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29547",
           os.path.join(ROOT, "tools", "cfg_parallel_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit-identical to the single-GPU loop on both ranks: True" in r.stdout


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("resolution,layers", [(512, 2), (1024, 3)])
def test_sequence_parallel_forward_is_bit_identical_to_the_single_gpu_forward(resolution, layers):
    """tools/ulysses_check.py exits 0 only when two denoise steps (both CFG branches, adapter included) split across the ranks leave latents that
    are torch.equal to the single-GPU ones on EVERY rank."""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29549",
           os.path.join(ROOT, "tools", "ulysses_check.py"), "--resolution", str(resolution), "--layers", str(layers), "--steps", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["bit_identical_on_all_ranks"] and res["finite"] and res["ranks"] == 2, res


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_training_step_under_ddp_matches_gradient_accumulation():
    """launch_training_task on 2 ranks (DDP, NCCL gradient all-reduce over the native backward) leaves identical parameters on both
    ranks, moved like one process accumulating the same two samples (tools/train_ddp_check.py)."""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29553",
           os.path.join(ROOT, "tools", "train_ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ranks_identical"] and res["grad_rel_l2_ddp_vs_accumulation"] < 2e-2, res


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs four GPUs")
def test_cfg_branch_per_half_with_sequence_parallel_halves_is_bit_identical():
    """parallel.make_cfg_sequence_groups(): two halves of the world, one CFG branch each, sequence-parallel inside a half: same latents as the single-GPU loop on
    every rank (tools/ulysses_check.py --cfg-split)."""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4", "--master-addr", "127.0.0.1", "--master-port", "29555",
           os.path.join(ROOT, "tools", "ulysses_check.py"), "--resolution", "512", "--layers", "2", "--steps", "2", "--cfg-split"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["bit_identical_on_all_ranks"] and res["finite"] and res["ranks"] == 4, res
