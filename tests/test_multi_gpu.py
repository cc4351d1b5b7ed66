"""Tests that need TWO GPUs (skipped on a 1-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).

CFG-parallel latency mode: one image on a pair of GPUs -- the positive branch of every denoise step on rank 0, the negative
one on rank 1, one in-place NCCL all-gather of the two predictions per step -- must give latents bit-identical to the single-GPU loop on
both ranks; and the data-parallel plumbing (NCCL weight broadcast, final gather) must hand every rank rank 0's weights."""
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
