#!/usr/bin/env python
"""Generates tests/golden/vl.pt: outputs of the installed transformers Qwen2.5-VL (driven as the reference's wrapper drives it,
oracle/vl_oracle.py) on the small seeded configuration, in this image:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_vl.py

Stored: final hidden states of edit_forward in fp32 and bf16 (with an image: 3-D positions; text only), the sequential-position
variant (what the reference's call degrades to under transformers 5.5), greedy token ids.  tests/test_vl_oracle.py re-runs the oracle
against these (a different transformers version on another box shows up there), tests/test_text_encoder_gpu.py checks the CUDA path.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vl_oracle as VO  # noqa: E402


def main():
    import transformers
    out = {"transformers": transformers.__version__, "cases": {}}
    m32, m16 = VO.hf_model(torch.float32), VO.hf_model(torch.bfloat16)
    for name, inp in (("image", VO.inputs(True)), ("text", VO.inputs(False))):
        h32, _ = VO.edit_forward(m32, inp)
        h16, _ = VO.edit_forward(m16, inp)
        hseq, _ = VO.edit_forward(m32, inp, mrope=False)
        new32, _ = VO.generate(m32, inp, 24)
        new16, _ = VO.generate(m16, inp, 24)
        out["cases"][name] = {"hidden_fp32": h32[0].clone(), "hidden_bf16": h16[0].float().clone(), "hidden_fp32_sequential": hseq[0].clone(),
                              "tokens_fp32": new32, "tokens_bf16": new16, "T": inp["input_ids"].shape[1]}
        print(name, h32.shape, "bf16 floor", ((h16.float() - h32).norm() / h32.norm()).item(), new32.tolist()[:8])
    torch.save(out, os.path.join(ROOT, "tests", "golden", "vl.pt"))


if __name__ == "__main__":
    main()
