#!/usr/bin/env python
"""Text-encoder leg of bench.py alone (7B-config Qwen2.5-VL, random weights): decode ms/token, per-kernel breakdown, roofline.
   python tools/te_probe.py [new_tokens]          PE_TE_EAGER=1: a few eager decode steps only (for an ncu launch list)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
if os.environ.get("PE_TE_EAGER") == "1":
    import math
    from physicedit_b200.text_encoder import QwenImageTextEncoder, VLConfig
    dev = torch.device("cuda", 0)
    cfg = VLConfig(v_depth=1)
    with torch.device("meta"):
        te = QwenImageTextEncoder(cfg)
    g = torch.Generator(device=dev).manual_seed(1)
    sd = {k: ((torch.rand(v.shape, generator=g, device=dev) * 2 - 1) * (1.0 / math.sqrt(max(math.prod(v.shape[1:]), 1))) if v.dim() >= 2 else torch.ones(v.shape, device=dev)).to(torch.bfloat16)
          for k, v in te.state_dict().items()}
    te.load_state_dict(sd, assign=True)
    te.cfg.eos_token_id = -1
    te.use_cuda_graph = False
    ids = torch.randint(1000, 100000, (1, 356))
    nb = int(os.environ.get("PE_TE_BATCH", "1"))
    te.generate_batch([dict(input_ids=ids.to(dev), attention_mask=torch.ones_like(ids).to(dev))] * nb, max_new_tokens=6)
    torch.cuda.synchronize()
    print("done")
else:
    print(json.dumps(bench.text_encoder_leg(torch.device("cuda", 0), 300.0, 13.0, new_tokens=int(sys.argv[1]) if len(sys.argv) > 1 else 160), indent=1))
