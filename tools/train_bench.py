#!/usr/bin/env python
"""The training-step leg of bench.py on its own (native fwd + bwd vs the reference under torch autograd, same GPU):
    python tools/train_bench.py [--layers 8] [--height 480 --width 832] [--text 512] [--rank 128]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=8)
ap.add_argument("--height", type=int, default=480)
ap.add_argument("--width", type=int, default=832)
ap.add_argument("--text", type=int, default=512)
ap.add_argument("--rank", type=int, default=128)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
torch.cuda.set_device(0)
print(json.dumps(bench.training_leg(torch.device("cuda", 0), layers=a.layers, H=a.height, W=a.width, T=a.text, rank=a.rank, iters=a.iters)))
