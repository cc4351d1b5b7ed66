#!/usr/bin/env bash
# Builder-run sweep over BASELINE configs #4 / #5 (and 256^2 for the launch-bound question): one bench line per resolution -> gpurun_out/r02_resolution_sweep.jsonl
set -u
out=gpurun_out/r02_resolution_sweep.jsonl; : > $out
for res in 256 512 1536 2048; do
  steps=4; [ $res -ge 1536 ] && steps=2
  timeout 600 python bench.py --resolution $res --steps $steps --warmup 3 --no-cpu-baseline --no-vae --no-text-encoder --no-stock-gpu 2>/dev/null | tail -1 >> $out
done
python - <<'PY'
import json
for l in open("gpurun_out/r02_resolution_sweep.jsonl"):
    b = json.loads(l)
    print(b["config"]["workload"][:12], "value", b["value"], "e2e", b["e2e"]["value"], "ms", b["ms_per_step"], "whole_step_frac", b["roofline"]["whole_step_frac"],
          "attn", b["roofline_by_kernel"]["attention"]["frac"], b["roofline_by_kernel"]["attention"]["share_of_step"], "launches", b["gpu_launches"])
PY
