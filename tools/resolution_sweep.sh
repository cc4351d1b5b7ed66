#!/usr/bin/env bash
# bench.py at the resolutions of BASELINE.json configs #4/#5 (edit image fixed at 4096 tokens); one JSON line per resolution
for r in 512 1536 2048; do
  python bench.py --resolution $r --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > /tmp/sweep_$r.json
  R=$r python - <<'PY'
import json, os
r = os.environ["R"]
d = json.load(open(f"/tmp/sweep_{r}.json"))
print(json.dumps({"res": int(r), "steps_per_s": d["value"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"]["value"],
                  "tflops": d["config"]["achieved_tflops_per_gpu"], "attention_share": d["kernel_time_share"]["attention"],
                  "sm_mhz": d["clocks"]["sm_mhz"]}))
PY
done
