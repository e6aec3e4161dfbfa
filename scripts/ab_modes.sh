#!/bin/bash
# A/B the encoder LayerNorm modes on the full bench workload
for m in 0 1 2; do
  python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --encoder-mode $m 2>/dev/null > gpurun_out/ab_mode$m.json
  python - $m <<'PY'
import sys, json
m = sys.argv[1]
d = json.load(open(f"gpurun_out/ab_mode{m}.json"))
print("mode", m, round(d["value"]), round(d["ms_per_step"], 1), d["clocks"]["sm_mhz"], [(s["kernel"][:12], round(s["ms_per_step"], 1)) for s in d["stages"][:3]])
PY
done
