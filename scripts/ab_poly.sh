#!/bin/bash
# in-step A/B of the attention exp2 split (the step is power-capped, so the fastest variant in isolation need not win)
for rep in 1 2; do
for pl in 6 0; do
  TASTE_FA_POLY=$pl python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/ab_poly$pl.json
  python - $pl <<'PY'
import sys, json
m = sys.argv[1]
d = json.load(open(f"gpurun_out/ab_poly{m}.json"))
print("poly", m, round(d["value"]), round(d["ms_per_step"], 2), d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"], [(s["kernel"][:12], round(s["ms_per_step"], 1)) for s in d["stages"][:2]])
PY
done; done
