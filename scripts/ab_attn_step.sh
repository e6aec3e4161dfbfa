#!/bin/bash
# In-step A/B of attention variants (the step is power-capped: what counts is the step time, not the kernel alone).
# usage: scripts/ab_attn_step.sh "VAR POLY TILES" ...
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  unset TASTE_FA_VAR TASTE_FA_POLY TASTE_FA_TILES
  if [ "$1" != "x" ]; then export TASTE_FA_VAR=$1 TASTE_FA_POLY=$2 TASTE_FA_TILES=$3; fi
  python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/ab_step.json
  python - "$cfg" <<'PY'
import json, sys
d = json.load(open('gpurun_out/ab_step.json'))
st = {s['kernel']: s['ms_per_step'] for s in d['stages']}
print(f"{sys.argv[1]:14s} step {d['ms_per_step']:.2f} ms  gemm {st['gemm_bf16_tcgen05']:.2f}  attention {st['attention_encoder']:.2f}  clock {d['clocks']['sm_mhz']:.0f} MHz  power {d['clocks']['power_w_max']:.0f} W")
PY
done
