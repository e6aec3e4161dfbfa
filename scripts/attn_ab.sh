#!/bin/bash
# A/B of the encoder attention variants in isolation (scripts/attn_bench.py): TASTE_FA_VAR / TASTE_FA_POLY / TASTE_FA_TILES
# usage: scripts/attn_ab.sh "VAR POLY [TILES]" ...   ("x x" = the default variant; TILES e.g. 3x64)
mkdir -p gpurun_out
[ $# -eq 0 ] && set -- "x x" "64 6"
for cfg in "$@"; do
  set -- $cfg
  unset TASTE_FA_VAR TASTE_FA_POLY TASTE_FA_TILES
  if [ "$1" != "x" ]; then export TASTE_FA_VAR=$1 TASTE_FA_POLY=$2; fi
  if [ -n "${3:-}" ]; then export TASTE_FA_TILES=$3; fi
  echo "== var $1 poly $2 tiles ${3:-default}"
  timeout 120 python scripts/attn_bench.py 2>&1 | grep -v "mma.sync"
done
