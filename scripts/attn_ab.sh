#!/bin/bash
# A/B of the encoder attention variants in isolation (scripts/attn_bench.py): TASTE_FA_VAR / TASTE_FA_POLY
# usage: scripts/attn_ab.sh "VAR POLY" ...   ("x x" = the default variant)
mkdir -p gpurun_out
[ $# -eq 0 ] && set -- "x x" "122 5"
for cfg in "$@"; do
  set -- $cfg
  if [ "$1" = "x" ]; then unset TASTE_FA_VAR TASTE_FA_POLY; else export TASTE_FA_VAR=$1 TASTE_FA_POLY=$2; fi
  echo "== var $1 poly $2"
  timeout 120 python scripts/attn_bench.py 2>&1 | grep -v "mma.sync"
done
