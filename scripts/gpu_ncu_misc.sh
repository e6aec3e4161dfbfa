#!/bin/bash
# ncu --set full of the HBM-bound kernels of one bench step (one launch each, from the 4th step), exported as small
# raw-page CSVs (the .ncu-rep files stay on the box: gpurun pulls at most 64 MiB).
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
for spec in "layernorm_kernel 110" "logmel_tile 3" "logmel_finish 3" "rvq_encode 3"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none -k regex:$1 -s $2 -c 1 -o /tmp/prof_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/raw_${1}_$TAG.csv 2>/dev/null
done
ls -la gpurun_out/raw_*_$TAG.csv
