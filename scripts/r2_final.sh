#!/bin/bash
# final evidence of the round on N GPUs: bash scripts/r2_final.sh N
set -u
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  bash scripts/gpu_check.sh r2final || exit 1
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2final_reference_arm.json 2> gpurun_out/r2final_reference_arm.err
  echo "reference arm exit $?"; cat gpurun_out/r2final_reference_arm.json | cut -c1-400
else
  for n in $N; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/r2final_bench_n$n.json 2> gpurun_out/r2final_bench_n$n.err
    echo "bench n=$n exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2final_bench_n$n.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['alt_precision']['value'] if d.get('alt_precision') else None)"
  done
fi
