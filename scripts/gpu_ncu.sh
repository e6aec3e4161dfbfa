#!/bin/bash
# ncu evidence for the final kernels: launch list of one bench step + full captures (GEMM pair kernel, attention).
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
K='regex:assemble_tokens|attention_|cast_bf16|embed_kernel|gemm_bf16|layernorm_kernel|logmel|rvq_|word_pool'
timeout 600 $CMD > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 690 -c 230 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 444 -c 4 \
    -o gpurun_out/prof_gemm_$TAG -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 100 -c 1 \
    -o gpurun_out/prof_attn_$TAG -f $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?"
# (the HBM-bound kernels: scripts/gpu_ncu_misc.sh - their reports stay on the box, gpurun pulls at most 64 MiB)
ls -la gpurun_out/*$TAG*.ncu-rep
