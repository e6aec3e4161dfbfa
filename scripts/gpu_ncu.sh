#!/bin/bash
# ncu evidence: launch list of one step + full captures of the top kernels.  Usage: gpu_ncu.sh <tag>
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
K='regex:assemble_tokens|attention_|cast_bf16|embed_kernel|gemm_bf16|layernorm_kernel|logmel|rvq_|word_pool'
timeout 600 $CMD > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 783 -c 261 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 444 -c 4 \
    -o gpurun_out/prof_gemm_$TAG -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_ -s 100 -c 1 \
    -o gpurun_out/prof_attn_$TAG -f $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?"
ls -la gpurun_out/*.ncu-rep
