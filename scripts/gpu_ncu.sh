#!/bin/bash
# ncu evidence for the current kernels: launch list of ONE bench step + --set full captures of one launch of every
# kernel family, exported as raw-page CSVs (small; the .ncu-rep files of GEMM and attention are kept as well).
#   bash scripts/gpu_ncu.sh <tag>        then, in the build container:  python scripts/summarize_ncu.py <tag>
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-alt-precision --no-config5"
L=232            # kernels of this library per step (bench.py: gpu_launches / steps)
K='regex:assemble_tokens|attention_|cast_bf16|embed_kernel|gemm_bf16|layernorm_kernel|logmel|rvq_|word_pool'
timeout 600 $CMD > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err || { echo "plain run failed"; tail -5 gpurun_out/bench_short.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s $((3 * L)) -c $L --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
# name / kernel regex / launches of that regex to skip (3 warm-up steps + position inside the 4th) / count
#   gemm: 147 per step (log-mel DFT, conv1, conv2, then qkv / out_proj / fc1 / fc2 of layer 0)
#   attention_tcgen05: 34 per step (32 encoder layers, then the aggregator's two ragged cross-attentions)
for spec in "gemm gemm_bf16 444 4" "attn attention_tcgen05 110 1" "attn_ragged attention_tcgen05 134 1" \
            "rvq_search rvq_search 3 1" "rvq_sgemm rvq_sgemm 6 2" "layernorm layernorm_kernel 130 1" "attn_mma attention_mma 6 1"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 \
      -o gpurun_out/prof_${1}_$TAG -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?"
  ncu -i gpurun_out/prof_${1}_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_${1}_$TAG.csv 2>/dev/null
done
ls -la gpurun_out/*$TAG*
# keep the transfer small (gpurun pulls at most 64 MiB): the raw CSVs carry what summarize_ncu.py needs
for n in rvq_sgemm layernorm attn_mma; do rm -f gpurun_out/prof_${n}_$TAG.ncu-rep; done
