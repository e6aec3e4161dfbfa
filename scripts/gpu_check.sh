#!/bin/bash
# Staged GPU check: each stage under its own timeout, logs into gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for t in test_layernorm test_logmel test_rvq test_word_pool test_map test_attention test_gemm; do
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$t" --timeout 300 > gpurun_out/pytest_$t.log 2>&1
  echo "$t exit $?" | tee -a gpurun_out/summary.txt
  tail -3 gpurun_out/pytest_$t.log
done
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/smoke.log
