#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt
./build/exp2_rate > gpurun_out/r2_exp2_rate.txt 2>&1; cat gpurun_out/r2_exp2_rate.txt
timeout 1500 python -m pytest tests/test_gpu_parity_big.py -m gpu -q --timeout 900 -s > gpurun_out/r2_pytest_parity.log 2>&1
echo "parity exit $?"; grep -E "^(batched|solo|default init):|passed|failed|Error" gpurun_out/r2_pytest_parity.log | cut -c1-1500
timeout 900 python scripts/attn_yardstick.py > gpurun_out/r2_attn_yardstick.json 2> gpurun_out/r2_attn_yardstick.err
echo "yardstick exit $?"; cat gpurun_out/r2_attn_yardstick.err | tail -8
timeout 1200 python -m pytest tests/test_gpu_tower.py tests/test_gpu_ingest.py -m gpu -q --timeout 600 > gpurun_out/r2_pytest_tower.log 2>&1
echo "tower exit $?"; tail -5 gpurun_out/r2_pytest_tower.log
