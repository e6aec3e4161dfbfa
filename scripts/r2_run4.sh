#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "attention" > gpurun_out/r2_pytest_attn.log 2>&1
echo "attn exit $?"; tail -5 gpurun_out/r2_pytest_attn.log
timeout 1200 python -m pytest tests/test_gpu_tower.py tests/test_gpu_parity_big.py -m gpu -q --timeout 600 > gpurun_out/r2_pytest_tower.log 2>&1
echo "tower exit $?"; tail -5 gpurun_out/r2_pytest_tower.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config5 --no-alt-precision > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
echo "bench exit $?"; tail -c 800 gpurun_out/r2_bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], d['clocks'])
for s in d['stages'][:12]: print(s['kernel'], round(s['ms_per_step'],3), round(s['achieved'],1), round(s['frac'],3))
PY
