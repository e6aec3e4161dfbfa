#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 -k "gemm" > gpurun_out/pytest_gemm.log 2>&1
echo "pytest gemm exit $?"; tail -5 gpurun_out/pytest_gemm.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "not gemm" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], d['clocks'])
for s in d['stages'][:6]: print(s['kernel'], round(s['ms_per_step'],2), round(s['achieved'],1), round(s['frac'],3))
PY
