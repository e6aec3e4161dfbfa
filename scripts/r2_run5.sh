#!/bin/bash
set -u
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 60 -x -k "rvq" > gpurun_out/r2_pytest_rvq.log 2>&1
rc=$?; echo "rvq exit $rc"; tail -25 gpurun_out/r2_pytest_rvq.log | cut -c1-300
[ $rc -eq 0 ] || exit 1
timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_gpu_properties.py -m gpu -q --timeout 120 -x > gpurun_out/r2_pytest_tower.log 2>&1
rc=$?; echo "tower exit $rc"; tail -8 gpurun_out/r2_pytest_tower.log | cut -c1-300
[ $rc -eq 0 ] || exit 1
timeout 600 python -m pytest tests/test_gpu_parity_big.py -m gpu -q --timeout 300 -x > gpurun_out/r2_pytest_parity.log 2>&1
rc=$?; echo "parity exit $rc"; tail -4 gpurun_out/r2_pytest_parity.log | cut -c1-300
[ $rc -eq 0 ] || exit 1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config5 --no-alt-precision > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
echo "bench exit $?"; tail -c 600 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_d.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], d['clocks'])
for s in d['stages'][:14]: print(s['kernel'], round(s['ms_per_step'],3), round(s['achieved'],1), round(s['frac'],3))
PY
