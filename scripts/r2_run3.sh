#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity_report.json
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?"; grep smoke gpurun_out/r2_smoke.log; tail -3 gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests/test_gpu_parity_big.py -m gpu -q --timeout 900 -s > gpurun_out/r2_pytest_parity.log 2>&1
echo "parity exit $?"; grep -E "^(batched|solo|default init) |passed|failed|Error" gpurun_out/r2_pytest_parity.log | cut -c1-700
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
echo "bench exit $?"; tail -c 800 gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_b.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], d['clocks'])
print('alt', d['alt_precision'])
PY
