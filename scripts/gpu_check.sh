#!/bin/bash
# GPU parity tests (fail-fast) + one bench run with a per-stage summary.  bash scripts/gpu_check.sh [tag]
set -u
TAG=${1:-check}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -6 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-300
[ $rc -eq 0 ] || exit 1
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], d['clocks'])
print('alt', {k: v for k, v in (d.get('alt_precision') or {}).items() if k != 'note'}); print('parity', d.get('parity_check')); print('config5', d.get('config5')); print('ragged', d.get('ragged_config3'))
for s in d['stages'][:14]: print(s['kernel'], round(s['ms_per_step'],3), round(s['achieved'],1), round(s['frac'],3))
PY
