#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tower.py tests/test_gpu_ingest.py -m gpu -q --timeout 600 -x > gpurun_out/r2_pytest_tower.log 2>&1
echo "tower exit $?"; tail -5 gpurun_out/r2_pytest_tower.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
echo "bench exit $?"; tail -c 1500 gpurun_out/r2_bench_a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_a.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], d['clocks'])
print('parity', d['parity_check']); print('config5', d['config5']); print('cpu', d['cpu_baseline'])
for s in d['stages'][:8]: print(s['kernel'], round(s['ms_per_step'],2), round(s['achieved'],1), round(s['frac'],3))
PY
timeout 900 python bench.py --workload corpus --utts 2048 > gpurun_out/r2_corpus_n1_2048.json 2> gpurun_out/r2_corpus_n1.err
echo "corpus exit $?"; tail -c 1500 gpurun_out/r2_corpus_n1.err; cat gpurun_out/r2_corpus_n1_2048.json
