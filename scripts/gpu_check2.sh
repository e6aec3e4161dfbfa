#!/bin/bash
# Full GPU pass: parity tests, bench line, ncu launch list of one step.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest_gpu exit $?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:taste -s 783 -c 261 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" | tee -a gpurun_out/summary.txt
