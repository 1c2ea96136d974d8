#!/bin/bash
set -x
mkdir -p gpurun_out/r2
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/r2/pytest_c3.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2/pytest_c3.log | tail -n 30
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2/bench_c3.json 2> gpurun_out/r2/bench_c3.err
tail -c 1000 gpurun_out/r2/bench_c3.err
cut -c1-2500 gpurun_out/r2/bench_c3.json
