#!/bin/bash
# round 2, call 2: new parity tests + bench with the new keys
set -x
mkdir -p gpurun_out/r2
( time python -m pytest tests -m gpu -x -q -s ) > gpurun_out/r2/pytest_c2.log 2>&1
tail -n 25 gpurun_out/r2/pytest_c2.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench_c2.json 2> gpurun_out/r2/bench_c2.err
tail -c 1500 gpurun_out/r2/bench_c2.err
cut -c1-3000 gpurun_out/r2/bench_c2.json
