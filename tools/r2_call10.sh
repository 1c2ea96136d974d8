#!/bin/bash
set -x
O=gpurun_out/r2
mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest_c10.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c10.log | tail -n 20
python tools/norm_bench.py > $O/norm_bench6.txt 2>&1; cat $O/norm_bench6.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench_c10.json 2> $O/bench_c10.err
tail -c 600 $O/bench_c10.err
cut -c1-2600 $O/bench_c10.json
