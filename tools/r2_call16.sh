#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest_c16.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c16.log | tail -n 12 | cut -c1-250
python tools/norm_bench.py > $O/norm_bench7.txt 2>&1; cat $O/norm_bench7.txt
python tools/other_configs.py > $O/other_configs2.jsonl 2> $O/other_configs2.err; cut -c1-600 $O/other_configs2.jsonl
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench_c16.json 2> $O/bench_c16.err
tail -c 400 $O/bench_c16.err; cut -c1-2700 $O/bench_c16.json
