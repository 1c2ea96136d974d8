#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest_c17.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c17.log | tail -n 12 | cut -c1-250
python tools/norm_bench.py > $O/norm_bench8.txt 2>&1; cat $O/norm_bench8.txt
python tools/other_configs.py > $O/other_configs3.jsonl 2> $O/other_configs3.err; cut -c1-700 $O/other_configs3.jsonl
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench_c17.json 2> $O/bench_c17.err
tail -c 400 $O/bench_c17.err; cut -c1-400 $O/bench_c17.json; python -c "
import json; l=json.loads(open('$O/bench_c17.json').read().strip().splitlines()[-1]); print(l['roofline']['by_kernel_ms'], l['e2e']['value'])"
