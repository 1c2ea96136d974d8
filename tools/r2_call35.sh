#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $O/pytest_c35.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c35.log | tail -n 6 | cut -c1-300
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench_c35.json 2> $O/bench_c35.err
tail -c 300 $O/bench_c35.err; python -c "
import json; l=json.loads(open('$O/bench_c35.json').read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['roofline']['by_kernel_ms'], l['e2e']['value'], l['roofline']['frac'], l.get('clocks'))"
