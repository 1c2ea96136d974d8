#!/bin/bash
# re-entry baseline: GPU tests, per-layer table, 1-GPU bench line at HEAD
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest_c21.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c21.log | tail -n 6 | cut -c1-300
bash tools/lb.sh > $O/lb_c21.txt 2>&1; cat $O/lb_c21.txt
python bench.py --steps 10 --warmup 3 > $O/bench_c21.json 2> $O/bench_c21.err
tail -c 300 $O/bench_c21.err; python -c "
import json; l=json.loads(open('$O/bench_c21.json').read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['roofline']['by_kernel_ms'], l['e2e']['value'], l.get('cpu_baseline'))"
