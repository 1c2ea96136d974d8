#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest_c19.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c19.log | tail -n 12 | cut -c1-300
bash tools/lb.sh > $O/lb_c19.txt 2>&1; cat $O/lb_c19.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench_c19.json 2> $O/bench_c19.err
tail -c 400 $O/bench_c19.err; cut -c1-400 $O/bench_c19.json; python -c "
import json; l=json.loads(open('$O/bench_c19.json').read().strip().splitlines()[-1]); print(l['roofline']['by_kernel_ms'], l['e2e']['value'], l['roofline'].get('traffic_stale'))"
