#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench_c29.json 2> $O/bench_c29.err
tail -c 300 $O/bench_c29.err; python -c "
import json; l=json.loads(open('$O/bench_c29.json').read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['roofline']['by_kernel_ms'], l['e2e']['value'], l['roofline']['frac'], l.get('clocks'))"
