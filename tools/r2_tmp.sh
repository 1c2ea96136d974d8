#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $O/pytest_c51.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c51.log | tail -n 8 | cut -c1-300
timeout 300 python tools/layer_bench.py --bwd 2>&1 | grep shape | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); k = r.get('kernels', {})
    print(r['shape'], {n: v[0] for n, v in k.items() if 'wgrad' in n or 'reduce' in n}, 'fwd+bwd', r.get('fwd_bwd_ms'))"
python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=l['roofline']['by_kernel_ms']; print('1 GPU', round(l['value'],1), round(l['ms_per_step'],2), k)"
