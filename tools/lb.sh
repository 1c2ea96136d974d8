#!/bin/bash
# compact per-layer kernel table (ms, TFLOP/s) from tools/layer_bench.py --bwd
cd "$(dirname "$0")/.."
timeout 400 python tools/layer_bench.py --bwd "$@" 2>&1 | grep shape | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); k = r.get('kernels', {})
    g = lambda n: k.get(n, [0, 0])
    print(r['shape'], 'wgrad', g('kc_wgrad_tc_kernel'), 'dgrad', g('kc_tc_kernel<dgrad>'), 'fwd', g('kc_tc_kernel<fwd>'), 'fwd+bwd', r.get('fwd_bwd_ms'))
"
