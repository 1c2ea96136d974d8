#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $O/pytest_c37.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c37.log | tail -n 6 | cut -c1-300
echo "--- stem from phi"; timeout 300 python tools/layer_bench.py --bwd 2>&1 | grep shape | head -2 | cut -c1-900
echo "--- stem fused"; KANCONV_STEM_FROM_PHI=0 timeout 300 python tools/layer_bench.py --bwd 2>&1 | grep shape | head -2 | cut -c1-900
