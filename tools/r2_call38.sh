#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $O/pytest_c38.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c38.log | tail -n 6 | cut -c1-300
echo "--- merged"; timeout 300 python tools/layer_bench.py --bwd 2>&1 | grep shape | head -1 | cut -c1-500
echo "--- not merged"; KANCONV_WGRAD_MERGE=0 timeout 300 python tools/layer_bench.py --bwd 2>&1 | grep shape | head -1 | cut -c1-500
