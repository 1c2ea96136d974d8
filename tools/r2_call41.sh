#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $O/pytest_c41.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c41.log | tail -n 6 | cut -c1-300
echo "--- fwd pairs"; timeout 300 bash tools/lb.sh 2>&1 | tee $O/lb_c41_pair.txt
echo "--- fwd single CTAs"; KANCONV_FWD_PAIR=0 timeout 300 bash tools/lb.sh 2>&1 | tee $O/lb_c41_single.txt
