#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $O/pytest_c30.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c30.log | tail -n 6 | cut -c1-300
echo "--- dgrad pairs"; timeout 300 bash tools/lb.sh 2>&1 | tee $O/lb_c30_pair.txt
echo "--- dgrad single CTAs"; KANCONV_DGRAD_PAIR=0 timeout 300 bash tools/lb.sh 2>&1 | tee $O/lb_c30_single.txt
