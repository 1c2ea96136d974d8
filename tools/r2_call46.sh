#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $O/pytest_c46.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c46.log | tail -n 8 | cut -c1-300
timeout 300 bash tools/lb.sh 2>&1 | tee $O/lb_c46.txt
