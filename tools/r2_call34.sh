#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
for s in 64,256,256,56; do python tools/trace_fwd.py --shape $s; done 2>&1 | cut -c1-300 | tee $O/trace_fwd_c34.txt
bash tools/lb.sh 2>&1 | tee $O/lb_c34.txt
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $O/pytest_c34.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c34.log | tail -n 6 | cut -c1-300
