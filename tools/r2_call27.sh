#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
echo "--- pairs"; bash tools/lb.sh 2>&1 | tee $O/lb_c27_pair.txt
echo "--- single CTAs"; KANCONV_WGRAD_PAIR=0 bash tools/lb.sh 2>&1 | tee $O/lb_c27_single.txt
for s in 64,256,256,56 16,64,64,224; do timeout 120 python tools/trace_wgrad.py --shape $s; done 2>&1 | tee $O/trace_wgrad_c27.txt
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $O/pytest_c27.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c27.log | tail -n 6 | cut -c1-300
