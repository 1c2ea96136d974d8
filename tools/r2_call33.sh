#!/bin/bash
export KANCONV_DEBUG=1
export KANCONV_DEBUG_DEFS="-DKANCONV_DEBUG_HALFB"
O=gpurun_out/r2; mkdir -p $O
for s in 64,256,256,56; do python tools/trace_fwd.py --shape $s; done 2>&1 | cut -c1-300 | tee $O/trace_fwd_c33.txt
bash tools/lb.sh 2>&1 | tee $O/lb_c33.txt
