#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
bash tools/lb.sh 2>&1 | tee $O/lb_c28.txt
bash tools/lb.sh 2>&1 | tee $O/lb_c28b.txt
