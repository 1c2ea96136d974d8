#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
for n in 0 1 2; do echo "--- KANCONV_FWD_NSUB=$n"; KANCONV_FWD_NSUB=$n timeout 300 bash tools/lb.sh 2>&1 | cut -d' ' -f1-4,9-14; done | tee $O/lb_c44_nsub.txt
