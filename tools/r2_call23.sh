#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
for s in 64,256,256,56 16,64,64,224 64,512,512,14 16,3,64,224 32,64,128,112; do python tools/trace_fwd.py --shape $s; done 2>&1 | tee $O/trace_fwd_c23.txt

