#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
for s in 64,256,256,56 16,64,64,224 64,512,512,14 16,3,64,224; do python tools/trace_wgrad.py --shape $s; done 2>&1 | tee $O/trace_wgrad_c22.txt
