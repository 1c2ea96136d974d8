#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
for s in 64,256,256,56 16,64,64,224; do
echo "=== $s pair"; timeout 120 python tools/trace_dgrad.py --shape $s | tail -5
echo "=== $s single"; KANCONV_DGRAD_PAIR=0 timeout 120 python tools/trace_dgrad.py --shape $s | tail -5
done 2>&1 | cut -c1-400 | tee $O/trace_dgrad_c31.txt
