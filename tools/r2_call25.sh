#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
timeout 120 python tools/mma_rate_2cta.py 2>&1 | tee $O/mma_rate_2cta.txt
