#!/bin/bash
export KANCONV_DEBUG=1
O=gpurun_out/r2; mkdir -p $O
timeout 200 python tools/mma_rate_bulk.py 2>&1 | tee $O/mma_rate_bulk.txt
