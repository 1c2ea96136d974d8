#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x -k "depthwise" 2>&1 | tail -4
timeout 300 python tools/dw_bench.py 64 2>&1 | tee $O/dw_bench_v2.txt | tail -11 | cut -c1-420
