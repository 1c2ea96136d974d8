#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
echo "--- clusters of 16 x 100 KB"; NORM_BENCH_SHAPES="64,64,224;64,128,112" timeout 200 python tools/norm_bench.py 2>&1 | tail -4 | cut -c1-400
echo "--- clusters of 8 x 200 KB"; KANCONV_NORM_BWD_CS16=0 NORM_BENCH_SHAPES="64,64,224;64,128,112" timeout 200 python tools/norm_bench.py 2>&1 | tail -4 | cut -c1-400
timeout 300 python -m pytest tests -m gpu -q -x -k "norm or layer or vgg or redzone" 2>&1 | tail -3
