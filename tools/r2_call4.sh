#!/bin/bash
set -x
mkdir -p gpurun_out/r2
python tools/norm_bench.py > gpurun_out/r2/norm_bench.txt 2>&1; cat gpurun_out/r2/norm_bench.txt
python tools/debug_vgg_parity.py > gpurun_out/r2/vgg_parity.txt 2>&1; cat gpurun_out/r2/vgg_parity.txt
python -m pytest tests/test_layers_gpu.py -m gpu -q -k "fused_norm or cluster_resident" 2>&1 | tail -n 5
