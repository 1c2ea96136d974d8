#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 400 python bench.py > $O/bench_1gpu_3d.json 2> $O/bench_1gpu_3d.log; tail -c 3000 $O/bench_1gpu_3d.json
