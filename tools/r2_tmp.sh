#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 140 python bench.py --steps 10 --warmup 3 > $O/bench_1gpu_final.json 2> $O/bench_1gpu_final.err; tail -c 600 $O/bench_1gpu_final.json
