#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -rs -s 2>&1 | grep -v "^$" > $O/gputest_log.txt; tail -3 $O/gputest_log.txt
timeout 500 python bench.py --steps 10 --warmup 3 > $O/bench_1gpu_head.json 2> $O/bench_1gpu_head.err; tail -c 1500 $O/bench_1gpu_head.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_head.json 2> $O/bench_ref_head.err; tail -c 600 $O/bench_ref_head.json
bash tools/ncu_round.sh r2 > $O/ncu_round.log 2>&1; tail -3 $O/ncu_round.log
