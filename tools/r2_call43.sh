#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 python tools/other_configs.py > $O/other_configs_c43.jsonl 2> $O/other_configs_c43.err; tail -3 $O/other_configs_c43.err; cut -c1-420 $O/other_configs_c43.jsonl
timeout 600 bash tools/bench_ngpu.sh 2 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline; cp gpurun_out/bench_2gpu.json $O/bench_2gpu_c43.json
python -c "
import json; l=json.load(open('$O/bench_2gpu_c43.json')); print('ddp_selfcheck', l.get('ddp_selfcheck'))"
