#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests/test_redzones_gpu.py -m gpu -q ) 2>&1 | tail -n 8
python tools/other_configs.py > $O/other_configs.jsonl 2> $O/other_configs.err; cat $O/other_configs.jsonl | cut -c1-900; tail -n 3 $O/other_configs.err
