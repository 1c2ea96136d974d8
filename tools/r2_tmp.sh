#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -rs -s 2>&1 | grep -v "^$" > $O/gputest_log.txt; tail -3 $O/gputest_log.txt
bash tools/ncu_round.sh r2 > $O/ncu_round.log 2>&1; tail -3 $O/ncu_round.log
