#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/pytest_c20.log 2>&1
grep -E "passed|failed|FAILED|Error" $O/pytest_c20.log | tail -n 6 | cut -c1-300
KANCONV_SAVE_PHI=0 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $O/bench_nophi.json 2> $O/bench_nophi.err
python -c "
import json; l=json.loads(open('$O/bench_nophi.json').read().strip().splitlines()[-1]); print('SAVE_PHI=0', l['value'], l['ms_per_step'], l['roofline']['by_kernel_ms'])"
bash tools/ncu_round.sh r2
