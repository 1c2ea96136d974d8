#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
( time python -m pytest tests -m gpu -q -s ) > $O/pytest_c13.log 2>&1
grep -E "passed|failed|FAILED|logits err|gradients over|Error" $O/pytest_c13.log | cut -c1-300 | tail -n 30
bash tools/sanitize.sh memcheck r2 2>&1 | tail -n 30
