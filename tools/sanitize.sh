#!/bin/bash
# one compute-sanitizer tool per gpurun call (B200_PROFILING.md): tools/sanitize.sh memcheck|racecheck|synccheck [round tag]
TOOL=${1:-memcheck}
TAG=${2:-r2}
O=gpurun_out/$TAG
mkdir -p $O
set -x
python tools/sanitize_target.py > $O/sanitize_plain.log 2>&1 || { tail -n 20 $O/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 50 python tools/sanitize_target.py > $O/sanitizer_$TOOL.log 2>&1
echo "exit code $?" >> $O/sanitizer_$TOOL.log
tail -n 25 $O/sanitizer_$TOOL.log
