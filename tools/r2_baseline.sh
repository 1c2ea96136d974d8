#!/bin/bash
# round 2, call 1: state of HEAD on a fresh box - GPU tests, per-layer kernel table, full ncu captures of the three
# tensor-core kernels on the two layer shapes VERDICT names (64->64 @224 and 256->256 @56).
set -x
mkdir -p gpurun_out/r2
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2/pytest_head.log 2>&1
bash tools/lb.sh > gpurun_out/r2/lb_head.txt 2>&1
for S in 16,64,64,224 32,256,256,56; do
  T=${S//,/_}
  python tools/one_layer.py --shape $S --bwd --iters 2 > gpurun_out/r2/plain_$T.log 2>&1 && \
  ncu --set full --clock-control none --import-source on \
      -k regex:'kc_tc_kernel|kc_wgrad_tc_kernel|kc_dgrad_persistent_kernel|kc_norm_bwd_kernel|kc_instnorm_fwd_kernel' -s 5 -c 5 \
      -o gpurun_out/r2/full_$T python tools/one_layer.py --shape $S --bwd --iters 2 > gpurun_out/r2/ncu_$T.log 2>&1
done
tail -n 3 gpurun_out/r2/pytest_head.log gpurun_out/r2/ncu_*.log
cat gpurun_out/r2/lb_head.txt
ls -la gpurun_out/r2
