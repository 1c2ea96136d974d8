#!/bin/bash
set -x
mkdir -p gpurun_out/r2
python tools/norm_bench.py > gpurun_out/r2/norm_bench2.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kc_norm_bwd_flat_cluster_kernel|kc_instnorm_fwd_cluster_kernel|kc_instnorm_fwd_kernel|kc_norm_bwd_kernel' \
    -o gpurun_out/r2/norm_kernels env NORM_BENCH_ITERS=1 python tools/norm_bench.py > gpurun_out/r2/ncu_norm.log 2>&1
tail -n 3 gpurun_out/r2/ncu_norm.log
