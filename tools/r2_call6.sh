#!/bin/bash
set -x
O=gpurun_out/r2
mkdir -p $O
python tools/norm_bench.py > $O/norm_bench3.txt 2>&1; cat $O/norm_bench3.txt
export NORM_BENCH_ITERS=1 NORM_BENCH_SHAPES="16,64,224;64,512,28"
python tools/norm_bench.py > $O/norm_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kc_norm_bwd_flat_cluster_kernel|kc_instnorm_fwd_cluster_kernel|kc_instnorm_fwd_kernel' \
    -o /tmp/norm_kernels python tools/norm_bench.py > $O/ncu_norm.log 2>&1
ncu -i /tmp/norm_kernels.ncu-rep --page raw --csv > $O/norm_kernels_raw.csv 2>/dev/null
ncu -i /tmp/norm_kernels.ncu-rep --page source --csv > /tmp/norm_src.csv 2>/dev/null
python tools/ncu_top.py /tmp/norm_kernels.ncu-rep 25 > $O/norm_top0.txt 2>&1
for i in 1 2 3 4 5 6 7; do KERNEL_INDEX=$i python tools/ncu_top.py /tmp/norm_kernels.ncu-rep 25 > $O/norm_top$i.txt 2>&1; done
ls -la /tmp/norm_kernels.ncu-rep $O | tail -n 15
