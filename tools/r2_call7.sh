#!/bin/bash
set -x
O=gpurun_out/r2
mkdir -p $O
python -m pytest tests/test_layers_gpu.py -m gpu -q -k "fused_norm or cluster_resident" 2>&1 | tail -n 4
python tools/norm_bench.py > $O/norm_bench4.txt 2>&1; cat $O/norm_bench4.txt
export NORM_BENCH_ITERS=1 NORM_BENCH_SHAPES="16,64,224;64,512,28"
python tools/norm_bench.py > $O/norm_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kc_norm_bwd_flat_cluster_kernel' -c 4 \
    -o /tmp/nbf python tools/norm_bench.py > $O/ncu_nbf.log 2>&1
ncu -i /tmp/nbf.ncu-rep --page raw --csv > $O/nbf_raw.csv 2>/dev/null
KERNEL_INDEX=0 python tools/ncu_top.py /tmp/nbf.ncu-rep 40 > $O/nbf_top_224.txt 2>&1
KERNEL_INDEX=4 python tools/ncu_top.py /tmp/nbf.ncu-rep 40 > $O/nbf_top_28.txt 2>&1
head -n 30 $O/nbf_top_224.txt
