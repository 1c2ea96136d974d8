#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + full captures of the tensor-core kernels.
set -x
CMD="python bench.py --batch 16 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 800 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
python tools/one_layer.py --shape 32,256,256,56 --bwd --iters 2 > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kc_tc_kernel|kc_wgrad_tc_kernel' -s 3 -c 3 -o gpurun_out/r1_tc_kernels python tools/one_layer.py --shape 32,256,256,56 --bwd --iters 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
ls -la gpurun_out/ | tail -8
