#!/bin/bash
# ncu evidence for profiles/: (1) launch list of a short bench run (every kernel, B=16), (2) DRAM traffic + duration of this
# library's kernels at the bench batch size, (3) full captures of the three tensor-core kernels on one layer shape.
set -x
CMD="python bench.py --batch 16 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 800 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/ncu_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'kc_' --csv --log-file gpurun_out/r1_traffic_b64.csv $CMD2 > gpurun_out/ncu_traffic.log 2>&1
python tools/one_layer.py --shape 32,256,256,56 --bwd --iters 2 > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kc_tc_kernel|kc_wgrad_tc_kernel|kc_dgrad_persistent_kernel' -s 3 -c 3 -o gpurun_out/r1_tc_kernels python tools/one_layer.py --shape 32,256,256,56 --bwd --iters 2 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_launches.log gpurun_out/ncu_traffic.log gpurun_out/ncu_full.log
ls -la gpurun_out/ | tail -n 10
