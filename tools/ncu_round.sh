#!/bin/bash
# ncu evidence for profiles/ (round tag = $1, default r2; raw outputs under gpurun_out/<tag>/, summarised by
# tools/summarize_profiles.py): (1) launch list of a short bench run (every kernel, B=16), (2) DRAM traffic + duration of this
# library's kernels at the bench batch size, (3) full captures of the tensor-core and normalisation kernels on the two layer
# shapes named by the round-1 verdict (64->64 @224 and 256->256 @56).  Each ncu command runs right after the same command
# exited 0 without ncu.
TAG=${1:-r2}
O=gpurun_out/$TAG
mkdir -p $O
set -x
python -c "import kanconv_b200 as K; print(K._lib.source_hash())" > $O/source_hash.txt
CMD="python bench.py --batch 16 --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline"
$CMD > $O/plain_b16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 800 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline"
$CMD2 > $O/plain_b64.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'kc_' --csv --log-file $O/traffic_b64.csv $CMD2 > $O/ncu_traffic.log 2>&1
for S in 16,64,64,224 32,256,256,56; do
  T=${S//,/_}
  python tools/one_layer.py --shape $S --bwd --iters 2 > $O/plain_$T.log 2>&1 && \
  ncu --set full --clock-control none --import-source on \
      -k regex:'kc_tc_kernel|kc_wgrad_tc_kernel|kc_dgrad_persistent_kernel|kc_norm_bwd|kc_instnorm_fwd' -s 5 -c 5 \
      -o $O/full_$T python tools/one_layer.py --shape $S --bwd --iters 2 > $O/ncu_full_$T.log 2>&1
done
tail -n 2 $O/ncu_launches.log $O/ncu_traffic.log $O/ncu_full_*.log
ls -la $O | tail -n 12
