#!/bin/bash
set -x
O=gpurun_out/r2
mkdir -p $O
python -m pytest tests/test_layers_gpu.py -m gpu -q -k "fused_norm or cluster_resident" 2>&1 | tail -n 4
python tools/norm_bench.py > $O/norm_bench5.txt 2>&1; cat $O/norm_bench5.txt
export NORM_BENCH_ITERS=1 NORM_BENCH_SHAPES="16,64,224;64,512,28"
python tools/norm_bench.py > $O/norm_plain.txt 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'kc_norm_bwd_flat_cluster_kernel|kc_norm_bwd_kernel|kc_dz_flat' --csv --log-file $O/nbf_metrics.csv python tools/norm_bench.py > $O/ncu_nbf.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2/nbf_metrics.csv')))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
col={h:i for i,h in enumerate(rows[hi])}
cur=None
for r in rows[hi+2:]:
    if len(r)!=len(rows[hi]): continue
    key=(r[col['ID']], r[col['Kernel Name']][:60])
    if key!=cur: print(); print(key, end=' '); cur=key
    print(r[col['Metric Name']].split('.')[0][-22:], r[col['Metric Value']], r[col['Metric Unit']], end=' | ')
print()
PY
