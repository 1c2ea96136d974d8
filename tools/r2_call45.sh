#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
run() { python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=l['roofline']['by_kernel_ms']; print('$1', round(l['value'],1), round(l['ms_per_step'],2), 'e2e', round(l['e2e']['value'],1), 'fwd', k.get('kc_tc_kernel<fwd>'), 'wgrad', k.get('kc_wgrad_tc_kernel'), 'dgrad', k.get('kc_tc_kernel<dgrad>'), 'nb', k.get('kc_norm_bwd_flat_kernel'), 'nf', k.get('kc_instnorm_fwd_kernel'), 'clk', l['clocks']['sm_mhz'])"; }
for i in 1 2 3; do run "streaming phi stores"; done 2>&1 | tee $O/ab_c45.txt
