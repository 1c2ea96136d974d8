#!/bin/bash
# N = 2: exposed cost of the gradient all-reduce and what moves it (tools/ddp_timeline.py)
O=gpurun_out/r2; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/ddp_timeline.py; }
echo "== default NCCL settings" > $O/ddp2.txt
run 29511 >> $O/ddp2.txt 2> $O/ddp2.err
echo "== NCCL_MAX_CTAS=8" >> $O/ddp2.txt
NCCL_MAX_CTAS=8 run 29512 >> $O/ddp2.txt 2>> $O/ddp2.err
echo "== NCCL_MAX_CTAS=2" >> $O/ddp2.txt
NCCL_MAX_CTAS=2 run 29513 >> $O/ddp2.txt 2>> $O/ddp2.err
cat $O/ddp2.txt | cut -c1-700
tail -n 5 $O/ddp2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
cut -c1-400 $O/bench_2gpu.json; python -c "
import json; l=json.loads(open('$O/bench_2gpu.json').read().strip().splitlines()[-1]); print(l.get('ddp_selfcheck'))"
python -m pytest tests/test_ddp_gpu.py -m gpu -q 2>&1 | tail -n 3
