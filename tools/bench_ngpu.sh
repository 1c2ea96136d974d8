#!/bin/bash
# usage: tools/bench_ngpu.sh N [extra bench args]   - runs bench.py on N GPUs of this node (torchrun for N > 1)
N=${1:-1}; shift
OUT=gpurun_out/bench_${N}gpu.json
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 "$@" > $OUT.log 2>&1
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $OUT.log 2>&1
fi
rc=$?
tail -1 $OUT.log > $OUT
python - <<PY
import json
try:
    r = json.load(open("$OUT"))
    print({k: r.get(k) for k in ("value", "ms_per_step", "n_gpus", "gpu_launches", "clocks")})
    print(r.get("e2e"))
except Exception as e:
    print("no JSON line:", e)
    print(open("$OUT.log").read()[-3000:])
PY
exit $rc
