#!/bin/bash
# rebuild libkanconv.so in-tree (nvcc, sm_100a); KANCONV_DEBUG=1 adds the micro-benchmarks / timeline trace
cd "$(dirname "$0")/.." && python -c "import kanconv_b200 as K; K.build(force=True)" 2>&1 | tail -2
