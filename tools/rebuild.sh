#!/bin/bash
cd /root/repo && python -c "import kanconv_b200 as K; K.build(force=True)" 2>&1 | tail -2
