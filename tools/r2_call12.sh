#!/bin/bash
python -m pytest tests/test_layers_gpu.py -m gpu -q -s -k "model_channel_counts" 2>&1 | grep -v "^$" | tail -n 30 | cut -c1-400
