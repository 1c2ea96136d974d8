#!/bin/bash
timeout 300 python -m pytest tests -m gpu -q -x -s -k "depthwise" 2>&1 | grep -v "^$" | tail -30
