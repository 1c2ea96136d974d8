#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
python tools/debug_vgg_parity.py VGG11 vgg11_forward > $O/vgg11_parity.txt 2>&1; cut -c1-200 $O/vgg11_parity.txt | head -75
python tools/debug_vgg_parity.py VGG16_kansmall vgg16_kansmall_128_forward > $O/vgg128_parity.txt 2>&1; cut -c1-200 $O/vgg128_parity.txt | grep -v "spline_conv\|base_conv.0.weight *full [0-9.e-]* *norm [0-9.]*e-0[5-9]" | head -60
