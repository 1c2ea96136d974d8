"""Run a few forward (and optionally backward) passes of ONE KAN conv layer - a short target for ncu."""
import argparse
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="16,64,64,224")
ap.add_argument("--bwd", action="store_true")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
n, cin, cout, hw = [int(v) for v in a.shape.split(",")]
torch.manual_seed(0)
m = K.KANConv2DLayer(cin, cout, 3, padding=1, base_activation=nn.SiLU).cuda()
x = torch.randn(n, cin, hw, hw, device="cuda", requires_grad=a.bwd)
for _ in range(a.iters):
    y = m(x)
    if a.bwd:
        y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
