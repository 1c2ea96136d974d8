"""Small forward + backward passes through every hand-written kernel family - the target of compute-sanitizer
(tools/sanitize.sh memcheck|racecheck|synccheck).  Shapes are tiny: the sanitizer slows kernels down 10-100x."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K  # noqa: E402

torch.manual_seed(0)
cases = [
    ("kan 3x3 tcgen05 fwd + persistent dgrad + wgrad + fused norm bwd", K.KANConv2DLayer(16, 32, 3, padding=1, base_activation=nn.SiLU), (2, 16, 20, 20), "bf16"),
    ("kan 1x1 (phi pre-pass + persistent GEMM forward)", K.KANConv2DLayer(16, 24, 1, base_activation=nn.SiLU), (2, 16, 9, 9), "bf16"),
    ("kan 3x3 stride 2 (strided epilogue, dz scatter)", K.KANConv2DLayer(8, 16, 3, padding=1, stride=2, base_activation=nn.SiLU), (2, 8, 15, 14), "bf16"),
    ("cheby (persistent dgrad, generic basis)", K.ChebyKANConv2DLayer(16, 16, 3, padding=1), (2, 16, 12, 12), "bf16"),
    ("gram (non-persistent dgrad with d/d beta partial rows)", K.GRAMKANConv2DLayer(8, 16, 3, padding=1), (2, 8, 12, 12), "bf16"),
    ("fastkan (RBF, input norm)", K.FastKANConv2DLayer(8, 16, 3, padding=1), (2, 8, 12, 12), "bf16"),
    ("kan 3x3 on 64x64 maps (cluster-resident norm forward / backward, cluster size > 1)", K.KANConv2DLayer(8, 16, 3, padding=1, base_activation=nn.SiLU), (1, 8, 64, 64), "bf16"),
    ("kan fp32 CUDA-core kernels", K.KANConv2DLayer(4, 6, 3, padding=1), (2, 4, 9, 7), "fp32"),
    ("KANLayer (GEMM + LayerNorm/PReLU kernels)", K.KANLayer(12, 9), (5, 12), "fp32"),
]
for name, m, shape, prec in cases:
    m = m.cuda()
    m.precision = prec
    x = torch.randn(*shape, device="cuda", requires_grad=True)
    y = m(x)
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(y).all()) and all(bool(torch.isfinite(p.grad).all()) for p in m.parameters() if p.grad is not None)
    print(("ok   " if ok else "BAD  ") + name, flush=True)
pool = K.functional.max_pool2d(torch.randn(2, 3, 8, 8, device="cuda", requires_grad=True), 2, 2)
pool.sum().backward()
torch.cuda.synchronize()
print("ok   max pool", flush=True)
print("launches:", K._lib.load().kc_launch_count())
