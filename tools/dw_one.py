"""One depthwise FastKAN layer forward + backward (ncu target for the kc_dw_* kernels).  usage: python tools/dw_one.py C HW STRIDE [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kanconv_b200 as K
c, hw, s = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
torch.manual_seed(0)
m = K.FastKANConv2DLayer(c, c, 3, groups=c, padding=1, stride=s, grid_size=5, norm_layer=torch.nn.BatchNorm2d).cuda().train()
m.precision = "fp32"
x = torch.randn(B, c, hw, hw, device="cuda", requires_grad=True)
g = None
for _ in range(3):
    y = m(x)
    g = torch.randn_like(y) if g is None else g
    y.backward(g)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
