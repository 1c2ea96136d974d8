"""Per-layer comparison of the TC path vs the FP32 path on the inputs each FastKAN layer sees inside MobileNetV2 (debug)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kanconv_b200 as K
from kanconv_b200.models import mobilenet_v2_kan
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
m = mobilenet_v2_kan(num_classes=10, width_mult=0.25, arch="kan_small", kan_conv="FastKAN", classifier_type="Linear", dropout=0.0).cuda().train()
K.set_precision("fp32")
ins = {}
hooks = []
for name, mod in m.named_modules():
    if isinstance(mod, K.FastKANConv2DLayer):
        hooks.append(mod.register_forward_hook(lambda mod, inp, out, name=name: ins.__setitem__(name, (inp[0].detach().clone(), out.detach().clone()))))
torch.manual_seed(1)
x = torch.randn(4, 3, 32, 32, device="cuda")
m(x)
for h in hooks: h.remove()
for name, mod in m.named_modules():
    if isinstance(mod, K.FastKANConv2DLayer):
        xi, yo = ins[name]
        errs = {}
        for prec in ("bf16",):
            mod.precision = prec
            try:
                xg = xi.clone().requires_grad_(True)
                y = mod(xg)
                errs[prec] = float((y - yo).abs().max() / yo.abs().max())
                g = torch.randn_like(y)
                y.backward(g)
                dx_tc = xg.grad.clone()
                gw_tc = mod.spline_conv[0].weight.grad.clone(); mod.spline_conv[0].weight.grad = None
                mod.precision = "fp32"
                xg2 = xi.clone().requires_grad_(True)
                y2 = mod(xg2); y2.backward(g)
                errs["dx"] = float((dx_tc - xg2.grad).abs().max() / xg2.grad.abs().max())
                errs["dw"] = float((gw_tc - mod.spline_conv[0].weight.grad).abs().max() / mod.spline_conv[0].weight.grad.abs().max())
                mod.spline_conv[0].weight.grad = None
            except Exception as e:
                errs[prec] = repr(e)[:80]
            mod.precision = None
        print(name, tuple(xi.shape), "->", tuple(yo.shape), "k", mod.kernel_size, "s", mod.stride, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in errs.items()})
