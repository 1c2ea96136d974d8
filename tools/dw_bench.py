"""Depthwise KAN convolution kernels (csrc/kc_dw.cu) against their HBM roofline: the `replace_depthwise=True` stages of
FastKAN-MobileNetV2 at 224x224 input (hidden widths 96@112 ... 960@7, grid_size 5 as in the models), forward / dX / dW timed
with CUDA events (functional.profile_*), algorithmic bytes as in the kernel header.  Also times the per-group route the
depthwise path replaces, at a small channel count (it launches 3+ kernels PER GROUP).
usage: python tools/dw_bench.py [batch, default 64]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kanconv_b200 as K
from kanconv_b200 import functional as KF

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
hbm = float(peaks["hbm_gbs"])
print("HBM peak used:", hbm, "GB/s")


def run(c, hw, stride, iters=5, force_loop=False):
    torch.manual_seed(0)
    m = K.FastKANConv2DLayer(c, c, 3, groups=c, padding=1, stride=stride, grid_size=5, norm_layer=torch.nn.BatchNorm2d).cuda().train()
    m.precision = "fp32"
    x = torch.randn(B, c, hw, hw, device="cuda", requires_grad=True)
    y = m(x)
    g = torch.randn_like(y)
    for _ in range(2):
        m(x).backward(g)
    KF.profile_begin()
    for _ in range(iters):
        m(x).backward(g)
    st = KF.profile_end()
    out = {}
    for k, v in st.items():
        if k.startswith("kc_dw") or force_loop:
            ms = v["ms"] / iters
            out[k] = (round(ms, 4), round(v["bytes"] / iters / ms / 1e6, 0) if v["bytes"] else None)
    return out


for c, hw, s in [(96, 112, 2), (144, 56, 1), (144, 56, 2), (192, 28, 1), (384, 14, 1), (576, 14, 2), (960, 7, 1)]:
    r = run(c, hw, s)
    line = {k: {"ms": v[0], "GB/s": v[1], "frac": None if not v[1] else round(v[1] / hbm, 3)} for k, v in r.items()}
    print(json.dumps({"B": B, "C": c, "hw": hw, "stride": s, **line}))
# the route the depthwise kernels replace: one forward / dgrad / wgrad launch PER GROUP (FP32 CUDA-core kernels)
dw_desc, KF._dw_desc = KF._dw_desc, (lambda *a, **k: None)
r = run(96, 56, 1, iters=2, force_loop=True)
KF._dw_desc = dw_desc
print("per-group route, 96 groups @56, ms per step by kernel:", {k: v[0] for k, v in r.items()}, "total", round(sum(v[0] for v in r.values()), 2))
r = run(96, 56, 1, iters=2)
print("depthwise route, same layer:", {k: v[0] for k, v in r.items()}, "total", round(sum(v[0] for v in r.values()), 3))
