"""Per-layer micro-benchmark of the KAN convolution ops (CUDA events, L2 flushed between iterations).

    python tools/layer_bench.py [--bwd] [--shapes vgg16|small]
"""
import argparse
import json
import sys
import os

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K  # noqa: E402
from kanconv_b200 import functional as KF  # noqa: E402

SHAPES = {
    "vgg16": [(16, 3, 64, 224), (16, 64, 64, 224), (32, 64, 128, 112), (32, 128, 128, 112), (64, 128, 256, 56),
              (64, 256, 256, 56), (64, 256, 512, 28), (64, 512, 512, 28), (64, 512, 512, 14)],
    "small": [(8, 64, 64, 56), (8, 128, 256, 28)],
}


def timeit(fn, iters, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="vgg16")
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", default="auto")
    a = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    K.set_precision(a.precision)
    for (n, cin, cout, hw) in SHAPES[a.shapes]:
        torch.manual_seed(0)
        m = K.KANConv2DLayer(cin, cout, 3, padding=1, base_activation=nn.SiLU).cuda()
        x = torch.randn(n, cin, hw, hw, device="cuda", requires_grad=a.bwd and cin > 3)
        wb = [c.weight for c in m.base_conv]
        ws = [c.weight for c in m.spline_conv]
        flops = 2.0 * n * hw * hw * cout * cin * 9 * 9
        with torch.no_grad():
            t_conv = timeit(lambda: KF.kan_conv(m._spec, x, None, None, wb, ws), a.iters, flush)
            z = KF.kan_conv(m._spec, x, None, None, wb, ws)
            t_norm = timeit(lambda: m._norm_act(z, m.layer_norm, 1, [p.weight for p in m.prelus]), a.iters, flush)
        rec = {"shape": [n, cin, cout, hw], "fwd_conv_ms": round(t_conv, 3), "fwd_conv_tflops": round(flops / t_conv / 1e9, 1),
               "norm_ms": round(t_norm, 3), "norm_gbs": round(2 * z.numel() * 4 / t_norm / 1e6, 1)}
        if a.bwd:
            g = torch.randn_like(z)
            def one():
                for p in m.parameters():
                    p.grad = None
                y = m(x)
                y.backward(g)
            one(); one()
            torch.cuda.synchronize()
            KF.profile_begin()
            one()
            prof = KF.profile_end()
            rec["kernels"] = {k: [round(v["ms"], 3), round(v["flops"] / v["ms"] / 1e9, 1) if v["flops"] else round(v["bytes"] / v["ms"] / 1e6, 1)]
                              for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
            def step():
                for p in m.parameters():
                    p.grad = None
                if x.grad is not None:
                    x.grad = None
                y = m(x)
                y.backward(g)
            t = timeit(step, max(2, a.iters // 2), flush)
            nd = 3 if cin > 3 else 2
            rec["fwd_bwd_ms"] = round(t, 3)
            rec["fwd_bwd_tflops"] = round(nd * flops / t / 1e9, 1)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
