"""Generate golden fixtures from the UNMODIFIED reference (dev container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports ``layers`` from /root/reference (read-only tree, never copied), builds each case under a
fixed seed, runs forward + backward in fp64 (the golden) and in fp32 (the reference's own working
precision) and stores inputs / weights / outputs / gradients in ``tests/golden/<case>.npz``.
The fixtures travel to the GPU box; /root/reference does not.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("KAN_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import layers as ref_layers  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

ACTS = {"gelu": nn.GELU, "silu": nn.SiLU, None: None}
NORMS = {"instance": nn.InstanceNorm2d, "batch": nn.BatchNorm2d, "batch3d": nn.BatchNorm3d}


def adversarial_(x, kind):
    """Overwrite a few entries with values at the basis' discontinuities / saturation points."""
    flat = x.view(-1)
    kind = kind[:-2] if kind[-2:] in ("1d", "2d", "3d") and kind != "kan1d" else kind
    if kind in ("kan", "kan1d", "kanlayer"):
        vals = [-2.2000000477, -1.8000000715, -1.0, -0.6000000238, 0.2000000179, 1.0, 2.2000000477,
                2.1999998, -2.3, 2.5, 7.0, -9.0, 0.0]
    elif kind == "cheby":
        vals = [0.0, 8.0, 9.0, 20.0, -20.0, -8.5, 1e-4]
    elif kind == "gram":
        vals = [0.0, 12.0, -12.0, 1e-3]
    else:
        vals = [0.0, 3.0, -3.0, 5.0]
    for i, v in enumerate(vals):
        flat[(i * 37 + 5) % flat.numel()] = v
    return x


CASES = [
    # name, kind, ctor kwargs, x shape, adversarial
    ("kan_small", "kan", dict(input_dim=4, output_dim=6, kernel_size=3, padding=1, base_activation="gelu"), (2, 4, 9, 7)),
    ("kan_silu_groups_s2", "kan", dict(input_dim=4, output_dim=4, kernel_size=3, padding=1, stride=2, groups=2,
                                       base_activation="silu"), (3, 4, 8, 10)),
    ("kan_g3k2_1x1", "kan", dict(input_dim=5, output_dim=3, kernel_size=1, padding=0, grid_size=3, spline_order=2,
                                  grid_range=[-2, 2], base_activation="silu"), (2, 5, 6, 5)),
    ("kan_affine_dil2", "kan", dict(input_dim=3, output_dim=4, kernel_size=3, padding=2, dilation=2,
                                    base_activation="gelu", affine=True), (2, 3, 8, 8)),
    ("kan_batchnorm", "kan", dict(input_dim=3, output_dim=5, kernel_size=3, padding=1, base_activation="silu",
                                  norm_layer="batch"), (4, 3, 6, 6)),
    ("kan_c8_16", "kan", dict(input_dim=8, output_dim=16, kernel_size=3, padding=1, base_activation="silu"), (2, 8, 12, 11)),
    ("cheby_small", "cheby", dict(input_dim=4, output_dim=6, kernel_size=3, padding=1, degree=3), (2, 4, 9, 7)),
    ("cheby_groups_s2", "cheby", dict(input_dim=4, output_dim=6, kernel_size=3, padding=1, stride=2, groups=2, degree=4),
     (2, 4, 8, 8)),
    ("gram_small", "gram", dict(input_dim=4, output_dim=6, kernel_size=3, padding=1, degree=3), (2, 4, 9, 7)),
    ("gram_groups_d4", "gram", dict(input_dim=4, output_dim=4, kernel_size=3, padding=1, groups=2, degree=4), (2, 4, 6, 6)),
    ("fast_small", "fast", dict(input_dim=4, output_dim=6, kernel_size=3, padding=1), (2, 4, 9, 7)),
    ("fast_bn_g5_1x1", "fast", dict(input_dim=6, output_dim=4, kernel_size=1, padding=0, grid_size=5, grid_range=[-1, 1],
                                    norm_layer="batch"), (3, 6, 5, 5)),
    ("fast_groups_s2", "fast", dict(input_dim=4, output_dim=4, kernel_size=3, padding=1, stride=2, groups=2), (2, 4, 8, 8)),
    # round 2: 1-D convolution layers (kan_layers.py:287-297) and the fully-connected KANLayer (kan_layers.py:8-114)
    ("kan1d_small", "kan1d", dict(input_dim=4, output_dim=6, kernel_size=3, padding=1, base_activation="gelu"), (2, 4, 11)),
    ("kan1d_groups_s2", "kan1d", dict(input_dim=4, output_dim=4, kernel_size=3, padding=0, stride=2, groups=2,
                                    base_activation="silu"), (3, 4, 12)),
    ("kanlayer_small", "kanlayer", dict(input_features=7, output_features=5, base_activation="gelu"), (6, 7)),
    ("kanlayer_silu_g3k2", "kanlayer", dict(input_features=12, output_features=9, grid_size=3, spline_order=2,
                                            grid_range=[-2, 2], base_activation="silu"), (5, 12)),
    # round 2: 3-D convolution layers (the *KANConv3DLayer bindings of the four families)
    ("kan3d_small", "kan3d", dict(input_dim=3, output_dim=5, kernel_size=3, padding=1, base_activation="gelu"), (2, 3, 5, 7, 6)),
    ("kan3d_groups_s2", "kan3d", dict(input_dim=4, output_dim=4, kernel_size=3, padding=1, stride=2, groups=2,
                                      base_activation="silu", affine=True), (2, 4, 6, 7, 8)),
    ("kan3d_k2_dil2_bn", "kan3d", dict(input_dim=2, output_dim=3, kernel_size=2, padding=1, dilation=2,
                                       base_activation="silu", norm_layer="batch3d"), (3, 2, 5, 6, 6)),
    ("cheby3d_small", "cheby3d", dict(input_dim=3, output_dim=4, kernel_size=3, padding=1, degree=3), (2, 3, 4, 6, 5)),
    ("gram3d_small", "gram3d", dict(input_dim=3, output_dim=4, kernel_size=3, padding=1, degree=3), (2, 3, 4, 6, 5)),
    ("fast3d_small", "fast3d", dict(input_dim=3, output_dim=4, kernel_size=3, padding=1), (2, 3, 4, 6, 5)),
    ("fast3d_s2_nopad", "fast3d", dict(input_dim=4, output_dim=4, kernel_size=3, padding=0, stride=2, groups=2), (2, 4, 7, 7, 9)),
    # round 2: the three-term-recurrence polynomial families; kind = "<family><rank>d"
    ("hermite_small", "hermite2d", dict(input_dim=4, output_dim=6, kernel_size=3, degree=3, padding=1), (2, 4, 9, 7)),
    ("hermite_groups_s2_silu", "hermite2d", dict(input_dim=4, output_dim=4, kernel_size=3, degree=4, padding=1, stride=2, groups=2,
                                                 base_activation="silu", affine=True), (3, 4, 8, 10)),
    ("gegenbauer_small", "gegenbauer2d", dict(input_dim=4, output_dim=6, kernel_size=3, degree=3, alpha_param=0.5, padding=1),
     (2, 4, 9, 7)),
    ("gegenbauer_d5_1x1", "gegenbauer2d", dict(input_dim=5, output_dim=3, kernel_size=1, degree=5, alpha_param=1.5, padding=0,
                                               base_activation="silu"), (2, 5, 6, 5)),
    ("laguerre_small", "laguerre2d", dict(input_dim=4, output_dim=6, kernel_size=3, degree=3, alpha=1.0, padding=1), (2, 4, 9, 7)),
    ("laguerre_bn_dil2", "laguerre2d", dict(input_dim=3, output_dim=4, kernel_size=3, degree=4, alpha=0.0, padding=2, dilation=2,
                                            norm_layer="batch"), (3, 3, 8, 8)),
    ("lucas_small", "lucas2d", dict(input_dim=4, output_dim=6, kernel_size=3, degree=3, padding=1), (2, 4, 9, 7)),
    ("fibonacci_small", "fibonacci2d", dict(input_dim=4, output_dim=6, kernel_size=3, degree=4, padding=1), (2, 4, 9, 7)),
    ("bessel_small", "bessel2d", dict(input_dim=4, output_dim=6, kernel_size=3, degree=3, padding=1), (2, 4, 9, 7)),
    ("taylor_small", "taylor2d", dict(input_dim=4, output_dim=6, kernel_size=3, degree=4, padding=1), (2, 4, 9, 7)),
    ("taylor_d1", "taylor2d", dict(input_dim=4, output_dim=4, kernel_size=3, degree=1, padding=1, groups=2), (2, 4, 6, 6)),
    ("legendre_small", "legendre2d", dict(input_dim=4, output_dim=6, kernel_size=3, degree=3, padding=1), (2, 4, 9, 7)),
    ("legendre_groups_s2", "legendre2d", dict(input_dim=4, output_dim=4, kernel_size=3, degree=4, padding=1, stride=2, groups=2),
     (3, 4, 8, 10)),
    ("jacobi_small", "jacobi2d", dict(input_dim=4, output_dim=6, kernel_size=3, degree=3, padding=1), (2, 4, 9, 7)),
    ("jacobi_a2_b05_silu", "jacobi2d", dict(input_dim=4, output_dim=4, kernel_size=3, degree=4, a=2.0, b=0.5, padding=1, groups=2,
                                            base_activation="silu"), (2, 4, 8, 8)),
    ("hermite1d_small", "hermite1d", dict(input_dim=4, output_dim=6, kernel_size=3, degree=3, padding=1), (2, 4, 11)),
    ("bessel3d_small", "bessel3d", dict(input_dim=3, output_dim=4, kernel_size=3, degree=2, padding=1), (2, 3, 4, 6, 5)),
    ("legendre3d_small", "legendre3d", dict(input_dim=3, output_dim=4, kernel_size=3, degree=2, padding=1), (2, 3, 4, 6, 5)),
    # round 2: non-finite inputs (NaN => NaN, +-Inf => NaN through 0 * Inf in the Cox-de Boor recursion).
    # One poisoned element per image: the fixture records which outputs / gradients the reference turns into NaN.
    ("kan_naninf", "kan", dict(input_dim=4, output_dim=6, kernel_size=3, padding=1, base_activation="silu"), (4, 4, 9, 7),
     {0: float("nan"), 1: float("inf"), 2: float("-inf")}),
    ("cheby_naninf", "cheby", dict(input_dim=4, output_dim=6, kernel_size=3, padding=1, degree=3), (3, 4, 9, 7),
     {0: float("nan"), 1: float("inf")}),
    ("gram_naninf", "gram", dict(input_dim=4, output_dim=6, kernel_size=3, padding=1, degree=3), (3, 4, 9, 7),
     {0: float("nan"), 1: float("-inf")}),
]

CTORS = {"kan": "KANConv2DLayer", "cheby": "ChebyKANConv2DLayer", "gram": "GRAMKANConv2DLayer",
         "fast": "FastKANConv2DLayer", "kan1d": "KANConv1DLayer", "kanlayer": "KANLayer",
         "kan3d": "KANConv3DLayer", "cheby3d": "ChebyKANConv3DLayer", "gram3d": "GRAMKANConv3DLayer",
         "fast3d": "FastKANConv3DLayer"}
for _f in ("Hermite", "Gegenbauer", "Laguerre", "Lucas", "Fibonacci", "Bessel", "Taylor", "Legendre", "Jacobi"):
    for _n in (1, 2, 3):
        CTORS[f"{_f.lower()}{_n}d"] = f"{_f}KANConv{_n}DLayer"


def build(kind, kw):
    kw = dict(kw)
    if "base_activation" in kw:
        kw["base_activation"] = ACTS[kw["base_activation"]]
    if "norm_layer" in kw:
        kw["norm_layer"] = NORMS[kw["norm_layer"]]
    return getattr(ref_layers, CTORS[kind])(**kw)


def run(module, x, g, dtype):
    m = module.to(dtype)
    if hasattr(m, "gram_poly"):
        try:
            type(m).gram_poly.cache_clear()
        except Exception:
            pass
    for p in m.parameters():
        p.grad = None
    xx = x.to(dtype).clone().requires_grad_(True)
    y = m(xx)
    y.backward(g.to(dtype))
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    return y.detach(), xx.grad.detach(), grads


def main():
    only = set(sys.argv[1:])            # optional: names of the cases to (re)generate; default = all + the checksums
    for case in CASES:
        name, kind, kw, xshape = case[:4]
        poison = case[4] if len(case) > 4 else {}
        if only and name not in only:
            continue
        torch.manual_seed(0)
        m = build(kind, kw)
        m.train()
        # make affine norm params / prelu non-trivial so their gradients are exercised
        with torch.no_grad():
            for k, p in m.named_parameters():
                if "layer_norm" in k:
                    p.add_(0.1 * torch.randn_like(p))
                if k == "beta_weights":
                    p.copy_(0.05 * torch.randn_like(p))
        torch.manual_seed(1)
        x = adversarial_(torch.randn(*xshape) * 1.2, kind)
        for img, val in poison.items():
            x[img].view(-1)[(img * 53 + 17) % x[img].numel()] = val
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        with torch.no_grad():
            yshape = m(x).shape
        m.load_state_dict(sd)                       # undo batch-norm running-stat update of the probe
        torch.manual_seed(2)
        g = torch.randn(*yshape)
        y32, dx32, gr32 = run(m, x, g, torch.float32)
        m.load_state_dict(sd)
        y64, dx64, gr64 = run(m, x, g, torch.float64)
        out = {"x": x.numpy(), "g": g.numpy(), "y64": y64.numpy(), "dx64": dx64.numpy(),
               "y32": y32.numpy(), "dx32": dx32.numpy(),
               "meta": np.frombuffer(json.dumps({"kind": kind, "kwargs": kw, "torch": torch.__version__}).encode(),
                                     dtype=np.uint8)}
        for k, v in sd.items():
            out["sd/" + k] = v.numpy()
        for k, v in gr64.items():
            if v is not None:
                out["grad64/" + k] = v.numpy()
        for k, v in gr32.items():
            if v is not None:
                out["grad32/" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: y{tuple(y64.shape)} |y|={float(torch.nan_to_num(y64).norm()):.6f} nan(y)={int(torch.isnan(y64).sum())}")
    if only:
        return

    # Known-answer checksums of BASELINE config 1 (SURVEY Appendix E recipe) - regenerated from the live reference.
    ck = {}
    for tag, ctor in [
        ("kan_gelu", lambda: ref_layers.KANConv2DLayer(3, 16, 3, spline_order=3, grid_size=5, padding=1)),
        ("kan_silu", lambda: ref_layers.KANConv2DLayer(3, 16, 3, spline_order=3, grid_size=5, padding=1,
                                                       base_activation=nn.SiLU)),
        ("cheby", lambda: ref_layers.ChebyKANConv2DLayer(8, 16, 3, degree=3, padding=1)),
        ("gram", lambda: ref_layers.GRAMKANConv2DLayer(8, 16, 3, degree=3, padding=1)),
        ("fast", lambda: ref_layers.FastKANConv2DLayer(8, 16, 3, padding=1)),
    ]:
        torch.manual_seed(0)
        m = ctor()
        cin = 3 if tag.startswith("kan") else 8
        torch.manual_seed(1)
        x = torch.randn(16, cin, 32, 32)
        torch.manual_seed(2)
        g = torch.randn(16, 16, 32, 32)
        y, dx, gr = run(m, x, g, torch.float64)
        ck[tag] = {"y_sum": float(y.sum()), "y_norm": float(y.norm()), "dx_sum": float(dx.sum()),
                   "dx_norm": float(dx.norm()),
                   "grads": {k: [float(v.sum()), float(v.norm())] for k, v in gr.items() if v is not None}}
        print(tag, ck[tag]["y_sum"], ck[tag]["y_norm"])
    with open(os.path.join(HERE, "config1_checksums.json"), "w") as f:
        json.dump(ck, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
