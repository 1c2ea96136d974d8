"""Shared helpers for the test-suite: golden-fixture loading and oracle construction."""
import functools
import glob
import json
import os

import numpy as np
import torch
import torch.nn as nn

from oracle import kan_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NORMS = {"instance": nn.InstanceNorm2d, "batch": nn.BatchNorm2d, "batch3d": nn.BatchNorm3d}
ORACLE_CTORS = {"kan": O.OracleKANConv2D, "cheby": O.OracleChebyKANConv2D, "gram": O.OracleGRAMKANConv2D,
                "fast": O.OracleFastKANConv2D, "kan1d": O.OracleKANConv1D, "kanlayer": O.OracleKANLayer,
                "kan3d": O.OracleKANConv3D, "cheby3d": O.OracleChebyKANConv3D, "gram3d": O.OracleGRAMKANConv3D,
                "fast3d": O.OracleFastKANConv3D}
# kind -> class name of the layer under test (same names in the reference's ``layers`` package and in kanconv_b200)
LAYER_NAMES = {"kan": "KANConv2DLayer", "cheby": "ChebyKANConv2DLayer", "gram": "GRAMKANConv2DLayer", "fast": "FastKANConv2DLayer",
               "kan1d": "KANConv1DLayer", "kanlayer": "KANLayer", "kan3d": "KANConv3DLayer", "cheby3d": "ChebyKANConv3DLayer",
               "gram3d": "GRAMKANConv3DLayer", "fast3d": "FastKANConv3DLayer"}
# three-term-recurrence polynomial families: kind = "<family><rank>d"
RECURRENCE_FAMILIES = ("hermite", "gegenbauer", "laguerre", "lucas", "fibonacci", "bessel", "taylor", "legendre", "jacobi")
for _f in RECURRENCE_FAMILIES:
    for _n in (1, 2, 3):
        _cls = O.OracleDegreeMajorPolyKANConv if _f in ("legendre", "jacobi") else O.OraclePolyKANConv
        ORACLE_CTORS[f"{_f}{_n}d"] = functools.partial(_cls, _f, _n)
        LAYER_NAMES[f"{_f}{_n}d"] = f"{_f.capitalize()}KANConv{_n}DLayer"


def layer_class(package, kind):
    return getattr(package, LAYER_NAMES[kind])


# whole-model fixtures (make_model_golden.py): a different record layout
MODEL_FIXTURES = {"mbv2_fastkan_forward", "mbv2_fastkan_rdw_forward", "vgg16_kansmall_forward", "vgg16_kansmall_128_forward", "vgg11_forward", "kan_mlp_forward"}


def golden_names():
    """Layer fixtures written by tests/golden/make_golden.py."""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if n not in MODEL_FIXTURES]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        meta = json.loads(bytes(z["meta"]).decode())
        self.name, self.kind, self.kwargs = name, meta["kind"], meta["kwargs"]
        self.x = torch.from_numpy(z["x"])
        self.g = torch.from_numpy(z["g"])
        self.sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
        self.y64, self.dx64 = torch.from_numpy(z["y64"]), torch.from_numpy(z["dx64"])
        self.y32, self.dx32 = torch.from_numpy(z["y32"]), torch.from_numpy(z["dx32"])
        self.grad64 = {k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad64/")}
        self.grad32 = {k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad32/")}

    def ctor_kwargs(self, for_oracle):
        kw = dict(self.kwargs)
        if "norm_layer" in kw:
            kw["norm_layer"] = NORMS[kw["norm_layer"]]
        if not for_oracle and "base_activation" in kw:
            kw["base_activation"] = {"gelu": nn.GELU, "silu": nn.SiLU, None: None}[kw["base_activation"]]
        return kw

    def oracle(self, dtype=torch.float64):
        m = ORACLE_CTORS[self.kind](**self.ctor_kwargs(True))
        m.load_state_dict(self.sd)
        return m.to(dtype).train()


def run_fwd_bwd(m, x, g):
    for p in m.parameters():
        p.grad = None
    xx = x.detach().clone().requires_grad_(True)
    y = m(xx)
    y.backward(g)
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    return y.detach(), xx.grad.detach(), grads


def rel_err(a, b):
    """max |a-b| / max |b| (scale-relative max error) over the finite elements of the expected tensor ``b``.
    Non-finite values must sit at exactly the same positions in both (NaN <-> NaN, +-Inf <-> the same Inf); otherwise inf."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    fin = torch.isfinite(b)
    if not bool(fin.all()):
        if not torch.equal(torch.isnan(a), torch.isnan(b)):
            return float("inf")
        inf = torch.isinf(b)
        if not torch.equal(torch.isinf(a), inf) or not torch.equal(a[inf], b[inf]):
            return float("inf")
        a, b = a[fin], b[fin]
        if b.numel() == 0:
            return 0.0
    elif not bool(torch.isfinite(a).all()):
        return float("inf")
    denom = float(b.abs().max())
    return float((a - b).abs().max()) / (denom if denom > 0 else 1.0)


def tol_violations(a, b, rtol=2e-2, atol=1e-3):
    """Fraction of elements outside BASELINE north_star's elementwise BF16 tolerance |a-b| <= atol + rtol*|b| (finite part)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    fin = torch.isfinite(b) & torch.isfinite(a)
    if int(fin.sum()) == 0:
        return 0.0
    bad = ((a - b).abs() > atol + rtol * b.abs()) & fin
    return float(bad.sum()) / float(fin.sum())
