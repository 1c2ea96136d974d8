"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol of include/kanconv.h,
descriptor structs match the header, modules mirror the reference's state_dict / RNG behaviour, errors are raised
like upstream, and the product refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn as nn

import kanconv_b200 as K
from kanconv_b200 import _lib as L
from _util import LAYER_NAMES, Golden, golden_names

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTORS = {k: getattr(K, v) for k, v in LAYER_NAMES.items()}


def test_library_exports_every_declared_symbol():
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "kanconv.h")).read()
    declared = set(re.findall(r"\b(kc_[a-z0-9_]+)\s*\(", header))
    declared -= {"kc_desc", "kc_norm_desc", "kc_rownorm_desc"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libkanconv.so does not export {name}"
    assert set(L.EXPORTED_SYMBOLS) == declared
    assert lib.kc_version() == L.KC_ABI_VERSION == int(re.search(r"#define KC_ABI_VERSION (\d+)", header).group(1))
    # the product library carries no debug / experiment exports (they exist only in a KANCONV_DEBUG=1 build)
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", L.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\b(kc_[a-z0-9_]+)\b", syms))
    if os.environ.get("KANCONV_DEBUG") != "1":
        assert not [s_ for s_ in exported if s_.startswith("kc_debug")], exported
    assert declared <= exported


def test_struct_layout_matches_header():
    assert ctypes.sizeof(L.KcDesc) == 20 * 4 + 2 * 8 + 4 * L.KC_MAX_PARAMS
    assert L.KcDesc.x_batch_stride.offset == 80
    assert ctypes.sizeof(L.KcNormDesc) == 40
    assert L.KcNormDesc.batch_stride.offset == 24
    assert ctypes.sizeof(L.KcRowNormDesc) == 20 and L.KcRowNormDesc.eps.offset == 16


def test_desc_validation_through_abi():
    lib = L.load()
    d = L.KcDesc()
    assert lib.kc_tc_supported(ctypes.byref(d)) == 0          # all-zero descriptor is invalid
    assert lib.kc_wgrad_workspace_bytes(ctypes.byref(d)) == 0
    rc = lib.kc_conv_fwd_f32(ctypes.byref(d), None, None, None, None, None, None, None)
    assert rc == L.KC_ERR_INVALID and b"kc_desc" in lib.kc_last_error()
    with pytest.raises(ValueError):
        L.check(rc, "kc_conv_fwd_f32")


@pytest.mark.parametrize("name", golden_names())
def test_state_dict_compatible_with_reference(name):
    gd = Golden(name)
    m = CTORS[gd.kind](**gd.ctor_kwargs(False))
    sd = m.state_dict()
    assert set(sd) == set(gd.sd)
    for k, v in gd.sd.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    m.load_state_dict(gd.sd)                                   # strict


@pytest.mark.parametrize("name", ["kan_small", "kan_silu_groups_s2", "cheby_small", "fast_small", "kan_c8_16", "kan3d_small",
                                  "cheby3d_small", "fast3d_small", "hermite_small", "gegenbauer_d5_1x1", "laguerre_small",
                                  "taylor_d1", "legendre_groups_s2", "jacobi_a2_b05_silu", "bessel3d_small"])
def test_same_seed_same_weights_as_reference(name):
    """Construction consumes the RNG like the reference ctor, so seed 0 reproduces the fixture's weights exactly."""
    gd = Golden(name)
    torch.manual_seed(0)
    m = CTORS[gd.kind](**gd.ctor_kwargs(False))
    for k, v in m.state_dict().items():
        assert torch.equal(v, gd.sd[k]), k


def test_reference_error_conventions():
    with pytest.raises(ValueError, match="groups must be a positive integer"):
        K.KANConv2DLayer(4, 4, 3, groups=0)
    with pytest.raises(ValueError, match="input_dim must be divisible by groups"):
        K.ChebyKANConv2DLayer(3, 4, 3, groups=2)
    with pytest.raises(ValueError, match="output_dim must be divisible by groups"):
        K.FastKANConv2DLayer(4, 3, 3, groups=2)
    with pytest.raises(ValueError):
        K.GRAMKANConv2DLayer(4, 3, 3, groups=2)


def test_no_cpu_fallback():
    m = K.KANConv2DLayer(3, 4, 3, padding=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 3, 5, 5))


def test_factory_mirrors_reference():
    import sys
    _same_padding = sys.modules["kanconv_b200.layers.kan_conv"]._calculate_same_padding
    f = K.CONV_KAN_FACTORY
    for key in ("KAN", "FastKAN", "GRAMKAN", "ChebyKAN", "conv"):
        assert key in f
    layer = f["KAN"](3, 8, 3)                                   # padding=None -> 'same'
    assert isinstance(layer, K.KANConv2DLayer) and layer.padding == 1
    assert _same_padding(3, 1) == 1 and _same_padding((3, 5), 1) == (1, 2)
    assert _same_padding(3, 2) == 2
    wrapped = f["ChebyKAN"](4, 4, 3, l1_decay=1e-4)
    assert type(wrapped).__name__ == "L1" and isinstance(wrapped.module, K.ChebyKANConv2DLayer)
    # extra kwargs are swallowed by **norm_kwargs and filtered against the norm signature, like upstream
    layer = f["KAN"](3, 8, 3, affine=True, degree=7, base_activation=nn.SiLU)
    assert layer.layer_norm[0].affine
    for key in ("WavKAN", "BersnsteinKAN", "FourierKAN", "ReLUKAN"):         # present, but outside the hot path
        with pytest.raises(NotImplementedError):
            f[key](3, 8, 3)
    # the three-term-recurrence families (kan_conv.py:120-157, 354-724): same defaults and quirks as upstream
    assert isinstance(f["LegendreKAN"](3, 8, 3), K.LegendreKANConv2DLayer)
    g = f["GegenbauerKAN"](4, 8, 3, alpha_param=0.5, affine=True, dilation=2)
    assert isinstance(g, K.GegenbauerKANConv2DLayer) and g.alpha_param == 0.5 and g.layer_norm[0].affine
    assert g.padding == 2 and g.dilation == 1                                   # upstream's builders drop `dilation`
    assert f["LaguerreKAN"](4, 8, 3).alpha == 1.0 and f["JacobiKAN"](4, 8, 3, b=2.0).b == 2.0
    assert f["TaylorKAN"](4, 8, 3, degree=4).poly_conv[0].weight.shape[1] == 4 * 4
    assert type(f["HermiteKAN"](4, 4, 3, l1_decay=1e-4)).__name__ == "L1"
    with pytest.raises(ValueError, match="degree must be at least 1"):
        f["FibonacciKAN"](4, 4, 3, degree=0)
    with pytest.raises(ValueError, match="alpha_param must be greater than -0.5"):
        K.GegenbauerKANConv2DLayer(4, 4, 3, 3, -0.5)
    assert isinstance(f["conv"](3, 8, 3), nn.Conv2d)


def test_modules_are_picklable():
    import pickle
    m = K.FastKANConv2DLayer(4, 4, 3, padding=1)
    m2 = pickle.loads(pickle.dumps(m))
    assert set(m2.state_dict()) == set(m.state_dict())
