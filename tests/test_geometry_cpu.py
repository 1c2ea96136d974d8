"""Host-side logic of the tensor-core path (no GPU needed): which shapes it accepts, buffer sizes, and the tile geometry the
kernels are launched with - TMEM / shared-memory budgets must hold for every layer of the BASELINE models."""
import ctypes

import pytest
import torch.nn as nn

import kanconv_b200 as K
from kanconv_b200 import functional as KF

lib = K._lib.load()
lib.kc_tc_geometry.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
FIELDS = ["nsub", "ntile", "n_nt", "na", "tps", "bst", "mtiles", "smem"]
SMEM_LIMIT = 227 * 1024


def _desc(n, cin, cout, h, w, k=3, pad=1, stride=1, dilation=1):
    m = K.KANConv2DLayer(cin, cout, k, padding=pad, stride=stride, dilation=dilation, base_activation=nn.SiLU)
    ho = (h + 2 * pad - dilation * (k - 1) - 1) // stride + 1
    wo = (w + 2 * pad - dilation * (k - 1) - 1) // stride + 1
    return KF._make_desc(m._spec, n, cin, h, w, cout, cin * h * w, cout * ho * wo), ho, wo


def _geom(d, which):
    out = (ctypes.c_longlong * 8)()
    assert lib.kc_tc_geometry(ctypes.byref(d), which, out) == 0, lib.kc_last_error()
    return dict(zip(FIELDS, list(out)))


VGG16 = [(64, 3, 64, 224), (64, 64, 64, 224), (64, 64, 128, 112), (64, 128, 128, 112), (64, 128, 256, 56), (64, 256, 256, 56),
         (64, 256, 512, 28), (64, 512, 512, 28), (64, 512, 512, 14)]


@pytest.mark.parametrize("n,cin,cout,hw", VGG16)
def test_vgg16_layer_geometry_respects_tmem_and_smem(n, cin, cout, hw):
    d, _, _ = _desc(n, cin, cout, hw, hw)
    assert lib.kc_tc_supported(ctypes.byref(d)) == 1
    f, g = _geom(d, 0), _geom(d, 1)
    assert f["nsub"] * f["ntile"] <= 512 and f["ntile"] % 16 == 0 and f["ntile"] <= 256
    assert f["ntile"] * f["n_nt"] >= cout
    assert f["smem"] <= SMEM_LIMIT and g["smem"] <= SMEM_LIMIT
    assert f["mtiles"] * f["nsub"] * 128 >= n * (hw + 1) * (hw + 1)             # the tiles cover the flat position range
    # the closed-form cubic dgrad runs the persistent kernel: 14 channels x 9 columns per 128-column tile, two accumulator sets
    assert g["ntile"] == 128 and g["nsub"] == 2 and g["n_nt"] == -(-cin // 14)
    assert g["tps"] in (1, 3, 9) and g["bst"] >= 2 and g["na"] in (2, 3)


def test_which_shapes_take_the_tensor_core_path():
    ok = lambda *a, **kw: lib.kc_tc_supported(ctypes.byref(_desc(*a, **kw)[0]))
    assert ok(2, 8, 16, 12, 12) == 1
    assert ok(2, 8, 16, 12, 12, k=1, pad=0) == 1
    assert ok(2, 8, 16, 12, 12, k=5, pad=2) == 1
    assert ok(2, 3, 32, 33, 31, stride=2) == 1                 # strided layers run on the stride-1 grid
    assert ok(2, 8, 16, 12, 12, pad=2, dilation=2) == 0        # dilation: FP32 CUDA-core kernels
    assert ok(2, 8, 16, 20, 20, k=9, pad=4) == 0               # filters larger than 8x8


def test_buffer_sizes():
    d, ho, wo = _desc(4, 32, 48, 10, 12)
    nbytes = lambda which: lib.kc_tc_bytes(ctypes.byref(d), which)
    L = 4 * (10 + 1) * (12 + 1)
    assert nbytes(2) == L * 48 * 2                              # dz_flat: flat positions x cout (multiple of 16) x bf16
    assert nbytes(4) == (32 + 16) * L * 16                      # phi: 32 spline planes (2 chunks of 16) + 16 base planes, 16 B rows
    assert nbytes(3) == nbytes(5) + nbytes(4)                   # workspace with / without the transient basis buffer
    assert nbytes(0) > 0 and nbytes(1) > 0 and nbytes(0) % 16 == 0 and nbytes(1) % 16 == 0
    # packed forward image: every (tap, k-core, cout) vector once; K = 32 channels x 8 basis + 32 base channels
    assert nbytes(0) == 9 * 48 * 16 * (32 + 4)


def test_pointwise_layers_ask_for_the_phi_buffer():
    d1, _, _ = _desc(2, 16, 24, 9, 9, k=1, pad=0)
    d3, _, _ = _desc(2, 16, 24, 9, 9)
    assert lib.kc_tc_fwd_needs_phi(ctypes.byref(d1)) == 1
    assert lib.kc_tc_fwd_needs_phi(ctypes.byref(d3)) == 0
    g = _geom(d1, 0)
    assert g["nsub"] == 2 and g["ntile"] == 32 and g["tps"] == 1        # persistent GEMM over phi, N tile = cout rounded to 16


def test_invalid_descriptor_reports_zero_bytes():
    d, _, _ = _desc(2, 8, 16, 12, 12)
    d.ho += 1                                                  # inconsistent geometry
    assert lib.kc_tc_supported(ctypes.byref(d)) == 0
    assert lib.kc_tc_bytes(ctypes.byref(d), 0) == 0
