"""Out-of-bounds WRITE detection without compute-sanitizer (it is closed on the GPU pool this repo is developed on:
profiles/r2_sanitizer_closed.log).  Every buffer the binding passes to the library is allocated through
functional._ALLOC; here that allocator puts 4 KiB poisoned guard bands before and after each buffer, the layers run forward
and backward through every kernel family on shapes that stress tile edges, and the guard bands must come back untouched.
(Reads cannot be caught this way; the parity tests on ragged shapes cover them indirectly: garbage read from outside a buffer
would change the results.)"""
import numpy as np
import pytest
import torch
import torch.nn as nn

import kanconv_b200 as K
from kanconv_b200 import functional as KF

pytestmark = pytest.mark.gpu
PAD, POISON = 4096, 0xA5


class RedZoneAllocator:
    def __init__(self):
        self.regions = []

    def __call__(self, *shape, dtype, device):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        raw = torch.full((2 * PAD + (nbytes + 255) // 256 * 256,), POISON, dtype=torch.uint8, device=device)
        self.regions.append((raw, nbytes))
        return raw[PAD:PAD + nbytes].view(dtype).view(shape)

    def check(self):
        torch.cuda.synchronize()
        for raw, nbytes in self.regions:
            assert bool((raw[:PAD] == POISON).all()), f"write BEFORE a {nbytes}-byte buffer"
            assert bool((raw[PAD + nbytes:] == POISON).all()), f"write AFTER a {nbytes}-byte buffer"
        return len(self.regions)


CASES = [
    ("kan 3x3", lambda: K.KANConv2DLayer(16, 32, 3, padding=1, base_activation=nn.SiLU), (2, 16, 20, 20)),
    ("kan ragged channels / image", lambda: K.KANConv2DLayer(5, 7, 3, padding=1, base_activation=nn.SiLU), (3, 5, 6, 9)),
    ("kan 3 -> 300 (two N tiles)", lambda: K.KANConv2DLayer(3, 300, 3, padding=1, base_activation=nn.SiLU), (1, 3, 17, 19)),
    ("kan 1x1", lambda: K.KANConv2DLayer(16, 24, 1, base_activation=nn.SiLU), (2, 16, 9, 9)),
    ("kan stride 2", lambda: K.KANConv2DLayer(8, 16, 3, padding=1, stride=2, base_activation=nn.SiLU), (2, 8, 15, 14)),
    ("kan valid conv", lambda: K.KANConv2DLayer(8, 12, 3, padding=0, base_activation=nn.SiLU), (2, 8, 30, 44)),
    ("kan groups", lambda: K.KANConv2DLayer(24, 24, 3, padding=1, groups=3, base_activation=nn.SiLU), (2, 24, 8, 8)),
    ("kan 64x64 maps (cluster norm kernels, cluster size > 1)", lambda: K.KANConv2DLayer(8, 20, 3, padding=1, base_activation=nn.SiLU), (1, 8, 64, 64)),
    ("kan 112x112 maps", lambda: K.KANConv2DLayer(8, 16, 3, padding=1, base_activation=nn.SiLU), (1, 8, 112, 112)),
    ("cheby", lambda: K.ChebyKANConv2DLayer(16, 16, 3, padding=1), (2, 16, 12, 12)),
    ("gram", lambda: K.GRAMKANConv2DLayer(8, 16, 3, padding=1), (2, 8, 12, 12)),
    ("fastkan", lambda: K.FastKANConv2DLayer(8, 16, 3, padding=1), (2, 8, 12, 12)),
    ("kan 1-D", lambda: K.KANConv1DLayer(4, 6, 3, padding=1), (2, 4, 11)),
    ("KANLayer", lambda: K.KANLayer(12, 9), (5, 12)),
]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("name,ctor,shape", CASES, ids=[c[0] for c in CASES])
def test_no_kernel_writes_outside_its_buffers(name, ctor, shape, precision):
    torch.manual_seed(0)
    m = ctor().cuda()
    m.precision = "auto" if precision == "bf16" else "fp32"
    x = torch.randn(*shape, device="cuda", requires_grad=True)
    alloc = RedZoneAllocator()
    old = KF._ALLOC
    KF._ALLOC = alloc
    KF.clear_pack_cache()
    try:
        y = m(x)
        y.backward(torch.randn_like(y))
        n = alloc.check()
    finally:
        KF._ALLOC = old
        KF.clear_pack_cache()
    assert n >= 3, "the layer allocated through functional._ALLOC"
    assert bool(torch.isfinite(y).all()) and bool(torch.isfinite(x.grad).all())


def test_max_pool_writes_stay_inside():
    alloc = RedZoneAllocator()
    old = KF._ALLOC
    KF._ALLOC = alloc
    try:
        x = torch.randn(2, 3, 9, 10, device="cuda", requires_grad=True)
        KF.max_pool2d(x, 2, 2).sum().backward()
        assert alloc.check() >= 3
    finally:
        KF._ALLOC = old


@pytest.mark.parametrize("name,ctor,shape", [CASES[0], CASES[2], CASES[7], CASES[8], CASES[10]], ids=[CASES[i][0] for i in (0, 2, 7, 8, 10)])
def test_repeated_runs_are_bit_identical(name, ctor, shape):
    """Stand-in for racecheck: a data race between pipeline stages (mbarrier protocol of the tcgen05 kernels, DSMEM exchange of
    the cluster kernels, partial-sum reductions) shows up as run-to-run differences.  Twelve forward + backward passes of the
    tensor-core path must agree bit for bit - outputs, dX and every parameter gradient."""
    torch.manual_seed(0)
    m = ctor().cuda()
    m.precision = "auto"
    x = torch.randn(*shape, device="cuda")
    g = None
    first = None
    for _ in range(12):
        for p in m.parameters():
            p.grad = None
        xr = x.clone().requires_grad_(True)
        y = m(xr)
        g = torch.randn_like(y) if g is None else g
        y.backward(g)
        cur = [y.detach().clone(), xr.grad.clone()] + [p.grad.clone() for p in m.parameters()]
        if first is None:
            first = cur
        else:
            assert all(torch.equal(a, b) for a, b in zip(first, cur))


def test_results_do_not_depend_on_the_contents_of_free_memory():
    """No kernel may consume bytes that no kernel wrote (compute-sanitizer's initcheck is not available on the GPU pool): the
    first six modules of KAN-VGG16 (four KAN convolution layers - stem route, CTA pairs, cluster / warp norm kernels - and two
    pooling stages) run forward + backward twice, once after the caching allocator's free blocks were filled with zeros and once
    after they were filled with 3000.0; every parameter gradient and the input gradient must be bit-identical.
    (tools/uninit_check.py is the NaN-poisoning variant, tools/garbage_check*.py the per-layer / per-prefix versions.)"""
    import kanconv_b200 as K
    from kanconv_b200.models.kan_vgg import vggkan
    dev = torch.device("cuda")
    K.set_precision("bf16")
    try:
        torch.manual_seed(0)
        model = vggkan(3, 10, arch="VGG16", classifier_type="Linear", expected_feature_shape=(1, 1), spline_order=3, grid_size=5).to(dev)
        feats = list(model.features.children())[:6]
        x = torch.randn(3, 3, 48, 48, device=dev)

        def run(val):
            torch.cuda.synchronize()
            junk = [torch.full((1 << 28,), val, device=dev)] + [torch.full((s,), val, device=dev) for s in (16, 256, 4096, 65536, 200000) for _ in range(8)]
            del junk
            model.zero_grad(set_to_none=True)
            xx = x.clone().requires_grad_(True)
            h = xx
            for f in feats:
                h = f(h)
            torch.manual_seed(1)
            (h * torch.randn_like(h)).sum().backward()
            torch.cuda.synchronize()
            out = {"dx": xx.grad.clone(), "y": h.detach().clone()}
            out.update({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
            return out

        a, b = run(0.0), run(3000.0)
        assert len(a) > 8
        for k in a:
            assert torch.equal(a[k], b[k]), k
    finally:
        K.set_precision("auto")
