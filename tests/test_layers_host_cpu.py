"""Host logic of the layer classes (CPU).  The CUDA entry points the layers call (``functional.kan_conv`` / ``kan_layer`` /
``norm_act``) are replaced by torch stand-ins, so that everything ABOVE the kernels is held to the reference's fixtures
without a GPU:
  * ``KANConvBase._kan_conv3d``: the volume convolution as kd depth-shifted calls of the 2-D op (slicing, weight taps, depth
    padding);
  * the three-term-recurrence families: coefficient tables (``recurrence_kan_layers.*_coef``) in the exact ``kc_desc.params``
    layout the CUDA functor reads, channel order, min-max normalisation of Legendre, output activations.
The stand-in for KC_BASIS_RECUR evaluates the recurrence from ``spec.params`` the way ``kc_recur_eval`` (kc_common.cuh) does;
``test_recur_functor_compiled_for_the_host`` compiles that very function for the host and compares it with the oracle.  The
CUDA kernels themselves are held to the same fixtures in ``test_layers_gpu.py``."""
import pytest
import torch
import torch.nn.functional as F

import kanconv_b200 as K
from kanconv_b200 import _lib as L
from kanconv_b200 import functional as KF
from oracle import kan_oracle as O
from _util import LAYER_NAMES, RECURRENCE_FAMILIES, Golden, golden_names, rel_err, run_fwd_bwd

CTORS = {k: getattr(K, v) for k, v in LAYER_NAMES.items()}
CASES = [n for n in golden_names() if "3d" in n or n.split("_")[0].rstrip("123d") in RECURRENCE_FAMILIES]


def _recur_polys(params, nb, t):
    """kc_recur_eval (kc_common.cuh) in torch: params = (pre, c0, a1, b1, A_2, B_2, C_2, ...)."""
    polys = [torch.full_like(t, params[1])]
    if nb > 1:
        polys.append(params[2] * t + params[3])
    for i in range(2, nb):
        A, B, C = params[4 + 3 * (i - 2): 7 + 3 * (i - 2)]
        polys.append((A * t + B) * polys[i - 1] + C * polys[i - 2])
    return polys


def _act(kind):
    return {L.ACT_IDENTITY: lambda t: t, L.ACT_GELU: F.gelu, L.ACT_SILU: F.silu}[kind]


def _kan_conv_standin(spec, x_base, x_basis, beta, w_base, w_basis, precision=None):
    xs = x_base if x_basis is None else x_basis
    groups = len(w_basis)
    cg = x_base.shape[1] // groups
    outs = []
    for g in range(groups):
        xb, xg = x_base[:, g * cg:(g + 1) * cg], xs[:, g * cg:(g + 1) * cg]
        if spec.basis == L.BASIS_BSPLINE:
            phi = O._expand(O.bspline_basis(xg, torch.tensor(spec.params, dtype=xg.dtype), spec.order))
        elif spec.basis == L.BASIS_CHEBY:
            phi = O._expand(O.cheby_basis(xg, spec.order))
        elif spec.basis == L.BASIS_GRAM:
            t = xg if spec.params else torch.tanh(xg)               # params = (1.0,): the caller squashed already
            phi = F.silu(torch.cat(O.gram_basis(t, spec.order, beta), dim=1))
        elif spec.basis in (L.BASIS_RECUR, L.BASIS_RECUR_DM):
            assert len(spec.params) == 4 + 3 * max(spec.nb - 2, 0)
            polys = _recur_polys(spec.params, spec.nb, xg if spec.params[0] != 0.0 else torch.tanh(xg))
            phi = torch.stack(polys, dim=2).flatten(1, 2) if spec.basis == L.BASIS_RECUR else torch.cat(polys, dim=1)
        else:
            phi = O._expand(O.rbf_basis(xg, torch.tensor(spec.params[:-1], dtype=xg.dtype), spec.params[-1]))
        z = F.conv2d(phi, w_basis[g], None, spec.stride, spec.padding, spec.dilation)
        if spec.has_base:
            z = z + F.conv2d(_act(spec.act)(xb), w_base[g], None, spec.stride, spec.padding, spec.dilation)
        outs.append(z)
    return torch.cat(outs, dim=1)


def _norm_act_standin(spec, z, gammas=(), betas=(), alphas=(), given_mean=None, given_rstd=None):
    cg = z.shape[1] // spec.groups
    outs = []
    for g in range(spec.groups):
        zg = z[:, g * cg:(g + 1) * cg]
        w, b = (gammas[g], betas[g]) if spec.affine else (None, None)
        if spec.norm == L.NORM_INSTANCE:
            zg = F.instance_norm(zg, None, None, w, b, True, 0.1, spec.eps)
        elif spec.norm == L.NORM_BATCH:
            zg = F.batch_norm(zg, None, None, w, b, True, 0.1, spec.eps)
        if spec.out_act == L.OUT_PRELU:
            zg = F.prelu(zg, alphas[g])
        elif spec.out_act == L.OUT_SILU:
            zg = F.silu(zg)
        outs.append(zg)
    y = torch.cat(outs, dim=1)
    stat = torch.zeros(spec.groups, cg)
    return y, stat, torch.ones_like(stat)


def _kan_layer_standin(spec, nspec, x, beta, w_base, w_basis, gammas, betas, alphas, given_mean, given_rstd, precision=None):
    return _norm_act_standin(nspec, _kan_conv_standin(spec, x, None, beta, w_base, w_basis), gammas, betas, alphas)


@pytest.mark.parametrize("name", CASES)
def test_layer_host_logic_matches_reference_fixture(name, monkeypatch):
    monkeypatch.setattr(KF, "kan_conv", _kan_conv_standin)
    monkeypatch.setattr(KF, "kan_layer", _kan_layer_standin)
    monkeypatch.setattr(KF, "norm_act", _norm_act_standin)
    gd = Golden(name)
    m = CTORS[gd.kind](**gd.ctor_kwargs(False))
    m.load_state_dict(gd.sd)
    m = m.double().train()
    y, dx, grads = run_fwd_bwd(m, gd.x.double(), gd.g.double())
    assert y.shape == gd.y64.shape
    assert rel_err(y, gd.y64) < 1e-6
    assert rel_err(dx, gd.dx64) < 1e-6
    assert set(grads) == set(gd.grad64)
    for k, v in gd.grad64.items():
        assert rel_err(grads[k], v) < 1e-6, k


def test_3d_layers_reject_wrong_rank():
    m = K.KANConv3DLayer(2, 2, 3, padding=1)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 2, 4, 4))


def _host_functor(tmp_path):
    import ctypes
    import os
    import subprocess
    import importlib
    B = importlib.import_module("kanconv_b200.build")
    if not B.have_nvcc():
        pytest.skip("nvcc not available")
    here = os.path.dirname(os.path.abspath(__file__))
    so = str(tmp_path / "librecur_host.so")
    subprocess.run([B._nvcc(), "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-I" + B.INCLUDE, "-I" + B.CSRC,
                    "-diag-suppress", "177", "-o", so, os.path.join(here, "native", "recur_host.cu")], check=True,
                   capture_output=True)
    return ctypes.CDLL(so)


def test_recur_functor_compiled_for_the_host(tmp_path):
    """kc_recur_eval (csrc/kc_common.cuh), compiled for the host, against the oracle's per-family recurrences: p_j(tanh x) to fp32
    rounding and d p_j / dx against autograd."""
    import ctypes
    from kanconv_b200.layers import recurrence_kan_layers as R
    lib = _host_functor(tmp_path)
    fams = [("hermite", {}, R.hermite_coef, 5), ("gegenbauer", {"alpha_param": 0.75}, lambda nb: R.gegenbauer_coef(nb, 0.75), 6),
            ("laguerre", {"alpha": 0.5}, lambda nb: R.laguerre_coef(nb, 0.5), 5), ("lucas", {}, R.lucas_coef, 7),
            ("fibonacci", {}, R.fibonacci_coef, 7), ("bessel", {}, R.bessel_coef, 4), ("taylor", {}, R.taylor_coef, 6),
            ("legendre", {}, R.legendre_coef, 7), ("jacobi", {"a": 2.0, "b": 0.5}, lambda nb: R.jacobi_coef(nb, 2.0, 0.5), 5)]
    torch.manual_seed(3)
    x = torch.cat([torch.randn(500) * 1.5, torch.tensor([0.0, 3.0, -3.0, 9.0, -20.0])]).float()
    n = x.numel()
    for fam, kw, coef, degree in fams:
        nb = degree if fam == "taylor" else degree + 1
        params = R.recur_params(coef(nb), nb, False)
        xd = x.double().requires_grad_(True)
        polys = torch.stack(O.recurrence_polys(fam, torch.tanh(xd), degree, **kw), dim=1)           # [n, nb]
        grads = torch.stack([torch.autograd.grad(polys[:, j].sum(), xd, retain_graph=True, allow_unused=True)[0]
                             if polys[:, j].requires_grad else torch.zeros_like(xd) for j in range(nb)], dim=1)
        phi = torch.empty(n, nb, dtype=torch.float32)
        dphi = torch.empty(n, nb, dtype=torch.float32)
        cp = (ctypes.c_float * len(params))(*params)
        rc = lib.recur_eval_host(cp, nb, ctypes.c_void_p(x.data_ptr()), n, ctypes.c_void_p(phi.data_ptr()),
                                 ctypes.c_void_p(dphi.data_ptr()))
        assert rc == 0
        assert rel_err(phi, polys) < 2e-6, fam
        assert rel_err(dphi, torch.nan_to_num(grads)) < 2e-6, fam
