"""Host logic of the 3-D layers (CPU): ``KANConvBase._kan_conv3d`` turns the volume convolution into kd depth-shifted calls of
the 2-D op.  Here the two CUDA entry points it calls (``functional.kan_conv`` / ``functional.norm_act``) are replaced by torch
stand-ins assembled from the oracle's basis functions, so that the slicing / weight-tap / depth-padding arithmetic is held to
the reference's 3-D fixtures without a GPU.  The CUDA kernels themselves are held to the same fixtures in
``test_layers_gpu.py`` (the ``*3d*`` cases of ``test_fp32_path_matches_reference_golden``)."""
import pytest
import torch
import torch.nn.functional as F

import kanconv_b200 as K
from kanconv_b200 import _lib as L
from kanconv_b200 import functional as KF
from oracle import kan_oracle as O
from _util import Golden, rel_err, run_fwd_bwd

CTORS = {"kan3d": K.KANConv3DLayer, "cheby3d": K.ChebyKANConv3DLayer, "gram3d": K.GRAMKANConv3DLayer,
         "fast3d": K.FastKANConv3DLayer}
CASES = ["kan3d_small", "kan3d_groups_s2", "kan3d_k2_dil2_bn", "cheby3d_small", "gram3d_small", "fast3d_small", "fast3d_s2_nopad"]


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


@pytest.mark.parametrize("name", CASES)
def test_depth_decomposition_matches_reference_3d_fixture(name, monkeypatch):
    monkeypatch.setattr(KF, "kan_conv", _kan_conv_standin)
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
