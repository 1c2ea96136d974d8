"""torch.autograd.Function wrappers around the C ABI of libkanconv.so.

Two differentiable ops make up every KAN convolution layer of the reference:

  * ``kan_conv``  - z = conv(act(x_base), W_base) + conv(basis(x_basis), W_basis)   (kan_layers.py:199-241 and siblings)
  * ``norm_act``  - y = out_act(norm(z))                                           (kan_layers.py:242-243)

Both call hand-written CUDA through ctypes (raw device pointers + the current CUDA stream); forward saves only the
layer INPUT (the basis expansion is recomputed in the backward kernels), so the 8x-expanded tensor and the ~30 autograd
intermediates of the reference never exist.  CPU tensors are rejected: there is no fallback path.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L

_TC_BACKWARD = os.environ.get("KANCONV_TC_BACKWARD", "1") != "0"   # debug switch: FP32 CUDA-core backward after a TC forward
# KANCONV_SAVE_PHI=0: do not keep the bf16 basis rows between forward and backward (saves 18 B per input element per
# layer of activation memory; the weight gradient then re-evaluates them in a pre-pass)
_SAVE_PHI = os.environ.get("KANCONV_SAVE_PHI", "1") != "0"
_PRECISION = "auto"      # "auto": tensor cores when the shape is supported, else CUDA-core FP32 | "bf16" | "fp32"


def set_precision(p: str) -> None:
    """Select the arithmetic path of kan_conv: 'auto' (default), 'bf16' (tcgen05 only, error if unsupported), 'fp32'."""
    global _PRECISION
    if p not in ("auto", "bf16", "fp32"):
        raise ValueError("precision must be 'auto', 'bf16' or 'fp32'")
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


@dataclass(frozen=True)
class ConvSpec:
    """Static description of one KAN convolution layer (all groups)."""
    basis: int
    act: int
    nb: int
    order: int
    params: Tuple[float, ...]
    kernel: Tuple[int, int]
    stride: Tuple[int, int]
    padding: Tuple[int, int]
    dilation: Tuple[int, int]
    groups: int

    @property
    def has_base(self) -> bool:
        return self.act != L.ACT_NONE

    def out_hw(self, h: int, w: int) -> Tuple[int, int]:
        ho = (h + 2 * self.padding[0] - self.dilation[0] * (self.kernel[0] - 1) - 1) // self.stride[0] + 1
        wo = (w + 2 * self.padding[1] - self.dilation[1] * (self.kernel[1] - 1) - 1) // self.stride[1] + 1
        return ho, wo


@dataclass(frozen=True)
class NormSpec:
    norm: int
    out_act: int
    groups: int
    affine: bool
    eps: float
    use_batch_stats: bool = True    # BatchNorm only: False -> normalise with the provided (running) statistics


# ---- optional per-call device timing (bench.py's roofline leg): CUDA events on the launching stream ----------------------
_PROFILE = None


def profile_begin() -> None:
    global _PROFILE
    _PROFILE = []


def profile_end():
    """-> {kernel name: {"ms", "calls", "flops", "bytes"}} summed over the calls since profile_begin()."""
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    out = {}
    for name, flops, nbytes, e0, e1 in rec or []:
        st = out.setdefault(name, {"ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0})
        st["ms"] += e0.elapsed_time(e1)
        st["calls"] += 1
        st["flops"] += flops
        st["bytes"] += nbytes
    return out


def _timed(name: str, flops: float, nbytes: float, call):
    if _PROFILE is None:
        return call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = call()
    e1.record()
    _PROFILE.append((name, flops, nbytes, e0, e1))
    return rc


def _conv_flops(d) -> float:
    wb = d.nb + (0 if d.act == L.ACT_NONE else 1)
    return 2.0 * d.n * d.ho * d.wo * d.cout * d.cin * wb * d.kh * d.kw


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: kanconv_b200 runs on CUDA tensors only (no CPU fallback); got device {t.device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype}")


def _make_desc(spec: ConvSpec, n, cin_g, h, w, cout_g, x_bs, z_bs) -> L.KcDesc:
    ho, wo = spec.out_hw(h, w)
    if ho <= 0 or wo <= 0:
        raise ValueError("kan_conv: kernel does not fit the input")
    d = L.KcDesc()
    d.basis, d.act = spec.basis, spec.act
    d.n, d.cin, d.h, d.w = n, cin_g, h, w
    d.cout, d.ho, d.wo = cout_g, ho, wo
    d.kh, d.kw = spec.kernel
    d.stride_h, d.stride_w = spec.stride
    d.pad_h, d.pad_w = spec.padding
    d.dil_h, d.dil_w = spec.dilation
    d.nb, d.order, d.nparams = spec.nb, spec.order, len(spec.params)
    d.x_batch_stride, d.z_batch_stride = x_bs, z_bs
    for i, v in enumerate(spec.params):
        d.params[i] = v
    return d


def _use_tc(lib, desc, precision: str) -> bool:
    if precision == "fp32":
        return False
    ok = bool(lib.kc_tc_supported(ctypes.byref(desc)))
    if precision == "bf16" and not ok:
        raise NotImplementedError("kan_conv: shape not supported by the tensor-core path: " + lib.kc_last_error().decode())
    return ok


class _KanConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec: ConvSpec, precision: str, x_base, x_basis, beta, *weights):
        lib = L.load()
        alias = x_basis is None
        xb = x_base.contiguous()
        xs = xb if alias else x_basis.contiguous()
        _require_cuda(xb, "kan_conv")
        _require_cuda(xs, "kan_conv")
        G = spec.groups
        n, c_total, h, w = xb.shape
        cg = c_total // G
        w_base = list(weights[:G]) if spec.has_base else [None] * G
        w_basis = list(weights[G:2 * G]) if spec.has_base else list(weights[:G])
        og = w_basis[0].shape[0]
        ho, wo = spec.out_hw(h, w)
        z = torch.empty((n, og * G, ho, wo), device=xb.device, dtype=torch.float32)
        stream = _stream()
        used_tc = []
        # basis rows saved for the weight gradient (the reference keeps the expanded basis alive for autograd, too)
        want_phi = _SAVE_PHI and _TC_BACKWARD and any(ctx.needs_input_grad[5:])
        phis = []
        for g in range(G):
            d = _make_desc(spec, n, cg, h, w, og, c_total * h * w, og * G * ho * wo)
            xbg, xsg, zg = xb[:, g * cg:(g + 1) * cg], xs[:, g * cg:(g + 1) * cg], z[:, g * og:(g + 1) * og]
            wbg = None if w_base[g] is None else w_base[g].contiguous()
            wsg = w_basis[g].contiguous()
            tc = _use_tc(lib, d, precision)
            used_tc.append(tc)
            if tc:
                nbytes = lib.kc_tc_bytes(ctypes.byref(d), 0)
                packed = torch.empty(nbytes, device=xb.device, dtype=torch.uint8)
                L.check(_timed("kc_pack_fwd_kernel", 0.0, 1.5 * nbytes, lambda: lib.kc_tc_pack_weights(
                    ctypes.byref(d), _ptr(wbg), _ptr(wsg), _ptr(packed), None, stream)), "kc_tc_pack_weights")
                phi = None
                if (want_phi or lib.kc_tc_fwd_needs_phi(ctypes.byref(d))) and lib.kc_tc_bytes(ctypes.byref(d), 4) > 0:
                    phi = torch.empty(lib.kc_tc_bytes(ctypes.byref(d), 4), device=xb.device, dtype=torch.uint8)
                phis.append(phi)
                L.check(_timed("kc_tc_kernel<fwd>", _conv_flops(d), 0.0, lambda: lib.kc_conv_fwd_tc(
                    ctypes.byref(d), _ptr(xbg), _ptr(xsg), _ptr(packed), _ptr(beta), _ptr(zg), _ptr(phi), stream)), "kc_conv_fwd_tc")
            else:
                phis.append(None)
                L.check(_timed("kc_fwd_simt_kernel", _conv_flops(d), 0.0, lambda: lib.kc_conv_fwd_f32(
                    ctypes.byref(d), _ptr(xbg), _ptr(xsg), _ptr(wbg), _ptr(wsg), _ptr(beta), _ptr(zg), stream)), "kc_conv_fwd_f32")
        ctx.spec, ctx.alias, ctx.precision, ctx.used_tc, ctx.phis = spec, alias, precision, used_tc, phis
        ctx.save_for_backward(xb, xs if not alias else None, beta, *weights)
        return z

    @staticmethod
    def backward(ctx, dz):
        lib = L.load()
        spec: ConvSpec = ctx.spec
        saved = ctx.saved_tensors
        xb, xs, beta = saved[0], saved[1], saved[2]
        weights = saved[3:]
        alias = ctx.alias
        if alias:
            xs = xb
        G = spec.groups
        n, c_total, h, w = xb.shape
        cg = c_total // G
        w_base = list(weights[:G]) if spec.has_base else [None] * G
        w_basis = list(weights[G:2 * G]) if spec.has_base else list(weights[:G])
        og = w_basis[0].shape[0]
        ho, wo = spec.out_hw(h, w)
        dz = dz.contiguous()
        stream = _stream()
        # inputs of forward: (spec, precision, x_base, x_basis, beta, *weights)
        need_dxb, need_dxs, need_dbeta = ctx.needs_input_grad[2], ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        need_w = list(ctx.needs_input_grad[5:])
        dbeta = torch.zeros_like(beta) if (spec.basis == L.BASIS_GRAM and beta is not None) else None
        run_dgrad = need_dxb or need_dxs or (dbeta is not None and need_dbeta)
        dx_base = dx_basis = None
        if run_dgrad:
            dx_base = torch.empty_like(xb)
            dx_basis = dx_base if alias else torch.empty_like(xs)
        dws: List[Optional[torch.Tensor]] = [None] * len(weights)
        for g in range(G):
            d = _make_desc(spec, n, cg, h, w, og, c_total * h * w, og * G * ho * wo)
            sl = slice(g * cg, (g + 1) * cg)
            xbg, xsg, dzg = xb[:, sl], xs[:, sl], dz[:, g * og:(g + 1) * og]
            wbg = None if w_base[g] is None else w_base[g].contiguous()
            wsg = w_basis[g].contiguous()
            tc_bwd = ctx.used_tc[g] and _TC_BACKWARD and lib.kc_tc_bytes(ctypes.byref(d), 1) > 0
            dzf = None
            if tc_bwd:
                dzf = torch.empty(lib.kc_tc_bytes(ctypes.byref(d), 2), device=xb.device, dtype=torch.uint8)
                L.check(_timed("kc_dz_flat_kernel", 0.0, 6.0 * dzg.numel(), lambda: lib.kc_tc_dz_flat(
                    ctypes.byref(d), _ptr(dzg), _ptr(dzf), stream)), "kc_tc_dz_flat")
            if run_dgrad:
                if tc_bwd:
                    nbytes = lib.kc_tc_bytes(ctypes.byref(d), 1)
                    packed_d = torch.empty(nbytes, device=xb.device, dtype=torch.uint8)
                    L.check(_timed("kc_pack_dgrad_kernel", 0.0, 1.5 * nbytes, lambda: lib.kc_tc_pack_weights(
                        ctypes.byref(d), _ptr(wbg), _ptr(wsg), None, _ptr(packed_d), stream)), "kc_tc_pack_weights")
                    L.check(_timed("kc_tc_kernel<dgrad>", _conv_flops(d), 0.0, lambda: lib.kc_conv_dgrad_tc(
                        ctypes.byref(d), None, _ptr(xbg), _ptr(xsg), _ptr(packed_d), _ptr(beta), _ptr(dx_base[:, sl]),
                        _ptr(dx_basis[:, sl]), _ptr(dbeta), _ptr(dzf), stream)), "kc_conv_dgrad_tc")
                else:
                    L.check(_timed("kc_dgrad_simt_kernel", _conv_flops(d), 0.0, lambda: lib.kc_conv_dgrad_f32(
                        ctypes.byref(d), _ptr(dzg), _ptr(xbg), _ptr(xsg), _ptr(wbg), _ptr(wsg), _ptr(beta),
                        _ptr(dx_base[:, sl]), _ptr(dx_basis[:, sl]), _ptr(dbeta), stream)), "kc_conv_dgrad_f32")
            wi_base, wi_basis = (g, G + g) if spec.has_base else (None, g)
            if (wi_base is not None and need_w[wi_base]) or need_w[wi_basis]:
                dwb = torch.empty_like(wbg) if wbg is not None else None
                dwsg = torch.empty_like(wsg)
                if tc_bwd and lib.kc_tc_bytes(ctypes.byref(d), 3) > 0:
                    phi = ctx.phis[g]
                    ws = torch.empty(lib.kc_tc_bytes(ctypes.byref(d), 5 if phi is not None else 3), device=xb.device, dtype=torch.uint8)
                    L.check(_timed("kc_wgrad_tc_kernel", _conv_flops(d), 0.0, lambda: lib.kc_conv_wgrad_tc(
                        ctypes.byref(d), _ptr(dzf), _ptr(xbg), _ptr(xsg), _ptr(beta), _ptr(phi), _ptr(dwb), _ptr(dwsg), _ptr(ws),
                        stream)), "kc_conv_wgrad_tc")
                    ctx.phis[g] = None
                else:
                    nbytes = lib.kc_wgrad_workspace_bytes(ctypes.byref(d))
                    ws = torch.empty(max(nbytes, 16), device=xb.device, dtype=torch.uint8)
                    L.check(_timed("kc_wgrad_simt_kernel", _conv_flops(d), 0.0, lambda: lib.kc_conv_wgrad_f32(
                        ctypes.byref(d), _ptr(dzg), _ptr(xbg), _ptr(xsg), _ptr(beta), _ptr(dwb), _ptr(dwsg), _ptr(ws), stream)),
                        "kc_conv_wgrad_f32")
                if wi_base is not None:
                    dws[wi_base] = dwb
                dws[wi_basis] = dwsg
        return (None, None, dx_base if need_dxb else None, (dx_basis if (need_dxs and not alias) else None),
                dbeta if need_dbeta else None, *dws)


def kan_conv(spec: ConvSpec, x_base: torch.Tensor, x_basis: Optional[torch.Tensor], beta: Optional[torch.Tensor],
             w_base: Sequence[torch.Tensor], w_basis: Sequence[torch.Tensor], precision: Optional[str] = None) -> torch.Tensor:
    """Pre-normalisation output of a KAN convolution layer.  ``x_basis=None`` means "same tensor as x_base"."""
    weights = (list(w_base) if spec.has_base else []) + list(w_basis)
    return _KanConvFn.apply(spec, precision or _PRECISION, x_base, x_basis, beta, *weights)


class _NormActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec: NormSpec, z, given_mean, given_rstd, *params):
        lib = L.load()
        z = z.contiguous()
        _require_cuda(z, "norm_act")
        G = spec.groups
        n, c_total, h, w = z.shape
        cg, hw = c_total // G, h * w
        # params layout: [gamma_0, beta_0, ..., gamma_{G-1}, beta_{G-1}] if affine, then [alpha_0..alpha_{G-1}] if PReLU
        gam = [params[2 * g] for g in range(G)] if spec.affine else [None] * G
        bet = [params[2 * g + 1] for g in range(G)] if spec.affine else [None] * G
        off = 2 * G if spec.affine else 0
        alp = [params[off + g] for g in range(G)] if spec.out_act == L.OUT_PRELU else [None] * G
        y = torch.empty_like(z)
        nstat = {L.NORM_NONE: 1, L.NORM_INSTANCE: n * cg, L.NORM_BATCH: cg}[spec.norm]
        mean = torch.empty((G, nstat), device=z.device, dtype=torch.float32)
        rstd = torch.empty((G, nstat), device=z.device, dtype=torch.float32)
        stream = _stream()
        for g in range(G):
            d = L.KcNormDesc()
            d.norm, d.out_act, d.n, d.c, d.hw, d.affine = spec.norm, spec.out_act, n, cg, hw, int(spec.affine)
            d.batch_stride, d.eps = c_total * hw, spec.eps
            scratch = None
            if spec.norm == L.NORM_BATCH:
                if spec.use_batch_stats:
                    scratch = torch.empty(2 * n * cg, device=z.device, dtype=torch.float32)
                else:
                    mean[g].copy_(given_mean[g * cg:(g + 1) * cg])
                    rstd[g].copy_(given_rstd[g * cg:(g + 1) * cg])
            L.check(_timed("kc_instnorm_fwd_kernel", 0.0, 8.0 * n * cg * hw, lambda: lib.kc_norm_act_fwd(
                ctypes.byref(d), _ptr(z[:, g * cg:(g + 1) * cg]), _ptr(gam[g]), _ptr(bet[g]), _ptr(alp[g]),
                _ptr(y[:, g * cg:(g + 1) * cg]), _ptr(mean[g]), _ptr(rstd[g]), _ptr(scratch), stream)), "kc_norm_act_fwd")
        ctx.spec = spec
        ctx.save_for_backward(z, mean, rstd, *params)
        ctx.mark_non_differentiable(mean, rstd)
        return y, mean, rstd

    @staticmethod
    def backward(ctx, dy, _dmean, _drstd):
        lib = L.load()
        spec: NormSpec = ctx.spec
        z, mean, rstd = ctx.saved_tensors[:3]
        params = ctx.saved_tensors[3:]
        if spec.norm == L.NORM_BATCH and not spec.use_batch_stats:
            raise NotImplementedError("norm_act: backward through BatchNorm in eval mode is not implemented")
        G = spec.groups
        n, c_total, h, w = z.shape
        cg, hw = c_total // G, h * w
        gam = [params[2 * g] for g in range(G)] if spec.affine else [None] * G
        bet = [params[2 * g + 1] for g in range(G)] if spec.affine else [None] * G
        off = 2 * G if spec.affine else 0
        alp = [params[off + g] for g in range(G)] if spec.out_act == L.OUT_PRELU else [None] * G
        dy = dy.contiguous()
        dz = torch.empty_like(z)
        grads: List[Optional[torch.Tensor]] = [None] * len(params)
        stream = _stream()
        for g in range(G):
            d = L.KcNormDesc()
            d.norm, d.out_act, d.n, d.c, d.hw, d.affine = spec.norm, spec.out_act, n, cg, hw, int(spec.affine)
            d.batch_stride, d.eps = c_total * hw, spec.eps
            partials = torch.empty(3 * n * cg + 2 * cg, device=z.device, dtype=torch.float32)
            dgam = torch.empty_like(gam[g]) if spec.affine else None
            dbet = torch.empty_like(bet[g]) if spec.affine else None
            dalp = torch.empty_like(alp[g]) if alp[g] is not None else None
            sl = slice(g * cg, (g + 1) * cg)
            L.check(_timed("kc_norm_bwd_kernel", 0.0, 12.0 * n * cg * hw, lambda: lib.kc_norm_act_bwd(
                ctypes.byref(d), _ptr(dy[:, sl]), _ptr(z[:, sl]), _ptr(mean[g]), _ptr(rstd[g]), _ptr(gam[g]), _ptr(bet[g]),
                _ptr(alp[g]), _ptr(dz[:, sl]), _ptr(dgam), _ptr(dbet), _ptr(dalp), _ptr(partials), stream)), "kc_norm_act_bwd")
            if spec.affine:
                grads[2 * g], grads[2 * g + 1] = dgam, dbet
            if dalp is not None:
                grads[off + g] = dalp
        return (None, dz, None, None, *grads)


def norm_act(spec: NormSpec, z: torch.Tensor, gammas: Sequence[torch.Tensor] = (), betas: Sequence[torch.Tensor] = (),
             alphas: Sequence[torch.Tensor] = (), given_mean: Optional[torch.Tensor] = None,
             given_rstd: Optional[torch.Tensor] = None):
    """y = out_act(norm(z)); returns (y, mean, rstd) with per-group statistics ([groups, n*c_g] or [groups, c_g])."""
    params: List[torch.Tensor] = []
    if spec.affine:
        for ga, be in zip(gammas, betas):
            params += [ga, be]
    if spec.out_act == L.OUT_PRELU:
        params += list(alphas)
    return _NormActFn.apply(spec, z, given_mean, given_rstd, *params)


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k: int, s: int):
        lib = L.load()
        _require_cuda(x, "max_pool2d")
        x = x.contiguous()
        n, c, h, w = x.shape
        if h < k or w < k:
            raise ValueError(f"max_pool2d: input {h}x{w} smaller than the window {k}")
        ho, wo = (h - k) // s + 1, (w - k) // s + 1
        y = torch.empty((n, c, ho, wo), device=x.device, dtype=torch.float32)
        idx = torch.empty((n, c, ho, wo), device=x.device, dtype=torch.uint8)
        L.check(_timed("kc_maxpool_fwd_kernel", 0.0, 4.0 * x.numel() + 5.0 * y.numel(), lambda: lib.kc_maxpool2d_fwd(
            _ptr(x), _ptr(y), _ptr(idx), n * c, h, w, k, s, ho, wo, _stream())), "kc_maxpool2d_fwd")
        ctx.save_for_backward(idx)
        ctx.geom = (n, c, h, w, k, s, ho, wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        (idx,) = ctx.saved_tensors
        n, c, h, w, k, s, ho, wo = ctx.geom
        dy = dy.contiguous()
        dx = torch.empty((n, c, h, w), device=dy.device, dtype=torch.float32)
        L.check(_timed("kc_maxpool_bwd_kernel", 0.0, 4.0 * dx.numel() + 5.0 * dy.numel(), lambda: lib.kc_maxpool2d_bwd(
            _ptr(dy), _ptr(idx), _ptr(dx), n * c, h, w, k, s, ho, wo, _stream())), "kc_maxpool2d_bwd")
        return dx, None, None


def max_pool2d(x: torch.Tensor, kernel_size: int, stride: Optional[int] = None) -> torch.Tensor:
    """nn.MaxPool2d(kernel_size, stride) (no padding / dilation / ceil_mode) on the library's own kernels."""
    return _MaxPoolFn.apply(x, int(kernel_size), int(stride if stride is not None else kernel_size))


class MaxPool2d(torch.nn.MaxPool2d):
    """``nn.MaxPool2d`` whose CUDA fp32 path runs kc_maxpool2d_fwd / _bwd (same constructor, repr and semantics); the
    configurations the kernels do not cover (padding, dilation, ceil_mode, return_indices, non-square) use ATen."""

    def forward(self, x):
        k, s = self.kernel_size, self.stride
        plain = (isinstance(k, int) and isinstance(s, int) and self.padding == 0 and self.dilation == 1 and not self.ceil_mode
                 and not self.return_indices and k <= 11)
        if plain and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4:
            return max_pool2d(x, k, s)
        return super().forward(x)
