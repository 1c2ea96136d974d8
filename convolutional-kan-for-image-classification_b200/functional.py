"""torch.autograd.Function wrappers around the C ABI of libkanconv.so.

Differentiable ops that make up the KAN layers of the reference:

  * ``kan_conv``       - z = conv(act(x_base), W_base) + conv(basis(x_basis), W_basis)   (kan_layers.py:199-241 and siblings)
  * ``norm_act``       - y = out_act(norm(z)), Instance / Batch norm over [n, c, hw]     (kan_layers.py:242-243)
  * ``layer_norm_act`` - y = out_act(LayerNorm(z)) over the features of a row            (kan_layers.py:110-112, KANLayer)
  * ``max_pool2d``     - nn.MaxPool2d between convolution stages                         (models/kan_vgg.py:121)

All of them call hand-written CUDA through ctypes (raw device pointers + the CUDA stream of the tensors' device).  CPU tensors
are rejected: there is no fallback path.

What is kept between forward and backward of ``kan_conv``: the layer input ``x`` (fp32, through save_for_backward) and - on the
tensor-core path, unless ``KANCONV_SAVE_PHI=0`` - ``phi``: the bf16 basis / base-activation rows the forward kernel evaluates
anyway, 18 B per input element (11 GB over the 13 layers of KAN-VGG16 at batch 64; the reference keeps the fp32 expansion,
36 B per element, plus ~30 autograd intermediates).  ``phi`` lives on ``ctx`` (not in save_for_backward, so saved-tensor hooks
do not see it) and is released by the first backward; a second backward through a retained graph re-evaluates the rows in a
pre-pass and gives the same gradients.  With ``KANCONV_SAVE_PHI=0`` the pre-pass always runs (+6 ms per VGG16 step).

Packed bf16 weight images are cached per parameter and re-packed only when the parameter's version counter changes (once per
optimizer step in training, never in inference).
"""
from __future__ import annotations

import ctypes
import os
import weakref
from dataclasses import dataclass, replace as _dc_replace
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L

_TC_BACKWARD = os.environ.get("KANCONV_TC_BACKWARD", "1") != "0"   # debug switch: FP32 CUDA-core backward after a TC forward
# KANCONV_SAVE_PHI=0: do not keep the bf16 basis rows between forward and backward (saves 18 B per input element per
# layer of activation memory; the weight gradient then re-evaluates them in a pre-pass)
_SAVE_PHI = os.environ.get("KANCONV_SAVE_PHI", "1") != "0"
# KANCONV_FUSED_NORM_BWD=0: debug switch, norm backward to fp32 dz + separate conversion to the bf16 flat layout
_FUSED_NORM_BWD = os.environ.get("KANCONV_FUSED_NORM_BWD", "1") != "0"
# norm + activation of a grouped layer as ONE launch over all channels instead of one launch per group (A/B switch)
_MERGE_NORM_GROUPS = os.environ.get("KANCONV_MERGE_NORM_GROUPS", "1") != "0"
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


# Every buffer the binding hands to the library is allocated through these two functions.  tests/test_redzones_gpu.py swaps
# _ALLOC for an allocator that surrounds each buffer with poisoned guard bands and checks them after the kernels ran
# (compute-sanitizer is closed on the GPU pool this repo is developed on, so out-of-bounds writes are hunted this way).
_ALLOC = torch.empty


def _empty(*shape, dtype, device):
    return _ALLOC(*shape, dtype=dtype, device=device)


def _empty_like(t: torch.Tensor):
    return _ALLOC(tuple(t.shape), dtype=t.dtype, device=t.device)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev: torch.device):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: kanconv_b200 runs on CUDA tensors only (no CPU fallback); got device {t.device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype}")


def _same_device(what: str, dev: torch.device, *tensors) -> None:
    """The library launches on the CURRENT device with raw pointers: every operand must live on the device the op runs on."""
    for t in tensors:
        if t is not None and t.device != dev:
            raise RuntimeError(f"{what}: all tensors must be on {dev}, got one on {t.device}")


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


# ---- packed-weight cache -------------------------------------------------------------------------------------------------
# key: (id of the root tensor of w_base, id of the root of w_basis, which image, descriptor bytes) -> (versions, packed).
# "Root" = the Parameter a view was taken from (1-D layers and KANLayer pass views); entries die with their parameter.
_PACKS: Dict[tuple, tuple] = {}
_PACK_STATS = {"hits": 0, "misses": 0}


def _root(t: Optional[torch.Tensor]):
    if t is None:
        return None
    return t._base if t._base is not None else t


def _pack_key(roots, which: int, d) -> tuple:
    return (tuple(0 if r is None else id(r) for r in roots), which, bytes(d))


def _drop_packs(rid: int) -> None:
    for k in [k for k in _PACKS if rid in k[0]]:
        _PACKS.pop(k, None)


def _packed_weights(lib, d, which: int, wb, ws, roots, versions, stream, kernel_name: str):
    """bf16 weight image of the forward (which=0) or dgrad (which=1) kernel; re-packed only when a version counter moved."""
    key = _pack_key(roots, which, d)
    hit = _PACKS.get(key)
    if hit is not None and hit[0] == versions and hit[1].device == ws.device:
        _PACK_STATS["hits"] += 1
        return hit[1]
    _PACK_STATS["misses"] += 1
    nbytes = lib.kc_tc_bytes(ctypes.byref(d), which)
    packed = hit[1] if (hit is not None and hit[1].numel() == nbytes and hit[1].device == ws.device) else \
        _empty(nbytes, device=ws.device, dtype=torch.uint8)
    L.check(_timed(kernel_name, 0.0, 1.5 * nbytes, lambda: lib.kc_tc_pack_weights(
        ctypes.byref(d), _ptr(wb), _ptr(ws), _ptr(packed) if which == 0 else None, _ptr(packed) if which == 1 else None,
        stream)), "kc_tc_pack_weights")
    if hit is None:
        for r in roots:
            if r is not None:
                weakref.finalize(r, _drop_packs, id(r))
    _PACKS[key] = (versions, packed)
    return packed


def pack_cache_stats() -> dict:
    return dict(_PACK_STATS, entries=len(_PACKS))


def clear_pack_cache() -> None:
    """Drop every cached weight image (needed only after writing weights through ``.data``, which bypasses version counters)."""
    _PACKS.clear()


# ---- depthwise layers (groups == channels) ---------------------------------------------------------------------------------
def _dw_desc(lib, spec: ConvSpec, precision: Optional[str], xb, beta, w_basis):
    """Descriptor of the one-launch depthwise kernels (csrc/kc_dw.cu) when the layer is a stack of single-channel groups they
    cover, else None.  ``precision == "bf16"`` (tensor cores forced) keeps the per-group route."""
    G = spec.groups
    n, c_total, h, w = xb.shape
    if G < 2 or c_total != G or w_basis[0].shape[0] != 1 or beta is not None or precision == "bf16":
        return None
    ho, wo = spec.out_hw(h, w)
    d = _make_desc(spec, n, 1, h, w, 1, c_total * h * w, G * ho * wo)
    return d if lib.kc_dwconv_supported(ctypes.byref(d), G) else None


def _dw_stack(ws):
    """Per-group filters [1, rows, kh, kw] -> one [G, rows * kh * kw] matrix in group order."""
    return torch.cat([w.reshape(1, -1) for w in ws], dim=0)


def _dw_bytes(d, G: int, passes_in: int, passes_out: int) -> float:
    return 4.0 * d.n * G * (passes_in * d.h * d.w + passes_out * d.ho * d.wo)


def _split_weights(spec: ConvSpec, weights):
    G = spec.groups
    w_base = list(weights[:G]) if spec.has_base else [None] * G
    w_basis = list(weights[G:2 * G]) if spec.has_base else list(weights[:G])
    return w_base, w_basis


def _conv_fwd(spec: ConvSpec, precision: str, xb, xs, beta, weights, want_phi: bool):
    """Forward of all groups.  -> (z, used_tc, phis, roots).  xb / xs contiguous, same device."""
    lib = L.load()
    dev = xb.device
    G = spec.groups
    n, c_total, h, w = xb.shape
    cg = c_total // G
    w_base, w_basis = _split_weights(spec, weights)
    og = w_basis[0].shape[0]
    ho, wo = spec.out_hw(h, w)
    used_tc, phis = [], []
    roots = [(_root(w_base[g]), _root(w_basis[g])) for g in range(G)]
    with torch.cuda.device(dev):
        stream = _stream(dev)
        z = _empty((n, og * G, ho, wo), device=dev, dtype=torch.float32)
        dw = _dw_desc(lib, spec, precision, xb, beta, w_basis)
        if dw is not None:          # groups == channels: all groups in one launch
            wb_all = _dw_stack(w_base) if spec.has_base else None
            ws_all = _dw_stack(w_basis)
            L.check(_timed("kc_dw_fwd_kernel", G * _conv_flops(dw), _dw_bytes(dw, G, 1, 1), lambda: lib.kc_dwconv_fwd_f32(
                ctypes.byref(dw), G, _ptr(xb), _ptr(xs), _ptr(wb_all), _ptr(ws_all), _ptr(z), stream)), "kc_dwconv_fwd_f32")
            return z, [False] * G, [None] * G, roots
        for g in range(G):
            d = _make_desc(spec, n, cg, h, w, og, c_total * h * w, og * G * ho * wo)
            xbg, xsg, zg = xb[:, g * cg:(g + 1) * cg], xs[:, g * cg:(g + 1) * cg], z[:, g * og:(g + 1) * og]
            wbg = None if w_base[g] is None else w_base[g].contiguous()
            wsg = w_basis[g].contiguous()
            tc = _use_tc(lib, d, precision)
            used_tc.append(tc)
            if tc:
                versions = (None if w_base[g] is None else w_base[g]._version, w_basis[g]._version)
                packed = _packed_weights(lib, d, 0, wbg, wsg, roots[g], versions, stream, "kc_pack_fwd_kernel")
                phi = None
                if (want_phi or lib.kc_tc_fwd_needs_phi(ctypes.byref(d))) and lib.kc_tc_bytes(ctypes.byref(d), 4) > 0:
                    phi = _empty(lib.kc_tc_bytes(ctypes.byref(d), 4), device=dev, dtype=torch.uint8)
                phis.append(phi)
                L.check(_timed("kc_tc_kernel<fwd>", _conv_flops(d), 0.0, lambda: lib.kc_conv_fwd_tc(
                    ctypes.byref(d), _ptr(xbg), _ptr(xsg), _ptr(packed), _ptr(beta), _ptr(zg), _ptr(phi), stream)),
                    "kc_conv_fwd_tc")
            else:
                phis.append(None)
                L.check(_timed("kc_fwd_simt_kernel", _conv_flops(d), 0.0, lambda: lib.kc_conv_fwd_f32(
                    ctypes.byref(d), _ptr(xbg), _ptr(xsg), _ptr(wbg), _ptr(wsg), _ptr(beta), _ptr(zg), stream)),
                    "kc_conv_fwd_f32")
    return z, used_tc, phis, roots


def _conv_bwd(spec: ConvSpec, used_tc, phis, roots, xb, xs, alias: bool, beta, weights, dz, dzf_of_group,
              need_dx: bool, need_dbeta: bool, need_w):
    """Backward of all groups.  ``dz`` = fp32 gradient of z, or None when ``dzf_of_group(g, desc)`` supplies the bf16 flat buffer
    of every tensor-core group (fused norm backward).  -> (dx_base, dx_basis, dbeta, dws)."""
    lib = L.load()
    dev = xb.device
    G = spec.groups
    n, c_total, h, w = xb.shape
    cg = c_total // G
    w_base, w_basis = _split_weights(spec, weights)
    og = w_basis[0].shape[0]
    ho, wo = spec.out_hw(h, w)
    gram = spec.basis == L.BASIS_GRAM and beta is not None
    run_dgrad = need_dx or (gram and need_dbeta)
    dx_base = dx_basis = dbeta = None
    dws: List[Optional[torch.Tensor]] = [None] * len(weights)
    with torch.cuda.device(dev):
        stream = _stream(dev)
        if run_dgrad:
            dx_base = _empty_like(xb)
            dx_basis = dx_base if alias else _empty_like(xs)
        dw = None if (any(used_tc) or dz is None) else _dw_desc(lib, spec, None, xb, beta, w_basis)
        if dw is not None:          # groups == channels: dX and dW of all groups in one launch each
            wb_all = _dw_stack(w_base) if spec.has_base else None
            ws_all = _dw_stack(w_basis)
            if run_dgrad:
                L.check(_timed("kc_dw_dgrad_kernel", G * _conv_flops(dw), _dw_bytes(dw, G, 2, 1), lambda: lib.kc_dwconv_dgrad_f32(
                    ctypes.byref(dw), G, _ptr(dz), _ptr(xb), _ptr(xs), _ptr(wb_all), _ptr(ws_all), _ptr(dx_base), _ptr(dx_basis),
                    stream)), "kc_dwconv_dgrad_f32")
            if any(need_w):
                dwb_all = _empty_like(wb_all) if wb_all is not None else None
                dws_all = _empty_like(ws_all)
                wsp = _empty(max(lib.kc_dwconv_wgrad_workspace_bytes(ctypes.byref(dw), G), 16), device=dev, dtype=torch.uint8)
                L.check(_timed("kc_dw_wgrad_kernel", G * _conv_flops(dw), _dw_bytes(dw, G, 1, 1), lambda: lib.kc_dwconv_wgrad_f32(
                    ctypes.byref(dw), G, _ptr(dz), _ptr(xb), _ptr(xs), _ptr(dwb_all), _ptr(dws_all), _ptr(wsp), stream)),
                    "kc_dwconv_wgrad_f32")
                for g in range(G):
                    if spec.has_base:
                        dws[g] = dwb_all[g].view_as(w_base[g])
                        dws[G + g] = dws_all[g].view_as(w_basis[g])
                    else:
                        dws[g] = dws_all[g].view_as(w_basis[g])
            return dx_base, dx_basis, dbeta, dws
        for g in range(G):
            d = _make_desc(spec, n, cg, h, w, og, c_total * h * w, og * G * ho * wo)
            sl = slice(g * cg, (g + 1) * cg)
            xbg, xsg = xb[:, sl], xs[:, sl]
            dzg = None if dz is None else dz[:, g * og:(g + 1) * og]
            wbg = None if w_base[g] is None else w_base[g].contiguous()
            wsg = w_basis[g].contiguous()
            tc_bwd = used_tc[g] and _TC_BACKWARD and lib.kc_tc_bytes(ctypes.byref(d), 1) > 0
            dzf = None
            if tc_bwd:
                dzf = dzf_of_group(g, d) if dzf_of_group is not None else None
                if dzf is None:
                    dzf = _empty(lib.kc_tc_bytes(ctypes.byref(d), 2), device=dev, dtype=torch.uint8)
                    L.check(_timed("kc_dz_flat_kernel", 0.0, 6.0 * dzg.numel(), lambda: lib.kc_tc_dz_flat(
                        ctypes.byref(d), _ptr(dzg), _ptr(dzf), stream)), "kc_tc_dz_flat")
            if run_dgrad:
                # GRAM: d/d beta_weights of this group (deterministic: per-block partial rows + fixed-order reduction)
                dbg = None
                if gram and need_dbeta:
                    dbg = _empty(lib.kc_dbeta_floats(ctypes.byref(d), 1 if tc_bwd else 0), device=dev, dtype=torch.float32)
                if tc_bwd:
                    versions = (None if w_base[g] is None else w_base[g]._version, w_basis[g]._version)
                    packed_d = _packed_weights(lib, d, 1, wbg, wsg, roots[g], versions, stream, "kc_pack_dgrad_kernel")
                    L.check(_timed("kc_tc_kernel<dgrad>", _conv_flops(d), 0.0, lambda: lib.kc_conv_dgrad_tc(
                        ctypes.byref(d), None, _ptr(xbg), _ptr(xsg), _ptr(packed_d), _ptr(beta), _ptr(dx_base[:, sl]),
                        _ptr(dx_basis[:, sl]), _ptr(dbg), _ptr(dzf), stream)), "kc_conv_dgrad_tc")
                else:
                    L.check(_timed("kc_dgrad_simt_kernel", _conv_flops(d), 0.0, lambda: lib.kc_conv_dgrad_f32(
                        ctypes.byref(d), _ptr(dzg), _ptr(xbg), _ptr(xsg), _ptr(wbg), _ptr(wsg), _ptr(beta),
                        _ptr(dx_base[:, sl]), _ptr(dx_basis[:, sl]), _ptr(dbg), stream)), "kc_conv_dgrad_f32")
                if dbg is not None:
                    part = dbg[:beta.numel()]
                    dbeta = part.clone() if dbeta is None else dbeta + part       # groups are added in order
            wi_base, wi_basis = (g, G + g) if spec.has_base else (None, g)
            if (wi_base is not None and need_w[wi_base]) or need_w[wi_basis]:
                dwb = _empty_like(wbg) if wbg is not None else None
                dwsg = _empty_like(wsg)
                if tc_bwd and lib.kc_tc_bytes(ctypes.byref(d), 3) > 0:
                    phi = phis[g]
                    ws = _empty(lib.kc_tc_bytes(ctypes.byref(d), 5 if phi is not None else 3), device=dev, dtype=torch.uint8)
                    L.check(_timed("kc_wgrad_tc_kernel", _conv_flops(d), 0.0, lambda: lib.kc_conv_wgrad_tc(
                        ctypes.byref(d), _ptr(dzf), _ptr(xbg), _ptr(xsg), _ptr(beta), _ptr(phi), _ptr(dwb), _ptr(dwsg), _ptr(ws),
                        stream)), "kc_conv_wgrad_tc")
                    phis[g] = None
                else:
                    nbytes = lib.kc_wgrad_workspace_bytes(ctypes.byref(d))
                    ws = _empty(max(nbytes, 16), device=dev, dtype=torch.uint8)
                    L.check(_timed("kc_wgrad_simt_kernel", _conv_flops(d), 0.0, lambda: lib.kc_conv_wgrad_f32(
                        ctypes.byref(d), _ptr(dzg), _ptr(xbg), _ptr(xsg), _ptr(beta), _ptr(dwb), _ptr(dwsg), _ptr(ws), stream)),
                        "kc_conv_wgrad_f32")
                if wi_base is not None:
                    dws[wi_base] = dwb
                dws[wi_basis] = dwsg
    return dx_base, dx_basis, dbeta, dws


class _KanConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec: ConvSpec, precision: str, x_base, x_basis, beta, *weights):
        alias = x_basis is None
        xb = x_base.contiguous()
        xs = xb if alias else x_basis.contiguous()
        _require_cuda(xb, "kan_conv")
        _require_cuda(xs, "kan_conv")
        _same_device("kan_conv", xb.device, xs, beta, *weights)
        # basis rows saved for the weight gradient (the reference keeps the expanded basis alive for autograd, too)
        want_phi = _SAVE_PHI and _TC_BACKWARD and any(ctx.needs_input_grad[5:])
        z, used_tc, phis, roots = _conv_fwd(spec, precision, xb, xs, beta, weights, want_phi)
        ctx.spec, ctx.alias, ctx.used_tc, ctx.phis, ctx.roots = spec, alias, used_tc, phis, roots
        ctx.save_for_backward(xb, xs if not alias else None, beta, *weights)
        return z

    @staticmethod
    def backward(ctx, dz):
        saved = ctx.saved_tensors
        xb, xs, beta = saved[0], saved[1], saved[2]
        weights = saved[3:]
        alias = ctx.alias
        if alias:
            xs = xb
        _same_device("kan_conv backward", xb.device, dz)
        # inputs of forward: (spec, precision, x_base, x_basis, beta, *weights)
        need_dxb, need_dxs, need_dbeta = ctx.needs_input_grad[2], ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        dx_base, dx_basis, dbeta, dws = _conv_bwd(ctx.spec, ctx.used_tc, ctx.phis, ctx.roots, xb, xs, alias, beta, weights,
                                                  dz.contiguous(), None, need_dxb or need_dxs, need_dbeta,
                                                  list(ctx.needs_input_grad[5:]))
        return (None, None, dx_base if need_dxb else None, (dx_basis if (need_dxs and not alias) else None),
                dbeta if need_dbeta else None, *dws)


def kan_conv(spec: ConvSpec, x_base: torch.Tensor, x_basis: Optional[torch.Tensor], beta: Optional[torch.Tensor],
             w_base: Sequence[torch.Tensor], w_basis: Sequence[torch.Tensor], precision: Optional[str] = None) -> torch.Tensor:
    """Pre-normalisation output of a KAN convolution layer.  ``x_basis=None`` means "same tensor as x_base"."""
    weights = (list(w_base) if spec.has_base else []) + list(w_basis)
    return _KanConvFn.apply(spec, precision or _PRECISION, x_base, x_basis, beta, *weights)


def _norm_params(spec: NormSpec, params):
    G = spec.groups
    # params layout: [gamma_0, beta_0, ..., gamma_{G-1}, beta_{G-1}] if affine, then [alpha_0..alpha_{G-1}] if PReLU
    gam = [params[2 * g] for g in range(G)] if spec.affine else [None] * G
    bet = [params[2 * g + 1] for g in range(G)] if spec.affine else [None] * G
    off = 2 * G if spec.affine else 0
    alp = [params[off + g] for g in range(G)] if spec.out_act == L.OUT_PRELU else [None] * G
    return gam, bet, alp, off


def _norm_desc(spec: NormSpec, n, cg, hw, c_total) -> L.KcNormDesc:
    d = L.KcNormDesc()
    d.norm, d.out_act, d.n, d.c, d.hw, d.affine = spec.norm, spec.out_act, n, cg, hw, int(spec.affine)
    d.batch_stride, d.eps = c_total * hw, spec.eps
    return d


# ---- grouped norms in one launch --------------------------------------------------------------------------------------------
# The norm kernels index CHANNELS; a group is only a separate set of parameter tensors (gamma_g, beta_g) and, for PReLU, one
# scalar alpha behind a single pointer.  Without PReLU all groups can therefore run as one launch over all channels with the
# affine parameters concatenated - what makes depthwise layers (hundreds of one-channel groups, each with its own norm module
# upstream: kan_layers.py:178-182) practical.  The statistics keep the per-group layout [G, ...] towards the callers.
def _norm_merged(spec: NormSpec) -> bool:
    return _MERGE_NORM_GROUPS and spec.groups > 1 and spec.out_act != L.OUT_PRELU


def _merged_params(spec: NormSpec, params):
    if not spec.affine:
        return []
    G = spec.groups
    return [torch.cat([params[2 * g].reshape(-1) for g in range(G)]), torch.cat([params[2 * g + 1].reshape(-1) for g in range(G)])]


def _stats_to_groups(t, spec: NormSpec, n: int, cg: int):
    """[1, nstat of all channels] -> [G, nstat of one group] (instance norm: planes are ordered [n][c])."""
    G = spec.groups
    if spec.norm == L.NORM_BATCH:
        return t.reshape(G, cg)
    if spec.norm == L.NORM_INSTANCE:
        return t.reshape(n, G, cg).permute(1, 0, 2).reshape(G, n * cg)
    return t.reshape(1, 1).expand(G, 1)


def _stats_from_groups(t, spec: NormSpec, n: int, cg: int):
    G = spec.groups
    if spec.norm == L.NORM_BATCH:
        return t.reshape(1, G * cg)
    if spec.norm == L.NORM_INSTANCE:
        return t.reshape(G, n, cg).permute(1, 0, 2).reshape(1, n * G * cg).contiguous()
    return t[:1].contiguous()


def _norm_fwd(spec: NormSpec, z, given_mean, given_rstd, params):
    """-> (y, mean, rstd); z contiguous."""
    if _norm_merged(spec):
        n, cg = z.shape[0], z.shape[1] // spec.groups
        y, mean, rstd = _norm_fwd(_dc_replace(spec, groups=1), z, given_mean, given_rstd, _merged_params(spec, params))
        return y, _stats_to_groups(mean, spec, n, cg), _stats_to_groups(rstd, spec, n, cg)
    lib = L.load()
    dev = z.device
    G = spec.groups
    n, c_total, h, w = z.shape
    cg, hw = c_total // G, h * w
    gam, bet, alp, _ = _norm_params(spec, params)
    given = spec.norm == L.NORM_BATCH and not spec.use_batch_stats
    if given and (given_mean is None or given_rstd is None):
        raise ValueError("norm_act: BatchNorm with use_batch_stats=False needs given_mean / given_rstd")
    nstat = {L.NORM_NONE: 1, L.NORM_INSTANCE: n * cg, L.NORM_BATCH: cg}[spec.norm]
    with torch.cuda.device(dev):
        stream = _stream(dev)
        y = _empty_like(z)
        if given:       # eval-mode BatchNorm: the running statistics are inputs of the kernel (scratch = NULL)
            mean = given_mean.detach().to(torch.float32).reshape(G, cg).contiguous()
            rstd = given_rstd.detach().to(torch.float32).reshape(G, cg).contiguous()
        else:
            mean = _empty((G, nstat), device=dev, dtype=torch.float32)
            rstd = _empty((G, nstat), device=dev, dtype=torch.float32)
        for g in range(G):
            d = _norm_desc(spec, n, cg, hw, c_total)
            scratch = None
            if spec.norm == L.NORM_BATCH and not given:
                scratch = _empty(2 * n * cg, device=dev, dtype=torch.float32)
            L.check(_timed("kc_instnorm_fwd_kernel", 0.0, 8.0 * n * cg * hw, lambda: lib.kc_norm_act_fwd(
                ctypes.byref(d), _ptr(z[:, g * cg:(g + 1) * cg]), _ptr(gam[g]), _ptr(bet[g]), _ptr(alp[g]),
                _ptr(y[:, g * cg:(g + 1) * cg]), _ptr(mean[g]), _ptr(rstd[g]), _ptr(scratch), stream)), "kc_norm_act_fwd")
    return y, mean, rstd


def _norm_bwd(spec: NormSpec, z, mean, rstd, params, dy):
    """-> (dz fp32, grads of params); dy contiguous."""
    if _norm_merged(spec):
        G, n, cg = spec.groups, z.shape[0], z.shape[1] // spec.groups
        dz, g1 = _norm_bwd(_dc_replace(spec, groups=1), z, _stats_from_groups(mean, spec, n, cg),
                           _stats_from_groups(rstd, spec, n, cg), _merged_params(spec, params), dy)
        grads: List[Optional[torch.Tensor]] = [None] * len(params)
        if spec.affine:
            for g in range(G):
                grads[2 * g] = g1[0][g * cg:(g + 1) * cg].view_as(params[2 * g])
                grads[2 * g + 1] = g1[1][g * cg:(g + 1) * cg].view_as(params[2 * g + 1])
        return dz, grads
    lib = L.load()
    dev = z.device
    given = int(spec.norm == L.NORM_BATCH and not spec.use_batch_stats)
    G = spec.groups
    n, c_total, h, w = z.shape
    cg, hw = c_total // G, h * w
    gam, bet, alp, off = _norm_params(spec, params)
    grads: List[Optional[torch.Tensor]] = [None] * len(params)
    with torch.cuda.device(dev):
        stream = _stream(dev)
        dz = _empty_like(z)
        for g in range(G):
            d = _norm_desc(spec, n, cg, hw, c_total)
            partials = _empty(3 * n * cg + 2 * cg, device=dev, dtype=torch.float32)
            dgam = _empty_like(gam[g]) if spec.affine else None
            dbet = _empty_like(bet[g]) if spec.affine else None
            dalp = _empty_like(alp[g]) if alp[g] is not None else None
            sl = slice(g * cg, (g + 1) * cg)
            L.check(_timed("kc_norm_bwd_kernel", 0.0, 12.0 * n * cg * hw, lambda: lib.kc_norm_act_bwd(
                ctypes.byref(d), _ptr(dy[:, sl]), _ptr(z[:, sl]), _ptr(mean[g]), _ptr(rstd[g]), _ptr(gam[g]), _ptr(bet[g]),
                _ptr(alp[g]), _ptr(dz[:, sl]), _ptr(dgam), _ptr(dbet), _ptr(dalp), _ptr(partials), given, stream)),
                "kc_norm_act_bwd")
            if spec.affine:
                grads[2 * g], grads[2 * g + 1] = dgam, dbet
            if dalp is not None:
                grads[off + g] = dalp
    return dz, grads


class _NormActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec: NormSpec, z, given_mean, given_rstd, *params):
        z = z.contiguous()
        _require_cuda(z, "norm_act")
        _same_device("norm_act", z.device, given_mean, given_rstd, *params)
        y, mean, rstd = _norm_fwd(spec, z, given_mean, given_rstd, params)
        ctx.spec = spec
        ctx.save_for_backward(z, mean, rstd, *params)
        ctx.mark_non_differentiable(mean, rstd)
        return y, mean, rstd

    @staticmethod
    def backward(ctx, dy, _dmean, _drstd):
        z, mean, rstd = ctx.saved_tensors[:3]
        params = ctx.saved_tensors[3:]
        _same_device("norm_act backward", z.device, dy)
        dz, grads = _norm_bwd(ctx.spec, z, mean, rstd, params, dy.contiguous())
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


class _KanLayerFn(torch.autograd.Function):
    """conv -> norm -> output activation of one KAN convolution layer as ONE autograd node (kan_layers.py:199-243).

    Same kernels as ``kan_conv`` followed by ``norm_act`` in the forward.  The point is the backward: on the tensor-core path
    the norm backward writes dz directly in the bf16 flat operand layout of the dgrad / wgrad kernels
    (``kc_norm_bwd_dz_flat``), so the fp32 dz tensor and the separate conversion pass never exist."""

    @staticmethod
    def forward(ctx, spec: ConvSpec, nspec: NormSpec, precision: str, x, beta, given_mean, given_rstd, n_weights: int, *params):
        weights, nparams = params[:n_weights], params[n_weights:]
        xb = x.contiguous()
        _require_cuda(xb, "kan_layer")
        _same_device("kan_layer", xb.device, beta, given_mean, given_rstd, *params)
        want_phi = _SAVE_PHI and _TC_BACKWARD and any(ctx.needs_input_grad[8:8 + n_weights])
        z, used_tc, phis, roots = _conv_fwd(spec, precision, xb, xb, beta, weights, want_phi)
        y, mean, rstd = _norm_fwd(nspec, z, given_mean, given_rstd, nparams)
        ctx.spec, ctx.nspec, ctx.used_tc, ctx.phis, ctx.roots, ctx.n_weights = spec, nspec, used_tc, phis, roots, n_weights
        ctx.save_for_backward(xb, beta, z, mean, rstd, *params)
        ctx.mark_non_differentiable(mean, rstd)
        return y, mean, rstd

    @staticmethod
    def backward(ctx, dy, _dmean, _drstd):
        lib = L.load()
        spec, nspec, nw = ctx.spec, ctx.nspec, ctx.n_weights
        xb, beta, z, mean, rstd = ctx.saved_tensors[:5]
        params = ctx.saved_tensors[5:]
        weights, nparams = params[:nw], params[nw:]
        dev = xb.device
        _same_device("kan_layer backward", dev, dy)
        dy = dy.contiguous()
        G = nspec.groups
        n, c_total, ho, wo = z.shape
        cg, hw = c_total // G, ho * wo
        _, _, alp, off = _norm_params(nspec, nparams)
        ngrads: List[Optional[torch.Tensor]] = [None] * len(nparams)
        fused = [False] * G

        def dzf_of_group(g, d):
            """bf16 flat dz of group g straight from (dy, z): the fused norm backward, when the shape allows it."""
            nd = _norm_desc(nspec, n, cg, hw, c_total)
            if not lib.kc_norm_bwd_dz_flat_supported(ctypes.byref(d), ctypes.byref(nd)):
                return None
            sl = slice(g * cg, (g + 1) * cg)
            dzf = _empty(lib.kc_tc_bytes(ctypes.byref(d), 2), device=dev, dtype=torch.uint8)
            partials = _empty(3 * n * cg, device=dev, dtype=torch.float32)
            dalp = _empty_like(alp[g]) if alp[g] is not None else None
            L.check(_timed("kc_norm_bwd_flat_kernel", 0.0, 10.0 * n * cg * hw, lambda: lib.kc_norm_bwd_dz_flat(
                ctypes.byref(d), ctypes.byref(nd), _ptr(dy[:, sl]), _ptr(z[:, sl]), _ptr(mean[g]), _ptr(rstd[g]), _ptr(alp[g]),
                _ptr(dzf), _ptr(dalp), _ptr(partials), _stream(dev))), "kc_norm_bwd_dz_flat")
            if dalp is not None:
                ngrads[off + g] = dalp
            fused[g] = True
            return dzf

        # which groups take the fused path is decided by the same predicates _conv_bwd uses
        n_in, cin_total, h, w = xb.shape
        d0 = _make_desc(spec, n_in, cin_total // spec.groups, h, w, cg, cin_total * h * w, c_total * hw)
        nd0 = _norm_desc(nspec, n, cg, hw, c_total)
        all_fused = _TC_BACKWARD and _FUSED_NORM_BWD and all(ctx.used_tc) and lib.kc_tc_bytes(ctypes.byref(d0), 1) > 0 and \
            lib.kc_tc_bytes(ctypes.byref(d0), 3) > 0 and \
            bool(lib.kc_norm_bwd_dz_flat_supported(ctypes.byref(d0), ctypes.byref(nd0)))
        dz = None
        with torch.cuda.device(dev):
            if not all_fused:
                dz, ngrads_full = _norm_bwd(nspec, z, mean, rstd, nparams, dy)
                ngrads = list(ngrads_full)
            need = ctx.needs_input_grad
            dx, _, dbeta, dws = _conv_bwd(spec, ctx.used_tc, ctx.phis, ctx.roots, xb, xb, True, beta, weights, dz,
                                          dzf_of_group if all_fused else None, need[3], need[4], list(need[8:8 + nw]))
        return (None, None, None, dx if need[3] else None, dbeta if need[4] else None, None, None, None, *dws, *ngrads)


def kan_layer(spec: ConvSpec, nspec: NormSpec, x: torch.Tensor, beta: Optional[torch.Tensor], w_base: Sequence[torch.Tensor],
              w_basis: Sequence[torch.Tensor], gammas: Sequence[torch.Tensor] = (), betas: Sequence[torch.Tensor] = (),
              alphas: Sequence[torch.Tensor] = (), given_mean: Optional[torch.Tensor] = None,
              given_rstd: Optional[torch.Tensor] = None, precision: Optional[str] = None):
    """out_act(norm(kan_conv(x))) as one differentiable op; returns (y, mean, rstd) like ``norm_act``."""
    weights = (list(w_base) if spec.has_base else []) + list(w_basis)
    nparams: List[torch.Tensor] = []
    if nspec.affine:
        for ga, be in zip(gammas, betas):
            nparams += [ga, be]
    if nspec.out_act == L.OUT_PRELU:
        nparams += list(alphas)
    return _KanLayerFn.apply(spec, nspec, precision or _PRECISION, x, beta, given_mean, given_rstd, len(weights), *weights, *nparams)


class _LayerNormActFn(torch.autograd.Function):
    """y = out_act(LayerNorm(z)) over the last axis - the tail of the reference's KANLayer (kan_layers.py:110-112)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, alpha, eps: float, out_act: int):
        lib = L.load()
        _require_cuda(z, "layer_norm_act")
        dev = z.device
        _same_device("layer_norm_act", dev, gamma, beta, alpha)
        feat = z.shape[-1]
        z2 = z.contiguous().reshape(-1, feat)
        d = L.KcRowNormDesc()
        d.rows, d.features, d.out_act, d.affine, d.eps = z2.shape[0], feat, out_act, int(gamma is not None), eps
        with torch.cuda.device(dev):
            y = _empty_like(z2)
            mean = _empty(z2.shape[0], device=dev, dtype=torch.float32)
            rstd = _empty_like(mean)
            L.check(lib.kc_layernorm_act_fwd(ctypes.byref(d), _ptr(z2), _ptr(gamma), _ptr(beta), _ptr(alpha), _ptr(y), _ptr(mean),
                                             _ptr(rstd), _stream(dev)), "kc_layernorm_act_fwd")
        ctx.desc_fields = (z2.shape[0], feat, out_act, int(gamma is not None), eps)
        ctx.shape = z.shape
        ctx.save_for_backward(z2, mean, rstd, gamma, beta, alpha)
        return y.reshape(z.shape)

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        z2, mean, rstd, gamma, beta, alpha = ctx.saved_tensors
        dev = z2.device
        _same_device("layer_norm_act backward", dev, dy)
        d = L.KcRowNormDesc()
        d.rows, d.features, d.out_act, d.affine, d.eps = ctx.desc_fields
        dy2 = dy.contiguous().reshape(z2.shape)
        with torch.cuda.device(dev):
            dz = _empty_like(z2)
            dgam = _empty_like(gamma) if gamma is not None else None
            dbet = _empty_like(beta) if beta is not None else None
            dalp = _empty_like(alpha) if alpha is not None else None
            partials = _empty(z2.shape[0], device=dev, dtype=torch.float32)
            L.check(lib.kc_layernorm_act_bwd(ctypes.byref(d), _ptr(dy2), _ptr(z2), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(beta),
                                             _ptr(alpha), _ptr(dz), _ptr(dgam), _ptr(dbet), _ptr(dalp), _ptr(partials),
                                             _stream(dev)), "kc_layernorm_act_bwd")
        return dz.reshape(ctx.shape), dgam, dbet, dalp, None, None


def layer_norm_act(z: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor],
                   alpha: Optional[torch.Tensor], eps: float = 1e-5, out_act: Optional[int] = None) -> torch.Tensor:
    """out_act(LayerNorm_{last axis}(z)) with per-feature gamma / beta; out_act defaults to PReLU when alpha is given."""
    if out_act is None:
        out_act = L.OUT_PRELU if alpha is not None else L.OUT_NONE
    return _LayerNormActFn.apply(z, gamma, beta, alpha, float(eps), int(out_act))


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k: int, s: int):
        lib = L.load()
        _require_cuda(x, "max_pool2d")
        x = x.contiguous()
        dev = x.device
        n, c, h, w = x.shape
        if h < k or w < k:
            raise ValueError(f"max_pool2d: input {h}x{w} smaller than the window {k}")
        ho, wo = (h - k) // s + 1, (w - k) // s + 1
        with torch.cuda.device(dev):
            y = _empty((n, c, ho, wo), device=dev, dtype=torch.float32)
            idx = _empty((n, c, ho, wo), device=dev, dtype=torch.uint8)
            L.check(_timed("kc_maxpool_fwd_kernel", 0.0, 4.0 * x.numel() + 5.0 * y.numel(), lambda: lib.kc_maxpool2d_fwd(
                _ptr(x), _ptr(y), _ptr(idx), n * c, h, w, k, s, ho, wo, _stream(dev))), "kc_maxpool2d_fwd")
        ctx.save_for_backward(idx)
        ctx.geom = (n, c, h, w, k, s, ho, wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        (idx,) = ctx.saved_tensors
        n, c, h, w, k, s, ho, wo = ctx.geom
        dev = idx.device
        _same_device("max_pool2d backward", dev, dy)
        dy = dy.contiguous()
        with torch.cuda.device(dev):
            dx = _empty((n, c, h, w), device=dev, dtype=torch.float32)
            L.check(_timed("kc_maxpool_bwd_kernel", 0.0, 4.0 * dx.numel() + 5.0 * dy.numel(), lambda: lib.kc_maxpool2d_bwd(
                _ptr(dy), _ptr(idx), _ptr(dx), n * c, h, w, k, s, ho, wo, _stream(dev))), "kc_maxpool2d_bwd")
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
