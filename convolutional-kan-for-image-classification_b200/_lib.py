"""ctypes binding of libkanconv.so (C ABI: include/kanconv.h).  No torch types cross this boundary - only raw device
pointers, sizes and the CUDA stream handle.  There is no fallback: a missing library is a hard error."""
import ctypes
import os
import threading

from . import build as _build

c_i32, c_i64, c_f32, c_vp, c_sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

KC_OK, KC_ERR_INVALID, KC_ERR_UNSUPPORTED, KC_ERR_CUDA = 0, -1, -2, -3
BASIS_BSPLINE, BASIS_CHEBY, BASIS_GRAM, BASIS_RBF, BASIS_RECUR, BASIS_RECUR_DM = 0, 1, 2, 3, 4, 5
ACT_NONE, ACT_IDENTITY, ACT_GELU, ACT_SILU = -1, 0, 1, 2
NORM_NONE, NORM_INSTANCE, NORM_BATCH = 0, 1, 2
OUT_NONE, OUT_PRELU, OUT_SILU = 0, 1, 2
KC_MAX_BASIS, KC_MAX_PARAMS = 16, 40
KC_ABI_VERSION = 2


class KcDesc(ctypes.Structure):
    _fields_ = [(n, c_i32) for n in ("basis", "act", "n", "cin", "h", "w", "cout", "ho", "wo", "kh", "kw", "stride_h",
                                     "stride_w", "pad_h", "pad_w", "dil_h", "dil_w", "nb", "order", "nparams")] + \
               [("x_batch_stride", c_i64), ("z_batch_stride", c_i64), ("params", c_f32 * KC_MAX_PARAMS)]


class KcNormDesc(ctypes.Structure):
    _fields_ = [(n, c_i32) for n in ("norm", "out_act", "n", "c", "hw", "affine")] + \
               [("batch_stride", c_i64), ("eps", c_f32)]


class KcRowNormDesc(ctypes.Structure):
    _fields_ = [(n, c_i32) for n in ("rows", "features", "out_act", "affine")] + [("eps", c_f32)]


_P = ctypes.POINTER
_SIGNATURES = {
    "kc_version": (ctypes.c_int, []),
    "kc_last_error": (ctypes.c_char_p, []),
    "kc_device_info": (ctypes.c_int, [_P(ctypes.c_int)] * 3),
    "kc_launch_count": (ctypes.c_longlong, []),
    "kc_conv_fwd_f32": (ctypes.c_int, [_P(KcDesc)] + [c_vp] * 7),
    "kc_conv_dgrad_f32": (ctypes.c_int, [_P(KcDesc)] + [c_vp] * 10),
    "kc_wgrad_workspace_bytes": (c_sz, [_P(KcDesc)]),
    "kc_conv_wgrad_f32": (ctypes.c_int, [_P(KcDesc)] + [c_vp] * 8),
    "kc_dwconv_supported": (ctypes.c_int, [_P(KcDesc), ctypes.c_int]),
    "kc_dwconv_fwd_f32": (ctypes.c_int, [_P(KcDesc), ctypes.c_int] + [c_vp] * 6),
    "kc_dwconv_dgrad_f32": (ctypes.c_int, [_P(KcDesc), ctypes.c_int] + [c_vp] * 8),
    "kc_dwconv_wgrad_workspace_bytes": (c_sz, [_P(KcDesc), ctypes.c_int]),
    "kc_dwconv_wgrad_f32": (ctypes.c_int, [_P(KcDesc), ctypes.c_int] + [c_vp] * 7),
    "kc_norm_act_fwd": (ctypes.c_int, [_P(KcNormDesc)] + [c_vp] * 9),
    "kc_norm_act_bwd": (ctypes.c_int, [_P(KcNormDesc)] + [c_vp] * 12 + [ctypes.c_int, c_vp]),
    "kc_dbeta_floats": (c_sz, [_P(KcDesc), ctypes.c_int]),
    "kc_layernorm_act_fwd": (ctypes.c_int, [_P(KcRowNormDesc)] + [c_vp] * 8),
    "kc_layernorm_act_bwd": (ctypes.c_int, [_P(KcRowNormDesc)] + [c_vp] * 13),
    "kc_tc_supported": (ctypes.c_int, [_P(KcDesc)]),
    "kc_tc_fwd_needs_phi": (ctypes.c_int, [_P(KcDesc)]),
    "kc_tc_bytes": (c_sz, [_P(KcDesc), ctypes.c_int]),
    "kc_tc_pack_weights": (ctypes.c_int, [_P(KcDesc)] + [c_vp] * 5),
    "kc_tc_dz_flat": (ctypes.c_int, [_P(KcDesc)] + [c_vp] * 3),
    "kc_conv_fwd_tc": (ctypes.c_int, [_P(KcDesc)] + [c_vp] * 7),
    "kc_conv_dgrad_tc": (ctypes.c_int, [_P(KcDesc)] + [c_vp] * 10),
    "kc_conv_wgrad_tc": (ctypes.c_int, [_P(KcDesc)] + [c_vp] * 9),
    "kc_maxpool2d_fwd": (ctypes.c_int, [c_vp] * 3 + [ctypes.c_longlong] + [ctypes.c_int] * 6 + [c_vp]),
    "kc_maxpool2d_bwd": (ctypes.c_int, [c_vp] * 3 + [ctypes.c_longlong] + [ctypes.c_int] * 6 + [c_vp]),
    "kc_norm_bwd_dz_flat_supported": (ctypes.c_int, [_P(KcDesc), _P(KcNormDesc)]),
    "kc_norm_bwd_dz_flat": (ctypes.c_int, [_P(KcDesc), _P(KcNormDesc)] + [c_vp] * 9),
    "kc_tc_geometry": (ctypes.c_int, [_P(KcDesc), ctypes.c_int, _P(ctypes.c_longlong)]),
    "kc_tc_selftest": (ctypes.c_int, [ctypes.c_int, _P(ctypes.c_float), c_vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lock = threading.Lock()
_lib = None


def library_path():
    return _build.LIB


def source_hash():
    """sha256 of the CUDA sources / headers / flags of the tree (what libkanconv.so must have been built from)."""
    return _build.source_hash()


def load():
    """Load and return the ctypes handle; raises if unavailable.  The library is (re)built first when it is missing, when
    KANCONV_REBUILD=1, or when the sources differ from the ones it was built from (content hash, build._stale); a box without nvcc
    cannot rebuild and raises instead of running a stale library; an ABI version mismatch is a hard error either way."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            path = _build.LIB
            if not os.path.exists(path) or os.environ.get("KANCONV_REBUILD") == "1":
                path = _build.build(force=True)
            elif _build._stale():
                if not _build.have_nvcc():
                    raise RuntimeError("libkanconv.so was built from different sources than the ones in the tree and nvcc is "
                                       "not available to rebuild it (python -m kanconv_b200.build)")
                path = _build.build(force=False)
            lib = ctypes.CDLL(path)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)      # AttributeError if the library does not export the ABI
                fn.restype, fn.argtypes = res, args
            if lib.kc_version() != KC_ABI_VERSION:
                raise RuntimeError("libkanconv.so ABI version mismatch")
            _lib = lib
    return _lib


def check(rc, what):
    if rc == KC_OK:
        return
    msg = load().kc_last_error().decode(errors="replace")
    if rc == KC_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    if rc == KC_ERR_INVALID:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")
