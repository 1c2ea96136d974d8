"""Per-shape timing of the normalisation kernels on the KAN-VGG16 planes (CUDA events, L2 flushed between iterations):
forward (cluster-resident kernel vs the generic two-pass kernel) and backward (fused norm backward + bf16 flat dz vs
norm backward + kc_dz_flat).  GB/s are ALGORITHMIC bytes: forward 8 B / element, fused backward 10 B, two-step 18 B."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K  # noqa: E402
from kanconv_b200 import _lib as L, functional as KF  # noqa: E402

lib = L.load()
SHAPES = [(64, 64, 224), (64, 128, 112), (64, 256, 56), (64, 512, 28), (64, 512, 14)]
if os.environ.get("NORM_BENCH_SHAPES"):        # e.g. "16,64,224;64,512,28"
    SHAPES = [tuple(int(v) for v in s_.split(",")) for s_ in os.environ["NORM_BENCH_SHAPES"].split(";")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


ITERS = int(os.environ.get('NORM_BENCH_ITERS', '5'))


def timeit(fn, iters=None):
    iters = iters or ITERS
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


for n, c, hw in SHAPES:
    z = torch.randn(n, c, hw, hw, device="cuda")
    dy = torch.randn_like(z)
    alpha = torch.tensor([0.25], device="cuda")
    spec = KF.NormSpec(L.NORM_INSTANCE, L.OUT_PRELU, 1, False, 1e-5)
    rec = {"shape": [n, c, hw, hw], "MB": round(z.numel() * 4 / 1e6, 1)}
    for tag, env in (("fwd_cluster", "1"), ("fwd_generic", "0")):
        os.environ["KANCONV_NORM_CLUSTER"] = env
        t = timeit(lambda: KF._norm_fwd(spec, z, None, None, [alpha]))
        rec[tag] = [round(t, 3), round(8.0 * z.numel() / t / 1e6)]
    os.environ["KANCONV_NORM_CLUSTER"] = "1"
    y, mean, rstd = KF._norm_fwd(spec, z, None, None, [alpha])
    # backward: conv descriptor of a 3x3 / pad 1 layer with c -> c channels gives the flat layout
    cs = KF.ConvSpec(basis=L.BASIS_BSPLINE, act=L.ACT_SILU, nb=8, order=3, params=tuple(float(-2.2 + 0.4 * i) for i in range(12)),
                     kernel=(3, 3), stride=(1, 1), padding=(1, 1), dilation=(1, 1), groups=1)
    d = KF._make_desc(cs, n, c, hw, hw, c, c * hw * hw, c * hw * hw)
    nd = KF._norm_desc(spec, n, c, hw * hw, c)
    dzf = torch.empty(lib.kc_tc_bytes(ctypes.byref(d), 2), device="cuda", dtype=torch.uint8)
    partials = torch.empty(3 * n * c + 2 * c, device="cuda")
    dalp = torch.empty(1, device="cuda")
    st = KF._stream(z.device)
    P = KF._ptr
    if lib.kc_norm_bwd_dz_flat_supported(ctypes.byref(d), ctypes.byref(nd)):
        t = timeit(lambda: L.check(lib.kc_norm_bwd_dz_flat(ctypes.byref(d), ctypes.byref(nd), P(dy), P(z), P(mean[0]), P(rstd[0]), P(alpha),
                                                          P(dzf), P(dalp), P(partials), st), "fused"))
        rec["bwd_fused"] = [round(t, 3), round(10.0 * z.numel() / t / 1e6)]
    dz = torch.empty_like(z)

    def two_step():
        L.check(lib.kc_norm_act_bwd(ctypes.byref(nd), P(dy), P(z), P(mean[0]), P(rstd[0]), None, None, P(alpha), P(dz), None, None,
                                    P(dalp), P(partials), 0, st), "bwd")
        L.check(lib.kc_tc_dz_flat(ctypes.byref(d), P(dz), P(dzf), st), "flat")
    t = timeit(two_step)
    rec["bwd_two_step"] = [round(t, 3), round(18.0 * z.numel() / t / 1e6)]
    print(json.dumps(rec), flush=True)
    del z, dy, dz, dzf, y
