"""tcgen05.mma rate of a CTA pair (cta_group::2, M = 256) vs one CTA (M = 128) for the no-swizzle layouts of this library
(debug micro-benchmark, needs a KANCONV_DEBUG=1 build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kanconv_b200 as K
lib = K._lib.load()
torch.zeros(1, device="cuda")
f1 = lib.kc_debug_mma_rate2
f1.argtypes = [ctypes.c_int] * 7 + [ctypes.POINTER(ctypes.c_float)]
f2 = lib.kc_debug_mma_rate_2cta
f2.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_float)]
def one(N, mn, nsub, iters=256):
    c = ctypes.c_float()
    assert f1(N, mn, iters, nsub, 0, 0, 0, ctypes.byref(c)) == 0, lib.kc_last_error()
    return round(c.value, 1)
def pair(N, mn, nsub, iters=256):
    out = (ctypes.c_float * 5)()
    assert f2(N, mn, iters, nsub, out) == 0, lib.kc_last_error()
    return round(out[0], 1), out[1], out[2], int(out[3]), int(out[4])
for mn in (0, 1):
    for N in (64, 128, 256):
        ns = min(4, 512 // N)
        p = pair(N, mn, ns)
        print(f"{'MN' if mn else 'K '}-major N={N}: one CTA M=128: {one(N, mn, ns)} cycles/MMA | CTA pair M=256: {p[0]} cycles/MMA "
              f"(accumulator check: leader {p[1]}, peer {p[2]}, expected {256 * 2 * 16}, {ns} accumulators; tmem bases {p[3]:#x} {p[4]:#x})", flush=True)
