"""tcgen05.mma execution rate under conv-kernel-like conditions (debug micro-benchmark)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kanconv_b200 as K
lib = K._lib.load()
f = lib.kc_debug_mma_rate2
f.argtypes = [ctypes.c_int] * 7 + [ctypes.POINTER(ctypes.c_float)]
torch.zeros(1, device="cuda")
def run(N, mn=0, nsub=1, commit=0, writers=0, shift=0, iters=256):
    c = ctypes.c_float()
    assert f(N, mn, iters, nsub, commit, writers, shift, ctypes.byref(c)) == 0, lib.kc_last_error()
    return round(c.value, 1)
for N, ns in ((64, 4), (128, 4), (144, 3), (256, 2)):
    print(f"N={N} (floor {N//2}): 1 acc {run(N)} | {ns} acc {run(N, 0, ns)} | +commit/step {run(N, 0, ns, 1)} | +shift 1 row {run(N, 0, ns, 1, 0, 1)}"
          f" | +4 writer warps {run(N, 0, ns, 1, 4, 1)} | +16 writer warps {run(N, 0, ns, 1, 16, 1)} | MN-major {run(N, 1, min(ns, 3), 1)}")
g = lib.kc_debug_mma_rate3
g.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_float)]
def run3(N, mn, reuse, writers=0, iters=256):
    c = ctypes.c_float()
    assert g(N, mn, iters, reuse, writers, ctypes.byref(c)) == 0, lib.kc_last_error()
    return round(c.value, 1)
print("A shared by three MMAs (wgrad pattern), cycles per MMA: plain | collector fill/use/lastuse | same with 16 writer warps")
for N in (64, 128, 160):
    for mn in (0, 1):
        print(f"  N={N} {'MN' if mn else 'K '}-major: {run3(N, mn, 0)} | {run3(N, mn, 1)} | {run3(N, mn, 0, 16)} -> {run3(N, mn, 1, 16)}")
