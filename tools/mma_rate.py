"""Raw tcgen05.mma execution rate, K-major vs MN-major no-swizzle operands (debug micro-benchmark)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kanconv_b200 as K
lib = K._lib.load()
f = lib.kc_debug_mma_rate2
f.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_float)]
torch.zeros(1, device="cuda")
def run(N, mn, iters=256):
    c = ctypes.c_float()
    assert f(N, mn, iters, ctypes.byref(c)) == 0, lib.kc_last_error()
    return round(c.value, 1)
for N in (64, 128, 144, 160, 256):
    print(f"N={N}: K-major {run(N, 0)} cycles/MMA | MN-major {run(N, 1)} cycles/MMA   (floor 128*N/256 = {N // 2})")
