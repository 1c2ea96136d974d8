"""Raw tcgen05.mma rate vs commit cadence (debug micro-benchmark)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kanconv_b200 as K
lib = K._lib.load()
f = lib.kc_debug_mma_rate
f.argtypes = [ctypes.c_int] * 10 + [ctypes.POINTER(ctypes.c_float)]
torch.zeros(1, device="cuda")
def run(N, a_lbo, nsub=1, commit_every=0, writers=0, shift=0, iters=128):
    c = ctypes.c_float()
    rc = f(N, a_lbo, 128, N * 16, 128, iters, shift, nsub, commit_every, writers, ctypes.byref(c))
    assert rc == 0, lib.kc_last_error()
    return round(c.value, 1)
for N in (64, 128, 256):
    for nsub in (1, 2) if N == 256 else (1, 4):
        print(f"N={N} nsub={nsub}: " + "  ".join(f"c/{ce}={run(N, 6544, nsub, ce)}" for ce in (0, 1, 2, 4, 8, 16, 32, 64)))
