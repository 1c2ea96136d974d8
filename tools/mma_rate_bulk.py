"""tcgen05.mma rate (one CTA, M = 128) while cp.async.bulk copies stream into another shared-memory region of the same CTA and
producer warps store 16-byte vectors: does the weight loader / the basis producers slow the MMAs down? (debug micro-benchmark)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kanconv_b200 as K
lib = K._lib.load()
torch.zeros(1, device="cuda")
f = lib.kc_debug_mma_rate2_bulk
f.argtypes = [ctypes.c_int] * 8 + [ctypes.POINTER(ctypes.c_float)]
def run(N, nsub, writers, streams, iters=2048):
    out = (ctypes.c_float * 2)()
    assert f(N, 0, iters, nsub, 1, writers, 1, streams, out) == 0, lib.kc_last_error()
    return round(out[0], 1), round(out[1], 1)
for N, ns in ((256, 2), (128, 4), (64, 4)):
    for writers in (0, 16):
        print(f"N={N} {ns} accumulators, {writers} writer warps: " + " | ".join(f"{st} bulk streams: {run(N, ns, writers, st)[0]} cyc/MMA, {run(N, ns, writers, st)[1]} B/clk" for st in (0, 1, 2, 4)), flush=True)
