"""cp.async.bulk latency/throughput micro-benchmark (debug)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kanconv_b200 as K
lib = K._lib.load()
f = lib.kc_debug_bulk_bench
f.argtypes = [ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_float)]
span = 64 << 20
src = torch.zeros(span, dtype=torch.uint8, device="cuda")
def run(bytes_, depth, nctas, same, iters=50):
    c = ctypes.c_float()
    rc = f(ctypes.c_void_p(src.data_ptr()), span, bytes_, depth, iters, nctas, same, ctypes.byref(c))
    assert rc == 0
    return c.value
for nctas in (1, 148):
    for same in (1, 0):
        for bytes_ in (2048, 8192, 32768):
            for depth in (1, 4):
                cyc = run(bytes_, depth, nctas, same)
                print(f"ctas={nctas:3d} same_addr={same} bytes={bytes_:6d} depth={depth}: {cyc:8.0f} cycles/round  -> {bytes_*depth/cyc:6.1f} B/cycle/SM")
