"""Timeline of one CTA of the dgrad kernel (debug).  Persistent kernel: per tile the MMA warp stamps [tile start, accumulator
free, then per chunk: a_full, b_full x stages]; the epilogue stamps [wait start, accumulator ready, tile written]."""
import argparse, ctypes, os, sys
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K
ap = argparse.ArgumentParser(); ap.add_argument("--shape", default="16,64,64,224"); a = ap.parse_args()
n, cin, cout, hw = [int(v) for v in a.shape.split(",")]
lib = K._lib.load()
m = K.KANConv2DLayer(cin, cout, 3, padding=1, base_activation=nn.SiLU).cuda()
x = torch.randn(n, cin, hw, hw, device="cuda", requires_grad=True)
y = m(x); y.backward(torch.ones_like(y)); torch.cuda.synchronize()
buf = torch.zeros(4 * 1024, dtype=torch.int64, device="cuda")
lib.kc_debug_trace.argtypes = [ctypes.c_void_p]
y = m(x)
lib.kc_debug_trace(ctypes.c_void_p(buf.data_ptr()))
y.backward(torch.ones_like(y)); torch.cuda.synchronize()
lib.kc_debug_trace(None)
t = buf.cpu().view(4, 1024)
t0 = int(t[t > 0].min())
for role, name in ((0, "producer t0"), (1, "mma w0"), (2, "loader"), (3, "epilogue t0")):
    ev = [int(v) - t0 for v in t[role] if v > 0]
    print(name, len(ev), ev[:60])
    if role in (1, 2) and len(ev) > 3:
        print("   gaps", [ev[i + 1] - ev[i] for i in range(min(len(ev) - 1, 70))])
