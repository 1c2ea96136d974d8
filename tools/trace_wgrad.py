"""Timeline trace of one CTA of kc_wgrad_tc_kernel (debug)."""
import argparse, ctypes, os, sys
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K
ap = argparse.ArgumentParser(); ap.add_argument("--shape", default="64,256,256,56"); a = ap.parse_args()
n, cin, cout, hw = [int(v) for v in a.shape.split(",")]
lib = K._lib.load()
m = K.KANConv2DLayer(cin, cout, 3, padding=1, base_activation=nn.SiLU).cuda()
x = torch.randn(n, cin, hw, hw, device="cuda", requires_grad=True)
def step():
    y = m(x); y.backward(torch.ones_like(y))
step(); torch.cuda.synchronize()
buf = torch.zeros(4 * 1024, dtype=torch.int64, device="cuda")
lib.kc_debug_trace_wgrad.argtypes = [ctypes.c_void_p]
lib.kc_debug_trace_wgrad(ctypes.c_void_p(buf.data_ptr()))
step(); torch.cuda.synchronize()
lib.kc_debug_trace_wgrad(None)
t = buf.cpu().view(4, 1024)
t0 = int(t[t > 0].min())
pe = [int(v) - t0 for v in t[0] if v > 0]
me = [int(v) - t0 for v in t[1] if v > 0]
print("producer events", len(pe), "mma events", len(me))
for b in range(20, 30):
    e = pe[4 * b:4 * b + 4]
    print(f"  blk {b}: start {e[0]:8d} fetch_issued +{e[1]-e[0]:5d} stage_free +{e[2]-e[1]:6d} stored +{e[3]-e[2]:5d}")
for b in range(20, 30):
    e = me[2 * b:2 * b + 3]
    print(f"  mma blk {b}: wait {e[0]:8d} got_full +{e[1]-e[0]:6d} issue+loop +{e[2]-e[1]:5d}")
