"""Timeline trace of one CTA of kc_fwd_tc_kernel (debug).  python tools/trace_fwd.py --shape n,cin,cout,hw"""
import argparse, ctypes, os, sys
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K
ap = argparse.ArgumentParser(); ap.add_argument("--shape", default="16,64,64,224"); ap.add_argument("--k", type=int, default=3); a = ap.parse_args()
n, cin, cout, hw = [int(v) for v in a.shape.split(",")]
lib = K._lib.load()
m = K.KANConv2DLayer(cin, cout, a.k, padding=a.k // 2, base_activation=nn.SiLU).cuda()
x = torch.randn(n, cin, hw, hw, device="cuda")
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    buf = torch.zeros(4 * 1024, dtype=torch.int64, device="cuda")
    lib.kc_debug_trace.argtypes = [ctypes.c_void_p]
    lib.kc_debug_trace(ctypes.c_void_p(buf.data_ptr()))
    m(x); torch.cuda.synchronize()
    lib.kc_debug_trace(None)
t = buf.cpu().view(4, 1024)
t0 = int(t[t > 0].min())
nsc = (cin + 7) // 8; nch = nsc + (cin // 8 + 7) // 8
names = ["producer tp=0", "mma", "loader", "producer tp=300"]
for role in (0,):
    ev = [int(v) - t0 for v in t[role] if v > 0]
    print(names[role], "events", len(ev))
    for q in range(min(nch, 6)):
        e = ev[4 * q:4 * q + 4]
        if len(e) == 4: print(f"  chunk {q}: start {e[0]:7d} buf_free +{e[1]-e[0]:6d} stores_done +{e[2]-e[1]:6d} arrive +{e[3]-e[2]:5d}")
ev = [int(v) - t0 for v in t[1] if v > 0]
print("mma events", len(ev))
i = 0
for q in range(min(nch, 6)):
    e = ev[i:i + 11]; i += 11
    if len(e) == 11: print(f"  chunk {q}: wait_afull {e[0]:7d} got +{e[1]-e[0]:6d} | b_full gaps " + " ".join(str(e[k+1]-e[k]) for k in range(1, 10)))
ev = [int(v) - t0 for v in t[2] if v > 0]
print("loader events", len(ev), "first 30 gaps:", " ".join(str(ev[k+1]-ev[k]) for k in range(min(30, len(ev)-1))))
print("total span", max(int(v) for v in t.flatten()) - t0)
# absolute timeline, steady state: loader b_empty-acquired (copy issue) vs MMA b_full-acquired, per tap
lo = [int(v) - t0 for v in t[2] if v > 0]
mm = [int(v) - t0 for v in t[1] if v > 0]
mma_tap = []
for q in range(len(mm) // 11):
    mma_tap += mm[11 * q + 2: 11 * q + 11]
print("tap: copy_issued  b_full_seen  (latency)   next_copy_issue - this b_full")
for k in range(36, min(54, len(lo), len(mma_tap))):
    print(f"  {k:3d}: {lo[k]:8d} {mma_tap[k]:8d}  ({mma_tap[k]-lo[k]:6d})")
print("mma first/last events", mm[:3], mm[-3:], "producer last", [int(v) - t0 for v in t[0] if v > 0][-2:])
print("epilogue (tid 0): producer done, acc ready, z written:", [int(v) - t0 for v in t[3] if v > 0])
