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
# summary of the persistent kernel: per tile the MMA warp stamps (tile start, accumulator set free, then a_full and b_full stamps)
mm = [int(v) for v in t[1] if v > 0]
ep = [int(v) for v in t[3] if v > 0]
nbc = ((cout + 15) // 16 * 16 // 8 + 3) // 4
for stages_per_chunk in (1, 3, 9):
    per = 2 + nbc * (1 + stages_per_chunk)
    if len(mm) >= 2 * per and (len(mm) % per == 0 or len(mm) == 1024):
        ntile = len(mm) // per
        waits = [mm[per * i + 1] - mm[per * i] for i in range(ntile)]
        tot = mm[per * (ntile - 1)] - mm[0]
        print("MMA warp 0: %d tiles traced, %d stamps per tile, %.0f cycles per tile, waiting for a drained accumulator set %.1f %% of the time"
              % (ntile, per, tot / max(1, ntile - 1), 100.0 * sum(waits[1:]) / max(1, tot)))
        print("  wait per tile:", waits[:12])
        break
# epilogue thread 0: (wait start, accumulator ready, tile written) per tile
if len(ep) >= 6:
    nt = len(ep) // 3
    w = [ep[3 * i + 1] - ep[3 * i] for i in range(nt)]
    e = [ep[3 * i + 2] - ep[3 * i + 1] for i in range(nt)]
    print("epilogue warp 0: %d tiles, waiting for the accumulators %.0f cycles per tile, draining a tile %.0f cycles" % (nt, sum(w[1:]) / max(1, nt - 1), sum(e) / nt))
