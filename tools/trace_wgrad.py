"""Timeline trace of one CTA of kc_wgrad_tc_kernel (needs a KANCONV_DEBUG=1 build): where the CTA's time goes.

role 0 = producer thread 0 (stamps: before / after the wait for a free stage), role 1 = MMA-issuing warp 0 (before / after the
wait for a full stage), role 2 = CTA life cycle (entry, set-up done, producer loop done, accumulators complete, end)."""
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
pe = [int(v) for v in t[0] if v > 0]
me = [int(v) for v in t[1] if v > 0]
le = [int(v) for v in t[2] if v > 0]
print("shape", a.shape, "producer stamps", len(pe), "mma stamps", len(me), "life stamps", len(le))
if len(le) >= 5:
    t0 = le[0]
    print("  CTA: setup %d, producer loop %d, accumulators ready at %d, epilogue %d, total %d cycles" % (
        le[1] - t0, le[2] - le[1], le[3] - t0, le[4] - le[3], le[4] - t0))
nb = len(me) // 2
if nb:
    waits = [me[2 * b + 1] - me[2 * b] for b in range(nb)]
    period = (me[2 * (nb - 1)] - me[0]) / max(1, nb - 1)
    print("  MMA warp 0: %d blocks, first full stage after %d cycles (from CTA entry), period %.0f cycles per block, waiting for a full stage %.1f %% of the loop"
          % (nb, me[1] - (le[0] if le else me[0]), period, 100.0 * sum(waits[1:]) / max(1, me[-1] - me[1])))
    print("  wait per block, blocks 20-40:", waits[20:40])
npb = len(pe) // 2
if npb:
    waits = [pe[2 * b + 1] - pe[2 * b] for b in range(npb)]
    print("  producer thread 0: %d blocks, waiting for a free stage %.1f %% of the loop" % (npb, 100.0 * sum(waits) / max(1, pe[-1] - pe[0])))
    print("  wait per block, blocks 20-40:", waits[20:40])
    issue = [pe[2 * b + 2] - pe[2 * b + 1] for b in range(npb - 1)]
    print("  issue+publish per block, blocks 20-40:", issue[20:40])
