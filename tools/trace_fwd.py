"""Timeline trace of one CTA of kc_tc_kernel<fwd> (needs a KANCONV_DEBUG=1 build): where the CTA's time goes.
python tools/trace_fwd.py --shape n,cin,cout,hw

role 0 = producer thread 0, role 1 = MMA-issuing warp 0 (per chunk: before / after the wait for the basis rows, then one stamp
per weight stage acquired), role 2 = weight loader, role 3 = CTA life cycle (entry, producer loop done, accumulators ready,
z written)."""
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
ev = [[int(v) for v in t[r] if v > 0] for r in range(4)]
le = ev[3]
print("shape", a.shape, "stamps", [len(e) for e in ev])
t0 = le[0]
if len(le) >= 4:
    print("  CTA: producer loop done at %d, accumulators ready at %d, epilogue %d, total %d cycles" % (le[1] - t0, le[2] - t0, le[3] - le[2], le[3] - t0))
mm = ev[1]
# chunk boundaries: the MMA warp stamps (before a_full wait, after) then one per weight stage; find the per-chunk count
nsc = (cin + 3) // 4; nbc = ((cin + 7) // 8 + 3) // 4; nch = nsc + nbc
per = len(mm) // nch if nch else 0
if per >= 3 and per * nch == len(mm):
    wa = sum(mm[per * q + 1] - mm[per * q] for q in range(1, nch))
    wb = sum(mm[per * q + 2 + k] - mm[per * q + 1 + k] for q in range(1, nch) for k in range(per - 2))
    span = mm[-1] - mm[per]
    print("  MMA warp 0: %d chunks x %d weight stages; first basis rows after %d cycles, first weights after %d; period %.0f cycles per chunk"
          % (nch, per - 2, mm[1] - t0, mm[2] - t0, span / max(1, nch - 1)))
    print("    of the loop: waiting for basis rows %.1f %%, acquiring weight stages (incl. issue of the previous stage) %.1f %%" % (100.0 * wa / span, 100.0 * wb / span))
    for q in range(nch // 2, min(nch, nch // 2 + 8)):
        nxt = mm[per * (q + 1)] if q + 1 < nch else mm[-1]
        print("    chunk %d: a_full wait %d, weight-stage gaps %s, last-stage issue %d" % (q, mm[per * q + 1] - mm[per * q],
              [mm[per * q + 2 + k] - mm[per * q + 1 + k] for k in range(per - 2)], nxt - mm[per * q + per - 1]))
else:
    print("  MMA stamps", len(mm), "chunks", nch)
pe = ev[0]
k0 = (len(pe) // 2) // 4 * 4
print("  producer thread 0: steady-state stamp gaps (chunk start, buffer free, stores done, arrived):", [pe[k + 1] - pe[k] for k in range(k0, min(k0 + 16, len(pe) - 1))])
print("  producer thread 0: last 16 stamp gaps (base-activation chunks):", [pe[k + 1] - pe[k] for k in range(max(0, len(pe) - 17), len(pe) - 1)])
print("  producer thread 0: first 12 stamp gaps", [pe[k + 1] - pe[k] for k in range(min(12, len(pe) - 1))], "last stamp at", pe[-1] - t0 if pe else None)
lo = ev[2]
if lo:
    print("  weight loader: first copy issued at %d, last at %d" % (lo[0] - t0, lo[-1] - t0))
