"""Summarise an ncu report's source page: top SASS instructions by stall samples.  usage: ncu_top.py rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
extra = sys.argv[3:]  # e.g. -k regex:wgrad
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# one "Address,..." header per kernel in the report; KERNEL_INDEX env picks which (default 0)
import os
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
which = int(os.environ.get("KERNEL_INDEX", "0"))
if os.environ.get("KERNEL_NAME"):      # first section whose "Kernel Name" line contains the substring
    want = os.environ["KERNEL_NAME"]
    which = next(k for k, i in enumerate(his) if i > 0 and rows[i - 1] and rows[i - 1][0] == "Kernel Name" and want in rows[i - 1][1])
hi = his[which]
print("kernel:", rows[hi - 1][1][:100] if hi > 0 else "?")
end = his[which + 1] if which + 1 < len(his) else len(rows)
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[col[s]]) for r in data) for s in stalls}
print("total samples", tot)
print("stall mix:", ", ".join(f"{k[6:]}={v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
data.sort(key=lambda r: -int(r[col["# Samples"]]))
for r in data[:topn]:
    s = int(r[col["# Samples"]])
    top = sorted(((int(r[col[k]]), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{s * 100.0 / tot:5.1f}%  {r[col['Source']].strip()[:70]:70s} exec={r[col['Instructions Executed']]:>9s} {top}")
