"""SASS evidence of the shipped library (no GPU needed): per kernel, the counts of the mnemonics that prove a Blackwell-native
kernel (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, cluster barriers -> UCGABAR,
legacy tensor path -> HMMA; UTCHMMA.2CTA = tcgen05.mma.cta_group::2, listed separately from the single-CTA UTCHMMA).  python tools/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kanconv_b200 as K  # noqa: E402

lib = K._lib.library_path()
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MNEM = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "LDGSTS", "HMMA", "HGMMA", "MUFU"]
cur, counts, order = None, collections.defaultdict(collections.Counter), []
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)[:90]
        order.append(cur)
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        for k in MNEM:
            if op.startswith(k):
                counts[cur][k] += 1
                break
        counts[cur]["_total"] += 1
print(f"library : {os.path.relpath(lib, ROOT)}")
print(f"sources : sha256 {K._lib.source_hash()}")
print("arch    : " + ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))))
print()
print(f"{'kernel':92s} " + " ".join(f"{k:>8s}" for k in MNEM) + "   instrs")
tot = collections.Counter()
for k in order:
    c = counts[k]
    tot.update(c)
    print(f"{k:92s} " + " ".join(f"{c[m]:8d}" for m in MNEM) + f" {c['_total']:8d}")
print(f"{'TOTAL':92s} " + " ".join(f"{tot[m]:8d}" for m in MNEM) + f" {tot['_total']:8d}")
print("\nlegacy tensor-core instructions (HMMA / HGMMA):", tot["HMMA"] + tot["HGMMA"])
