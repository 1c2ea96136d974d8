"""Refresh profiles/<tag>_ncu_summary.md from the raw outputs of tools/ncu_round.sh: runs tools/summarize_profiles.py <tag> and
replaces the generated tables (everything from "## launch list" up to "## Reading") and the source hash in the
"State of the code" paragraph; the hand-written reading below the tables is kept.
usage: python tools/splice_summary.py [tag] ["description of the code state"]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
desc = sys.argv[2] if len(sys.argv) > 2 else None
tables = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_profiles.py"), tag], check=True,
                        capture_output=True, text=True).stdout
path = os.path.join(ROOT, "profiles", tag + "_ncu_summary.md")
s = open(path).read()
a, b = s.index("## launch list"), s.index("## Reading")
s = s[:a] + tables.rstrip("\n") + "\n\n\n" + s[b:]
h = open(os.path.join(ROOT, "gpurun_out", tag, "source_hash.txt")).read().strip()
s = re.sub(r"sha256 `[0-9a-f]{64}`", f"sha256 `{h}`", s, count=1)
if desc:
    s = re.sub(r"State of the code: .*?, sha256", f"State of the code: {desc}, sha256", s, count=1, flags=re.S)
open(path, "w").write(s)
print("spliced", path, "hash", h[:12])
