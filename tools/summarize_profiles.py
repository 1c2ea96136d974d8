"""Turn the raw ncu outputs of tools/ncu_round.sh (gpurun_out/<round>/) into the committed summaries under profiles/:
<round>_ncu_launches_b16.csv (copied), <round>_traffic_b64.json (with the hash of the CUDA sources the capture was taken with; bench.py
refuses a capture of other kernels) and the tables of <round>_ncu_summary.md (printed to stdout as markdown).
usage: python tools/summarize_profiles.py [round tag, default r2]"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
OUT, PROF = os.path.join(ROOT, "gpurun_out", TAG), os.path.join(ROOT, "profiles")


def rows_of(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    return hdr, {h: i for i, h in enumerate(hdr)}, [r for r in rows[hi + 2:] if len(r) == len(hdr)]


def short(n):
    n = n.replace("void ", "").replace("<unnamed>::", "").split("(")[0]
    # template arguments -> readable names; CTA-pair (cta_group::2) instantiations keep a "<pair>" suffix, and the forward of the
    # stem / pointwise layers (pre-pass + persistent GEMM, FAM 2) is listed with the forward kernel it replaces
    import re
    m = re.match(r"(kc_tc_kernel|kc_dgrad_persistent_kernel|kc_wgrad_tc_kernel)<(\d+)(?:, (0|1|true|false))?>", n)
    if m:
        base, a, b = m.group(1), m.group(2), m.group(3) in ("1", "true")
        if base == "kc_tc_kernel":
            n = "kc_tc_kernel<" + ("fwd" if a == "0" else "dgrad") + ">"
        elif base == "kc_dgrad_persistent_kernel" and a == "2":
            n = "kc_tc_kernel<fwd><from phi>"
        else:
            n = f"{base}<{a}>"
        if b:
            n += "<pair>"
    return n[:80]


shutil.copy(os.path.join(OUT, "launches.csv"), os.path.join(PROF, TAG + "_ncu_launches_b16.csv"))
hdr, col, data = rows_of(os.path.join(PROF, TAG + "_ncu_launches_b16.csv"))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    v = float(r[col["Metric Value"]].replace(",", "")); u = r[col["Metric Unit"]]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    a = agg[short(r[col["Kernel Name"]])]; a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
print(f"## launch list: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} ms\n\n| share | launches | mean us | kernel |\n|---|---|---|---|")
mine = 0.0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    mine += v[1] if k.startswith("kc_") else 0.0
    if v[1] / tot > 0.004:
        print(f"| {100 * v[1] / tot:.1f} % | {v[0]} | {v[1] / v[0]:.1f} | `{k}` |")
print(f"\nkernels of this library: {100 * mine / tot:.1f} % of device time\n")

hdr, col, data = rows_of(os.path.join(OUT, "traffic_b64.csv"))
launch = collections.OrderedDict()
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in data:
    d = launch.setdefault(r[col["ID"]], {"name": short(r[col["Kernel Name"]])})
    d[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", "")) * scale[r[col["Metric Unit"]]]
L = list(launch.values())
nsteps = sum(1 for l in L if l["name"].startswith("kc_tc_kernel<fwd>")) / 13.0          # 13 conv layers per step
agg = collections.defaultdict(lambda: {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
for l in L:
    a = agg[l["name"]]; a["n"] += 1; a["ms"] += l.get("gpu__time_duration.sum", 0); a["rd"] += l.get("dram__bytes_read.sum", 0); a["wr"] += l.get("dram__bytes_write.sum", 0)
out = {}
print(f"## DRAM traffic per step at batch 64 ({nsteps:.0f} steps captured)\n\n| kernel | launches/step | ms/step under ncu | DRAM read GB/step | DRAM write GB/step |\n|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    out[n] = {"launches_per_step": round(a["n"] / nsteps, 3), "ms_per_step_under_ncu": round(a["ms"] / nsteps, 3),
              "dram_read_bytes_per_step": round(a["rd"] / nsteps), "dram_write_bytes_per_step": round(a["wr"] / nsteps),
              "dram_bytes_per_launch": round((a["rd"] + a["wr"]) / a["n"])}
    o = out[n]
    print(f"| `{n}` | {o['launches_per_step']:.0f} | {o['ms_per_step_under_ncu']:.2f} | {o['dram_read_bytes_per_step'] / 1e9:.2f} | {o['dram_write_bytes_per_step'] / 1e9:.2f} |")
import kanconv_b200 as K
srchash = open(os.path.join(OUT, "source_hash.txt")).read().strip() if os.path.exists(os.path.join(OUT, "source_hash.txt")) else ""
srchash = srchash or K._lib.source_hash()
json.dump({"source_hash": srchash, "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:kc_, "
           "python bench.py --steps 1 --warmup 3 --no-cpu-baseline (KAN-VGG16 @224, batch 64)", "steps_captured": nsteps, "kernels": out},
          open(os.path.join(PROF, TAG + "_traffic_b64.json"), "w"), indent=1)

for rep in sorted(p_ for p_ in os.listdir(OUT) if p_.endswith(".ncu-rep")):
    title = rep[:-8]
    rep = os.path.join(OUT, rep)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ["gpu__time_duration.sum", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]
    names = [short(r[hdr.index("Kernel Name")]) for r in rows[2:]]
    print(f"\n## ncu --set full, {title}\n\n| metric | " + " | ".join(names) + " |\n|---|" + "---|" * len(names))
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} [{rows[1][i]}] | " + " | ".join(r[i] for r in rows[2:]) + " |")
