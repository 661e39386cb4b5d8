"""Runs HERE after tools/profile_remote.sh: turns gpurun_out/<tag>_* into the tracked summaries under profiles/."""
import csv, collections, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[2:]


for f in ("bench.json", "launches_bench.csv"):
    src = os.path.join(G, f"{tag}_{f}")
    if os.path.exists(src):
        open(os.path.join(P, f"{tag}_{f}"), "w").write(open(src).read())
# launch shares
rows = list(csv.reader(open(os.path.join(G, f"{tag}_launches_bench.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[hi + 2:]:
    if len(r) <= vi: continue
    a = agg.setdefault((r[ki].split("(")[0][-36:], r[gi]), [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
lines = [f"# {tag}: kernel time shares from the ncu launch list of `python bench.py --steps 2 --warmup 1` (cold-cache, serialised)", "```"]
for (n, g), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    lines.append(f"{n:38s} grid {g:>14s} x{c:4d} total {t/1e3:10.1f} us  avg {t/c/1e3:9.1f} us  {t/tot*100:5.1f}%")
lines.append("```")
open(os.path.join(P, f"{tag}_launch_shares.md"), "w").write("\n".join(lines) + "\n")
traffic = {}
for kern in ("nn", "post"):
    rep = os.path.join(G, f"{tag}_prof_{kern}.ncu-rep")
    if not os.path.exists(rep): continue
    s1 = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "3"], capture_output=True, text=True).stdout
    s2 = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "30"], capture_output=True, text=True).stdout
    s1 = s1[:s1.index("total warp instructions")] if "total warp instructions" in s1 else s1
    open(os.path.join(P, f"{tag}_ncu_{kern}.md"), "w").write(
        f"# {tag}: ncu --set full --clock-control none of k_icp_{kern} inside bench.py (512 scans, ~14.1 M points per launch)\n```\n{s1}\n{s2}```\n")
    h, rws = raw(rep)
    for r in rws:
        name = r[h.index("Kernel Name")]
        rd = float(r[h.index("dram__bytes_read.sum")]); wr = float(r[h.index("dram__bytes_write.sum")])
        u = rows and None
        ur = h.index("dram__bytes_read.sum")
        traffic.setdefault(name.split("(")[0], []).append(rd + wr)
# units row tells Mbyte/Gbyte; recompute in bytes from the units row
for kern in ("nn", "post"):
    rep = os.path.join(G, f"{tag}_prof_{kern}.ncu-rep")
    if not os.path.exists(rep): continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rws = list(csv.reader(out.splitlines())); h, units = rws[0], rws[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rws[2:]:
        tot_b = sum(float(r[h.index(m)]) * scale[units[h.index(m)]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic[f"{r[h.index('Kernel Name')].split('(')[0]}@{r[h.index('launch__grid_size')]}"] = tot_b
json.dump({k: v for k, v in traffic.items() if "@" in k}, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_launch_shares.md")).read())
print({k: v for k, v in traffic.items() if "@" in k})
