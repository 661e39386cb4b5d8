"""profiles/roofline_traffic.json from an ncu launch list that carries dram__bytes_read.sum / dram__bytes_write.sum /
gpu__time_duration.sum per launch (tools/profile_remote.sh):  python tools/traffic_from_ncu.py launches.csv out.json
Groups by (kernel base name, grid size): mean DRAM bytes and mean duration per launch, number of launches."""
import csv, json, re, sys, collections
src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src, errors="replace")) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
H = rows[hdr]
ci = {k: H.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
gi = H.index("Grid Size") if "Grid Size" in H else None
per = collections.defaultdict(dict)
for r in rows[hdr + 1:]:
    if len(r) <= ci["Metric Value"]:
        continue
    try:
        v = float(r[ci["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ci["Metric Unit"]].lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0}.get(unit, 1.0)
    name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[ci["Kernel Name"]])).replace("void ", "").split("::")[-1].strip()
    grid = int(re.sub(r"[^0-9,]", "", r[gi]).split(",")[0]) if gi is not None and r[gi] else 0
    d = per[(r[ci["ID"]], name, grid)]
    d[r[ci["Metric Name"]]] = v * scale
groups = collections.defaultdict(lambda: [0, 0.0, 0.0])
for (_, name, grid), d in per.items():
    g = groups[(name, grid)]
    g[0] += 1
    g[1] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    g[2] += d.get("gpu__time_duration.sum", 0.0)
out = {"source": src, "launch_groups": [{"kernel": k, "grid": g, "launches": n, "mean_bytes": b / n, "mean_ms": t / n}
                                          for (k, g), (n, b, t) in sorted(groups.items(), key=lambda kv: -kv[1][2])]}
json.dump(out, open(dst, "w"), indent=1)
tot = sum(t for _, _, t in groups.values())
for rec in out["launch_groups"][:14]:
    print(f"{rec['kernel']:28s} grid {rec['grid']:7d} x{rec['launches']:5d}  {rec['mean_ms']:8.4f} ms  {rec['mean_bytes']/1e6:9.1f} MB/launch  {100*rec['mean_ms']*rec['launches']/tot:5.1f}% of time")
