"""Per-kernel hottest source lines of an .ncu-rep by executed warp instructions. Usage: ncu_lines.py report [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
kern = None; cur = None; hdr = None; agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'Function Name': kern = r[1].split('(')[0][-40:]; agg.setdefault(kern, {}); hdr = None; continue
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r; ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed') if 'Thread Instructions Executed' in hdr else None
        ss = hdr.index('Warp Stall Sampling (All Samples)'); continue
    if hdr is None or kern is None or len(r) <= ie or not r[0].isdigit(): continue
    try: v = int(r[ie]); sm = int(r[ss] or 0); tv = int(r[te]) if te is not None and r[te] else 0
    except ValueError: continue
    a = agg[kern].setdefault((cur, int(r[0]), r[1].strip()[:90]), [0, 0, 0]); a[0] += v; a[1] += sm; a[2] += tv
for k, d in agg.items():
    tot = sum(a[0] for a in d.values()) or 1; tots = sum(a[1] for a in d.values()) or 1
    print(f"=== {k}: {tot} warp instructions")
    for (f, l, s), (v, sm, tv) in sorted(d.items(), key=lambda kv: -kv[1][0])[:nl]:
        print(f"inst {v/tot*100:5.1f}%  stall {sm/tots*100:5.1f}%  act {tv/max(v,1):4.1f}  {f}:{l}  {s}")
