"""Print a compact summary of an .ncu-rep (raw page + hottest source lines). Usage: ncu_summary.py report [n_lines]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"{w:70s} {r[i]:>18s} {units[i]}")
    st = []
    for h in hdr:
        if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
            v = r[hdr.index(h)]
            if v and float(v) > 0.1: st.append((float(v), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    print("stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; hdr = None; agg = {}; tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; ie = hdr.index('Instructions Executed'); ss = hdr.index('Warp Stall Sampling (All Samples)'); continue
    if hdr is None or len(r) <= ie or not r[0].isdigit(): continue
    try: v = int(r[ie]); sm = int(r[ss] or 0)
    except ValueError: continue
    key = (cur, int(r[0]), r[1].strip()[:95]); a = agg.setdefault(key, [0, 0]); a[0] += v; a[1] += sm; tot += v
tots = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions {tot}")
for (f, l, s), (v, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:nl]:
    print(f"stall {sm/tots*100:5.1f}%  inst {v/tot*100:5.1f}%  {f}:{l}  {s}")
