"""BASELINE.md section 7 ("Results") from profiles/r2_bench.json (+ profiles/r2_scale.jsonl): python tools/results_md.py"""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
d = json.loads(open(os.path.join(P, "r2_bench.json")).read().strip().splitlines()[-1])
c = d["configs"]
cb = d.get("cpu_baseline", {})
L = []
L.append("## 7. Results (round 2, one B200 unless stated; `profiles/r2_bench.json`, driver copies in `BENCH_r02.json` / `SCALE_r02.json`)\n")
L.append("| config | metric | B200 (inputs resident) | B200 end to end (host buffers through the C ABI) | roofline fraction (HBM, measured peak %.0f GB/s) | CPU port of the reference loop, same run |" % d["roofline"]["peak"])
L.append("|---|---|---|---|---|---|")
q1 = cb.get("q1_ann_vs_exact", {})
L.append("| C1 / C2 / C4 P2Plane ICP, 512 scans x 27.6 k points vs 1 M-point map, 10 iterations | points/s | **%.0f M** (%.1f ms per step, %.0f scans/s) | **%.0f M** | %.3f (stage-1 search), %.3f (whole iteration); DRAM traffic %.2f GB per launch vs %.2f GB algorithmic | %.0f k points/s on %d core (%s) |" % (
    d["value"] / 1e6, d["ms_per_step"], d["scans_per_s"], d["e2e"]["value"] / 1e6, d["roofline"]["frac"], d["roofline"]["iteration_frac"],
    (d["roofline"]["traffic"] or 0) / 1e9, d["roofline"]["algorithmic_bytes_per_launch"] / 1e9, cb.get("value", 0) / 1e3, cb.get("cores", 1), "literal ANN kd-tree"))
t = d["track"]
L.append("| C2 proper: one scan (%d points), `ScanMatch` | ms | %.3f ms kernels (%d launches) | %.3f ms wall | %.4f | ~%.0f ms (1 core) |" % (
    t["scan_points"], t["kernel_ms"], t["launches_per_scan_match"], t["e2e_ms"], t["roofline_frac"], t["scan_points"] / max(cb.get("value", 1), 1) * 1e3))
s4 = c["C4_strong"]
L.append("| C4 as written: 4096 scans fixed, `locreg_align_batch_sharded` | points/s | %.0f M (%.1f ms) | %.0f M | %.3f | |" % (
    s4["value"] / 1e6, s4["ms_per_step"], s4["e2e"]["value"] / 1e6, s4["roofline"]["frac"]))
c3 = c["C3"]
L.append("| C3 direct NDT, %d-point scan vs %d M-point map (%d voxels) | points/s | %.0f M (%.3f ms per scan) | %.0f M | %.3f (%.0f B per point-iteration) | %.2f M points/s (oracle AlignNdt, 1 core) |" % (
    c3["config"]["scan_points"], round(c3["map_build"]["points"] / 1e6), c3["config"]["voxels"], c3["value"] / 1e6, c3["ms_per_step"], c3["e2e"]["value"] / 1e6,
    c3["roofline"]["frac"], c3["roofline"]["bytes_per_point_iteration"], c3.get("cpu_baseline", {}).get("value", 0) / 1e6))
c5 = c["C5"]
L.append("| C5 relocalisation, %d hypotheses x (10 iterations + score), `locreg_relocalise_sharded` | hypotheses/s | %.0f (%.1f s) | %.0f | %.4f | %.2f hypotheses/s (1 core) |" % (
    c5["config"]["hypotheses"], c5["value"], c5["ms_per_step"] / 1e3, c5["e2e"]["value"], c5["roofline"]["frac"], c5.get("cpu_baseline", {}).get("value", 0)))
mb = d["map_build"]
L.append("| ICP index (1 M points) | ms | first build %.1f ms, rebuild %.1f ms wall / %.1f ms device; %.0f B per point (%.0f B without lists: %.0f M points/s) | | | kd-tree build ~0.5 s |" % (
    mb["wall_ms"], mb["rebuild_wall_ms"], mb["rebuild_device_ms"], mb["index_bytes_per_point"], mb["knn_lists_0"]["index_bytes_per_point"], mb["knn_lists_0"]["points_per_s_64_scans"] / 1e6))
if "icp_index_20M" in c3:
    i20 = c3["icp_index_20M"]
    L.append("| ICP index (20 M points) | ms | rebuild %.0f ms wall / %.0f ms device; %.1f GB | | | |" % (i20["rebuild_wall_ms"], i20["rebuild_device_ms"], i20["index_bytes"] / 1e9))
if "lio_keyframe" in c:
    lk = c["lio_keyframe"]
    ks = [k for k in lk if k.startswith("capacity_")]
    L.append("| Lio key-frame step, incremental NDT (device LRU cache) | ms per key frame | " + "; ".join("%s: %.2f ms kernels, %.1f ms wall, %d voxels" % (k, lk[k]["cache_update_kernel_ms"], lk[k]["keyframe_wall_ms"], lk[k]["voxels_cached"]) for k in ks) + " | | | |")
L.append("")
L.append("Clocks during every timed region: %s MHz of %s, throttle reasons %s.  Q1 (reference's always-on ANN pruning vs the exact search this library implements): %.0f %% of the 5-NN rows differ (%.0f %% of the entries); the final poses differ by %.1e rad / %.1e m." % (
    d["clocks"]["sm_mhz"], d["clocks"]["sm_max_mhz"], d["clocks"]["reasons"] or "none", 100 * q1.get("nn_rows_differing", 0), 100 * q1.get("nn_entries_differing", 0),
    q1.get("final_pose_delta_rad_max", 0), q1.get("final_pose_delta_m_max", 0)))
sc = os.path.join(P, "r2_scale.jsonl")
if os.path.exists(sc):
    L.append("")
    L.append("Scaling (`profiles/r2_scale.jsonl`, builder run; the driver's own curve is `SCALE_r02.json`):\n")
    L.append("| GPUs | headline batch ICP, weak (points/s; end to end) | C4 strong, 4096 scans (points/s; end to end) | C5 relocalisation, strong (hypotheses/s; s per 65 536; collective) |")
    L.append("|---|---|---|---|")
    for l in open(sc):
        l = l.strip()
        if not l.startswith("{"):
            continue
        r = json.loads(l); rc = r.get("configs", {})
        L.append("| %d | %.2f G; %.2f G | %.2f G; %.2f G | %.0f; %.2f s; %s |" % (r["n_gpus"], r["value"] / 1e9, r["e2e"]["value"] / 1e9,
                 rc.get("C4_strong", {}).get("value", 0) / 1e9, rc.get("C4_strong", {}).get("e2e", {}).get("value", 0) / 1e9,
                 rc.get("C5", {}).get("value", 0), rc.get("C5", {}).get("ms_per_step", 0) / 1e3, rc.get("C5", {}).get("config", {}).get("collective", "")))
txt = "\n".join(L) + "\n"
bp = os.path.join(ROOT, "BASELINE.md")
s = open(bp).read()
s = s[:s.index("## 7. Results")] + txt
open(bp, "w").write(s)
print(txt)
