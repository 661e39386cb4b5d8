"""Development probe (run under gpurun): timings of the main entry points on one B200 -> gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L  # noqa: E402
from loc_lib_b200 import synth  # noqa: E402

out = {}
w = synth.World(200.0)
t = time.time(); m = w.sample_map(1_000_000); out["gen_map_s"] = time.time() - t
S = int(os.environ.get("PROBE_SCANS", "296"))
gt = w.poses(S)
t = time.time(); buf, counts = w.scan_batch(gt); out["gen_scans_s"] = time.time() - t
init = synth.perturb_poses(gt)
scans = [buf[i, :counts[i]] for i in range(S)]
clouds = np.concatenate(scans)
offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
out["points_per_scan"] = float(np.mean(counts))

for name, method in (("p2plane", 2), ("p2p", 0)):
    for loop_mode in (0, 1):
        r = L.IcpRegistration(L.IcpOptions(method_=method, max_iteration_=10, eps_=0.0, loop_mode=loop_mode))
        t = time.time(); r.SetInputTarget(m); wall = time.time() - t
        out[f"{name}_set_target_ms"] = (r.last_timing()[0], wall * 1e3)
        ts, ws = [], []
        for i in range(12):
            t = time.time(); r.ScanMatch(scans[i % 4], init[i % 4]); ws.append((time.time() - t) * 1e3)
            ts.append(r.last_timing()[0])
        out[f"{name}_single_loop{loop_mode}_kernel_ms"] = sorted(ts[2:])[len(ts[2:]) // 2]
        out[f"{name}_single_loop{loop_mode}_wall_ms"] = sorted(ws[2:])[len(ws[2:]) // 2]
        out[f"{name}_single_loop{loop_mode}_launches"] = r.last_timing()[1]
        if loop_mode == 0:
            for rep in range(3):
                t = time.time(); poses, res = r.ScanMatchBatch(clouds, offsets, init); wall = time.time() - t
                out[f"{name}_batch{S}_kernel_ms"] = r.last_timing()[0]
                out[f"{name}_batch{S}_wall_ms"] = wall * 1e3
            out[f"{name}_batch_scans_per_s"] = S / (out[f"{name}_batch{S}_kernel_ms"] * 1e-3)
            out[f"{name}_batch_points_per_s"] = float(counts.sum()) / (out[f"{name}_batch{S}_kernel_ms"] * 1e-3)
            err = [np.linalg.norm(poses[i][4:] - gt[i][4:]) for i in range(S)]
            out[f"{name}_batch_median_trans_err_m"] = float(np.median(err))
        r.close()

nd = L.NdtRegistration(L.NdtOptions(max_iteration_=10, eps_=0.0))
t = time.time(); nd.SetInputTarget(m); out["ndt_set_target_ms"] = (nd.last_timing()[0], (time.time() - t) * 1e3)
ts = []
for i in range(8):
    nd.ScanMatch(scans[i % 4], init[i % 4]); ts.append(nd.last_timing()[0])
out["ndt_single_kernel_ms"] = sorted(ts[2:])[len(ts[2:]) // 2]
poses, res = nd.ScanMatchBatch(clouds, offsets, init)
out[f"ndt_batch{S}_kernel_ms"] = nd.last_timing()[0]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
