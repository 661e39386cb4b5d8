"""Per-kernel-class device time of one batch ScanMatch (search / fit / solve), for experiment builds:
LOCREG_SO=liblocreg_x.so python tools/icp_breakdown.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L
from loc_lib_b200 import synth
method = int(os.environ.get("METHOD", "2")); S = int(os.environ.get("S", "256")); cell = float(os.environ.get("CELL", "0.5"))
w = synth.World(200.0); m = w.sample_map(1_000_000); gt = w.poses(S)
buf, counts = w.scan_batch(gt); init = synth.perturb_poses(gt)
clouds = np.concatenate([buf[i, :counts[i]] for i in range(S)]); offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
r = L.IcpRegistration(L.IcpOptions(method_=method, max_iteration_=10, eps_=0.0, knn_cell_size=cell))
r.SetInputTarget(m)
for i in range(2):
    poses, res = r.ScanMatchBatch(clouds, offsets, init)
total_ms = r.last_timing()[0]
r.profile(True)
poses, res = r.ScanMatchBatch(clouds, offsets, init)
prof = r.profile(False)
err = np.median(np.linalg.norm(poses[:, 4:] - gt[:, 4:], axis=1))
print(os.environ.get("LOCREG_SO", "liblocreg.so"), f"cell={cell} S={S} pts={offsets[-1]} total {total_ms:.2f} ms -> {S/total_ms*1e3:.0f} scans/s, {offsets[-1]/total_ms/1e3:.1f} Mpts/s | per launch: " +
      ", ".join(f"{k} {v[0]/max(v[1],1):.3f} ms x{v[1]}" for k, v in prof.items()), f"| median t err {err:.4f} m")
