"""ncu target: one batch ScanMatch of S scans (ICP pipeline)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L
from loc_lib_b200 import synth
method = int(os.environ.get("METHOD", "2")); S = int(os.environ.get("S", "64"))
w = synth.World(200.0); m = w.sample_map(1_000_000); gt = w.poses(S)
buf, counts = w.scan_batch(gt); init = synth.perturb_poses(gt)
clouds = np.concatenate([buf[i, :counts[i]] for i in range(S)]); offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
r = L.IcpRegistration(L.IcpOptions(method_=method, max_iteration_=10, eps_=0.0))
r.SetInputTarget(m)
for i in range(int(os.environ.get("REPS", "2"))):
    poses, res = r.ScanMatchBatch(clouds, offsets, init)
print(r.last_timing(), S / (r.last_timing()[0] * 1e-3), "scans/s")
