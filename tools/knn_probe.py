"""Throughput of the stand-alone warp-cooperative k-NN kernel (locreg_knn) on transformed scan points."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import loc_lib_b200 as L
from loc_lib_b200 import synth
w = synth.World(200.0); m = w.sample_map(1_000_000)
S = 128
gt = w.poses(S); buf, counts = w.scan_batch(gt); init = synth.perturb_poses(gt)
def quat_R(p):
    x, y, z, ww = p[:4]
    return np.array([[1-2*(y*y+z*z), 2*(x*y-z*ww), 2*(x*z+y*ww)], [2*(x*y+z*ww), 1-2*(x*x+z*z), 2*(y*z-x*ww)], [2*(x*z-y*ww), 2*(y*z+x*ww), 1-2*(x*x+y*y)]])
for name, poses in (("init", init), ("gt", gt)):
    q = np.concatenate([(buf[i, :counts[i], :3].astype(np.float64) @ quat_R(poses[i]).T + poses[i][4:]).astype(np.float32) for i in range(S)])
    q4 = np.zeros((len(q), 4), np.float32); q4[:, :3] = q
    for cell in (0.5, 0.35):
        r = L.IcpRegistration(L.IcpOptions(method_=2, knn_cell_size=cell)); r.SetInputTarget(m)
        for k in (5, 1):
            for rep in range(3):
                r.Knn(q4, k); ms = r.last_timing()[0]
            print(f"{name} cell={cell} k={k}: {len(q4)} queries in {ms:.3f} ms -> {len(q4)/ms/1e6:.2f} Gq/s")
        r.close()
