"""What spatial order would buy the stage-1 search of a BATCH ("next" (0)): the stand-alone exact k-NN
kernels (locreg_knn: the unseeded stage 1 + stage 2, as iteration 0 of a batch runs them) on the 14.1 M transformed points
of 512 scans, (a) in scan order (consecutive rays of a ring), (b) the same queries sorted by the Morton code of their fine
cell, so that the lanes of a warp share a neighbourhood list.  Same queries, same results (checked), different order."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L
from loc_lib_b200 import synth
S = int(os.environ.get("S", "512"))
w = synth.World(200.0); m = w.sample_map(1_000_000)
gt = w.poses(S); buf, counts = w.scan_batch(gt); init = synth.perturb_poses(gt)
def quat_R(p):
    x, y, z, ww = p[:4]
    return np.array([[1-2*(y*y+z*z), 2*(x*y-z*ww), 2*(x*z+y*ww)], [2*(x*y+z*ww), 1-2*(x*x+z*z), 2*(y*z-x*ww)], [2*(x*z-y*ww), 2*(y*z+x*ww), 1-2*(x*x+y*y)]])
def part(v):  # spread the low 10 bits of v to every third bit
    v = v.astype(np.uint64) & 0x3FF
    v = (v | (v << 16)) & 0x30000FF; v = (v | (v << 8)) & 0x300F00F; v = (v | (v << 4)) & 0x30C30C3; v = (v | (v << 2)) & 0x9249249
    return v
r = L.IcpRegistration(L.IcpOptions(method_=2)); r.SetInputTarget(m)
for name, poses in (("initial poses (0.3 m / 2 deg off)", init), ("ground truth poses", gt)):
    q = np.concatenate([(buf[i, :counts[i], :3].astype(np.float64) @ quat_R(poses[i]).T + poses[i][4:]).astype(np.float32) for i in range(S)])
    q4 = np.zeros((len(q), 4), np.float32); q4[:, :3] = q
    c = np.floor(q / 0.5).astype(np.int64) + 512
    key = part(c[:, 0]) | (part(c[:, 1]) << 1) | (part(c[:, 2]) << 2)
    order = np.argsort(key, kind="stable")
    res = {}
    for tag, qq in (("scan order", q4), ("Morton order of the fine cell", q4[order])):
        for rep in range(3):
            idx = r.Knn(qq, 5); ms = r.last_timing()[0]
        res[tag] = idx
        print(f"{name}, {tag}: {len(qq)} queries in {ms:.3f} ms -> {len(qq)/ms/1e6:.2f} G queries/s", flush=True)
    assert np.array_equal(res["scan order"][order], res["Morton order of the fine cell"])
    u, cnt = np.unique(key, return_counts=True)
    print(f"   {len(u)} occupied fine cells, {len(q)/len(u):.1f} queries per cell on average, median {np.median(cnt):.0f}")
