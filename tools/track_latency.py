"""Single-scan tracking latency (config 2): kernel ms / wall ms / launches of ScanMatch, P2Plane and P2P and NDT."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L
from loc_lib_b200 import synth
w = synth.World(200.0); m = w.sample_map(1_000_000); gt = w.poses(4)
scans = [w.scan(g, seed=synth.SEED_SCAN + i) for i, g in enumerate(gt)]; init = synth.perturb_poses(gt)
for name, reg in (("p2plane", L.IcpRegistration(L.IcpOptions(method_=2, max_iteration_=10, eps_=0.0))),
                  ("p2p", L.IcpRegistration(L.IcpOptions(method_=0, max_iteration_=10, eps_=0.0))),
                  ("ndt", L.NdtRegistration(L.NdtOptions(max_iteration_=10, eps_=0.0)))):
    reg.SetInputTarget(m)
    k, wl = [], []
    for i in range(40):
        t0 = time.perf_counter(); reg.ScanMatch(scans[i % 4], init[i % 4], want_cloud=True); wl.append((time.perf_counter() - t0) * 1e3)
        k.append(reg.last_timing()[0])
    if name != "ndt" and os.environ.get("LOCREG_PROFILE_TRACE"):
        reg.profile(True); reg.ScanMatch(scans[0], init[0]); print(reg.profile(False))
    print(f"{name}: kernel {np.median(k[8:]):.3f} ms, wall {np.median(wl[8:]):.3f} ms, launches {reg.last_timing()[1]}, pts {len(scans[0])}")
