"""ncu target: a few single-scan ScanMatch calls in multi-launch mode (one k_eval + k_finalize per iteration)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L
from loc_lib_b200 import synth
method = int(os.environ.get("METHOD", "2"))
loop_mode = int(os.environ.get("LOOP", "1"))
w = synth.World(200.0); m = w.sample_map(1_000_000); gt = w.poses(2)
scans = [w.scan(g) for g in gt]; init = synth.perturb_poses(gt)
if method == 3:
    r = L.NdtRegistration(L.NdtOptions(max_iteration_=10, eps_=0.0, loop_mode=loop_mode))
else:
    r = L.IcpRegistration(L.IcpOptions(method_=method, max_iteration_=10, eps_=0.0, loop_mode=loop_mode))
r.SetInputTarget(m)
for i in range(int(os.environ.get("REPS", "3"))):
    r.ScanMatch(scans[i % 2], init[i % 2], want_cloud=False)
print(r.last_timing())
