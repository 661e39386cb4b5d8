"""Phase times of the persistent single-scan ICP kernel (block 0's %globaltimer stamps): LOCREG_PERSIST_STAMPS=1 is set here."""
import os, sys
os.environ["LOCREG_PERSIST_STAMPS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L
from loc_lib_b200 import synth
w = synth.World(200.0); m = w.sample_map(1_000_000); gt = w.poses(2)
scan = w.scan(gt[0]); init = synth.perturb_poses(gt)
reg = L.IcpRegistration(L.IcpOptions(method_=2, max_iteration_=10, eps_=0.0))
reg.SetInputTarget(m)
for i in range(3):
    print("--- ScanMatch", i, file=sys.stderr)
    reg.ScanMatch(scan, init[0], want_cloud=False)
print(reg.last_timing())
