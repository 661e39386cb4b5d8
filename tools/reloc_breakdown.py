"""Per-kernel-class device time of one relocalisation (HYP hypotheses of one scan)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import loc_lib_b200 as L
from loc_lib_b200 import synth
n_hyp = int(os.environ.get("HYP", "2048"))
w = synth.World(200.0); m = w.sample_map(1_000_000); gt = w.poses(3)[2]; scan = w.scan(gt)
hyp = bench.reloc_hypotheses(gt, 65536)
hyp = hyp[:: len(hyp) // n_hyp][:n_hyp]
r = L.IcpRegistration(L.IcpOptions(method_=2, max_iteration_=10, eps_=0.0)); r.SetInputTarget(m)
r.Relocalise(scan, hyp[:64])
r.profile(True)
pose, idx, score, _, _ = r.Relocalise(scan, hyp)
prof = r.profile(False)
print(f"hyp={n_hyp} pts/launch={n_hyp*len(scan)} total {r.last_timing()[0]:.1f} ms | per launch: " +
      ", ".join(f"{k} {v[0]/max(v[1],1):.3f} ms x{v[1]}" for k, v in prof.items()), "| best", idx, score, np.linalg.norm(pose[4:]-gt[4:]))
