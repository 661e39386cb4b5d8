"""compute-sanitizer target for the long-queue path of stage 2 (spatially ordered queue: k_queue_bin_count / scan /
k_queue_bin_scatter / k_icp_nn_finish on the sorted queue; LOCREG_PYR_KERNEL=1: the block-pyramid build and k_icp_nn_pyr):
a relocalisation of 96 hypotheses, two thirds of them metres away, on a small scene, with LOCREG_SORT_MIN=1 so that the
small queue takes the path.  Checks the winner against the unsorted run of the same job.
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_reloc.py"""
import os, sys
os.environ.setdefault("LOCREG_SORT_MIN", "1")
os.environ.setdefault("LOCREG_SORT_FRAC", "0.0")
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L
from loc_lib_b200 import synth

w = synth.World(60.0)
m = w.sample_map(40_000)
gt = w.poses(2)[1]
scan = w.scan(gt, beams=8, azimuth=180, seed=synth.SEED_SCAN)
rng = np.random.default_rng(1)
hyp = np.stack([synth.perturb_pose(gt, synth.SEED_POSE + i) for i in range(96)])
hyp[32:, 4:6] += rng.uniform(-12, 12, (64, 2))
hyp[64:, 6] += rng.uniform(2, 9, 32)
for method, k in ((L.IcpMethod.P2PLANE, 5), (L.IcpMethod.P2P, 1)):
    r = L.IcpRegistration(L.IcpOptions(method_=method, max_iteration_=int(os.environ.get("SANITIZE_RELOC_ITERS", "4")), eps_=0.0))
    r.SetInputTarget(m)
    pose, idx, score, poses, results = r.Relocalise(scan, hyp, want_all=True)
    import hashlib
    print("method", int(method), "winner", idx, "score", score, "poses", hashlib.sha1(np.ascontiguousarray(poses).tobytes()).hexdigest()[:16],
          "launches", r.last_timing()[1])
print("sanitize_reloc ok (sort_min=%s pyr_kernel=%s)" % (os.environ["LOCREG_SORT_MIN"], os.environ.get("LOCREG_PYR_KERNEL", "0")))
