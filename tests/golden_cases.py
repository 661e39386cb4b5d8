"""What the golden fixture tests/golden/registration_small.npz holds and how it is computed (by the oracle).

The reference ships no golden vectors and cannot be built here, so these are REGRESSION pins of the
oracle (itself pinned by the KATs and the numpy cross-checks of tests/test_oracle.py), frozen so that a later change
to the oracle or the generator cannot silently move the parity target.  tests/golden/make_golden.py writes the file;
tests/test_oracle.py checks the oracle against it on CPU; tests/test_gpu_parity.py checks the CUDA path against it.
"""
import numpy as np

import oracle_py as O

ITERS = 5
EXACT_KEYS = ["knn1", "knn5", "p2plane_gate", "p2p_gate", "ndt_keys", "ndt_npts", "ndt_mu", "ndt_hits",
              "p2plane_counts", "p2p_counts", "ndt_counts"]
CLOSE_KEYS = ["p2plane_H", "p2plane_B", "p2p_H", "p2p_B", "ndt_H", "ndt_B", "ndt_info", "p2plane_trace", "p2p_trace",
              "ndt_trace"]


def queries(scan, pose):
    R = O.pose_matrix(pose)
    return (scan[:, :3].astype(np.float64) @ R.T + pose[4:]).astype(np.float32)


def compute(map_cloud, scan, init):
    out = {}
    q = queries(scan, init)
    for name, method in (("p2plane", O.P2PLANE), ("p2p", O.P2P)):
        ref = O.OracleIcp(method=method, max_iteration=ITERS, eps=0.0, nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1,
                          max_plane_distance=0.05, max_nn_distance=0.3)
        ref.set_target(map_cloud)
        if method == O.P2PLANE:
            out["knn1"] = ref.knn(q, 1)
            out["knn5"] = ref.knn(q, 5)
        _, H, B, res, gate, _ = ref.compute_hb(scan, init)
        out[name + "_H"], out[name + "_B"], out[name + "_gate"] = H, B, gate
        out[name + "_counts"] = np.array([res["n_effective"], res["n_inlier"]], np.int64)
        _, _, _, trace = ref.align(scan, init, want_cloud=False)
        out[name + "_trace"] = trace
    ndt = O.OracleNdt(max_iteration=ITERS, eps=0.0, skip_nonfinite=1)
    ndt.set_target(map_cloud)
    out["ndt_keys"], out["ndt_mu"], out["ndt_info"], out["ndt_npts"] = ndt.voxels()
    H, B, res, hits = ndt.compute_hb(scan, init)
    out["ndt_H"], out["ndt_B"], out["ndt_hits"] = H, B, hits
    out["ndt_counts"] = np.array([res["n_effective"], res["n_inlier"]], np.int64)
    _, _, _, trace = ndt.align(scan, init, want_cloud=False)
    out["ndt_trace"] = trace
    return out
