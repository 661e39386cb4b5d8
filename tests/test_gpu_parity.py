"""-m gpu parity tests: the CUDA path through the C ABI against the CPU oracle on identical seeded inputs.

Gates (BASELINE.json north_star): NN indices bit-exact under the (float32 dis2, index) order; gate masks identical;
H and B within 1e-6 relative; final poses within 1e-5 rad / 1e-4 m.
"""
import numpy as np
import pytest

import oracle_py as O
from conftest import pose_delta

pytestmark = pytest.mark.gpu

H_TOL = 1e-6
ROT_TOL, TRANS_TOL = 1e-5, 1e-4


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def icp_pair(scene):
    import loc_lib_b200 as L
    gpu = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=10, eps_=0.0))
    gpu.SetInputTarget(scene.map)
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=10, eps=0.0, nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
    ref.set_target(scene.map)
    return gpu, ref


def test_knn_bit_exact(scene, icp_pair):
    gpu, ref = icp_pair
    R = O.pose_matrix(scene.init[0])
    q = (scene.scan[:, :3].astype(np.float64) @ R.T + scene.init[0][4:]).astype(np.float32)
    rng = np.random.default_rng(7)
    far = rng.uniform(-60, 60, (4000, 3)).astype(np.float32)
    far[:, 2] = rng.uniform(-4, 25, 4000)
    for queries in (q, far):
        for k in (1, 5):
            assert np.array_equal(gpu.Knn(queries, k), ref.knn(queries, k))


def test_knn_far_and_out_of_bounds_both_job_sizes(scene, icp_pair):
    """Far-off queries (above the roofs, outside the map, hundreds of metres away) through BOTH stage-2 forms: a small
    probe runs the warp-per-query kernel (what a single scan uses), a large one the thread-per-query kernel (batches)."""
    gpu, ref = icp_pair
    rng = np.random.default_rng(21)
    above = rng.uniform(-45, 45, (1500, 3)).astype(np.float32)
    above[:, 2] = rng.uniform(25, 45, 1500)
    outside = rng.uniform(-120, 120, (1500, 3)).astype(np.float32)
    way_out = rng.uniform(-900, 900, (24, 3)).astype(np.float32)
    near = scene.map[rng.integers(0, len(scene.map), 3000), :3] + rng.normal(0, 0.5, (3000, 3)).astype(np.float32)
    small = np.concatenate([above, outside, way_out, near]).astype(np.float32)
    for k in (1, 5):
        assert np.array_equal(gpu.Knn(small, k), ref.knn(small, k))
    # > 2 * SMs * 256 queries switch the probe (like a batch job) to the thread-per-query kernel
    more = scene.map[rng.integers(0, len(scene.map), 80_000), :3] + rng.normal(0, 0.4, (80_000, 3)).astype(np.float32)
    big = np.concatenate([small, more.astype(np.float32)])
    for k in (1, 5):
        assert np.array_equal(gpu.Knn(big, k), ref.knn(big, k))


def test_knn_matches_brute_force(scene, icp_pair):
    gpu, _ = icp_pair
    rng = np.random.default_rng(11)
    q = scene.map[rng.integers(0, len(scene.map), 300), :3] + rng.normal(0, 0.3, (300, 3)).astype(np.float32)
    assert np.array_equal(gpu.Knn(q.astype(np.float32), 5), O.bfnn(scene.map, q.astype(np.float32), 5))


@pytest.mark.parametrize("method", ["P2PLANE", "P2P", "P2LINE"])
def test_hb_and_gates(scene, method):
    import loc_lib_b200 as L
    m = getattr(L.IcpMethod, method)
    # tight thresholds so that every gate has points on both sides
    gpu = L.IcpRegistration(L.IcpOptions(method_=m, max_plane_distance_=0.004, max_nn_distance_=0.08, max_line_distance_=0.05))
    gpu.SetInputTarget(scene.map)
    ref = O.OracleIcp(method=getattr(O, method), max_plane_distance=0.004, max_nn_distance=0.08, max_line_distance=0.05,
                      nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
    ref.set_target(scene.map)
    for pose in (scene.init[0], scene.gt[0]):
        ok, H, B = gpu.CaculateMatrixHAndB(scene.scan, pose)
        rok, rH, rB, rres, rgate, rnn = ref.compute_hb(scene.scan, pose)
        k = 1 if method == "P2P" else 5
        gate, nn = gpu.DebugPoints(scene.scan, pose, k)
        assert np.array_equal(nn, rnn)
        assert np.array_equal(gate, rgate)
        assert len(set(rgate.tolist())) >= 2
        assert rel(H, rH) < H_TOL and rel(B, rB) < H_TOL
        assert bool(ok) == bool(rok)
        assert gpu.last_result["n_effective"] == rres["n_effective"]
        assert gpu.last_result["n_inlier"] == rres["n_inlier"]


@pytest.mark.parametrize("loop_mode", [0, 1])
@pytest.mark.parametrize("method", ["P2PLANE", "P2P", "P2LINE"])
def test_scan_match_pose(scene, method, loop_mode):
    import loc_lib_b200 as L
    gpu = L.IcpRegistration(L.IcpOptions(method_=getattr(L.IcpMethod, method), max_iteration_=10, eps_=0.0,
                                         loop_mode=loop_mode))
    gpu.SetInputTarget(scene.map)
    ref = O.OracleIcp(method=getattr(O, method), max_iteration=10, eps=0.0, nn_mode=O.NN_EXACT_TIEBREAK,
                      skip_nonfinite=1)
    ref.set_target(scene.map)
    for i in range(2):
        ok, cloud, pose = gpu.ScanMatch(scene.scans[i], scene.init[i])
        rpose, rcloud, rres, _ = ref.align(scene.scans[i], scene.init[i])
        dr, dt = pose_delta(pose, rpose)
        assert ok and dr < ROT_TOL and dt < TRANS_TOL
        assert gpu.last_result["iters"] == rres["iters"] and gpu.last_result["updates"] == rres["updates"]
        assert gpu.last_result["n_inlier"] == rres["n_inlier"]
        # transformed cloud: float32 transform of the (slightly different) double pose
        assert np.abs(cloud[:, :3] - rcloud[:, :3]).max() < 2e-4
        assert np.array_equal(cloud[:, 3], scene.scans[i][:, 3])


def test_convergence_break_and_default_options(scene):
    import loc_lib_b200 as L
    gpu = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE))  # eps 1e-2, 20 iterations
    gpu.SetInputTarget(scene.map)
    ref = O.OracleIcp(method=O.P2PLANE, nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
    ref.set_target(scene.map)
    _, _, pose = gpu.ScanMatch(scene.scan, scene.init[0], want_cloud=False)
    rpose, _, rres, _ = ref.align(scene.scan, scene.init[0], want_cloud=False)
    assert gpu.last_result["converged"] == rres["converged"] == 1
    assert gpu.last_result["iters"] == rres["iters"]
    dr, dt = pose_delta(pose, rpose)
    assert dr < ROT_TOL and dt < TRANS_TOL


def test_transform_cloud_bit_exact(scene, icp_pair):
    gpu, _ = icp_pair
    pts = np.zeros((len(scene.scan), 8), np.float32)  # pcl::PointXYZI layout: 32-byte stride
    pts[:, :3] = scene.scan[:, :3]
    pts[:, 4] = scene.scan[:, 3]
    pts[5, 0] = np.nan
    out = gpu.TransformCloud(pts, scene.gt[0])
    ref = O.transform_cloud(pts, scene.gt[0])
    assert np.array_equal(out, ref, equal_nan=True)


def test_edge_cases(scene):
    import loc_lib_b200 as L
    gpu = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=3))
    gpu.SetInputTarget(scene.map)
    # empty scan: every evaluation fails (effective_num < min_effective_pts), pose stays the prediction
    ok, cloud, pose = gpu.ScanMatch(np.zeros((0, 4), np.float32), scene.init[0])
    assert ok and np.allclose(pose, scene.init[0]) and gpu.last_result["degenerate"] == 1
    assert gpu.last_result["iters"] == 3 and gpu.last_result["updates"] == 0
    # scan with non-finite points: skipped (deviation D1), rest unchanged
    s = scene.scan.copy()
    s[::7, 1] = np.nan
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=3, nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
    ref.set_target(scene.map)
    _, _, pose = gpu.ScanMatch(s, scene.init[0], want_cloud=False)
    rpose, _, _, _ = ref.align(s, scene.init[0], want_cloud=False)
    dr, dt = pose_delta(pose, rpose)
    assert dr < ROT_TOL and dt < TRANS_TOL
    # duplicates in the map (quirk Q3) and a tiny map (< 5 leaves)
    dup = np.concatenate([scene.map[:5000], scene.map[:2500]])
    g2 = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE))
    g2.SetInputTarget(dup)
    r2 = O.OracleIcp(method=O.P2PLANE, nn_mode=O.NN_EXACT_TIEBREAK)
    r2.set_target(dup)
    q = dup[::50, :3] + np.float32(0.01)
    assert np.array_equal(g2.Knn(q, 5), r2.knn(q, 5))
    g2.SetInputTarget(scene.map[:3])
    assert np.all(g2.Knn(q, 5)[:, 3:] == -1)
    ok, H, B = g2.CaculateMatrixHAndB(scene.scan, scene.init[0])
    assert not ok and np.all(H == 0)


def test_batch_matches_single(scene):
    import loc_lib_b200 as L
    gpu = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=10, eps_=0.0))
    gpu.SetInputTarget(scene.map)
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=10, eps=0.0, nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
    ref.set_target(scene.map)
    clouds = np.concatenate(scene.scans)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in scene.scans])]).astype(np.int64)
    poses, results = gpu.ScanMatchBatch(clouds, offsets, scene.init)
    for i in range(len(scene.scans)):
        rpose, _, rres, _ = ref.align(scene.scans[i], scene.init[i], want_cloud=False)
        dr, dt = pose_delta(poses[i], rpose)
        assert dr < ROT_TOL and dt < TRANS_TOL
        assert results[i]["iters"] == rres["iters"] and results[i]["n_inlier"] == rres["n_inlier"]


def test_relocalise(scene):
    import loc_lib_b200 as L
    from loc_lib_b200 import synth
    gpu = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=6, eps_=0.0))
    gpu.SetInputTarget(scene.map)
    hyp = np.stack([synth.perturb_pose(scene.gt[0], 1000 + i, 1.5, 6.0) for i in range(48)])
    hyp[17] = scene.init[0]
    best_pose, best_idx, best_score, scores, poses = gpu.Relocalise(scene.scan, hyp, want_all=True)
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=6, eps=0.0, nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
    ref.set_target(scene.map)
    rscores = []
    for i in (3, 17, 30):
        rpose, _, _, _ = ref.align(scene.scan, hyp[i], want_cloud=False)
        _, _, _, rres, _, _ = ref.compute_hb(scene.scan, rpose, False, False)
        sc = rres["sum_sq_res"] / rres["n_inlier"] if rres["n_inlier"] else np.inf
        rscores.append(sc)
        dr, dt = pose_delta(poses[i], rpose)
        assert dr < 1e-4 and dt < 1e-3  # far hypotheses are ill-conditioned; the contract applies to the winner
        assert abs(scores[i] - sc) <= 1e-5 * sc
    assert best_idx == int(np.argmin(np.float32(scores)))
    assert np.isclose(best_score, scores[best_idx])
    dr, dt = pose_delta(best_pose, scene.gt[0])
    assert dr < 5e-3 and dt < 0.05


# ---------------------------------------------------------------------------------------------- NDT
@pytest.fixture(scope="module")
def ndt_pair(scene):
    import loc_lib_b200 as L
    gpu = L.NdtRegistration(L.NdtOptions(max_iteration_=10, eps_=0.0))
    gpu.SetInputTarget(scene.map)
    ref = O.OracleNdt(max_iteration=10, eps=0.0, skip_nonfinite=1)
    ref.set_target(scene.map)
    return gpu, ref


def test_ndt_voxels(ndt_pair):
    gpu, ref = ndt_pair
    k, mu, info, npts = gpu.Voxels()
    rk, rmu, rinfo, rn = ref.voxels()
    assert np.array_equal(k, rk) and np.array_equal(npts, rn)
    assert np.array_equal(mu, rmu)  # same summation order, no FMA: bit-exact
    assert np.abs(info - rinfo).max() <= 1e-9 * np.abs(rinfo).max()


@pytest.mark.parametrize("nearby", [0, 1])
def test_ndt_hb_and_pose(scene, nearby):
    import loc_lib_b200 as L
    gpu = L.NdtRegistration(L.NdtOptions(max_iteration_=10, eps_=0.0, nearby_type_=nearby))
    gpu.SetInputTarget(scene.map)
    ref = O.OracleNdt(max_iteration=10, eps=0.0, nearby6=nearby, skip_nonfinite=1)
    ref.set_target(scene.map)
    ok, H, B = gpu.CaculateMatrixHAndB(scene.scan, scene.init[0])
    rH, rB, rres, rhits = ref.compute_hb(scene.scan, scene.init[0])
    hits, _ = gpu.DebugPoints(scene.scan, scene.init[0], 0)
    assert np.array_equal(hits, rhits)
    assert rel(H, rH) < H_TOL and rel(B, rB) < H_TOL
    assert gpu.last_result["n_inlier"] == rres["n_inlier"]
    for loop_mode in (0, 1):
        g = L.NdtRegistration(L.NdtOptions(max_iteration_=10, eps_=0.0, nearby_type_=nearby, loop_mode=loop_mode))
        g.SetInputTarget(scene.map)
        _, cloud, pose = g.ScanMatch(scene.scan, scene.init[0])
        rpose, rcloud, rr, _ = ref.align(scene.scan, scene.init[0])
        dr, dt = pose_delta(pose, rpose)
        assert dr < ROT_TOL and dt < TRANS_TOL
        assert g.last_result["iters"] == rr["iters"]


def test_inc_ndt_cache_hb_and_pose(scene):
    """Incremental NDT: clouds added one by one into a small LRU cache (evictions), voxel dump, weighted H / B, poses."""
    import loc_lib_b200 as L
    clouds = [scene.map[:40_000], scene.map[30_000:80_000], scene.map[70_000:]]
    for nearby in (0, 1):
        gpu = L.NdtRegistration(L.NdtOptions(method_=L.NdtMethod.INCREMENTAL_NDT, capacity_=2500, max_iteration_=8, eps_=0.0,
                                             nearby_type_=nearby))
        ref = O.OracleIncNdt(capacity=2500, max_iteration=8, eps=0.0, nearby6=nearby, skip_nonfinite=1)
        for c in clouds:
            gpu.SetInputTarget(c)
            ref.set_target(c)
            k, mu, info, npts = gpu.Voxels()
            rk, rmu, rinfo, rn = ref.voxels()
            assert len(rk) <= 2499
            assert np.array_equal(k, rk) and np.array_equal(npts, rn) and np.array_equal(mu, rmu)
            assert np.abs(info - rinfo).max() <= 1e-9 * np.abs(rinfo).max()
        ok, H, B = gpu.CaculateMatrixHAndB(scene.scan, scene.init[0])
        rH, rB, rres, rhits = ref.compute_hb(scene.scan, scene.init[0])
        hits, _ = gpu.DebugPoints(scene.scan, scene.init[0], 0)
        assert np.array_equal(hits, rhits)
        assert rel(H, rH) < H_TOL and rel(B, rB) < H_TOL
        assert gpu.last_result["n_effective"] == rres["n_effective"]
        for loop_mode in (0, 1):
            g = L.NdtRegistration(L.NdtOptions(method_=L.NdtMethod.INCREMENTAL_NDT, max_iteration_=8, eps_=0.0,
                                               nearby_type_=nearby, loop_mode=loop_mode))
            r = O.OracleIncNdt(max_iteration=8, eps=0.0, nearby6=nearby, skip_nonfinite=1)
            g.SetInputTarget(scene.map)
            r.set_target(scene.map)
            _, cloud, pose = g.ScanMatch(scene.scan, scene.init[0])
            rpose, rcloud, rr = r.align(scene.scan, scene.init[0])
            dr, dt = pose_delta(pose, rpose)
            assert dr < ROT_TOL and dt < TRANS_TOL and g.last_result["iters"] == rr["iters"]
            assert np.abs(cloud[:, :3] - rcloud[:, :3]).max() < 2e-4
    # too few residuals: the loop stops at once, the pose IS written (= the prediction)  (ndt_registration.cpp:349-353)
    far = scene.scan.copy()
    far[:, :3] += np.float32(5000.0)
    _, _, pose = g.ScanMatch(far, scene.init[0], want_cloud=False)
    assert g.last_result["degenerate"] == 1 and g.last_result["pose_written"] == 1 and g.last_result["iters"] == 1
    assert np.array_equal(pose, scene.init[0])


def test_voxel_grid_large_extent_and_overflow(icp_pair):
    """pcl::VoxelGrid over an index space far larger than the cloud (a far outlier in a key frame): filtered exactly like
    the oracle up to PCL's limit of 2^31 - 1 voxels, returned unfiltered beyond it (voxel_grid.hpp)."""
    gpu, _ = icp_pair
    rng = np.random.default_rng(5)
    pts = np.zeros((20000, 4), np.float32)
    pts[:, :3] = rng.normal(0, 8, (20000, 3))
    pts[:, 3] = rng.random(20000)
    far = pts.copy()
    far[0, :3] = [600.0, -550.0, 40.0]      # 0.5 m leaf: ~1300 x 1200 x 200 = 3e8 voxels (> 2^28, < 2^31)
    got, exp = gpu.VoxelFilter(far, 0.5), O.filter_voxel_grid(far, 0.5)
    assert len(got) == len(exp) and np.array_equal(got, exp)
    far[0, :3] = [9000.0, -9000.0, 900.0]   # 18000 x 18000 x 1900 voxels: the int index would overflow
    got, exp = gpu.VoxelFilter(far, 0.5), O.filter_voxel_grid(far, 0.5)
    assert len(exp) == len(far) and np.array_equal(got, exp) and np.array_equal(got, far)


@pytest.mark.parametrize("capacity,n_keys,n_pts,seed", [(4, 6, 60, 0), (9, 12, 400, 1), (33, 40, 3000, 2), (33, 200, 3000, 3),
                                                         (120, 150, 5000, 4), (2, 5, 50, 5), (50, 30, 2000, 6), (700, 900, 40000, 7)])
def test_inc_ndt_device_lru_adversarial(capacity, n_keys, n_pts, seed):
    """The device-side LRU of the incremental NDT cache against the oracle's literal std::list (ndt_registration.cpp:150-183):
    tiny capacities, voxels evicted and re-inserted within one cloud, runs of equal keys, non-finite points in between."""
    import loc_lib_b200 as L
    rng = np.random.default_rng(seed)
    gpu = L.NdtRegistration(L.NdtOptions(method_=L.NdtMethod.INCREMENTAL_NDT, capacity_=capacity))
    ref = O.OracleIncNdt(capacity=capacity, skip_nonfinite=1)
    for cloud in range(6):
        base = cloud * (n_keys // 3)
        ks = []
        while len(ks) < n_pts:
            ks += [int(base + rng.integers(0, n_keys))] * int(rng.integers(1, 4))
        ks = np.array(ks[:n_pts])
        pts = np.zeros((n_pts, 4), np.float32)
        pts[:, 0] = (ks % 37) + rng.random(n_pts) * 0.98 + 0.01   # voxel (k % 37, k // 37, 0), 1 m voxels
        pts[:, 1] = (ks // 37) + rng.random(n_pts) * 0.98 + 0.01
        pts[:, 2] = rng.random(n_pts) * 0.98 + 0.01
        pts[rng.integers(0, n_pts, n_pts // 50), rng.integers(0, 3, n_pts // 50)] = np.nan
        gpu.SetInputTarget(pts)
        ref.set_target(pts)
        k, mu, info, npts = gpu.Voxels()
        rk, rmu, rinfo, rn = ref.voxels()
        assert len(rk) <= capacity - 1
        assert np.array_equal(k, rk), cloud
        assert np.array_equal(npts, rn), cloud
        assert np.array_equal(mu, rmu), cloud
        assert np.abs(info - rinfo).max() <= 1e-9 * np.abs(rinfo).max()


def test_prefilters_match_oracle(scene, icp_pair):
    gpu, _ = icp_pair
    for width in (8, 4):  # pcl::PointXYZI (32 B) and float4
        pts = np.zeros((len(scene.map), width), np.float32)
        pts[:, :3] = scene.map[:, :3]
        pts[:, 3] = np.arange(len(pts)) % 251 if width == 4 else 1.0
        if width == 8:
            pts[:, 4] = np.arange(len(pts)) % 251
        pts[::97, 1] = np.nan
        pts[5, 0] = np.inf
        assert np.array_equal(gpu.RemoveNanPoint(pts), O.filter_remove_nan(pts))
        lo, hi = np.float32([-10, -5, -1]), np.float32([12, 20, 3])
        box = gpu.BoxFilter(pts, lo, hi)
        assert np.array_equal(box, O.filter_crop_box(pts, lo, hi)) and 0 < len(box) < len(pts)
        for leaf in (0.5, 1.0, 2.5):
            vg = gpu.VoxelFilter(pts, leaf)
            ref = O.filter_voxel_grid(pts, leaf)
            assert vg.shape == ref.shape and np.array_equal(vg, ref) and 0 < len(vg) < len(pts)
    assert len(gpu.VoxelFilter(np.zeros((0, 4), np.float32), 1.0)) == 0
    assert len(gpu.RemoveNanPoint(np.full((7, 4), np.nan, np.float32))) == 0


def test_loc_tracker_device_local_map_equals_host_path(scene):
    """Loc's loop (constant-velocity prediction, re-crop near the box edge) with the global map resident on the device
    gives exactly the poses of the host path: CropBox on the host, SetInputTarget, ScanMatch."""
    import loc_lib_b200 as L
    opts = L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=6, eps_=0.0)
    half, margin = (25.0, 25.0, 25.0), 12.0
    dev = L.IcpRegistration(opts)
    trk = L.LocTracker(dev, scene.map, scene.init[0], half_size=half, margin=margin)
    host = L.IcpRegistration(opts)
    origin = np.float32(scene.init[0][4:])
    hs = np.float32(half)
    local = O.filter_crop_box(scene.map, -hs + origin, hs + origin)
    assert trk.n_local == len(local) and 0 < len(local) < len(scene.map)
    host.SetInputTarget(local)
    predict, last = scene.init[0].copy(), scene.init[0].copy()
    resets = 1
    for step in range(len(scene.scans)):
        # the scene's scans come from different places: move each into the frame of the first so that the walk is short
        scan = scene.scans[0] if step % 2 == 0 else scene.scans[0][::2]
        result, _ = trk.Update(scan)
        _, _, ref = host.ScanMatch(scan, predict, want_cloud=False)
        assert np.array_equal(result, ref)
        predict = L.se3_mul(L.se3_mul(ref, L.se3_inv(last)), ref)
        last = ref
        edge = np.stack([-hs + origin, hs + origin], 1)
        if any(not (abs(ref[4 + i] - edge[i, 0]) > margin and abs(ref[4 + i] - edge[i, 1]) > margin) for i in range(3)):
            origin = np.float32(ref[4:])
            host.SetInputTarget(O.filter_crop_box(scene.map, -hs + origin, hs + origin))
            resets += 1
        assert np.array_equal(trk.predict, predict)
    assert trk.resets == resets


@pytest.mark.parametrize("kind", ["icp", "inc_ndt"])
def test_lio_tracker_device_local_map_equals_host_path(scene, kind):
    """Lio's loop (lio.cpp:238-307): first scan = local map; later scans matched from the constant-velocity prediction;
    key frames transformed, pushed into a sliding window (3 here, so the pop-and-rebuild branch runs), local map
    voxel-filtered and re-indexed - all on the device.  Poses and the local map must equal the host path bit for bit:
    oracle transform + oracle voxel grid + SetInputTarget + ScanMatch."""
    import loc_lib_b200 as L
    if kind == "icp":
        make = lambda: L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=6, eps_=0.0))
    else:
        make = lambda: L.NdtRegistration(L.NdtOptions(method_=L.NdtMethod.INCREMENTAL_NDT, max_iteration_=6, eps_=0.0))
    max_kfs, leaf, kf_dist = 3, 0.5, 0.3
    dev, host = make(), make()
    trk = L.LioTracker(dev, num_kfs_in_local_map=max_kfs, kf_distance=kf_dist, kf_angle_deg=10.0, local_map_leaf=leaf)
    base = scene.scans[0]
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    kfs, local = [], None
    predict, last, last_kf = ident.copy(), ident.copy(), ident.copy()
    n_kf = 0
    for step in range(9):
        scan = base.copy()
        scan[:, 0] -= np.float32(0.2 * step)  # the sensor moves 0.2 m along x per scan
        if step % 3 == 2:
            scan = scan[::2].copy()
        pose, is_kf = trk.AddCloud(scan)
        if step == 0:
            ref, ref_kf = ident.copy(), True
        else:
            _, _, ref = host.ScanMatch(O.filter_remove_nan(scan), predict, want_cloud=False)
            predict = L.se3_mul(L.se3_mul(ref, L.se3_inv(last)), ref)
            last = ref
            delta = L.se3_mul(L.se3_inv(last_kf), ref)
            ref_kf = np.linalg.norm(delta[4:]) > kf_dist
        assert np.array_equal(pose, ref), step
        assert is_kf == ref_kf, step
        if ref_kf:
            last_kf = ref.copy()
            kf = O.transform_cloud_d(scan, ref)  # Lio transforms key frames with the DOUBLE matrix (lio.cpp:278)
            kfs.append(kf)
            if len(kfs) > max_kfs:
                kfs.pop(0)
                local = np.concatenate(kfs)
            else:
                local = kf if local is None else np.concatenate([local, kf])
            local = O.filter_voxel_grid(local, leaf)
            host.SetInputTarget(kf if kind == "inc_ndt" else local)
            n_kf += 1
            assert trk.n_local == len(local), step
            got = dev.GetLocalMap()
            assert np.array_equal(got[:, :3], local[:, :3]), step
    assert n_kf > max_kfs + 1 and trk.keyframes == n_kf  # the window slid at least twice


def test_ndt_degenerate_early_return(scene, ndt_pair):
    """det(H)==0 on the first iteration: result_pose keeps the caller's value (quirk Q11)."""
    gpu, ref = ndt_pair
    far = scene.scan.copy()
    far[:, :3] += np.float32(5000.0)  # no voxel anywhere near
    keep = np.array([0.0, 0.0, 0.70710678, 0.70710678, 1.0, 2.0, 3.0])
    _, cloud, pose = gpu.ScanMatch(far, scene.init[0], result_pose_init=keep)
    rpose, rcloud, rres, _ = ref.align(far, scene.init[0], pose_out_init=keep)
    assert gpu.last_result["pose_written"] == rres["pose_written"] == 0
    assert np.array_equal(pose, keep) and np.array_equal(rpose, keep)
    assert np.abs(cloud[:, :3] - rcloud[:, :3]).max() < 1e-2


# ---------------------------------------------------------------------------------------------- golden fixture
def test_gpu_matches_golden_fixture():
    """tests/golden/registration_small.npz (frozen oracle outputs, tests/golden/make_golden.py) vs the CUDA path."""
    import os
    import golden_cases as G
    import loc_lib_b200 as L
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "registration_small.npz"))
    m, scan, init = g["map"], g["scan"], g["init"]
    q = G.queries(scan, init)
    for name, method, k in (("p2plane", L.IcpMethod.P2PLANE, 5), ("p2p", L.IcpMethod.P2P, 1)):
        gpu = L.IcpRegistration(L.IcpOptions(method_=method, max_iteration_=G.ITERS, eps_=0.0, max_plane_distance_=0.05,
                                             max_nn_distance_=0.3))
        gpu.SetInputTarget(m)
        assert np.array_equal(gpu.Knn(q, k), g["knn%d" % k])
        ok, H, B = gpu.CaculateMatrixHAndB(scan, init)
        gate, nn = gpu.DebugPoints(scan, init, k)
        assert np.array_equal(gate, g[name + "_gate"]) and np.array_equal(nn, g["knn%d" % k])
        assert rel(H, g[name + "_H"]) < H_TOL and rel(B, g[name + "_B"]) < H_TOL
        assert [gpu.last_result["n_effective"], gpu.last_result["n_inlier"]] == g[name + "_counts"].tolist()
        _, _, pose = gpu.ScanMatch(scan, init, want_cloud=False)
        dr, dt = pose_delta(pose, g[name + "_trace"][-1])
        assert dr < ROT_TOL and dt < TRANS_TOL
    ndt = L.NdtRegistration(L.NdtOptions(max_iteration_=G.ITERS, eps_=0.0))
    ndt.SetInputTarget(m)
    k_, mu, info, npts = ndt.Voxels()
    assert np.array_equal(k_, g["ndt_keys"]) and np.array_equal(npts, g["ndt_npts"]) and np.array_equal(mu, g["ndt_mu"])
    ok, H, B = ndt.CaculateMatrixHAndB(scan, init)
    hits, _ = ndt.DebugPoints(scan, init, 0)
    assert np.array_equal(hits, g["ndt_hits"])
    assert rel(H, g["ndt_H"]) < H_TOL and rel(B, g["ndt_B"]) < H_TOL
    _, _, pose = ndt.ScanMatch(scan, init, want_cloud=False)
    dr, dt = pose_delta(pose, g["ndt_trace"][-1])
    assert dr < ROT_TOL and dt < TRANS_TOL


@pytest.mark.gpu
def test_relocalise_long_queue_paths_agree_bit_for_bit():
    """Stage 2 of long queues: arrival order, spatial order (k_queue_bin_count / scatter) and the block-pyramid ball query
    (k_icp_nn_pyr) must give the same neighbours, hence bit-identical poses of EVERY hypothesis.  Two Gauss-Newton
    iterations + the score pass: all before the tracked iterations, whose margin shortcut keeps a neighbour SET in the
    order of its last full search (which stage-2 kernel leaves which margins behind then shows up in the last bits of a
    plane).  The switches are read once per process, so each variant runs tools/sanitize_reloc.py (96 hypotheses, two
    thirds far off, LOCREG_SORT_MIN=1 so that the small queue takes the long-queue path) in its own process."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    # "waves": 96 hypotheses in waves of 17 (0.0005 GiB of neighbour scratch) instead of one
    for name, env in (("arrival", {"LOCREG_SORT": "0", "LOCREG_SORT_MIN": "1"}), ("sorted", {}), ("pyramid", {"LOCREG_PYR_KERNEL": "1"}),
                      ("waves", {"LOCREG_RELOC_WAVE_GIB": "0.0005"})):
        e = dict(os.environ); e.update(env); e["SANITIZE_RELOC_ITERS"] = "2"
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_reloc.py")], env=e, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = [l.split(" launches")[0] for l in out.stdout.splitlines() if l.startswith("method")]
        assert len(lines) == 2, out.stdout
        outs[name] = (lines, [int(l.split("launches ")[1]) for l in out.stdout.splitlines() if l.startswith("method")])
    assert outs["arrival"][0] == outs["sorted"][0] == outs["pyramid"][0] == outs["waves"][0], outs
    assert outs["sorted"][1][0] > outs["arrival"][1][0]  # the sort kernels really ran
    assert outs["waves"][1][0] > 2 * outs["arrival"][1][0]  # ... and so did six waves
