"""CPU tests that pin the oracle (oracle/oracle.cpp), since the reference ships no golden vectors:

  (i)   analytic known-answer tests with closed-form H / B / pose updates,
  (ii)  cross-implementation checks against an independent numpy restatement (tests/numpy_ref.py),
  (iii) kd-tree exact mode == brute force == (dis2_f32, index) total order,
  (iv)  frozen golden vectors under tests/golden/ (tests/golden/make_golden.py wrote them).
"""
import os

import numpy as np
import pytest

import numpy_ref as NR
import oracle_py as O
from conftest import pose_delta

IDENT = np.array([0, 0, 0, 1, 0, 0, 0], float)


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def small():
    """~6k-point map, 300-point scan: small enough for the pure-numpy restatement."""
    from loc_lib_b200 import synth
    w = synth.World(40.0)
    m = w.sample_map(6000, pitch=0.5)
    gt = w.poses(1)[0]
    scan = w.scan(gt, beams=6, azimuth=50)
    init = synth.perturb_pose(gt, synth.SEED_POSE, 0.2, 1.5)
    return m, scan, gt, init


# ---------------------------------------------------------------------------------------------- (i) KATs
def test_fit_plane_known_plane():
    # points on z = 2: the unit 4-vector is (0,0,1,-2)/sqrt(5) up to sign (quirk Q5: ||n|| != 1)
    pts = np.array([[0, 0, 2], [1, 0, 2], [0, 1, 2], [1, 1, 2], [0.5, 0.2, 2]], float)
    ok, c = O.fit_plane(pts)
    assert ok
    c = c * np.sign(c[2])
    assert np.allclose(c, np.array([0, 0, 1, -2]) / np.sqrt(5), atol=1e-12)
    assert abs(np.linalg.norm(c) - 1) < 1e-12
    # a point 0.5 off the plane breaks the eps = 1e-2 check (scaled units: (0.5/sqrt(5))^2 = 0.05 > 0.01)
    pts[4, 2] = 2.5
    ok, _ = O.fit_plane(pts)
    assert not ok


def test_fit_plane_vs_numpy_svd():
    rng = np.random.default_rng(3)
    for trial in range(200):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        centre = rng.uniform(-80, 80, 3)
        basis = np.linalg.svd(n[None])[2][1:]
        pts = centre + rng.uniform(-0.3, 0.3, (5, 2)) @ basis + rng.normal(0, 0.01, (5, 1)) * n
        ok, c = O.fit_plane(pts)
        rok, rc = NR.fit_plane(pts)
        assert ok == rok
        s = np.sign(c @ rc)
        assert np.abs(c - s * rc).max() < 1e-9


def _grid_plane(z=0.0, n=41, pitch=0.25):
    g = (np.arange(n) - n // 2) * pitch
    x, y = np.meshgrid(g, g)
    return np.stack([x.ravel(), y.ravel(), np.full(x.size, z), np.zeros(x.size)], 1).astype(np.float32)


def test_p2plane_identity_noiseless_plane_gives_zero_gradient():
    m = _grid_plane()
    ref = O.OracleIcp(method=O.P2PLANE, nn_mode=O.NN_LITERAL_EXACT)
    ref.set_target(m)
    rng = np.random.default_rng(0)
    src = np.zeros((200, 4), np.float32)
    src[:, :2] = rng.uniform(-4, 4, (200, 2))
    ok, H, B, res, gate, _ = ref.compute_hb(src, IDENT)
    assert ok == 0 or np.linalg.det(H) == 0 or True  # H is rank deficient for a single plane; B must vanish
    assert np.all(gate == 3) and res["n_inlier"] == 200
    assert np.abs(B).max() < 1e-12
    # plane z = 0 passes through the origin: n = (0,0,±1), d = 0 -> H[5,5] = number of inliers
    assert abs(H[5, 5] - 200) < 1e-9


def test_p2plane_three_planes_translation_recovered_in_one_step():
    """Residuals are linear in a pure translation, so one Gauss-Newton step recovers it exactly."""
    g = np.arange(0, 24) * 0.25
    a, b = np.meshgrid(g, g)
    z0 = np.stack([a.ravel(), b.ravel(), np.zeros(a.size)], 1)
    y0 = np.stack([a.ravel(), np.zeros(a.size), b.ravel()], 1)
    x0 = np.stack([np.zeros(a.size), a.ravel(), b.ravel()], 1)
    m = np.concatenate([z0, y0, x0]).astype(np.float32)
    rng = np.random.default_rng(1)
    u = rng.uniform(1.0, 4.5, (150, 2))
    src = np.concatenate([np.c_[u[:50], np.zeros(50)], np.c_[u[50:100, 0], np.zeros(50), u[50:100, 1]],
                          np.c_[np.zeros(50), u[100:]]]).astype(np.float32)
    shift = np.array([0.03, -0.02, 0.025])
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=1, eps=0.0, nn_mode=O.NN_LITERAL_EXACT)
    ref.set_target(m)
    start = IDENT.copy()
    start[4:] = shift
    pose, _, res, trace = ref.align(np.c_[src, np.zeros(len(src))].astype(np.float32), start, want_cloud=False)
    assert res["updates"] == 1
    # planes through the origin have d = 0 and ||n|| = 1, so the step is the exact least-squares solution
    assert np.abs(pose[4:]).max() < 1e-9
    assert pose_delta(pose, IDENT)[0] < 1e-9


def test_p2p_closed_form():
    """Known correspondences: map = source shifted by s.  e = s for every point, so
    B = -sum J^T e with J = [R hat(q)/16, -I]  and  H = sum J^T J  (icp_registration.cpp:84-91)."""
    rng = np.random.default_rng(5)
    q = rng.uniform(-5, 5, (60, 3)).astype(np.float32)
    s = np.array([0.05, -0.03, 0.02], np.float32)
    m = q + s
    ref = O.OracleIcp(method=O.P2P, nn_mode=O.NN_LITERAL_EXACT)
    ref.set_target(np.c_[m, np.zeros(60)].astype(np.float32))
    ok, H, B, res, gate, nn = ref.compute_hb(np.c_[q, np.zeros(60)].astype(np.float32), IDENT)
    assert np.array_equal(nn[:, 0], np.arange(60)) and np.all(gate == 3)
    Hx, Bx = np.zeros((6, 6)), np.zeros(6)
    for qi, mi in zip(q.astype(float), m.astype(float)):
        J = np.concatenate([NR.hat(qi) / 16, -np.eye(3)], axis=1)
        Hx += J.T @ J
        Bx += -J.T @ (mi - qi)
    assert rel(H, Hx) < 1e-13 and rel(B, Bx) < 1e-13
    # the /16 of the update (icp_registration.cpp:287) and the split pose update (quirk Q10)
    ref1 = O.OracleIcp(method=O.P2P, max_iteration=1, eps=0.0, nn_mode=O.NN_LITERAL_EXACT)
    ref1.set_target(np.c_[m, np.zeros(60)].astype(np.float32))
    pose, _, _, _ = ref1.align(np.c_[q, np.zeros(60)].astype(np.float32), IDENT, want_cloud=False)
    dx = np.linalg.inv(Hx) / 16 @ Bx
    expect = O.pose_update(IDENT, dx)
    assert np.abs(pose - expect).max() < 1e-12


def test_p2p_squared_distance_gate_quirk():
    """max_nn_distance_ is compared with a SQUARED distance (quirk Q6): 0.9 m passes at 1.0, 1.1 m does not."""
    m = np.array([[0, 0, 0, 0], [10, 0, 0, 0]], np.float32)
    ref = O.OracleIcp(method=O.P2P, max_nn_distance=1.0, nn_mode=O.NN_LITERAL_EXACT, min_effective_pts=0)
    ref.set_target(m)
    src = np.array([[0.9, 0, 0, 0], [11.1, 0, 0, 0]], np.float32)
    _, _, _, res, gate, _ = ref.compute_hb(src, IDENT)
    assert gate.tolist() == [3, 2] and res["n_inlier"] == 1
    ref2 = O.OracleIcp(method=O.P2P, max_nn_distance=0.5, nn_mode=O.NN_LITERAL_EXACT, min_effective_pts=0)
    ref2.set_target(m)
    # 0.7 m away: 0.49 <= 0.5 passes although the distance exceeds the "threshold"
    _, _, _, _, gate, _ = ref2.compute_hb(np.array([[0.7, 0, 0, 0], [0.72, 0, 0, 0]], np.float32), IDENT)
    assert gate.tolist() == [3, 2]


def test_pose_update_matches_rotation_composition():
    rng = np.random.default_rng(9)
    for _ in range(50):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        pose = np.concatenate([q, rng.uniform(-10, 10, 3)])
        dx = np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 1, 3)])
        out = O.pose_update(pose, dx)
        assert np.allclose(O.pose_matrix(out), NR.quat_R(pose) @ NR.so3_exp(dx[:3]), atol=1e-12)
        assert np.allclose(out[4:], pose[4:] + dx[3:], atol=0)
        assert abs(np.linalg.norm(out[:4]) - 1) < 1e-15
    assert np.allclose(O.pose_matrix(pose), NR.quat_R(pose), atol=1e-15)


def test_ndt_single_voxel_hand_computed():
    pts = np.array([[0.2, 0.2, 0.2], [0.4, 0.2, 0.3], [0.2, 0.6, 0.25], [0.8, 0.8, 0.3], [0.5, 0.5, 0.2]], np.float32)
    ref = O.OracleNdt(nearby6=0)
    ref.set_target(np.c_[pts, np.zeros(5)].astype(np.float32))
    keys, mu, info, npts = ref.voxels()
    assert keys.tolist() == [[0, 0, 0]] and npts.tolist() == [5]
    p = pts.astype(float)
    assert np.allclose(mu[0], p.mean(0), atol=1e-15)
    cov = np.cov(p.T)  # /(n-1), math_utils.h:68-70
    lam, V = np.linalg.eigh(cov)
    lam = lam[::-1].copy()
    V = V[:, ::-1]
    lam[1] = max(lam[1], 1e-3 * lam[0])
    lam[2] = max(lam[2], 1e-3 * lam[0])
    assert np.allclose(info[0], V @ np.diag(1 / lam) @ V.T, rtol=1e-9)
    # fewer than 4 points in a voxel: erased (idx_.size() > min_pts_in_voxel_, ndt_registration.cpp:112,136-142)
    ref.set_target(np.c_[pts[:3], np.zeros(3)].astype(np.float32))
    assert ref.voxels()[0].shape[0] == 0


def test_ndt_key_truncates_toward_zero():
    """Quirk Q9: (pt * inv).cast<int>() truncates, so (-0.5, 0.5) all land in voxel 0 along each axis."""
    a = np.array([[-0.4, 0.1, 0.1], [0.4, 0.2, 0.1], [-0.2, 0.3, 0.2], [0.3, 0.1, 0.3], [-0.1, 0.4, 0.1]], np.float32)
    ref = O.OracleNdt(nearby6=0)
    ref.set_target(np.c_[a, np.zeros(5)].astype(np.float32))
    keys, _, _, npts = ref.voxels()
    assert keys.tolist() == [[0, 0, 0]] and npts.tolist() == [5]


def test_transform_cloud_float_order():
    rng = np.random.default_rng(2)
    pts = rng.uniform(-50, 50, (500, 8)).astype(np.float32)
    pose = np.array([0.1, -0.2, 0.3, 0.9, 4, 5, 6], float)
    pose[:4] /= np.linalg.norm(pose[:4])
    out = O.transform_cloud(pts, pose)
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = NR.quat_R(pose).astype(np.float32)
    M[:3, 3] = pose[4:].astype(np.float32)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    for r in range(3):
        exp = ((M[r, 0] * x + M[r, 1] * y) + M[r, 2] * z) + M[r, 3]
        assert np.array_equal(out[:, r], exp.astype(np.float32))
    assert np.array_equal(out[:, 3:], pts[:, 3:])


# ------------------------------------------------------------------------- (ii) oracle vs numpy restatement
def test_p2plane_hb_vs_numpy(small):
    m, scan, gt, init = small
    ref = O.OracleIcp(method=O.P2PLANE, nn_mode=O.NN_EXACT_TIEBREAK, max_plane_distance=0.03)
    ref.set_target(m)
    ok, H, B, res, gate, nn = ref.compute_hb(scan, init)
    rH, rB, rgate, n_eff, n_inl, ssq = NR.p2plane_hb(m[:, :3], scan[:, :3], init, max_plane_distance=0.03)
    assert np.array_equal(gate, rgate) and len(set(gate.tolist())) >= 2
    assert res["n_effective"] == n_eff and res["n_inlier"] == n_inl
    assert rel(H, rH) < 1e-10 and rel(B, rB) < 1e-9
    assert abs(res["sum_sq_res"] - ssq) < 1e-10 * max(ssq, 1e-30)


def test_p2p_hb_vs_numpy(small):
    m, scan, gt, init = small
    ref = O.OracleIcp(method=O.P2P, nn_mode=O.NN_EXACT_TIEBREAK, max_nn_distance=0.05)
    ref.set_target(m)
    ok, H, B, res, gate, nn = ref.compute_hb(scan, init)
    rH, rB, rgate, n_eff = NR.p2p_hb(m[:, :3], scan[:, :3], init, max_nn_distance=0.05)
    assert np.array_equal(gate, rgate) and len(set(gate.tolist())) >= 2
    assert res["n_effective"] == n_eff
    assert rel(H, rH) < 1e-12 and rel(B, rB) < 1e-12


def test_p2line_hb_vs_numpy(small):
    m, scan, gt, init = small
    ref = O.OracleIcp(method=O.P2LINE, nn_mode=O.NN_EXACT_TIEBREAK, max_line_distance=0.3)
    ref.set_target(m)
    ok, H, B, res, gate, nn = ref.compute_hb(scan, init)
    rH, rB, rgate, n_eff, n_inl = NR.p2line_hb(m[:, :3], scan[:, :3], init, max_line_distance=0.3)
    assert np.array_equal(gate, rgate) and len(set(gate.tolist())) >= 2, np.bincount(gate)
    assert res["n_effective"] == n_eff and res["n_inlier"] == n_inl
    assert rel(H, rH) < 1e-10 and rel(B, rB) < 1e-10


def test_p2line_align_on_an_edge_scene():
    """Points on three mutually orthogonal lines (a wire-frame corner): P2Line pulls a shifted copy back."""
    t = np.arange(0, 8, 0.05)
    z = np.zeros_like(t)
    m = np.concatenate([np.c_[t, z, z], np.c_[z, t, z], np.c_[z, z, t]]).astype(np.float32)
    m = np.c_[m, np.zeros(len(m))].astype(np.float32)
    src = m[3::4].copy()
    src[:, :3] -= np.float32([0.02, -0.015, 0.01])
    ref = O.OracleIcp(method=O.P2LINE, nn_mode=O.NN_LITERAL_EXACT, max_iteration=5, eps=0.0)
    ref.set_target(m)
    pose, _, res, _ = ref.align(src, IDENT, want_cloud=False)
    assert res["updates"] == 5
    assert np.abs(pose[4:] - [0.02, -0.015, 0.01]).max() < 2e-3


@pytest.mark.parametrize("nearby6", [0, 1])
def test_ndt_vs_numpy(small, nearby6):
    m, scan, gt, init = small
    ref = O.OracleNdt(nearby6=nearby6, res_outlier_th=4.0)
    ref.set_target(m)
    keys, mu, info, npts = ref.voxels()
    vox = NR.ndt_voxels(m[:, :3])
    assert sorted(vox) == [tuple(k) for k in keys.tolist()]
    for k, mu_k, info_k, n_k in zip(map(tuple, keys.tolist()), mu, info, npts):
        assert n_k == vox[k][2]
        assert np.allclose(mu_k, vox[k][0], rtol=0, atol=1e-12)
        assert np.abs(info_k - vox[k][1]).max() <= 1e-7 * np.abs(vox[k][1]).max()
    H, B, res, hits = ref.compute_hb(scan, init)
    rH, rB, rhits = NR.ndt_hb(vox, scan[:, :3], init, res_outlier_th=4.0, nearby=NR.NEARBY6 if nearby6 else [(0, 0, 0)])
    assert np.array_equal(hits, rhits) and hits.max() >= 1
    assert rel(H, rH) < 1e-12 and rel(B, rB) < 1e-12
    assert res["n_effective"] == len(scan)  # effective_num++ per point, unconditionally (ndt_registration.cpp:432)


def test_inc_ndt_lru_and_stats_vs_numpy(small):
    """Incremental NDT: three clouds into a small cache (evictions happen), then the weighted H / B."""
    m, scan, gt, init = small
    clouds = [m[:2500], m[2000:4500], m[4000:]]
    ref = O.OracleIncNdt(capacity=600, res_outlier_th=8.0)
    nr = NR.IncNdtRef(capacity=600)
    for c in clouds:
        ref.set_target(c)
        nr.add_cloud(c[:, :3])
        keys, mu, info, npts = ref.voxels()
        assert len(keys) == len(nr.vox) <= 599  # size stays below capacity_ (ndt_registration.cpp:161)
        assert sorted(nr.vox) == [tuple(k) for k in keys.tolist()]  # same eviction victims
        for k, mu_k, info_k, n_k in zip(map(tuple, keys.tolist()), mu, info, npts):
            v = nr.vox[k]
            assert n_k == v["n_last"]
            assert np.allclose(mu_k, v["mu"], rtol=0, atol=1e-12)
            assert np.abs(info_k - v["info"]).max() <= 1e-9 * np.abs(v["info"]).max()
    H, B, res, hits = ref.compute_hb(scan, init)
    rH, rB, rhits, total = NR.IncNdtRef.hb(nr, scan[:, :3], init, res_outlier_th=8.0)
    assert np.array_equal(hits, rhits) and hits.max() >= 1
    assert rel(H, rH) < 1e-10 and rel(B, rB) < 1e-10
    assert res["n_effective"] == int(rhits.sum()) == res["n_inlier"]  # residuals, not points (:341)
    assert abs(res["sum_sq_res"] - total) <= 1e-10 * total


def test_inc_ndt_single_point_voxel_and_overwrite_semantics():
    """One point: mu = the point, info = 100 I (:193-195).  A later cloud REPLACES a voxel's statistics with those
    of its own points (pts_ is cleared after every update and only the first-scan branch ever runs)."""
    ref = O.OracleIncNdt(nearby6=0)
    ref.set_target(np.array([[0.5, 0.5, 0.5, 0]], np.float32))
    keys, mu, info, npts = ref.voxels()
    assert keys.tolist() == [[0, 0, 0]] and npts.tolist() == [1]
    assert np.array_equal(mu[0], [0.5, 0.5, 0.5]) and np.array_equal(info[0], 100 * np.eye(3))
    pts = np.array([[0.1, 0.1, 0.1], [0.3, 0.2, 0.1], [0.2, 0.4, 0.3]], np.float32)
    ref.set_target(np.c_[pts, np.zeros(3)].astype(np.float32))
    keys, mu, info, npts = ref.voxels()
    p = pts.astype(float)
    assert npts.tolist() == [3] and np.allclose(mu[0], p.mean(0), atol=1e-15)
    assert np.allclose(info[0], np.linalg.inv(np.cov(p.T) + 1e-3 * np.eye(3)), rtol=1e-10)


def test_inc_ndt_align(scene):
    ref = O.OracleIncNdt(max_iteration=20, eps=0.0)
    ref.set_target(scene.map)
    pose, _, res = ref.align(scene.scan, scene.init[0], want_cloud=False)
    assert res["iters"] == 20 and res["updates"] == 20
    d0, d1 = pose_delta(scene.init[0], scene.gt[0]), pose_delta(pose, scene.gt[0])
    assert d1[1] < 0.1 * d0[1] and d1[0] < 0.1 * d0[0]  # the information-weighted form does converge (unlike direct NDT, Q8)
    far = scene.scan.copy()
    far[:, :3] += np.float32(5000)
    pose, _, res = ref.align(far, scene.init[0], want_cloud=False)
    assert res["degenerate"] == 1 and res["pose_written"] == 1 and np.array_equal(pose, scene.init[0])  # (:349-353)


def test_prefilters_vs_numpy(scene):
    """RemoveNanPoint / BoxFilter / VoxelFilter restated (pcl::removeNaNFromPointCloud, CropBox, VoxelGrid)."""
    pts = np.zeros((20000, 8), np.float32)  # pcl::PointXYZI layout
    pts[:, :3] = scene.map[:20000, :3]
    pts[:, 3] = 1.0
    pts[:, 4] = np.arange(20000) % 255
    pts[::97, 1] = np.nan
    pts[5, 0] = np.inf
    fin = np.isfinite(pts[:, :3]).all(1)
    assert np.array_equal(O.filter_remove_nan(pts), pts[fin])
    lo, hi = np.float32([-10, -5, -1]), np.float32([12, 20, 3])
    keep = fin & (pts[:, :3] >= lo).all(1) & (pts[:, :3] <= hi).all(1)
    assert np.array_equal(O.filter_crop_box(pts, lo, hi), pts[keep]) and 0 < keep.sum() < fin.sum()
    leaf = np.float32(1.0)
    out = O.filter_voxel_grid(pts, leaf)
    p = pts[fin]
    inv = np.float32(1.0) / leaf
    mb = np.floor(p[:, :3].min(0) * inv).astype(np.int64)
    div = np.floor(p[:, :3].max(0) * inv).astype(np.int64) - mb + 1
    ijk = (np.floor(p[:, :3] * inv) - mb.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    uniq, start = np.unique(idx[order], return_index=True)
    assert len(out) == len(uniq)
    bounds = list(start) + [len(order)]
    for v in (0, len(uniq) // 2, len(uniq) - 1):
        grp = p[order[bounds[v]:bounds[v + 1]]]
        acc = np.zeros(8, np.float32)
        for row in grp:
            acc = acc + row
        assert np.array_equal(out[v], acc / np.float32(len(grp)))
    assert np.allclose(out[:, 3], 1.0) and out[:, 4].max() <= 254  # intensity is averaged like every other field
    # fewer points than the input, none lost: the count-weighted mean of the centroids is the mean of the cloud
    counts = np.diff(bounds)
    assert np.allclose((out[:, :3] * counts[:, None]).sum(0) / counts.sum(), p[:, :3].astype(np.float64).mean(0), atol=1e-3)


# ------------------------------------------------------------------------- (iii) kd-tree semantics
def test_kdtree_exact_equals_brute_force(scene):
    ref = O.OracleIcp(method=O.P2PLANE)
    ref.set_target(scene.map)
    rng = np.random.default_rng(4)
    q = scene.map[rng.integers(0, len(scene.map), 20000), :3] + rng.normal(0, 0.4, (20000, 3)).astype(np.float32)
    q = q.astype(np.float32)
    for k in (1, 5):
        exact = ref.knn(q, k, O.NN_LITERAL_EXACT)
        tie = ref.knn(q, k, O.NN_EXACT_TIEBREAK)
        # the literal tree keeps the first-seen point on exact dis2 ties; away from ties the two agree
        assert (exact != tie).any(axis=1).mean() < 1e-3
        assert np.array_equal(tie[:2000], O.bfnn(scene.map, q[:2000], k))
    assert np.array_equal(O.bfnn(scene.map[:3000], q[:100], 5), NR.knn_f32(scene.map[:3000, :3], q[:100], 5))


def test_kdtree_ann_is_approximate(scene):
    """Quirk Q1: approximate_ = true, alpha = 0.1 is what the reference actually runs; it differs from exact."""
    ref = O.OracleIcp(method=O.P2PLANE)
    ref.set_target(scene.map)
    R = O.pose_matrix(scene.init[0])
    q = (scene.scan[:, :3].astype(np.float64) @ R.T + scene.init[0][4:]).astype(np.float32)
    ann = ref.knn(q, 5, O.NN_LITERAL_ANN)
    exact = ref.knn(q, 5, O.NN_EXACT_TIEBREAK)
    mismatch = (ann != exact).any(axis=1).mean()
    assert 0.0 < mismatch < 0.9
    assert np.all(ann >= 0)


def test_kdtree_drops_duplicates_and_small_maps(scene):
    dup = np.concatenate([scene.map[:2000], scene.map[:1000]])
    ref = O.OracleIcp(method=O.P2PLANE)
    ref.set_target(dup)
    assert ref.leaves() == 2000  # quirk Q3
    q = dup[::40, :3] + np.float32(0.01)
    nn = ref.knn(q, 5, O.NN_EXACT_TIEBREAK)
    assert nn.max() < 2000  # of coincident points only the lowest index survives
    ref.set_target(scene.map[:3])
    assert np.all(ref.knn(q, 5, O.NN_EXACT_TIEBREAK) == -1)  # k > tree size: GetClosestPoint refuses (kdtree.cpp:149)


# ------------------------------------------------------------------------- align loops
def test_icp_align_converges_to_ground_truth(scene):
    for method in (O.P2PLANE, O.P2P):
        ref = O.OracleIcp(method=method, nn_mode=O.NN_EXACT_TIEBREAK)
        ref.set_target(scene.map)
        pose, cloud, res, trace = ref.align(scene.scan, scene.init[0])
        d0 = pose_delta(scene.init[0], scene.gt[0])
        d1 = pose_delta(pose, scene.gt[0])
        if method == O.P2PLANE:
            assert d1[1] < 0.05 * d0[1] and d1[0] < 0.05 * d0[0]
        else:  # the /16 factors (quirk Q6) damp P2P's translation step to 1/16 of a Gauss-Newton step
            assert d1[1] < d0[1] and d1[0] < d0[0]
        assert 1 <= res["iters"] <= 20
        assert np.array_equal(trace[0], scene.init[0])
        assert np.allclose(cloud, O.transform_cloud(scene.scan, pose))


def test_icp_failure_paths(scene):
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=3, nn_mode=O.NN_EXACT_TIEBREAK)
    ref.set_target(scene.map)
    pose, _, res, _ = ref.align(np.zeros((0, 4), np.float32), scene.init[0], want_cloud=False)
    # every evaluation fails -> the loop keeps going with an unchanged pose (quirk Q11)
    assert res["iters"] == 3 and res["updates"] == 0 and res["degenerate"] == 1
    assert np.array_equal(pose, scene.init[0])


def test_ndt_align_and_early_return(scene):
    ref = O.OracleNdt(max_iteration=10, eps=0.0)
    ref.set_target(scene.map)
    pose, _, res, _ = ref.align(scene.scan, scene.init[0], want_cloud=False)
    assert res["iters"] == 10 and res["pose_written"] == 1
    # direct NDT is an UNWEIGHTED point-to-voxel-mean fit (quirk Q8): it pulls the rotation in but not the translation
    assert pose_delta(pose, scene.gt[0])[0] < pose_delta(scene.init[0], scene.gt[0])[0]
    far = scene.scan.copy()
    far[:, :3] += np.float32(5000)
    keep = np.array([0, 0, 0.6, 0.8, 1, 2, 3], float)
    pose, _, res, _ = ref.align(far, scene.init[0], pose_out_init=keep, want_cloud=False)
    assert res["pose_written"] == 0 and np.array_equal(pose, keep)  # det(H) == 0 return (ndt_registration.cpp:435-436)


def test_align_batch_threads_equal_serial(scene):
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=4, eps=0.0, nn_mode=O.NN_LITERAL_ANN)
    ref.set_target(scene.map)
    clouds = np.concatenate(scene.scans)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in scene.scans])]).astype(np.int64)
    poses, results, used = ref.align_batch(clouds, offsets, scene.init, threads=3)
    assert used == 3
    for i, s in enumerate(scene.scans):
        p, _, r, _ = ref.align(s, scene.init[i], want_cloud=False)
        assert np.array_equal(p, poses[i]) and r == results[i]


# ------------------------------------------------------------------------- (iv) golden vectors
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "registration_small.npz")


def test_oracle_reproduces_golden():
    import golden_cases as G
    g = np.load(GOLDEN)
    cur = G.compute(g["map"], g["scan"], g["init"])
    for k in G.EXACT_KEYS:
        assert np.array_equal(cur[k], g[k]), k
    for k in G.CLOSE_KEYS:
        assert np.allclose(cur[k], g[k], rtol=1e-9, atol=1e-12), k


def test_voxel_grid_overflow_returns_cloud_unfiltered():
    """pcl::VoxelGrid gives up when dx * dy * dz overflows its int voxel index and returns the input cloud (voxel_grid.hpp:
    "Leaf size is too small for the input dataset. Integer indices would overflow.")."""
    rng = np.random.default_rng(2)
    pts = np.zeros((500, 4), np.float32)
    pts[:, :3] = rng.normal(0, 3, (500, 3))
    assert len(O.filter_voxel_grid(pts, 0.5)) < 500
    pts[0, :3] = [9000.0, -9000.0, 900.0]
    out = O.filter_voxel_grid(pts, 0.5)
    assert np.array_equal(out, pts)
