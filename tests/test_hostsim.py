"""CPU tests of the CUDA kernels' per-element LOGIC: the __host__ __device__ bodies of loc_lib_b200/csrc/*.cuh
compiled with g++ into tests/hostsim (test-only, serial) and compared with the oracle on the same seeded inputs.
The real kernels are checked by tests/test_gpu_parity.py on the B200 (-m gpu); this suite lets the exact k-NN
termination, the dedupe, the fast plane fit, the gates and the Gauss-Newton update be debugged without a GPU."""
import numpy as np
import pytest

import hostsim_py as HS
import numpy_ref as NR
import oracle_py as O
from conftest import pose_delta


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def hs_map(scene):
    return HS.HsMap(scene.map)


@pytest.fixture(scope="module")
def ref_icp(scene):
    r = O.OracleIcp(method=O.P2PLANE, nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
    r.set_target(scene.map)
    return r


def _queries(scene):
    R = O.pose_matrix(scene.init[0])
    return (scene.scan[:, :3].astype(np.float64) @ R.T + scene.init[0][4:]).astype(np.float32)


@pytest.mark.parametrize("hint", [0, 0xFFFFFFFD, 0xFFFFFFFF], ids=["lists+mid", "lists", "blocks"])
@pytest.mark.parametrize("cell", [0.5, 0.3, 1.3])
def test_knn_bit_exact_near_and_far(scene, ref_icp, cell, hint):
    """Every stage-2 route: the mid level's lists (default), the corner lists + fine shells (no mid level: what a
    map too large for a second set of lists falls back to), block tables only."""
    m = HS.HsMap(scene.map, cell=cell, capacity_hint=hint)
    rng = np.random.default_rng(7)
    far = rng.uniform(-60, 60, (1500, 3)).astype(np.float32)
    far[:, 2] = rng.uniform(-4, 25, 1500)
    way_out = rng.uniform(-400, 400, (40, 3)).astype(np.float32)  # beyond kBruteForceShell cells: linear-scan path
    for q in (_queries(scene)[:3000], far, way_out):
        for k in (1, 5):
            assert np.array_equal(m.knn(q, k), ref_icp.knn(q, k))


def test_far_queries_use_the_coarse_level_not_a_linear_scan(scene, ref_icp):
    """Queries metres away from every map point (relocalisation hypotheses, points outside the map) must be answered
    by the coarse level's shells; the linear scan is only the last resort hundreds of cells away."""
    rng = np.random.default_rng(5)
    q = rng.uniform(-45, 45, (600, 3)).astype(np.float32)
    q[:, 2] = rng.uniform(25, 40, 600)  # 5 - 20 m above the highest roof
    for hint, expect_linear in ((0, False), (0xFFFFFFFE, True)):
        m = HS.HsMap(scene.map, capacity_hint=hint)
        HS.knn_stats()
        got = m.knn(q, 5)
        st = HS.knn_stats()
        assert np.array_equal(got, ref_icp.knn(q, 5))
        assert (st["linear_scans"] > 0) == expect_linear, st
    out = np.array([[500, 0, 0], [-300, 700, 10]], np.float32)  # beyond the coarse reach: linear scan, still exact
    m = HS.HsMap(scene.map)
    HS.knn_stats()
    assert np.array_equal(m.knn(out, 5), ref_icp.knn(out, 5))
    assert HS.knn_stats()["linear_scans"] == 2


@pytest.mark.parametrize("mode", [1, 2], ids=["after-mid-list", "all-of-stage-2"])
@pytest.mark.parametrize("cell,hint", [(0.5, 0), (0.3, 0), (1.3, 0xFFFFFFFD), (0.5, 0xFFFFFFFF)])
def test_knn_pyramid_ball_query_is_exact(scene, ref_icp, cell, hint, mode):
    """Stage 2 as a depth-first ball query through the block pyramid (PyrWalk - what k_icp_nn_pyr runs per lane): near,
    far, out-of-bounds and way-out queries, unseeded and seeded, lattice ties; never a shell, never a linear scan."""
    m = HS.HsMap(scene.map, cell=cell, capacity_hint=hint)
    m.set_pyr_mode(mode)
    rng = np.random.default_rng(17)
    far = rng.uniform(-60, 60, (1500, 3)).astype(np.float32)
    far[:, 2] = rng.uniform(-4, 40, 1500)
    way_out = rng.uniform(-900, 900, (60, 3)).astype(np.float32)
    near = _queries(scene)[:3000]
    HS.knn_stats()
    for q in (near, far, way_out):
        for k in (1, 5):
            assert np.array_equal(m.knn(q, k), ref_icp.knn(q, k))
    st = HS.knn_stats()
    assert st["pyr_queries"] > 0 and st["ring_queries"] == 0 and st["linear_scans"] == 0, st
    for q in (near, far):
        exp = ref_icp.knn(q, 5)
        for sigma in (0.02, 0.3, 5.0):
            assert np.array_equal(m.knn_seeded(q, (q + rng.normal(0, sigma, q.shape)).astype(np.float32)), exp)


def test_knn_pyramid_ties_and_tiny_maps():
    g = np.arange(-6, 7, dtype=np.float32) * np.float32(0.25)
    x, y, z = np.meshgrid(g, g, g[:3])
    pts = np.stack([x.ravel(), y.ravel(), z.ravel(), np.zeros(x.size)], 1).astype(np.float32)
    rng = np.random.default_rng(0)
    pts = pts[rng.permutation(len(pts))]
    q = np.concatenate([np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.float32)[::3] + np.float32(0.125),
                        rng.uniform(-30, 30, (300, 3)).astype(np.float32)])
    for n in (len(pts), 7, 3, 1):  # fewer points than K: the result is padded, as the shells' is
        m = HS.HsMap(pts[:n], cell=0.5)
        exp = m.knn(q, 5)
        if n >= 5:
            assert np.array_equal(exp, O.bfnn(pts[:n], q, 5))
        for mode in (1, 2):
            m.set_pyr_mode(mode)
            assert np.array_equal(m.knn(q, 5), exp)
            assert np.array_equal(m.knn_seeded(q, q + np.float32(0.25)), exp)
        m.set_pyr_mode(0)
    # a map far from the origin on negative and positive sides (top nodes on both sides of zero)
    big = np.concatenate([pts[:, :3] + np.float32([-700, 300, 5]), pts[:, :3] + np.float32([900, -20, -3])]).astype(np.float32)
    big = np.concatenate([big, np.zeros((len(big), 1), np.float32)], 1)
    m = HS.HsMap(big, cell=0.5)
    qq = np.concatenate([q + np.float32([-700, 300, 5]), q + np.float32([900, -20, -3]), q]).astype(np.float32)
    exp = O.bfnn(big, qq, 5)
    for mode in (0, 1, 2):
        m.set_pyr_mode(mode)
        assert np.array_equal(m.knn(qq, 5), exp)


@pytest.mark.parametrize("k", [1, 5])
def test_knn_tracked_shortcut_is_exact(scene, hs_map, ref_icp, k):
    """Queries that moved less than their margin since their last full search only re-sort their k neighbours
    (knn_track_try).  A converging sequence (steps of 30 cm down to 0.1 mm, as Gauss-Newton iterations produce) must
    give the exact result at every step, and most late steps must take the shortcut."""
    q0 = _queries(scene)[:3000]
    rng = np.random.default_rng(11)
    direction = rng.normal(0, 1, 3); direction /= np.linalg.norm(direction)
    offsets = [0.3, 0.1, 0.03, 0.01, 0.003, 0.001, 0.0003, 0.0001, 0.0]
    # a common translation plus a small per-point part (a rotation moves every point differently)
    steps = np.stack([(q0 + direction * d + rng.normal(0, 0.2 * d, q0.shape)).astype(np.float32) for d in offsets])
    got, skipped = hs_map.knn_tracked(steps, k)
    for s in range(len(offsets)):
        assert np.array_equal(got[s], ref_icp.knn(steps[s], k)), s
    assert skipped > 2.0 * len(q0), skipped  # of 8 tracked steps per point, clearly more than two use the shortcut
    # adversarial: jumps larger than any margin right after tiny ones must fall back to the search
    jumpy = np.stack([steps[8], steps[7], steps[0], steps[8], steps[2], steps[8]])
    got, _ = hs_map.knn_tracked(jumpy, k)
    for s in range(len(jumpy)):
        assert np.array_equal(got[s], ref_icp.knn(jumpy[s], k)), s


def test_knn_seeded_search_is_still_exact(scene, hs_map, ref_icp):
    """Seeds (the previous Gauss-Newton iteration's neighbours) only tighten the threshold: near, far and useless
    seeds must all give the exact result, without duplicates."""
    q = _queries(scene)[:4000]
    rng = np.random.default_rng(3)
    exp = ref_icp.knn(q, 5)
    for sigma in (0.0, 0.02, 0.3, 5.0):
        seed_q = (q + rng.normal(0, sigma, q.shape)).astype(np.float32)
        got = hs_map.knn_seeded(q, seed_q)
        assert np.array_equal(got, exp)
    seed_q = q.copy()
    seed_q[::2] = np.nan  # no seeds for every other query
    assert np.array_equal(hs_map.knn_seeded(q, seed_q), exp)


def test_knn_ties_resolved_by_index():
    """Lattice map: many exactly equal float32 distances; the total order (dis2, index) must decide."""
    g = np.arange(-6, 7, dtype=np.float32) * np.float32(0.25)
    x, y, z = np.meshgrid(g, g, g[:3])
    pts = np.stack([x.ravel(), y.ravel(), z.ravel(), np.zeros(x.size)], 1).astype(np.float32)
    rng = np.random.default_rng(0)
    pts = pts[rng.permutation(len(pts))]
    q = np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.float32)[::3] + np.float32(0.125)
    m = HS.HsMap(pts, cell=0.5)
    assert np.array_equal(m.knn(q, 5), NR.knn_f32(pts[:, :3], q, 5))
    assert np.array_equal(m.knn(q, 5), O.bfnn(pts, q, 5))
    assert np.array_equal(m.knn_seeded(q, q + np.float32(0.25)), O.bfnn(pts, q, 5))


def test_map_stats_and_duplicates(scene):
    dup = np.concatenate([scene.map[:5000], scene.map[:2500]])
    dup[100] = np.nan  # non-finite target points are dropped (deviation D1)
    m = HS.HsMap(dup)
    st = m.stats()
    assert st["n_pts"] == 7499 and st["n_unique"] == 5000  # point 100 survives through its copy at 5100
    r = O.OracleIcp(method=O.P2PLANE)
    keep = np.concatenate([dup[:100], dup[101:]])
    r.set_target(keep)
    assert r.leaves() == 5000
    q = scene.map[::50, :3][:100] + np.float32(0.01)
    got = m.knn(q, 5)
    exp = r.knn(q, 5, O.NN_EXACT_TIEBREAK)
    exp = np.where(exp >= 100, exp + 1, exp)  # indices refer to the caller's cloud, which still holds point 100
    assert np.array_equal(got, exp)
    tiny = HS.HsMap(scene.map[:3])
    assert np.all(tiny.knn(q, 5)[:, 3:] == -1) and np.all(tiny.knn(q, 5)[:, :3] >= 0)


def test_plane_fit_fast_path_matches_svd():
    rng = np.random.default_rng(3)
    n_fast = 0
    for trial in range(3000):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        centre = rng.uniform(-100, 100, 3)
        basis = np.linalg.svd(n[None])[2][1:]
        spread = rng.choice([0.05, 0.3, 1.0])
        pts = centre + rng.uniform(-spread, spread, (5, 2)) @ basis + rng.normal(0, rng.choice([0, 0.01, 0.1]), (5, 1)) * n
        pts = pts.astype(np.float32).astype(np.float64)  # map points are float32
        c = np.zeros(4)
        ok = HS.lib().hs_plane_fit5_fast(np.ascontiguousarray(pts).ctypes.data, c.ctypes.data)
        svd = HS.plane_svd5(pts)
        _, rc = NR.fit_plane(pts)
        assert np.abs(svd * np.sign(svd @ rc) - rc).max() < 1e-8
        if ok:
            n_fast += 1
            assert np.abs(c * np.sign(c @ rc) - rc).max() < 1e-8
    assert n_fast > 2500  # the fast path must be the common case
    # collinear points: the fast path must refuse, the SVD still returns a unit vector
    line = np.outer(np.arange(5.0), [1.0, 2.0, 0.5])
    c = np.zeros(4)
    assert HS.lib().hs_plane_fit5_fast(np.ascontiguousarray(line).ctypes.data, c.ctypes.data) == 0
    assert abs(np.linalg.norm(HS.plane_svd5(line)) - 1) < 1e-12


def test_small_linear_algebra():
    rng = np.random.default_rng(5)
    for _ in range(200):
        A = rng.normal(size=(6, 6))
        H = A @ A.T + 1e-3 * np.eye(6)
        b = rng.normal(size=6)
        Hu = np.array([H[r, c] for r in range(6) for c in range(r, 6)])
        dx = np.zeros(6)
        assert HS.lib().hs_gn_solve6(Hu.ctypes.data, b.ctypes.data, dx.ctypes.data) == 1
        assert np.allclose(dx, np.linalg.solve(H, b), rtol=1e-9)
        C = rng.normal(size=(3, 3))
        S = C @ C.T
        lam, Q = HS.sym3_eigen(np.array([S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]]))
        assert np.allclose(lam, np.linalg.eigvalsh(S)[::-1], rtol=1e-10, atol=1e-14)
        assert np.allclose(Q @ np.diag(lam) @ Q.T, S, atol=1e-12)
    Z = np.zeros(21)
    assert HS.lib().hs_gn_solve6(Z.ctypes.data, np.zeros(6).ctypes.data, np.zeros(6).ctypes.data) == 0  # det == 0


@pytest.mark.parametrize("method", ["P2PLANE", "P2P", "P2LINE"])
def test_hb_gates_match_oracle(scene, hs_map, method):
    mid = getattr(O, method)
    ref = O.OracleIcp(method=mid, max_plane_distance=0.004, max_nn_distance=0.08, max_line_distance=0.05,
                      nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
    ref.set_target(scene.map)
    prm = HS.params(max_nn_distance=0.08, max_plane_distance=0.004, max_line_distance=0.05)
    scan = scene.scan.copy()
    scan[::97, 0] = np.nan
    for pose in (scene.init[0], scene.gt[0]):
        H, B, res, gate, nn = hs_map.icp_hb(mid, prm, scan, pose)
        rok, rH, rB, rres, rgate, rnn = ref.compute_hb(scan, pose)
        assert np.array_equal(nn, rnn) and np.array_equal(gate, rgate)
        assert len(set(rgate.tolist())) >= 3
        assert rel(H, rH) < 1e-9 and rel(B, rB) < 1e-9
        assert res["n_effective"] == rres["n_effective"] and res["n_inlier"] == rres["n_inlier"]
        assert abs(res["sum_sq_res"] - rres["sum_sq_res"]) <= 1e-9 * rres["sum_sq_res"]


@pytest.mark.parametrize("method", ["P2PLANE", "P2P", "P2LINE"])
def test_align_matches_oracle(scene, hs_map, method):
    mid = getattr(O, method)
    for eps, iters in ((0.0, 6), (1e-2, 20)):
        ref = O.OracleIcp(method=mid, max_iteration=iters, eps=eps, nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
        ref.set_target(scene.map)
        pose, st = hs_map.icp_align(mid, HS.params(eps=eps, max_iteration=iters), scene.scans[1], scene.init[1])
        rpose, _, rres, _ = ref.align(scene.scans[1], scene.init[1], want_cloud=False)
        dr, dt = pose_delta(pose, rpose)
        assert dr < 1e-7 and dt < 1e-6
        assert st["iters"] == rres["iters"] and st["updates"] == rres["updates"] and st["converged"] == rres["converged"]


def test_ndt_build_and_loop_match_oracle(scene):
    hs = HS.HsNdt(scene.map)
    ref = O.OracleNdt(max_iteration=6, eps=0.0, skip_nonfinite=1)
    ref.set_target(scene.map)
    k, mu, info, npts = hs.voxels()
    rk, rmu, rinfo, rn = ref.voxels()
    assert np.array_equal(k, rk) and np.array_equal(npts, rn) and np.array_equal(mu, rmu)
    assert np.abs(info - rinfo).max() <= 1e-9 * np.abs(rinfo).max()
    for nearby in (1, 7):
        r2 = O.OracleNdt(max_iteration=6, eps=0.0, nearby6=int(nearby == 7), skip_nonfinite=1)
        r2.set_target(scene.map)
        prm = HS.ndt_params(eps=0.0, max_iteration=6, n_nearby=nearby)
        H, B, res, hits = hs.hb(prm, scene.scan, scene.init[0])
        rH, rB, rres, rhits = r2.compute_hb(scene.scan, scene.init[0])
        assert np.array_equal(hits, rhits) and rel(H, rH) < 1e-9 and rel(B, rB) < 1e-9
        assert res["n_inlier"] == rres["n_inlier"] and res["n_effective"] == rres["n_effective"]
        pose, st = hs.align(prm, scene.scan, scene.init[0])
        rpose, _, rr, _ = r2.align(scene.scan, scene.init[0], want_cloud=False)
        dr, dt = pose_delta(pose, rpose)
        assert dr < 1e-7 and dt < 1e-6 and st["iters"] == rr["iters"]
    far = scene.scan.copy()
    far[:, :3] += np.float32(5000)
    keep = np.array([0, 0, 0.6, 0.8, 1, 2, 3], float)
    pose, st = hs.align(HS.ndt_params(), far, scene.init[0], pose_out_init=keep)
    assert st["pose_written"] == 0 and np.array_equal(pose, keep)


def test_inc_ndt_bodies_match_oracle(scene):
    hs = HS.HsIncNdt(scene.map)
    for nearby in (1, 7):
        ref = O.OracleIncNdt(nearby6=int(nearby == 7), skip_nonfinite=1)
        ref.set_target(scene.map)
        k, mu, info, npts = hs.voxels()
        rk, rmu, rinfo, rn = ref.voxels()
        assert np.array_equal(k, rk) and np.array_equal(npts, rn) and np.array_equal(mu, rmu)
        assert np.abs(info - rinfo).max() <= 1e-9 * np.abs(rinfo).max()
        H, B, res, hits = hs.hb(HS.ndt_params(n_nearby=nearby), scene.scan, scene.init[0])
        rH, rB, rres, rhits = ref.compute_hb(scene.scan, scene.init[0])
        assert np.array_equal(hits, rhits) and rel(H, rH) < 1e-9 and rel(B, rB) < 1e-9
        assert res["n_effective"] == rres["n_effective"] == int(rhits.sum())
        assert abs(res["sum_sq_res"] - rres["sum_sq_res"]) <= 1e-9 * rres["sum_sq_res"]


def test_hostsim_reproduces_golden():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "registration_small.npz"))
    import golden_cases as G
    m = HS.HsMap(g["map"])
    q = G.queries(g["scan"], g["init"])
    assert np.array_equal(m.knn(q, 1), g["knn1"]) and np.array_equal(m.knn(q, 5), g["knn5"])
    prm = HS.params(max_nn_distance=0.3, max_plane_distance=0.05, eps=0.0, max_iteration=G.ITERS)
    for name, mid in (("p2plane", O.P2PLANE), ("p2p", O.P2P)):
        H, B, res, gate, _ = m.icp_hb(mid, prm, g["scan"], g["init"])
        assert np.array_equal(gate, g[name + "_gate"])
        assert rel(H, g[name + "_H"]) < 1e-9 and rel(B, g[name + "_B"]) < 1e-9
        pose, _ = m.icp_align(mid, prm, g["scan"], g["init"])
        dr, dt = pose_delta(pose, g[name + "_trace"][-1])
        assert dr < 1e-7 and dt < 1e-6
