"""-m gpu tests at BASELINE.json's full sizes (1 M-point map, 32 x 940-ray scans, 10 iterations), where the oracle is
too slow to run on everything: size-independent properties plus oracle spot checks on samples.

  * k-NN: a sample of the transformed scan points against brute force over the whole 1 M-point map (bit-exact);
  * batch == single: every scan of a batch gets the same pose as its own ScanMatch call - bit-identical through the
    per-iteration pipeline (loop_mode 1: per-tile partials summed in tile order either way), and to 1e-9 through the
    one-launch persistent kernel (loop_mode 0, the default: the same partial rows summed in another fixed order);
  * registration recovers the ground truth the scans were generated from (the synthetic world is the fixture);
  * idempotence: restarting from the converged pose moves it by less than the convergence threshold;
  * relocalisation: argmin == numpy argmin of the returned scores, ties to the lowest index; the sharded form
    (strided shards + packed-key min) returns the same winner as the single call;
  * H is symmetric PSD and B == -J^T r consistent between compute_hb and the first Gauss-Newton step.
"""
import numpy as np
import pytest

import oracle_py as O
from conftest import pose_delta

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    import loc_lib_b200 as L
    from loc_lib_b200 import synth

    class F:
        pass
    f = F()
    f.world = synth.World(200.0)
    f.map = f.world.sample_map(1_000_000)
    f.S = 24
    f.gt = f.world.poses(f.S)
    buf, counts = f.world.scan_batch(f.gt)
    f.scans = [buf[i, :counts[i]].copy() for i in range(f.S)]
    f.init = synth.perturb_poses(f.gt)
    f.reg = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=10, eps_=0.0))
    f.reg.SetInputTarget(f.map)
    return f


def test_fullsize_knn_sample_vs_brute_force(full):
    R = O.pose_matrix(full.init[0])
    q = (full.scans[0][:, :3].astype(np.float64) @ R.T + full.init[0][4:]).astype(np.float32)
    got = full.reg.Knn(q, 5)
    assert got.min() >= 0 and got.max() < len(full.map)
    sel = np.random.default_rng(0).choice(len(q), 1500, replace=False)
    assert np.array_equal(got[sel], O.bfnn(full.map, q[sel], 5))
    # every returned set is sorted by (float32 dis2, index)
    d = q[:, None, :] - full.map[got, :3]
    d2 = d[..., 0] * d[..., 0] + (d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])
    assert np.all((d2[:, 1:] > d2[:, :-1]) | ((d2[:, 1:] == d2[:, :-1]) & (got[:, 1:] > got[:, :-1])))


def test_fullsize_batch_equals_single_and_recovers_ground_truth(full):
    clouds = np.concatenate(full.scans)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in full.scans])]).astype(np.int64)
    import loc_lib_b200 as L
    poses, results = full.reg.ScanMatchBatch(clouds, offsets, full.init)
    piped = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=10, eps_=0.0, loop_mode=L.LOOP_GRAPH))
    piped.SetInputTarget(full.map)
    for i in range(full.S):
        _, _, single = piped.ScanMatch(full.scans[i], full.init[i], want_cloud=False)
        assert np.array_equal(single, poses[i])
        assert piped.last_result == results[i]
        _, _, one_launch = full.reg.ScanMatch(full.scans[i], full.init[i], want_cloud=False)
        dr, dt = pose_delta(one_launch, poses[i])
        assert dr < 1e-9 and dt < 1e-9
        assert full.reg.last_timing()[1] <= 2  # the loop is ONE launch (+ nothing else: no cloud requested)
        r1 = full.reg.last_result
        assert (r1["iters"], r1["updates"], r1["n_effective"], r1["n_inlier"]) == \
            (results[i]["iters"], results[i]["updates"], results[i]["n_effective"], results[i]["n_inlier"])
        assert results[i]["iters"] == 10 and results[i]["degenerate"] == 0
        dr, dt = pose_delta(poses[i], full.gt[i])
        assert dr < 2e-3 and dt < 0.03  # 2 cm range noise, 1 cm map jitter
    assert np.median([pose_delta(p, g)[1] for p, g in zip(poses, full.gt)]) < 0.006


def test_fullsize_single_scan_of_many_tiles_through_the_one_launch_kernel(full):
    """A 128-beam scan (~230 k points = ~900 tiles, six tiles per block of the persistent grid) through the one-launch
    single-scan kernel - grid-wide and in-block stage 2, the rotating queue counters - against the per-iteration pipeline
    (a batch of one): the same neighbours and planes, sums in another fixed order."""
    scan = full.world.scan(full.gt[3], beams=128, azimuth=1953)
    assert len(scan) > 150_000
    _, cloud, pose = full.reg.ScanMatch(scan, full.init[3])
    assert full.reg.last_timing()[1] <= 3  # the loop is one launch (+ the cloud transform)
    res1 = dict(full.reg.last_result)
    offsets = np.array([0, len(scan)], np.int64)
    poses, res = full.reg.ScanMatchBatch(scan, offsets, full.init[3:4])
    dr, dt = pose_delta(pose, poses[0])
    assert dr < 1e-9 and dt < 1e-9, (dr, dt)
    assert res1["iters"] == res[0]["iters"] == 10 and res1["n_effective"] == res[0]["n_effective"]
    assert res1["n_inlier"] == res[0]["n_inlier"]
    dr, dt = pose_delta(pose, full.gt[3])
    assert dt < 0.05 and dr < 5e-3


def test_fullsize_pipelined_batch_equals_plain_batch(full):
    """locreg_align_batch from PINNED host memory overlaps chunked copies with compute; same poses, bit for bit."""
    import torch
    reps = 2  # 48 scans, 1.3 M points: above the pipelining threshold
    clouds = np.concatenate(full.scans * reps)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in full.scans * reps])]).astype(np.int64)
    init = np.concatenate([full.init] * reps)
    plain, res_plain = full.reg.ScanMatchBatch(clouds, offsets, init)
    pinned = torch.from_numpy(clouds).pin_memory().numpy()
    piped, res_piped = full.reg.ScanMatchBatch(pinned, offsets, init)
    assert np.array_equal(plain, piped) and res_plain == res_piped
    assert np.array_equal(plain[:full.S], plain[full.S:])
    # 32-byte stride (pcl::PointXYZI layout) through the same path
    wide = torch.zeros((len(clouds), 8), dtype=torch.float32).pin_memory().numpy()
    wide[:, :3] = clouds[:, :3]
    piped32, _ = full.reg.ScanMatchBatch(wide, offsets, init)
    assert np.array_equal(plain, piped32)
    # from 4 M points on the batch is cut into four chunks that run concurrently on their own streams, each on its own
    # slice of the per-job scratch (locreg.cu, align_batch_resident): the same poses again
    reps = 7  # 168 scans, 4.6 M points
    clouds = np.concatenate(full.scans * reps)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in full.scans * reps])]).astype(np.int64)
    init = np.concatenate([full.init] * reps)
    assert len(clouds) >= (4 << 20)
    pinned = torch.from_numpy(clouds).pin_memory().numpy()
    piped4, res4 = full.reg.ScanMatchBatch(pinned, offsets, init)
    for r in range(reps):
        assert np.array_equal(piped4[r * full.S:(r + 1) * full.S], plain[:full.S])
        assert res4[r * full.S:(r + 1) * full.S] == res_plain[:full.S]


def test_fullsize_idempotent_at_convergence_and_oracle_spot_check(full):
    _, _, pose = full.reg.ScanMatch(full.scans[3], full.init[3], want_cloud=False)
    _, _, again = full.reg.ScanMatch(full.scans[3], pose, want_cloud=False)
    dr, dt = pose_delta(again, pose)
    assert dr < 2e-4 and dt < 2e-3
    # one full-size scan against the oracle (exact-NN mode): the north-star tolerances
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=10, eps=0.0, nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
    ref.set_target(full.map)
    rpose, _, rres, _ = ref.align(full.scans[3], full.init[3], want_cloud=False)
    dr, dt = pose_delta(pose, rpose)
    assert dr < 1e-5 and dt < 1e-4
    ok, H, B = full.reg.CaculateMatrixHAndB(full.scans[3], full.init[3])
    rok, rH, rB, _, rgate, _ = ref.compute_hb(full.scans[3], full.init[3], True, False)
    gate, _ = full.reg.DebugPoints(full.scans[3], full.init[3], 5)
    assert np.array_equal(gate, rgate)
    assert np.linalg.norm(H - rH) < 1e-6 * np.linalg.norm(rH) and np.linalg.norm(B - rB) < 1e-6 * np.linalg.norm(rB)
    assert np.allclose(H, H.T) and np.linalg.eigvalsh(H).min() > -1e-9 * np.linalg.eigvalsh(H).max()


def test_fullsize_relocalise_argmin_and_sharding(full):
    from loc_lib_b200 import dist as D
    from loc_lib_b200 import synth
    hyp = np.stack([synth.perturb_pose(full.gt[5], 500 + i, 2.0, 8.0) for i in range(96)])
    hyp[40] = full.init[5]
    hyp[71] = full.init[5]  # exact duplicate: the lower index must win
    pose, idx, score, scores, poses = full.reg.Relocalise(full.scans[5], hyp, want_all=True)
    f32 = np.float32(scores)
    assert idx == int(np.lexsort((np.arange(len(f32)), f32))[0])
    assert np.array_equal(poses[40], poses[71]) and scores[40] == scores[71]
    assert pose_delta(poses[40], full.gt[5])[1] < 0.03
    # the score (mean squared gated residual) can prefer a hypothesis that keeps fewer, tighter inliers; the
    # tie rule is checked on a truncated set whose winner is the duplicated hypothesis
    sub = np.concatenate([hyp[40:41], hyp[:8], hyp[71:72]])
    _, sidx, _, sscores, _ = full.reg.Relocalise(full.scans[5], sub, want_all=True)
    if np.float32(sscores[0]) == np.float32(sscores).min():
        assert sidx == 0 and sscores[0] == sscores[-1]
    best = None
    for rank in range(3):  # what three ranks would compute; the packed-key minimum is the all-reduce's result
        p, gi, sc, _, _ = full.reg.Relocalise(full.scans[5], hyp[rank::3])
        key = D.pack_score(sc, rank + gi * 3)
        best = key if best is None else min(best, key)
    assert D.unpack_score(best)[1] == idx


# ---------------------------------------------------------------------------------------------- config 5 at scale
def test_fullsize_relocalise_4096_hypotheses_with_oracle_spot_checks(full):
    """A 16 x 16 xy grid (0.5 m pitch) x 16 yaws around the truth = 4096 hypotheses of one scan through the wave logic of
    locreg_relocalise: argmin == numpy's over the returned scores, the winner sits at the truth, and a sample of
    hypotheses - near, far, rotated by 180 degrees - agrees with the oracle run on the same start pose."""
    gt = full.gt[7]
    g = (np.arange(16) - 7.5) * 0.5
    hyp = []
    ax, ay, az, aw = gt[:4]
    for k in range(16):
        a = k * np.pi / 8
        bz, bw = np.sin(a / 2), np.cos(a / 2)
        q = np.array([ax * bw + ay * bz, ay * bw - ax * bz, az * bw + aw * bz, aw * bw - az * bz])
        for y in g:
            for x in g:
                hyp.append(np.concatenate([q, [gt[4] + x, gt[5] + y, gt[6]]]))
    hyp = np.array(hyp)
    assert len(hyp) == 4096
    pose, idx, score, scores, poses = full.reg.Relocalise(full.scans[7], hyp, want_all=True)
    f32 = np.float32(scores)
    assert idx == int(np.lexsort((np.arange(len(f32)), f32))[0])
    assert np.array_equal(pose, poses[idx]) and score == scores[idx]
    assert idx < 256 and pose_delta(pose, gt)[1] < 0.03  # yaw 0, and at the truth
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=10, eps=0.0, nn_mode=O.NN_EXACT_TIEBREAK, skip_nonfinite=1)
    ref.set_target(full.map)
    for i in (idx, 0, 8 * 256 + 100, 4095):
        rpose, _, _, _ = ref.align(full.scans[7], hyp[i], want_cloud=False)
        _, _, _, rres, _, _ = ref.compute_hb(full.scans[7], rpose, False, False)
        dr, dt = pose_delta(poses[i], rpose)
        # the north-star gate for the winner; hypotheses that end in a wrong basin are ill-conditioned (near-singular H)
        tol = (1e-5, 1e-4) if i == idx else (1e-3, 1e-2)
        assert dr < tol[0] and dt < tol[1], (i, dr, dt)
        if i == idx:
            assert abs(scores[i] - rres["sum_sq_res"] / rres["n_inlier"]) <= 1e-5 * scores[i]


# ---------------------------------------------------------------------------------------------- config 3 at scale
@pytest.fixture(scope="module")
def ndt_full():
    import loc_lib_b200 as L
    from loc_lib_b200 import synth

    class F:
        pass
    f = F()
    f.world = synth.World(900.0)
    f.map = f.world.sample_map(20_000_000)
    f.gt = f.world.poses(3)
    f.scans = [f.world.scan(g, beams=128, azimuth=1953, seed=synth.SEED_SCAN + i) for i, g in enumerate(f.gt)]
    f.init = synth.perturb_poses(f.gt)
    f.reg = L.NdtRegistration(L.NdtOptions(max_iteration_=10, eps_=0.0))
    f.reg.SetInputTarget(f.map)
    f.ref = O.OracleNdt(max_iteration=10, eps=0.0, skip_nonfinite=1)
    f.ref.set_target(f.map)
    return f


def test_fullsize_ndt_grid_of_the_20m_point_map(ndt_full):
    """SetDirectNdtTargetCloud on BASELINE config 3's map: every voxel of the oracle's grid, bit-exact keys / counts /
    means (same summation order, no FMA), information matrices to 1e-9."""
    k, mu, info, npts = ndt_full.reg.Voxels()
    rk, rmu, rinfo, rn = ndt_full.ref.voxels()
    assert len(k) == len(rk) > 1_000_000

    def order(keys):
        return np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    a, b = order(k), order(rk)
    assert np.array_equal(k[a], rk[b]) and np.array_equal(npts[a], rn[b])
    assert np.array_equal(mu[a], rmu[b])
    scale = np.abs(rinfo[b]).reshape(len(rk), -1).max(axis=1)
    err = np.abs(info[a] - rinfo[b]).reshape(len(rk), -1).max(axis=1)
    assert np.all(err <= 1e-9 * scale)


def test_fullsize_ndt_128_beam_scan_hb_and_pose(ndt_full):
    """AlignNdt of a 128 x 1953-ray scan (~225 k points) against the 20 M-point map: per-point hit masks equal, H / B to
    1e-6, final pose to the north-star tolerances - CENTER and NEARBY6 - and the batch kernel (one CTA per scan) against
    single ScanMatch calls."""
    import loc_lib_b200 as L
    f = ndt_full
    assert 200_000 < len(f.scans[0]) < 250_000
    ok, H, B = f.reg.CaculateMatrixHAndB(f.scans[0], f.init[0])
    rH, rB, rres, rhits = f.ref.compute_hb(f.scans[0], f.init[0])
    hits, _ = f.reg.DebugPoints(f.scans[0], f.init[0], 0)
    assert np.array_equal(hits, rhits) and f.reg.last_result["n_inlier"] == rres["n_inlier"]
    assert np.linalg.norm(H - rH) < 1e-6 * np.linalg.norm(rH) and np.linalg.norm(B - rB) < 1e-6 * np.linalg.norm(rB)
    singles = []
    for i in range(3):
        _, _, pose = f.reg.ScanMatch(f.scans[i], f.init[i], want_cloud=False)
        singles.append(pose)
        if i == 0:
            rpose, _, rr, _ = f.ref.align(f.scans[0], f.init[0], want_cloud=False)
            dr, dt = pose_delta(pose, rpose)
            assert dr < 1e-5 and dt < 1e-4 and f.reg.last_result["iters"] == rr["iters"] == 10
    # k_align_batch<NdtProblem>: the same three registrations as one batch
    clouds = np.concatenate(f.scans)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in f.scans])]).astype(np.int64)
    poses, results = f.reg.ScanMatchBatch(clouds, offsets, f.init)
    for i in range(3):
        dr, dt = pose_delta(poses[i], singles[i])
        assert dr < 1e-9 and dt < 1e-9  # another summation order (one CTA per scan), the same algorithm
        assert results[i]["iters"] == 10 and results[i]["pose_written"] == 1
    # CENTER mode on the same map
    c = L.NdtRegistration(L.NdtOptions(max_iteration_=10, eps_=0.0, nearby_type_=L.NdtNearbyType.CENTER))
    c.SetInputTarget(f.map)
    cref = O.OracleNdt(max_iteration=10, eps=0.0, nearby6=0, skip_nonfinite=1)
    cref.set_target(f.map)
    _, _, pose = c.ScanMatch(f.scans[1], f.init[1], want_cloud=False)
    rpose, _, _, _ = cref.align(f.scans[1], f.init[1], want_cloud=False)
    dr, dt = pose_delta(pose, rpose)
    assert dr < 1e-5 and dt < 1e-4
