"""compute-sanitizer target: every kernel family on a small scene (all methods, both loop modes, batch, relocalisation,
stage-2 queues, incremental NDT, filters, the k-NN probe).  Run as
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_small.py
Sizes are tiny on purpose (the tools slow kernels down 10-100x); far-off queries are included so that the stage-2 search
and its queue append - the barrier-free code - run."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L
from loc_lib_b200 import synth

w = synth.World(60.0)
m = w.sample_map(40_000)
gt = w.poses(3)
scans = [w.scan(g, beams=8, azimuth=180, seed=synth.SEED_SCAN + i) for i, g in enumerate(gt)]
init = synth.perturb_poses(gt)
far = init[0].copy(); far[4:] += [6.0, -5.0, 1.0]  # queries outside the lists: stage 2
clouds = np.concatenate(scans)
offsets = np.concatenate([[0], np.cumsum([len(s) for s in scans])]).astype(np.int64)
done = []
for method in (L.IcpMethod.P2PLANE, L.IcpMethod.P2P, L.IcpMethod.P2LINE):
    for loop_mode in (L.LOOP_PERSISTENT, L.LOOP_GRAPH):
        r = L.IcpRegistration(L.IcpOptions(method_=method, max_iteration_=5, eps_=0.0, loop_mode=loop_mode))
        r.SetInputTarget(m)
        r.ScanMatch(scans[0], init[0])
        r.ScanMatch(scans[0], far, want_cloud=False)
        done.append(f"icp{method}/loop{loop_mode}")
    r.CaculateMatrixHAndB(scans[1], init[1])
    r.ScanMatchBatch(clouds, offsets, init)
    r.Relocalise(scans[0], np.stack([init[0], far, init[1]]), want_all=True)
    r.Knn(m[:500, :3] + 0.01, 5 if method != L.IcpMethod.P2P else 1)
    r.DebugPoints(scans[0], init[0], 5 if method != L.IcpMethod.P2P else 1)
for nearby in (L.NdtNearbyType.CENTER, L.NdtNearbyType.NEARBY6):
    for loop_mode in (L.LOOP_PERSISTENT, L.LOOP_GRAPH):
        r = L.NdtRegistration(L.NdtOptions(max_iteration_=5, eps_=0.0, nearby_type_=nearby, loop_mode=loop_mode))
        r.SetInputTarget(m)
        r.ScanMatch(scans[0], init[0])
        done.append(f"ndt{nearby}/loop{loop_mode}")
    r.ScanMatchBatch(clouds, offsets, init)
    r.Relocalise(scans[0], np.stack([init[0], far]), want_all=True)
    r.CaculateMatrixHAndB(scans[1], init[1])
r = L.NdtRegistration(L.NdtOptions(max_iteration_=5, eps_=0.0, method_=L.NdtMethod.INCREMENTAL_NDT, capacity_=3000))
for part in (m[:15_000], m[10_000:30_000], m[25_000:]):
    r.SetInputTarget(part)
r.ScanMatch(scans[0], init[0])
done.append("inc_ndt")
# the device-side LRU at a tiny capacity: voxels evicted and re-inserted within one cloud (the reuse-distance count runs)
r = L.NdtRegistration(L.NdtOptions(method_=L.NdtMethod.INCREMENTAL_NDT, capacity_=40))
rng = np.random.default_rng(0)
for c in range(3):
    k = np.repeat(rng.integers(0, 90, 700) + 20 * c, rng.integers(1, 4, 700))
    p = np.zeros((len(k), 4), np.float32)
    p[:, 0] = (k % 10) + rng.random(len(k)) * 0.9; p[:, 1] = (k // 10) + rng.random(len(k)) * 0.9; p[:, 2] = rng.random(len(k)) * 0.9
    p[::97, 1] = np.nan
    r.SetInputTarget(p)
r.Voxels()
done.append("inc_ndt_lru")
# Lio's key-frame block on the device (double-precision transform, window, voxel grid, target)
for kind in (L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=3)),
             L.NdtRegistration(L.NdtOptions(method_=L.NdtMethod.INCREMENTAL_NDT, capacity_=2000, max_iteration_=3))):
    for i in range(4):
        kind.AddKeyFrame(scans[i % 3], gt[i % 3], max_keyframes=2, leaf=0.5)
    kind.ScanMatch(scans[0], init[0], want_cloud=False)
done.append("lio_keyframes")
pts = np.concatenate([clouds, np.full((3, 4), np.nan, np.float32)])
r2 = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE))
r2.RemoveNanPoint(pts); r2.BoxFilter(clouds, clouds[:, :3].min(0) / 2, clouds[:, :3].max(0) / 2); r2.VoxelFilter(clouds, 1.0)
done.append("filters")
print("sanitize_small ok:", ", ".join(done))
