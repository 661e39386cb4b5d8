"""Independent numpy restatement of the reference loops, used ONLY to cross-check the C++ oracle (tests/).

Written from the reference source text, not from oracle/oracle.cpp:
  * CaculateMatrixHAndBP2P       LocUtils/src/model/matching/3d/icp/icp_registration.cpp:57-103
  * CaculateMatrixHAndBP2Plane   icp_registration.cpp:161-213  (+ math::FitPlane, math_utils.h:112-136)
  * SetDirectNdtTargetCloud      LocUtils/src/model/matching/3d/ndt/ndt_registration.cpp:87-148
  * AlignNdt loop body           ndt_registration.cpp:399-433
  * BfnnRegistration             LocUtils/src/model/search_point/bfnn/bfnn.cpp:24-50  (exact k-NN definition)
numpy.linalg.svd / inv stand in for Eigen's JacobiSVD / inverse(); brute force stands in for the kd-tree.
"""
import numpy as np


def quat_R(p):
    x, y, z, w = p[:4]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], float)


def so3_exp(w):
    th = np.linalg.norm(w)
    K = hat(w)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def knn_f32(map_xyz, q_xyz, k):
    """Exact k-NN under (float32 dis2 = dx*dx + (dy*dy + dz*dz), index) ascending; -1 padded."""
    m = np.asarray(map_xyz, np.float32)
    out = np.full((len(q_xyz), k), -1, np.int32)
    for i, q in enumerate(np.asarray(q_xyz, np.float32)):
        d = q[None, :] - m
        d2 = d[:, 0] * d[:, 0] + (d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        order = np.lexsort((np.arange(len(m)), d2))[:k]
        out[i, :len(order)] = order
    return out


def fit_plane(pts, eps=1e-2):
    A = np.concatenate([np.asarray(pts, float), np.ones((len(pts), 1))], axis=1)
    _, _, vt = np.linalg.svd(A, full_matrices=False)
    c = vt[-1]
    ok = bool(np.all((A @ c) ** 2 <= eps))
    return ok, c


def p2plane_hb(map_xyz, src_xyz, pose, max_plane_distance=0.1, eps=1e-2):
    R, t = quat_R(pose), np.asarray(pose[4:], float)
    H, B = np.zeros((6, 6)), np.zeros(6)
    gates = np.zeros(len(src_xyz), np.uint8)
    n_eff = n_inl = 0
    ssq = 0.0
    m64 = np.asarray(map_xyz, np.float32).astype(np.float64)
    for i, q32 in enumerate(np.asarray(src_xyz, np.float32)):
        q = q32.astype(np.float64)
        qs = R @ q + t
        nn = knn_f32(map_xyz, qs.astype(np.float32)[None], 5)[0]
        if (nn >= 0).sum() <= 3:
            continue
        ok, c = fit_plane(m64[nn[nn >= 0]], eps)
        if not ok:
            gates[i] = 1
            continue
        n_eff += 1
        dis = c[:3] @ qs + c[3]
        if abs(dis) > max_plane_distance:
            gates[i] = 2
            continue
        J = np.concatenate([-c[:3] @ R @ hat(q), c[:3]])
        H += np.outer(J, J)
        B += -J * dis
        gates[i] = 3
        n_inl += 1
        ssq += dis * dis
    return H, B, gates, n_eff, n_inl, ssq


def p2p_hb(map_xyz, src_xyz, pose, max_nn_distance=1.0):
    R, t = quat_R(pose), np.asarray(pose[4:], float)
    H, B = np.zeros((6, 6)), np.zeros(6)
    gates = np.zeros(len(src_xyz), np.uint8)
    n_eff = 0
    m64 = np.asarray(map_xyz, np.float32).astype(np.float64)
    for i, q32 in enumerate(np.asarray(src_xyz, np.float32)):
        if not np.all(np.isfinite(q32)):
            continue
        q = q32.astype(np.float64)
        qs = R @ q + t
        nn = knn_f32(map_xyz, qs.astype(np.float32)[None], 1)[0]
        e = m64[nn[0]] - qs
        if e @ e > max_nn_distance:  # squared distance against the un-squared threshold (icp_registration.cpp:75)
            gates[i] = 2
            continue
        J = np.concatenate([R @ hat(q) / 16, -np.eye(3)], axis=1)
        H += J.T @ J
        B += -J.T @ e
        gates[i] = 3
        n_eff += 1
    return H, B, gates, n_eff


def fit_line(pts, eps):
    pts = np.asarray(pts, float)
    origin = np.zeros(3)
    for row in pts:
        origin = origin + row
    origin = origin / len(pts)
    Y = pts - origin
    _, _, vt = np.linalg.svd(Y, full_matrices=True)
    d = vt[0]
    ok = bool(np.all((np.cross(d, Y) ** 2).sum(1) <= eps))
    return ok, origin, d


def p2line_hb(map_xyz, src_xyz, pose, max_line_distance=0.5):
    """CaculateMatrixHAndBP2Line (icp_registration.cpp:105-159) + math::FitLine (math_utils.h:138-163)."""
    R, t = quat_R(pose), np.asarray(pose[4:], float)
    H, B = np.zeros((6, 6)), np.zeros(6)
    gates = np.zeros(len(src_xyz), np.uint8)
    n_eff = n_inl = 0
    m64 = np.asarray(map_xyz, np.float32).astype(np.float64)
    for i, q32 in enumerate(np.asarray(src_xyz, np.float32)):
        q = q32.astype(np.float64)
        qs = R @ q + t
        nn = knn_f32(map_xyz, qs.astype(np.float32)[None], 5)[0]
        if (nn >= 0).sum() != 5:
            continue
        ok, p0, d = fit_line(m64[nn], max_line_distance)
        if not ok:
            gates[i] = 1
            continue
        n_eff += 1
        e = hat(d) @ (qs - p0)
        if np.linalg.norm(e) > max_line_distance:
            gates[i] = 2
            continue
        J = np.concatenate([-hat(d) @ R @ hat(q), hat(d)], axis=1)
        H += J.T @ J
        B += -J.T @ e
        gates[i] = 3
        n_inl += 1
    return H, B, gates, n_eff, n_inl


def ndt_voxels(map_xyz, voxel_size=1.0, min_pts=3):
    inv = 1.0 / voxel_size
    m = np.asarray(map_xyz, np.float32).astype(np.float64)
    keys = np.trunc(m * inv).astype(np.int64)  # C++ double->int conversion truncates toward zero
    vox = {}
    for i, k in enumerate(map(tuple, keys)):
        vox.setdefault(k, []).append(i)
    out = {}
    for k, idx in vox.items():
        if len(idx) <= min_pts:
            continue
        p = m[idx]
        mu = np.zeros(3)
        for row in p:
            mu = mu + row
        mu = mu / len(idx)
        d = p - mu
        cov = d.T @ d / (len(idx) - 1)
        u, s, vt = np.linalg.svd(cov)
        s = s.copy()
        s[1] = max(s[1], 1e-3 * s[0])
        s[2] = max(s[2], 1e-3 * s[0])
        out[k] = (mu, vt.T @ np.diag(1.0 / s) @ u.T, len(idx))
    return out


NEARBY6 = [(0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1)]


def ndt_hb(vox, src_xyz, pose, voxel_size=1.0, res_outlier_th=20.0, nearby=NEARBY6):
    R, t = quat_R(pose), np.asarray(pose[4:], float)
    inv = 1.0 / voxel_size
    H, B = np.zeros((6, 6)), np.zeros(6)
    hits = np.zeros(len(src_xyz), np.uint8)
    for i, q32 in enumerate(np.asarray(src_xyz, np.float32)):
        q = q32.astype(np.float64)
        qs = R @ q + t
        key = np.trunc(qs * inv).astype(np.int64)
        J = np.concatenate([-R @ hat(q), np.eye(3)], axis=1)
        for off in nearby:
            v = vox.get(tuple(key + np.array(off)))
            if v is None:
                continue
            e = qs - v[0]
            res = e @ v[1] @ e
            if np.isnan(res) or res > res_outlier_th:
                continue
            H += J.T @ J  # the information matrix only gates (ndt_registration.cpp:426-427)
            B += -J.T @ e
            hits[i] += 1
    return H, B, hits


class IncNdtRef:
    """SetIncNdtTargetCloud / UpdateVoxel (first-scan branch, the only reachable one) / AlignIncNdt loop body
    (ndt_registration.cpp:150-236, 289-347) with an OrderedDict as the LRU list (last = most recent)."""

    def __init__(self, voxel_size=1.0, capacity=100000):
        from collections import OrderedDict
        self.inv = 1.0 / voxel_size
        self.capacity = capacity
        self.vox = OrderedDict()  # key -> dict(mu, info, pts)

    def add_cloud(self, xyz):
        m = np.asarray(xyz, np.float32).astype(np.float64)
        active = []
        for pt in m:
            key = tuple(np.trunc(pt * self.inv).astype(np.int64))
            if key not in self.vox:
                self.vox[key] = dict(mu=None, info=None, pts=[pt])
                if len(self.vox) >= self.capacity:
                    self.vox.popitem(last=False)  # evict the least recently touched voxel
            else:
                self.vox[key]["pts"].append(pt)
                self.vox.move_to_end(key)
            if key not in active:
                active.append(key)
        for key in active:
            v = self.vox.get(key)
            if v is None or not v["pts"]:
                continue
            p = np.array(v["pts"])
            if len(p) > 1:
                mu = np.zeros(3)
                for row in p:
                    mu = mu + row
                mu = mu / len(p)
                d = p - mu
                v["mu"] = mu
                v["info"] = np.linalg.inv(d.T @ d / (len(p) - 1) + 1e-3 * np.eye(3))
            else:
                v["mu"] = p[0]
                v["info"] = 1e2 * np.eye(3)
            v["n_last"] = len(p)
            v["pts"] = []

    def hb(self, src_xyz, pose, res_outlier_th=20.0, nearby=NEARBY6):
        R, t = quat_R(pose), np.asarray(pose[4:], float)
        H, B = np.zeros((6, 6)), np.zeros(6)
        hits = np.zeros(len(src_xyz), np.uint8)
        total = 0.0
        for i, q32 in enumerate(np.asarray(src_xyz, np.float32)):
            q = q32.astype(np.float64)
            qs = R @ q + t
            key = np.trunc(qs * self.inv).astype(np.int64)
            J = np.concatenate([-R @ hat(q), np.eye(3)], axis=1)
            for off in nearby:
                v = self.vox.get(tuple(key + np.array(off)))
                if v is None:
                    continue
                e = qs - v["mu"]
                res = e @ v["info"] @ e
                if np.isnan(res) or res > res_outlier_th:
                    continue
                H += J.T @ v["info"] @ J
                B += -J.T @ v["info"] @ e
                total += res
                hits[i] += 1
        return H, B, hits, total
