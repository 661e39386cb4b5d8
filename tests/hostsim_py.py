"""ctypes binding of tests/hostsim (serial, TEST-ONLY build of the CUDA kernels' per-element bodies)."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DIR = os.path.join(_ROOT, "tests", "hostsim")
_SO = os.path.join(_DIR, "_build", "libhostsim.so")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _DIR])
        L = C.CDLL(_SO)
        vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
        L.hs_map_create.restype = vp
        L.hs_map_create.argtypes = [vp, sz, sz, C.c_float, C.c_uint]
        L.hs_map_destroy.argtypes = [vp]
        L.hs_map_set_pyr_mode.argtypes = [vp, i32]
        L.hs_map_stats.argtypes = [vp, vp]
        L.hs_knn.argtypes = [vp, vp, sz, sz, i32, vp]
        L.hs_knn_stats.argtypes = [vp]
        L.hs_knn_seeded.argtypes = [vp, vp, vp, sz, sz, vp]
        L.hs_knn_tracked.argtypes = [vp, vp, sz, sz, i32, vp]
        L.hs_knn_tracked.restype = sz
        L.hs_icp_hb.argtypes = [vp, i32, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp, vp]
        L.hs_icp_align.restype = i32
        L.hs_icp_align.argtypes = [vp, i32, vp, vp, sz, sz, vp, vp, vp]
        L.hs_plane_svd5.argtypes = [vp, vp]
        L.hs_plane_fit5_fast.restype = i32
        L.hs_plane_fit5_fast.argtypes = [vp, vp]
        L.hs_sym3_eigen.argtypes = [vp, vp, vp]
        L.hs_gn_solve6.restype = i32
        L.hs_gn_solve6.argtypes = [vp, vp, vp]
        _LIB = L
    return _LIB


def _cloud(a):
    a = np.ascontiguousarray(a, np.float32)
    return a, a.shape[0], a.strides[0]


STAT_NAMES = ["fast_queries", "fast_candidates", "corner_queries", "corner_lists", "corner_candidates", "ring_queries",
              "ring_block_probes", "ring_candidates", "linear_scans", "skipped_searches", "pyr_queries", "pyr_probes",
              "pyr_candidates", "pyr_child_tests", "pyr_cells", "blocks_pruned"]


def knn_stats():
    """Search-stage counters accumulated since the last call (hostsim is built with -DLR_STATS)."""
    out = np.zeros(16, np.uint64)
    lib().hs_knn_stats(out.ctypes.data)
    return {k: int(out[i]) for i, k in enumerate(STAT_NAMES)}


def params(max_nn_distance=1.0, max_plane_distance=0.1, plane_fit_eps=1e-2, eps=1e-2, max_iteration=20,
           min_effective_pts=10, max_line_distance=0.5):
    return np.array([max_nn_distance, max_plane_distance, plane_fit_eps, eps, max_iteration, min_effective_pts,
                     max_line_distance], np.float64)


class HsMap:
    def __init__(self, cloud, cell=0.5, capacity_hint=0):
        a, n, s = _cloud(cloud)
        self._h = lib().hs_map_create(a.ctypes.data, n, s, cell, capacity_hint)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hs_map_destroy(self._h)
            self._h = None

    def set_pyr_mode(self, mode):
        """Stage 2 of the search: 0 = shells, 1 = block-pyramid ball query after the mid level's list, 2 = instead of it."""
        lib().hs_map_set_pyr_mode(self._h, mode)

    def stats(self):
        out = np.zeros(4, np.uint32)
        lib().hs_map_stats(self._h, out.ctypes.data)
        return dict(capacity=int(out[0]), n_pts=int(out[1]), n_unique=int(out[2]), n_cells=int(out[3]))

    def knn(self, q, k):
        a, n, s = _cloud(q)
        out = np.empty((n, k), np.int32)
        lib().hs_knn(self._h, a.ctypes.data, n, s, k, out.ctypes.data)
        return out

    def knn_tracked(self, q_steps, k):
        """k-NN of a sequence of query sets (steps, n, 3) with k_icp_nn's skip-the-scan bookkeeping carried from step to
        step; returns (idx (steps, n, k), number of searches the shortcut replaced)."""
        q = np.ascontiguousarray(q_steps, np.float32)
        steps, n, _ = q.shape
        out = np.empty((steps, n, k), np.int32)
        skipped = lib().hs_knn_tracked(self._h, q.ctypes.data, steps, n, k, out.ctypes.data)
        return out, int(skipped)

    def knn_seeded(self, q, seed_q):
        """5-NN of q, the search seeded with the 5-NN of seed_q (same shape)."""
        a, n, s = _cloud(q)
        b, nb, sb = _cloud(seed_q)
        assert (n, s) == (nb, sb)
        out = np.empty((n, 5), np.int32)
        lib().hs_knn_seeded(self._h, a.ctypes.data, b.ctypes.data, n, s, out.ctypes.data)
        return out

    def icp_hb(self, method, prm, src, pose7):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        H = np.zeros(36)
        B = np.zeros(6)
        counts = np.zeros(2, np.int64)
        ssq = np.zeros(1)
        gate = np.zeros(n, np.uint8)
        k = 1 if method == 0 else 5
        nn = np.zeros((n, k), np.int32)
        lib().hs_icp_hb(self._h, method, prm.ctypes.data, a.ctypes.data, n, s, pose7.ctypes.data, H.ctypes.data,
                        B.ctypes.data, counts.ctypes.data, ssq.ctypes.data, gate.ctypes.data, nn.ctypes.data)
        return H.reshape(6, 6).T.copy(), B, dict(n_effective=int(counts[0]), n_inlier=int(counts[1]),
                                                 sum_sq_res=float(ssq[0])), gate, nn

    def icp_align(self, method, prm, src, pose7):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        out = np.zeros(7)
        st = np.zeros(3, np.int32)
        it = lib().hs_icp_align(self._h, method, prm.ctypes.data, a.ctypes.data, n, s, pose7.ctypes.data,
                                out.ctypes.data, st.ctypes.data)
        return out, dict(iters=it, updates=int(st[0]), converged=int(st[1]), degenerate=int(st[2]))


def plane_svd5(pts):
    p = np.ascontiguousarray(pts, np.float64)
    c = np.zeros(4)
    lib().hs_plane_svd5(p.ctypes.data, c.ctypes.data)
    return c


def sym3_eigen(S6):
    s = np.ascontiguousarray(S6, np.float64)
    lam = np.zeros(3)
    Q = np.zeros((3, 3))
    lib().hs_sym3_eigen(s.ctypes.data, lam.ctypes.data, Q.ctypes.data)
    return lam, Q


def ndt_params(res_outlier_th=20.0, eps=1e-2, max_iteration=20, min_effective_pts=10, min_pts_in_voxel=3, n_nearby=7):
    return np.array([res_outlier_th, eps, max_iteration, min_effective_pts, min_pts_in_voxel, n_nearby], np.float64)


class HsNdt:
    def __init__(self, cloud, voxel_size=1.0, min_pts=3):
        L = lib()
        vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
        L.hs_ndt_create.restype = vp
        L.hs_ndt_create.argtypes = [vp, sz, sz, C.c_double, i32]
        L.hs_ndt_destroy.argtypes = [vp]
        L.hs_ndt_num_voxels.restype = sz
        L.hs_ndt_num_voxels.argtypes = [vp]
        L.hs_ndt_get_voxels.argtypes = [vp, vp, vp, vp, vp]
        L.hs_ndt_hb.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp]
        L.hs_ndt_align.restype = i32
        L.hs_ndt_align.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp]
        a, n, s = _cloud(cloud)
        self._h = L.hs_ndt_create(a.ctypes.data, n, s, voxel_size, min_pts)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hs_ndt_destroy(self._h)
            self._h = None

    def voxels(self):
        nv = lib().hs_ndt_num_voxels(self._h)
        keys = np.zeros((nv, 3), np.int32)
        mu = np.zeros((nv, 3))
        info = np.zeros((nv, 3, 3))
        npts = np.zeros(nv, np.int32)
        lib().hs_ndt_get_voxels(self._h, keys.ctypes.data, mu.ctypes.data, info.ctypes.data, npts.ctypes.data)
        return keys, mu, info, npts

    def hb(self, prm, src, pose7):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        H = np.zeros(36)
        B = np.zeros(6)
        counts = np.zeros(2, np.int64)
        ssq = np.zeros(1)
        hits = np.zeros(n, np.uint8)
        lib().hs_ndt_hb(self._h, prm.ctypes.data, a.ctypes.data, n, s, pose7.ctypes.data, H.ctypes.data,
                        B.ctypes.data, counts.ctypes.data, ssq.ctypes.data, hits.ctypes.data)
        return H.reshape(6, 6).T.copy(), B, dict(n_effective=int(counts[0]), n_inlier=int(counts[1]),
                                                 sum_sq_res=float(ssq[0])), hits

    def align(self, prm, src, pose7, pose_out_init=None):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        out = np.array([0, 0, 0, 1, 0, 0, 0], np.float64) if pose_out_init is None else np.array(pose_out_init, np.float64)
        st = np.zeros(4, np.int32)
        it = lib().hs_ndt_align(self._h, prm.ctypes.data, a.ctypes.data, n, s, pose7.ctypes.data, out.ctypes.data,
                                st.ctypes.data)
        return out, dict(iters=it, updates=int(st[0]), converged=int(st[1]), degenerate=int(st[2]),
                         pose_written=int(st[3]))


class HsIncNdt(HsNdt):
    """Incremental-NDT voxel table as ONE SetIncNdtTargetCloud call on an empty cache leaves it (no evictions): voxel
    statistics and the weighted per-point body come from the kernels' host build."""

    def __init__(self, cloud, voxel_size=1.0):
        L = lib()
        vp, sz = C.c_void_p, C.c_size_t
        L.hs_inc_ndt_create.restype = vp
        L.hs_inc_ndt_create.argtypes = [vp, vp, vp, sz, vp, sz, C.c_double]
        L.hs_inc_ndt_hb.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp, vp, vp, vp]
        L.hs_ndt_destroy.argtypes = [vp]
        L.hs_ndt_num_voxels.restype = sz
        L.hs_ndt_num_voxels.argtypes = [vp]
        L.hs_ndt_get_voxels.argtypes = [vp, vp, vp, vp, vp]
        a, n, s = _cloud(cloud)
        keys = np.trunc(a[:, :3].astype(np.float64) * (1.0 / voxel_size)).astype(np.int32)
        order = np.lexsort((np.arange(n), keys[:, 2], keys[:, 1], keys[:, 0]))  # group by key, arrival order inside
        ks = keys[order]
        first = np.ones(n, bool)
        first[1:] = (ks[1:] != ks[:-1]).any(1)
        starts = np.concatenate([np.flatnonzero(first), [n]]).astype(np.uint32)
        self._keys = np.ascontiguousarray(ks[first], np.int32)
        self._members = np.ascontiguousarray(order, np.uint32)
        self._starts = starts
        self._cloud = a
        self._h = L.hs_inc_ndt_create(self._keys.ctypes.data, starts.ctypes.data, self._members.ctypes.data, len(self._keys),
                                      a.ctypes.data, s, voxel_size)

    def hb(self, prm, src, pose7):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        H = np.zeros(36)
        B = np.zeros(6)
        counts = np.zeros(2, np.int64)
        ssq = np.zeros(1)
        hits = np.zeros(n, np.uint8)
        lib().hs_inc_ndt_hb(self._h, prm.ctypes.data, a.ctypes.data, n, s, pose7.ctypes.data, H.ctypes.data,
                            B.ctypes.data, counts.ctypes.data, ssq.ctypes.data, hits.ctypes.data)
        return H.reshape(6, 6).T.copy(), B, dict(n_effective=int(counts[0]), n_inlier=int(counts[1]),
                                                 sum_sq_res=float(ssq[0])), hits
