"""ctypes binding of the CPU oracle (oracle/oracle.h).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")

P2P, P2LINE, P2PLANE = 0, 1, 2
NN_LITERAL_ANN, NN_LITERAL_EXACT, NN_EXACT_TIEBREAK, NN_BRUTE_FORCE = 0, 1, 2, 3


class IcpOptions(C.Structure):
    _fields_ = [("max_iteration", C.c_int32), ("max_nn_distance", C.c_double), ("max_plane_distance", C.c_double),
                ("max_line_distance", C.c_double), ("min_effective_pts", C.c_int32), ("eps", C.c_double),
                ("method", C.c_int32), ("nn_mode", C.c_int32), ("skip_nonfinite", C.c_int32)]


class NdtOptions(C.Structure):
    _fields_ = [("max_iteration", C.c_int32), ("voxel_size", C.c_double), ("min_effective_pts", C.c_int32),
                ("min_pts_in_voxel", C.c_int32), ("eps", C.c_double), ("res_outlier_th", C.c_double),
                ("nearby6", C.c_int32), ("skip_nonfinite", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("iters", C.c_int32), ("updates", C.c_int32), ("converged", C.c_int32), ("degenerate", C.c_int32),
                ("n_effective", C.c_int64), ("n_inlier", C.c_int64), ("sum_sq_res", C.c_double),
                ("pose_written", C.c_int32), ("pad_", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "pad_"}


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")])
        L = C.CDLL(_SO)
        vp, sz, i32, dbl = C.c_void_p, C.c_size_t, C.c_int, C.c_double
        L.oracle_icp_default_options.argtypes = [C.POINTER(IcpOptions)]
        L.oracle_ndt_default_options.argtypes = [C.POINTER(NdtOptions)]
        L.oracle_icp_create.restype = vp
        L.oracle_icp_create.argtypes = [C.POINTER(IcpOptions)]
        L.oracle_icp_destroy.argtypes = [vp]
        L.oracle_icp_set_target.argtypes = [vp, vp, sz, sz]
        L.oracle_icp_tree_leaves.restype = sz
        L.oracle_icp_tree_leaves.argtypes = [vp]
        L.oracle_icp_knn.argtypes = [vp, vp, sz, sz, i32, i32, vp]
        L.oracle_icp_compute_hb.argtypes = [vp, vp, sz, sz, vp, vp, vp, C.POINTER(Result), vp, vp]
        L.oracle_icp_align.argtypes = [vp, vp, sz, sz, vp, vp, vp, C.POINTER(Result), vp]
        L.oracle_icp_align_batch.argtypes = [vp, vp, vp, sz, vp, sz, vp, vp, i32]
        L.oracle_fit_plane.argtypes = [vp, i32, vp, dbl]
        L.oracle_bfnn.argtypes = [vp, sz, sz, vp, sz, sz, i32, vp]
        L.oracle_ndt_create.restype = vp
        L.oracle_ndt_create.argtypes = [C.POINTER(NdtOptions)]
        L.oracle_ndt_destroy.argtypes = [vp]
        L.oracle_ndt_set_target.argtypes = [vp, vp, sz, sz]
        L.oracle_ndt_num_voxels.restype = sz
        L.oracle_ndt_num_voxels.argtypes = [vp]
        L.oracle_ndt_get_voxels.argtypes = [vp, vp, vp, vp, vp]
        L.oracle_ndt_compute_hb.argtypes = [vp, vp, sz, sz, vp, vp, vp, C.POINTER(Result), vp]
        L.oracle_ndt_align.argtypes = [vp, vp, sz, sz, vp, vp, vp, C.POINTER(Result), vp]
        L.oracle_inc_ndt_create.restype = vp
        L.oracle_inc_ndt_create.argtypes = [C.POINTER(NdtOptions), sz]
        L.oracle_inc_ndt_destroy.argtypes = [vp]
        L.oracle_inc_ndt_add_cloud.argtypes = [vp, vp, sz, sz]
        L.oracle_inc_ndt_num_voxels.restype = sz
        L.oracle_inc_ndt_num_voxels.argtypes = [vp]
        L.oracle_inc_ndt_get_voxels.argtypes = [vp, vp, vp, vp, vp]
        L.oracle_inc_ndt_compute_hb.argtypes = [vp, vp, sz, sz, vp, vp, vp, C.POINTER(Result), vp]
        L.oracle_inc_ndt_align.argtypes = [vp, vp, sz, sz, vp, vp, vp, C.POINTER(Result), vp]
        L.oracle_filter_remove_nan.restype = sz
        L.oracle_filter_remove_nan.argtypes = [vp, sz, sz, vp]
        L.oracle_filter_crop_box.restype = sz
        L.oracle_filter_crop_box.argtypes = [vp, sz, sz, vp, vp, vp]
        L.oracle_filter_voxel_grid.restype = sz
        L.oracle_filter_voxel_grid.argtypes = [vp, sz, sz, C.c_float, vp]
        L.oracle_transform_cloud.argtypes = [vp, sz, sz, vp, vp]
        L.oracle_transform_cloud_d.argtypes = [vp, sz, sz, vp, vp]
        L.oracle_pose_update.argtypes = [vp, vp]
        L.oracle_pose_matrix.argtypes = [vp, vp]
        _LIB = L
    return _LIB


def _cloud(a):
    a = np.ascontiguousarray(a, np.float32)
    assert a.ndim == 2 and a.shape[1] >= 3
    return a, a.shape[0], a.strides[0]


def icp_options(**kw):
    o = IcpOptions()
    lib().oracle_icp_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def ndt_options(**kw):
    o = NdtOptions()
    lib().oracle_ndt_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class OracleIcp:
    """IcpRegistration restated (icp_registration.cpp)."""

    def __init__(self, **opts):
        self.opt = icp_options(**opts)
        self._h = lib().oracle_icp_create(C.byref(self.opt))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_icp_destroy(self._h)
            self._h = None

    def set_target(self, cloud):
        a, n, s = _cloud(cloud)
        self._target = a
        return lib().oracle_icp_set_target(self._h, a.ctypes.data, n, s)

    def leaves(self):
        return lib().oracle_icp_tree_leaves(self._h)

    def knn(self, q, k, mode=NN_EXACT_TIEBREAK):
        a, n, s = _cloud(q)
        out = np.empty((n, k), np.int32)
        lib().oracle_icp_knn(self._h, a.ctypes.data, n, s, k, mode, out.ctypes.data)
        return out

    def compute_hb(self, src, pose7, want_gate=True, want_nn=True):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        H = np.zeros(36)
        B = np.zeros(6)
        res = Result()
        k = 1 if self.opt.method == P2P else 5
        gate = np.zeros(n, np.uint8) if want_gate else None
        nn = np.full((n, k), -1, np.int32) if want_nn else None
        ok = lib().oracle_icp_compute_hb(self._h, a.ctypes.data, n, s, pose7.ctypes.data, H.ctypes.data,
                                         B.ctypes.data, C.byref(res), gate.ctypes.data if want_gate else None,
                                         nn.ctypes.data if want_nn else None)
        return ok, H.reshape(6, 6).T.copy(), B, res.as_dict(), gate, nn

    def align(self, src, pose7, want_cloud=True):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        out_pose = np.zeros(7)
        out = np.zeros_like(a) if want_cloud else None
        res = Result()
        trace = np.zeros((self.opt.max_iteration + 1, 7))
        lib().oracle_icp_align(self._h, a.ctypes.data, n, s, pose7.ctypes.data, out_pose.ctypes.data,
                               out.ctypes.data if want_cloud else None, C.byref(res), trace.ctypes.data)
        return out_pose, out, res.as_dict(), trace


    def align_batch(self, clouds, offsets, poses, threads=0):
        a, n, s = _cloud(clouds)
        offsets = np.ascontiguousarray(offsets, np.int64)
        S = len(offsets) - 1
        pin = np.ascontiguousarray(poses, np.float64).reshape(S, 7)
        pout = np.zeros((S, 7))
        res = (Result * S)()
        used = lib().oracle_icp_align_batch(self._h, a.ctypes.data, offsets.ctypes.data, s, pin.ctypes.data, S,
                                            pout.ctypes.data, res, threads)
        return pout, [r.as_dict() for r in res], used


class OracleNdt:
    """NdtRegistration (direct) restated (ndt_registration.cpp)."""

    def __init__(self, **opts):
        self.opt = ndt_options(**opts)
        self._h = lib().oracle_ndt_create(C.byref(self.opt))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_ndt_destroy(self._h)
            self._h = None

    def set_target(self, cloud):
        a, n, s = _cloud(cloud)
        return lib().oracle_ndt_set_target(self._h, a.ctypes.data, n, s)

    def voxels(self):
        nv = lib().oracle_ndt_num_voxels(self._h)
        keys = np.zeros((nv, 3), np.int32)
        mu = np.zeros((nv, 3))
        info = np.zeros((nv, 3, 3))
        npts = np.zeros(nv, np.int32)
        lib().oracle_ndt_get_voxels(self._h, keys.ctypes.data, mu.ctypes.data, info.ctypes.data, npts.ctypes.data)
        return keys, mu, info, npts

    def compute_hb(self, src, pose7):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        H = np.zeros(36)
        B = np.zeros(6)
        res = Result()
        hits = np.zeros(n, np.uint8)
        lib().oracle_ndt_compute_hb(self._h, a.ctypes.data, n, s, pose7.ctypes.data, H.ctypes.data, B.ctypes.data,
                                    C.byref(res), hits.ctypes.data)
        return H.reshape(6, 6).T.copy(), B, res.as_dict(), hits

    def align(self, src, pose7, pose_out_init=None, want_cloud=True):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        out_pose = np.array([0, 0, 0, 1, 0, 0, 0], np.float64) if pose_out_init is None else \
            np.array(pose_out_init, np.float64)
        out = np.zeros_like(a) if want_cloud else None
        res = Result()
        trace = np.zeros((self.opt.max_iteration + 1, 7))
        lib().oracle_ndt_align(self._h, a.ctypes.data, n, s, pose7.ctypes.data, out_pose.ctypes.data,
                               out.ctypes.data if want_cloud else None, C.byref(res), trace.ctypes.data)
        return out_pose, out, res.as_dict(), trace


class OracleIncNdt:
    """NdtRegistration (INCREMENTAL_NDT) restated: set_target ADDS a cloud to the LRU voxel cache."""

    def __init__(self, capacity=100000, **opts):
        self.opt = ndt_options(**opts)
        self._h = lib().oracle_inc_ndt_create(C.byref(self.opt), capacity)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_inc_ndt_destroy(self._h)
            self._h = None

    def set_target(self, cloud):
        a, n, s = _cloud(cloud)
        return lib().oracle_inc_ndt_add_cloud(self._h, a.ctypes.data, n, s)

    def voxels(self):
        nv = lib().oracle_inc_ndt_num_voxels(self._h)
        keys = np.zeros((nv, 3), np.int32)
        mu = np.zeros((nv, 3))
        info = np.zeros((nv, 3, 3))
        npts = np.zeros(nv, np.int32)
        lib().oracle_inc_ndt_get_voxels(self._h, keys.ctypes.data, mu.ctypes.data, info.ctypes.data, npts.ctypes.data)
        return keys, mu, info, npts

    def compute_hb(self, src, pose7):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        H = np.zeros(36)
        B = np.zeros(6)
        res = Result()
        hits = np.zeros(n, np.uint8)
        lib().oracle_inc_ndt_compute_hb(self._h, a.ctypes.data, n, s, pose7.ctypes.data, H.ctypes.data, B.ctypes.data,
                                        C.byref(res), hits.ctypes.data)
        return H.reshape(6, 6).T.copy(), B, res.as_dict(), hits

    def align(self, src, pose7, want_cloud=True):
        a, n, s = _cloud(src)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        out_pose = np.zeros(7)
        out = np.zeros_like(a) if want_cloud else None
        res = Result()
        lib().oracle_inc_ndt_align(self._h, a.ctypes.data, n, s, pose7.ctypes.data, out_pose.ctypes.data,
                                   out.ctypes.data if want_cloud else None, C.byref(res), None)
        return out_pose, out, res.as_dict()


def fit_plane(pts, eps=1e-2):
    pts = np.ascontiguousarray(pts, np.float64)
    c = np.zeros(4)
    ok = lib().oracle_fit_plane(pts.ctypes.data, pts.shape[0], c.ctypes.data, eps)
    return bool(ok), c


def bfnn(map_cloud, q, k):
    m, n, ms = _cloud(map_cloud)
    qq, nq, qs = _cloud(q)
    out = np.empty((nq, k), np.int32)
    lib().oracle_bfnn(m.ctypes.data, n, ms, qq.ctypes.data, nq, qs, k, out.ctypes.data)
    return out


def filter_remove_nan(cloud):
    a, n, s = _cloud(cloud)
    out = np.zeros_like(a)
    return out[:lib().oracle_filter_remove_nan(a.ctypes.data, n, s, out.ctypes.data)]


def filter_crop_box(cloud, min3, max3):
    a, n, s = _cloud(cloud)
    out = np.zeros_like(a)
    lo, hi = np.ascontiguousarray(min3, np.float32), np.ascontiguousarray(max3, np.float32)
    return out[:lib().oracle_filter_crop_box(a.ctypes.data, n, s, lo.ctypes.data, hi.ctypes.data, out.ctypes.data)]


def filter_voxel_grid(cloud, leaf):
    a, n, s = _cloud(cloud)
    out = np.zeros_like(a)
    return out[:lib().oracle_filter_voxel_grid(a.ctypes.data, n, s, leaf, out.ctypes.data)]


def transform_cloud(src, pose7):
    a, n, s = _cloud(src)
    pose7 = np.ascontiguousarray(pose7, np.float64)
    out = np.zeros_like(a)
    lib().oracle_transform_cloud(a.ctypes.data, n, s, pose7.ctypes.data, out.ctypes.data)
    return out


def transform_cloud_d(src, pose7):
    """pcl::transformPointCloud with a double matrix (Lio's key frames): double arithmetic, one cast to float."""
    a, n, s = _cloud(src)
    pose7 = np.ascontiguousarray(pose7, np.float64)
    out = np.zeros_like(a)
    lib().oracle_transform_cloud_d(a.ctypes.data, n, s, pose7.ctypes.data, out.ctypes.data)
    return out


def pose_update(pose7, dx6):
    p = np.array(pose7, np.float64)
    d = np.ascontiguousarray(dx6, np.float64)
    lib().oracle_pose_update(p.ctypes.data, d.ctypes.data)
    return p


def pose_matrix(pose7):
    p = np.ascontiguousarray(pose7, np.float64)
    R = np.zeros((3, 3))
    lib().oracle_pose_matrix(p.ctypes.data, R.ctypes.data)
    return R
