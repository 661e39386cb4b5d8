"""ctypes binding of the deterministic synthetic world / LiDAR generator (include/locreg_synth.h).

The reference ships no data, so BASELINE.json's configs are defined on this
generator.  Clouds are (n, 4) float32 arrays (x, y, z, tag); poses are (7,)
float64 arrays [qx qy qz qw tx ty tz] (Sophus::SE3d::data() layout).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SEED_WORLD = 0x5EED0001
SEED_SCAN = 0x5EED0002
SEED_POSE = 0x5EED0003


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsynth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.synth_world_create.restype = C.c_void_p
        L.synth_world_create.argtypes = [C.c_double, C.c_uint64]
        L.synth_world_destroy.argtypes = [C.c_void_p]
        L.synth_world_num_boxes.restype = C.c_size_t
        L.synth_world_num_boxes.argtypes = [C.c_void_p]
        L.synth_world_sample_map.restype = C.c_size_t
        L.synth_world_sample_map.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_uint64, C.c_void_p]
        L.synth_world_scan.restype = C.c_size_t
        L.synth_world_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        L.synth_world_scan_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint64,
                                             C.c_void_p, C.c_void_p, C.c_int]
        L.synth_world_poses.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]
        L.synth_perturb_pose.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_void_p]
        _LIB = L
    return _LIB


class World:
    """Piecewise-planar world of side W metres (ground, buildings on a jittered 40 m grid, perimeter wall)."""

    def __init__(self, W=200.0, seed=SEED_WORLD):
        self._h = _lib().synth_world_create(float(W), int(seed))
        self.W = float(W)

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().synth_world_destroy(self._h)
            self._h = None

    def sample_map(self, n_map, pitch=0.2, sigma=0.01, seed=SEED_WORLD):
        out = np.empty((int(n_map), 4), np.float32)
        n = _lib().synth_world_sample_map(self._h, int(n_map), pitch, sigma, int(seed), out.ctypes.data)
        return out[:n]

    def scan(self, pose7, beams=32, azimuth=940, seed=SEED_SCAN):
        pose7 = np.ascontiguousarray(pose7, np.float64)
        out = np.empty((beams * azimuth, 4), np.float32)
        n = _lib().synth_world_scan(self._h, pose7.ctypes.data, beams, azimuth, int(seed), out.ctypes.data)
        return out[:n].copy()

    def scan_batch(self, poses7, beams=32, azimuth=940, seed=SEED_SCAN, threads=0, out=None):
        """Returns (buf (S, beams*azimuth, 4) float32, counts (S,) int32); scan s = buf[s, :counts[s]]."""
        poses7 = np.ascontiguousarray(poses7, np.float64)
        S = poses7.shape[0]
        if out is None:
            out = np.zeros((S, beams * azimuth, 4), np.float32)
        counts = np.zeros(S, np.int32)
        _lib().synth_world_scan_batch(self._h, poses7.ctypes.data, S, beams, azimuth, int(seed), out.ctypes.data,
                                      counts.ctypes.data, threads)
        return out, counts

    def poses(self, n, seed=SEED_POSE):
        out = np.empty((int(n), 7), np.float64)
        _lib().synth_world_poses(self._h, int(n), int(seed), out.ctypes.data)
        return out


def perturb_pose(gt7, seed, max_trans=0.3, max_rot_deg=2.0):
    gt7 = np.ascontiguousarray(gt7, np.float64)
    out = np.empty(7, np.float64)
    _lib().synth_perturb_pose(gt7.ctypes.data, int(seed), max_trans, np.deg2rad(max_rot_deg), out.ctypes.data)
    return out


def perturb_poses(gt, seed=SEED_POSE, max_trans=0.3, max_rot_deg=2.0):
    return np.stack([perturb_pose(g, seed + 7919 * i, max_trans, max_rot_deg) for i, g in enumerate(gt)])
