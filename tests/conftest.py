import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_count():
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        if lib.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        return n.value if lib.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


GPU_COUNT = _gpu_count()
HAS_GPU = GPU_COUNT > 0


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_everything():
    import __graft_entry__ as g
    g.build()


class Scene:
    """Small seeded world shared by the tests: map cloud, one scan, ground-truth and perturbed pose."""

    def __init__(self, W=80.0, n_map=120_000, beams=16, azimuth=500, n_poses=4):
        from loc_lib_b200 import synth
        self.world = synth.World(W)
        self.map = self.world.sample_map(n_map)
        self.gt = self.world.poses(n_poses)
        self.scans = [self.world.scan(g, beams=beams, azimuth=azimuth, seed=synth.SEED_SCAN + i)
                      for i, g in enumerate(self.gt)]
        self.init = synth.perturb_poses(self.gt, synth.SEED_POSE)
        self.scan = self.scans[0]


@pytest.fixture(scope="session")
def scene():
    return Scene()


def pose_delta(a, b):
    """(rotation angle [rad], translation distance [m]) between two 7-double poses."""
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    d = abs(float(np.dot(a[:4], b[:4]))) / (np.linalg.norm(a[:4]) * np.linalg.norm(b[:4]))
    return 2.0 * np.arccos(min(1.0, d)), float(np.linalg.norm(a[4:] - b[4:]))
