"""Writes tests/golden/registration_small.npz: seeded inputs + the oracle's outputs (see tests/golden_cases.py).
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import golden_cases as G  # noqa: E402
from loc_lib_b200 import synth  # noqa: E402

w = synth.World(60.0)
m = w.sample_map(24_000, pitch=0.35)
gt = w.poses(1)[0]
scan = w.scan(gt, beams=12, azimuth=160)
init = synth.perturb_pose(gt, synth.SEED_POSE, 0.25, 1.5)
out = G.compute(m, scan, init)
np.savez_compressed(os.path.join(HERE, "registration_small.npz"), map=m, scan=scan, init=init, gt=gt, **out)
print({k: v.shape for k, v in out.items()}, "map", m.shape, "scan", scan.shape)
