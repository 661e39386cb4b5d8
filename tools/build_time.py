"""SetInputTarget wall / device time for repeated builds (ICP index, 1M points) and Loc-style re-crops."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L
from loc_lib_b200 import synth
w = synth.World(200.0); m = w.sample_map(1_000_000)
r = L.IcpRegistration(L.IcpOptions(method_=2))
for i in range(4):
    t = time.perf_counter(); r.SetInputTarget(m); wall = (time.perf_counter() - t) * 1e3
    print(f"SetInputTarget #{i}: wall {wall:.1f} ms, device span {r.last_timing()[0]:.1f} ms")
r.SetGlobalMap(m)
for i in range(3):
    t = time.perf_counter(); n = r.ResetLocalMap(10.0 * i, 0, 0, half_size=(60, 60, 60)); wall = (time.perf_counter() - t) * 1e3
    print(f"ResetLocalMap #{i}: {n} pts, wall {wall:.1f} ms")
