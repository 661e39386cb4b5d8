"""SetInputTarget wall / device time for repeated builds of the ICP search index, its size, and Loc-style re-crops.
N=1000000 LISTS=1 python tools/build_time.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import loc_lib_b200 as L
from loc_lib_b200 import synth
n = int(os.environ.get("N", "1000000")); lists = int(os.environ.get("LISTS", "1"))
w = synth.World(200.0); m = w.sample_map(n)
r = L.IcpRegistration(L.IcpOptions(method_=2, knn_lists=lists))
for i in range(4):
    t = time.perf_counter(); r.SetInputTarget(m); wall = (time.perf_counter() - t) * 1e3
    b, p, l = r.index_info()
    print(f"SetInputTarget #{i}: wall {wall:.1f} ms, device span {r.last_timing()[0]:.1f} ms | index {b/1e6:.0f} MB = {b/max(p,1):.0f} B/point, {l} lists")
r.SetGlobalMap(m)
for i in range(3):
    t = time.perf_counter(); k = r.ResetLocalMap(10.0 * i, 0, 0, half_size=(60, 60, 60)); wall = (time.perf_counter() - t) * 1e3
    print(f"ResetLocalMap #{i}: {k} pts, wall {wall:.1f} ms")
