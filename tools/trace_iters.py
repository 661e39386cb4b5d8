"""Per-iteration kernel-class times of one batch ScanMatch from LOCREG_PROFILE_TRACE (classes: 0 search stage 1,
3 stage 2, 1 fit + normal equations, 2 solve):  S=512 python tools/trace_iters.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
env = dict(os.environ, LOCREG_PROFILE_TRACE="1", LOCREG_CHUNKS="1")
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "icp_breakdown.py")], env=env, capture_output=True, text=True)
rows = [l.split() for l in out.stderr.splitlines() if l.startswith("locreg-trace")]
vals = [(int(r[2]), float(r[3])) for r in rows]
names = {0: "nn", 3: "stage2", 1: "fit+post", 2: "solve"}
it, line, tot = 0, {}, 0.0
for c, ms in vals:
    line[c] = line.get(c, 0.0) + ms
    if c == 2:
        s = sum(line.values()); tot += s
        print(f"iter {it:2d}: " + "  ".join(f"{names[k]} {line.get(k, 0):.3f}" for k in (0, 3, 1, 2)) + f"  | {s:.3f} ms")
        it += 1; line = {}
print(f"sum of kernels {tot:.3f} ms")
print(out.stdout.strip())
