"""Development aid: constraint-graph degree vs colours in use, and solver time with the persistent colouring against a
from-scratch colouring every step.   usage: python tools/color_stats.py mixed_262144"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, edyn_b200 as E
wl = sys.argv[1]
scene = bench.make_scene(wl)
nall = len(scene["bodies"]["kind"])
for flags in (0, 1):
    w = E.scenes.build_world(scene, max_manifolds=bench.capacity(wl, nall), flags=flags)
    w.step(150); w.sync(); w.reset_timers(); w.step(20); w.sync()
    st = w.stats()
    c = w.contacts()
    act = c["num"] > 0
    dyn = scene["bodies"]["kind"] == 0
    deg = np.bincount(c["pairs"][act].reshape(-1), minlength=nall)[:nall]
    deg = deg[dyn]
    print(f"{wl} flags={flags}: colours {st['contact_colors']}  max degree {deg.max()}  mean {deg.mean():.2f}  p99 {np.percentile(deg, 99):.0f}  "
          f"step {st['last_step_ms'] / 20:.3f} ms  solve {st['solve_ms']:.3f} ms  active manifolds {int(act.sum())}")
    w.close()
