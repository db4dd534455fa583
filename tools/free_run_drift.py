"""Development aid: relative position error between the device and the free-running oracle (same Gauss-Seidel order
injected each step, no state ever copied) every 50 steps -- how long a settling pile stays within 1e-4.
usage: python tools/free_run_drift.py boxes 16 | chains 64"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import edyn_b200 as E
from oracle import oracle as O
kind, n = sys.argv[1], int(sys.argv[2])
scene = E.scenes.boxes_on_plane(n) if kind == "boxes" else E.scenes.hinge_chains(n, n)
w = E.scenes.build_world(scene, max_manifolds=8 * len(scene["bodies"]["kind"]))
o = O.OracleWorld(vel_iters=scene["settings"]["velocity_iterations"], pos_iters=scene["settings"]["position_iterations"], threads=os.cpu_count())
o.add_bodies(scene["bodies"])
if scene["hinges"]:
    h = scene["hinges"]; o.add_hinges(h["a"], h["b"], h["pivot_a"], h["pivot_b"], h["axis_a"], h["axis_b"])
if scene["exclusions"] is not None:
    o.add_exclusions(*scene["exclusions"])
nd = scene["dynamic"]
for s in range(1000):
    w.run_phases(E.world.PH_ALL)
    hi, pr = w.solver_order()
    o.run_phases(O.PH_BROAD | O.PH_NARROW | O.PH_ISLANDS); o.set_order(hi, pr); o.run_phases(O.PH_SOLVE)
    if s % 50 == 49:
        g, c = w.download_state(aabb=False), o.state()
        err = np.abs(g["pos"][:nd] - c["pos"][:nd]).max(axis=1)
        scale = np.abs(c["pos"][:nd]).max()
        print(f"step {s + 1:4d}  max rel {err.max() / scale:.3e}  bodies beyond 1e-4: {(err / scale > 1e-4).sum():5d} / {nd}   max |v| {np.abs(c['linvel'][:nd]).max():.3f}", flush=True)
