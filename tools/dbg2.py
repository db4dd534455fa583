import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, edyn_b200 as E
from oracle import oracle as O
scene = E.scenes.spheres_in_box(16, 8, 16)
w = E.scenes.build_world(scene)
o = O.OracleWorld(vel_iters=10, pos_iters=3); o.add_bodies(scene["bodies"])
def ps(p): return {tuple(x) for x in p.tolist()}
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    w.run_phases(1); o.run_phases(1)
    gp, op = ps(w.pairs()), ps(o.pairs())
    if gp != op:
        print("step", s, "pair diff: gpu-only", sorted(gp - op)[:6], "oracle-only", sorted(op - gp)[:6], len(gp), len(op))
        break
    w.run_phases(14)
    hi, pr = w.solver_order(); o.run_phases(6); o.set_order(hi, pr); o.run_phases(8)
    g = w.download_state()
    c = o.state()
    e = np.abs(g["pos"] - c["pos"]).max()
    if e > 1e-5: print("step", s, "pos err", e)
    o.set_state(g["pos"], g["orn"], g["linvel"], g["angvel"])
    gc = w.contacts(); o.set_contacts(gc["pairs"], gc["num"], gc["pts"], gc["att"], gc["lifetime"])
print("done", s, w.stats())
